#!/bin/bash
# run on a GPU box: prune every probe of gpurun_out/tileprobe with cusparseLt TILE (and STRIP for the bf16 ones)
D=gpurun_out/tileprobe
python tests/golden/make_tile_probe.py gen $D
while read name dt; do
  oracle/_ref/cusparselt_ref prunefile 1024 1024 tile $dt $D/$name.in.bin $D/$name.tile.bin || echo "FAILED $name"
done < $D/list.txt
oracle/_ref/cusparselt_ref prunefile 1024 1024 strip bf16 $D/full_bf16.in.bin $D/full_bf16.strip.bin
oracle/_ref/cusparselt_ref prunefile 1024 1024 strip bf16 $D/rnd64_bf16.in.bin $D/rnd64_bf16.strip.bin
oracle/_ref/cusparselt_ref prunefile 1024 1024 strip f16 $D/wide.in.bin $D/wide.strip.bin
ls -la $D

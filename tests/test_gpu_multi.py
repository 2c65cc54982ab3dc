"""The path's only collective on real hardware: spfy_mg_allgather / spfy_mg_broadcast_many over NCCL, one process per
GPU (torchrun).  Needs two GPUs on the box -- skipped on a single-GPU box; the host-side partitioning logic is
covered by the world-size-2 gloo tests in tests/test_multigpu_gloo.py."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_output_gather_on_two_gpus(cuda):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "mg_worker.py")],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "MG_WORKER_OK" in r.stdout, (r.stdout + r.stderr)[-3000:]


def test_mg_entry_points_fail_cleanly_without_a_communicator(spfy, cuda):
    capi = spfy.capi
    with pytest.raises(spfy.SpfyError) as e:
        capi.spfy_mg_allgather(None, None, None, 16, None)
    assert e.value.code == capi.E_INVALID
    assert capi.spfy_mg_world(None) == 0 and capi.spfy_mg_rank(None) == -1

#!/usr/bin/env python3
"""Opcode histogram per kernel of the built library (cuobjdump -sass), tensor / TMA / TMEM opcodes first.
    python tools/sass_histogram.py > profiles/rNN_sass_histogram.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "sparsify.me_b200", "lib", "libsparsifyme_b200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "UTCOMMA", "UTCMMA", "UTCCP", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP",
        "UTMAPF", "SYNCS", "ELECT", "HMMA", "FFMA", "FFMA2", "LDS", "STS", "LDG", "STG", "ATOMG", "RED"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    cur, hist = None, collections.OrderedDict()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            hist[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            hist[cur][m.group(1).split(".")[0]] += 1
    names = subprocess.run(["c++filt"], input="\n".join(hist), capture_output=True, text=True).stdout.splitlines()
    print("# SASS opcode histogram of sparsify.me_b200/lib/libsparsifyme_b200.so (cuobjdump -sass, sm_100a)")
    print("# UTC*MMA = tcgen05.mma, UTCCP = tcgen05.cp, LDTM / STTM = tcgen05.ld / .st, UTMALDG / UTMASTG = TMA tensor load / store,")
    print("# UBLKCP = cp.async.bulk, UTMAPF = TMA prefetch, SYNCS = mbarrier, ELECT = elect.sync")
    for (fn, c), name in zip(hist.items(), names):
        name = re.sub(r"\(anonymous namespace\)::", "", name)
        name = name if len(name) <= 120 else name[:117] + "..."
        parts = [f"{k}={c[k]}" for k in KEYS if c.get(k)]
        parts += [f"{k}={v}" for k, v in c.items() if k.startswith("UTC") and k not in KEYS]
        print(f"{name}\n    total={sum(c.values())}  " + "  ".join(parts))


if __name__ == "__main__":
    main()

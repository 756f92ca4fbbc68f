"""Per-kernel SASS mnemonic counts of the built library (cuobjdump -sass): which kernels use the FP64 tensor pipe
(DMMA), cp.async (LDGSTS), TMA bulk copies (UBLKCP) and mbarriers (SYNCS).  Usage: python tools/sass_counts.py > profiles/sass_mnemonics_rNN.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "qaintensor.jl_b200", "lib", "libqaintensor_cuda.so")
KEYS = ("DMMA", "LDGSTS", "UBLKCP", "SYNCS", "DFMA", "FFMA", "LDS", "STS", "BAR", "ATOM", "RED", "LDG", "STG", "MUFU")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
cur, counts = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if cur and m:
        for k in KEYS:
            if m.group(1).startswith(k):
                counts[cur][k] += 1
                break
names = list(counts)
dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
print("# SASS mnemonic counts per kernel of lib/libqaintensor_cuda.so (cuobjdump -sass, sm_100a).")
print("# DMMA = FP64 tensor-pipe MMA (mma.sync.m8n8k4.f64; tcgen05 has no f64 kind), LDGSTS = cp.async global->shared,")
print("# UBLKCP = TMA bulk copy (cp.async.bulk), SYNCS = mbarrier ops, BAR = CTA barriers, ATOM / RED = atomics.")
for n, d in zip(names, dem):
    c = counts[n]
    d = re.sub(r"\(.*", "", d.replace("(anonymous namespace)::", "")).replace("void ", "").replace("qtn::", "")
    print("%-72s %s" % (d[:72], " ".join("%s=%d" % (k, c[k]) for k in KEYS if c[k])))

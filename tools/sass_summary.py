"""SASS instruction counts per kernel of the shipped library (cuobjdump -sass) -> profiles/<name>.  Evidence that the FP64
tensor path (DMMA), the TMA loads (UTMALDG / UBLKCP), mbarriers (SYNCS) and proxy fences are in the product binary."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "dynadjust_b200", "libgadj.so")
out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r2b_sass_summary.txt")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
KEYS = ["DMMA", "UTMALDG", "UBLKCP", "SYNCS", "FENCE.VIEW.ASYNC", "RED", "ATOM", "SHFL", "LDS", "STS", "LDG", "STG", "LDL", "STL", "BAR.SYNC", "WARPSYNC", "MEMBAR", "NANOSLEEP"]
rows = []
cur, cnt, total, k = None, None, 0, 0
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        if cur is not None:
            rows.append((cur, total, cnt))
        cur = names[k]
        k += 1
        cnt, total = collections.Counter(), 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur is not None:
        total += 1
        op = m.group(1)
        for key in KEYS:
            if op == key or op.startswith(key + ".") or (key in ("RED", "ATOM") and op.startswith(key)):
                cnt[key] += 1
                break
if cur is not None:
    rows.append((cur, total, cnt))
with open(out, "w") as f:
    f.write("SASS instruction counts per kernel of dynadjust_b200/libgadj.so (cuobjdump -sass, sm_100a; tools/sass_summary.py;\n"
            "build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3)\n"
            "DMMA = FP64 tensor-core MMA (mma.sync.m8n8k4.f64); UTMALDG = cp.async.bulk.tensor (TMA tile load); UBLKCP = cp.async.bulk (1-D bulk copy);\n"
            "SYNCS = mbarrier operations; FENCE.VIEW.ASYNC = fence.proxy.async; RED / ATOM = atomics (FP64 add in the Schur scatter and the substitutions,\n"
            "system-scope adds in the multi-GPU barrier); LDL / STL = local memory (spills).  gemm_tile_kernel<LOADER, TM, TN, consumer warps, CTAs/SM>:\n"
            "LOADER 0 = the TMA product path, 1 = the plain-load debug loader.\n\n")
    for name, total, cnt in rows:
        short = re.sub(r"\(.*", "", name.replace("(anonymous namespace)::", "").replace("<unnamed>::", "")).replace("gadj::", "").replace("void ", "")
        f.write("%-46s total %6d  %s\n" % (short, total, "  ".join("%s=%d" % (k2, cnt[k2]) for k2 in KEYS if cnt[k2])))
print(open(out).read())

"""SASS opcode histogram per kernel of a built library (cuobjdump -sass | c++filt):
python tools/sass_histogram.py [audiossl_b200/libatst_b200.so] > profiles/rNN_sass_opcode_histogram.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "audiossl_b200/libatst_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True, check=True).stdout
WATCH = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "REDG", "SHFL", "BAR.SYNC",
         "MUFU", "F2FP", "MEMBAR.ALL.GPU", "STL", "LDL"]
kernels, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kernels[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        kernels[cur]["instr"] += 1
        for w in WATCH:
            if op == w or op.startswith(w + ".") or (w == "BAR.SYNC" and op.startswith("BAR.SYNC")):
                kernels[cur][w] += 1
names = subprocess.run(["c++filt"], input="\n".join(kernels), stdout=subprocess.PIPE, text=True).stdout.splitlines()
print("SASS opcode histogram per kernel of %s (cuobjdump -sass, sm_100a)" % lib)
print("UTCHMMA = tcgen05.mma (kind::tf32), LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = cp.async.bulk.tensor load / store,")
print("UBLKCP = cp.async.bulk (1-D TMA), SYNCS = mbarrier ops, HMMA = legacy mma.sync (the N > 256 attention fallback only),")
print("F2FP = fp16 packing of the gelu' side stream, STL / LDL = register spills; gemm2_tf32_kernel<A_MN, B_MN, ECLS>: ECLS 0 generic,")
print("1 plain, 2 residual, 3 GELU without side stream, 4 fp16 gelu' forward, 5 fp16 gelu' backward\n")
for (k, c), n in zip(kernels.items(), names):
    n = re.sub(r"\(anonymous namespace\)::", "", n)
    print("%-84s instr %6d  %s" % (n[:84], c["instr"], "  ".join("%s %d" % (w, c[w]) for w in WATCH if c[w])))

"""Build the C-ABI libraries in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m audiossl_b200.build [--force] [-v] [--debug]

  libatst_b200.so          the product: TF32 tcgen05 path (include/atst_b200.h)
  libatst_b200_precise.so  validation build of the same sources with -DATST_PRECISE: producers keep fp32 and every
                           tensor-core product is an error-compensated 3xTF32 product (fp32-equivalent, ~3x the GEMM
                           time); selected at run time with audiossl_b200.set_precision("3xtf32")
  libatst_b200_debug.so    (--debug only) adds probe.cu and the bring-up entry points of include/atst_b200_debug.h
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libatst_b200.so")
SOURCES = ["err.cu", "capi.cu", "gemm_tcgen05.cu", "gemm2_tcgen05.cu", "mel.cu", "attention.cu", "attention_tc.cu",
           "attention_bwd_tc.cu", "layernorm.cu", "elementwise.cu", "augment.cu", "loss_optim.cu"]
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
         "-Xptxas=-v"]
VARIANTS = {
    "": dict(lib="libatst_b200.so", defs=[], extra=[]),
    "precise": dict(lib="libatst_b200_precise.so", defs=["-DATST_PRECISE"], extra=[]),
    "debug": dict(lib="libatst_b200_debug.so", defs=["-DATST_DEBUG_ABI"], extra=["probe.cu"]),
}
# experiment builds: python -m audiossl_b200.build --exp NAME=-DMACRO=1[,-D...] writes libatst_b200_exp_NAME.so


def lib_path(variant=""):
    return os.path.join(HERE, VARIANTS[variant]["lib"])


def _stale(lib):
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    for f in os.listdir(CSRC):
        if os.path.getmtime(os.path.join(CSRC, f)) > t:
            return True
    inc = os.path.join(HERE, "..", "include")
    return any(os.path.getmtime(os.path.join(inc, f)) > t for f in os.listdir(inc))


def build_variant(variant="", force=False, verbose=False):
    v = VARIANTS[variant]
    lib = lib_path(variant)
    if not force and not _stale(lib):
        return lib
    objdir = os.path.join(HERE, "build", variant or "default")
    os.makedirs(objdir, exist_ok=True)
    objs, procs = [], []
    for s in SOURCES + v["extra"]:
        o = os.path.join(objdir, s.replace(".cu", ".o"))
        objs.append(o)
        cmd = ["nvcc", *FLAGS, *v["defs"], "-c", os.path.join(CSRC, s), "-o", o]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed for %s (%s):\n%s" % (s, variant or "default", out))
        if verbose:
            print(out)
    r = subprocess.run(["nvcc", "-shared", "-o", lib, *objs, "-lcudart"], stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout)
    return lib


def build(force=False, verbose=False, debug=False):
    """the product library and the 3xTF32 validation build (the GPU tests use both)"""
    path = build_variant("", force, verbose)
    build_variant("precise", force, verbose)
    if debug:
        build_variant("debug", force, verbose)
    return path


if __name__ == "__main__":
    if "--exp" in sys.argv:
        name, defs = sys.argv[sys.argv.index("--exp") + 1].split("=", 1)
        VARIANTS["exp_" + name] = dict(lib="libatst_b200_exp_%s.so" % name, defs=defs.split(","), extra=[])
        print(build_variant("exp_" + name, True, "-v" in sys.argv))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, debug="--debug" in sys.argv))

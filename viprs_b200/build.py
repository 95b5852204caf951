"""
In-tree build of libviprs_b200.so (nvcc, sm_100a only).  `python -m viprs_b200.build` or
`__graft_entry__.build()`.  The .so lands in viprs_b200/_C/ (git-ignored, shipped to the GPU box).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, os.environ.get("VIPRS_B200_OUT", "_C"))     # VIPRS_B200_OUT: variant builds for A/B timing
LIB = os.path.join(OUT_DIR, "libviprs_b200.so")
SOURCES = ["ld.cu", "api.cu", "slab_f32.cu", "slab_f64.cu", "mix_f32.cu", "mix_f64.cu", "grid_f32.cu", "grid_f64.cu", "em.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _newer_than_lib():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "viprs_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a and link the C-ABI shared library."""
    if not force and not _newer_than_lib():
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "nvcc")
    extra = ["-DVB_TRACE"] if os.environ.get("VIPRS_B200_BUILD_TRACE") else []
    extra += os.environ.get("VIPRS_B200_BUILD_DEFS", "").split()
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(OUT_DIR, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((cmd, subprocess.Popen(cmd)))
        objs.append(obj)
    for cmd, p in procs:
        if p.wait() != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

"""Build libpof.so in-tree with nvcc for sm_100a (B200).

    python -m planar_optical_flow_b200.build [--force] [--verbose]

The library is a plain C-ABI shared object (include/pof.h); it does not link
against torch or Python.  nvcc cross-compiles without a GPU, so this runs in
the CPU-only build container; the resulting .so is git-ignored but travels to
the GPU box with the repo snapshot.
"""
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
INCLUDE = os.path.join(ROOT, "include")
LIB_PATH = os.path.join(PKG_DIR, "libpof.so")

SOURCES = ["pof_api.cu", "pof_cutout.cu", "pof_gate.cu", "pof_nms.cu", "pof_backbone.cu", "pof_conv_tc.cu", "pof_corr.cu", "pof_cutout_legacy.cu", "pof_bnact.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--fmad=true",            # explicit __f*_rn / __d*_rn intrinsics are used wherever contraction must not happen
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "-Xptxas", "-v",
    "--shared", "-cudart", "shared",
]


def find_nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found; libpof.so cannot be built (there is no CPU fallback)")


def _stale():
    if not os.path.isfile(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(INCLUDE, "pof.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build_variant(name, defines):
    """Tuning aid: build libpof with extra -D flags into _variants/libpof_<name>.so."""
    vdir = os.path.join(PKG_DIR, "_variants")
    os.makedirs(vdir, exist_ok=True)
    out = os.path.join(vdir, "libpof_%s.so" % name)
    cmd = [find_nvcc()] + [f for f in NVCC_FLAGS if f not in ("-Xptxas", "-v")] + ["-D%s" % d for d in defines]
    cmd += ["-I", INCLUDE, "-I", CSRC, "-o", out] + [os.path.join(CSRC, s) for s in SOURCES]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        raise RuntimeError(proc.stdout)
    return out


def build(force=False, verbose=False):
    """Compile every CUDA source of the package into libpof.so.  Returns its path.

    Safe under several processes at once (the ranks of a torchrun job importing the package together): the build is
    serialised by a lock file, staleness is re-checked under the lock, and the binary is replaced atomically."""
    if not force and not _stale():
        return LIB_PATH
    import fcntl

    with open(os.path.join(PKG_DIR, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not _stale():          # another process built it while this one waited
                return LIB_PATH
            tmp = "%s.tmp.%d" % (LIB_PATH, os.getpid())
            cmd = [find_nvcc()] + NVCC_FLAGS + ["-I", INCLUDE, "-I", CSRC, "-o", tmp]
            cmd += [os.path.join(CSRC, s) for s in SOURCES]
            proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            log = os.path.join(PKG_DIR, "build.log")
            with open(log, "w") as f:
                f.write(" ".join(cmd) + "\n" + proc.stdout)
            if proc.returncode != 0:
                sys.stderr.write(proc.stdout)
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("nvcc failed (exit %d); see %s" % (proc.returncode, log))
            os.replace(tmp, LIB_PATH)
            if verbose:
                sys.stdout.write(proc.stdout)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)

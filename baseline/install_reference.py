"""Make the UNMODIFIED reference available to `bench.py --impl reference` on the GPU box.

    python baseline/install_reference.py            (run in the build container; __graft_entry__.build() calls it)

The reference (huzjkevin/planar_optical_flow at /root/reference, read-only, absent on the GPU box) is a plain Python
source tree with an EMPTY setup.py, so the contract's

    python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref /root/reference

cannot succeed (setuptools has no metadata to build a wheel from; tried first, outcome recorded in
baseline/_ref/INSTALL_LOG.txt).  The fallback is what an install of a pure-Python package amounts to: the `src/` package and
`config/` are copied byte for byte into `baseline/_ref/` - git-ignored (nothing of the reference enters this repository's
history) but not gpurun-ignored, so it travels to the box.  `oracle/ref_shim.py` then imports it from there with the same
three environment shims it uses for /root/reference (stub matplotlib, np.int / np.float aliases, direct module import);
no reference file is edited.
"""
import filecmp
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SOURCE = os.environ.get("POF_REFERENCE_SOURCE", "/root/reference")


def install(verbose=True):
    """Returns the path of the installed tree, or None when there is no reference mount to install from."""
    if not os.path.isfile(os.path.join(SOURCE, "src", "utils", "utils.py")):
        return DEST if os.path.isfile(os.path.join(DEST, "src", "utils", "utils.py")) else None
    os.makedirs(DEST, exist_ok=True)
    log = []
    marker = os.path.join(DEST, "INSTALL_LOG.txt")
    if not os.path.isfile(marker):
        tmp = os.path.join("/tmp", "pof_ref_pip_%d" % os.getpid())
        shutil.copytree(SOURCE, tmp, dirs_exist_ok=True)            # pip wants to write build files next to setup.py
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--find-links", "/opt/wheelhouse",
               "--target", os.path.join(tmp, "_target"), tmp]
        try:
            proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
            got = os.listdir(os.path.join(tmp, "_target")) if os.path.isdir(os.path.join(tmp, "_target")) else []
            usable = any(g == "src" for g in got)
            log.append("pip install (copy under /tmp, --no-deps): exit %d, installed %s -> %s\n%s" % (
                proc.returncode, got, "usable" if usable else "NOT usable (empty setup.py: no package is declared)",
                proc.stdout[-1500:]))
        except Exception as e:      # noqa: BLE001
            log.append("pip install could not run: %r" % (e,))
        shutil.rmtree(tmp, ignore_errors=True)
    n = 0
    for sub in ("src", "config"):
        for dirpath, _, files in os.walk(os.path.join(SOURCE, sub)):
            rel = os.path.relpath(dirpath, SOURCE)
            os.makedirs(os.path.join(DEST, rel), exist_ok=True)
            for f in files:
                if f.endswith((".py", ".yaml", ".yml")):
                    a, b = os.path.join(dirpath, f), os.path.join(DEST, rel, f)
                    if not (os.path.isfile(b) and filecmp.cmp(a, b, shallow=False)):
                        shutil.copyfile(a, b)
                    n += 1
    if log:
        with open(marker, "w") as f:
            f.write("\n".join(log) + "\nfallback: %d files of src/ and config/ copied unmodified from %s\n" % (n, SOURCE))
    if verbose:
        print("reference tree available at %s (%d files)" % (DEST, n))
    return DEST


if __name__ == "__main__":
    install()

"""Recipe for oracle/_ref/: the UNMODIFIED reference (eigenfoo/littlemcmc, pure Python) installed from /root/reference
so that it can travel to the GPU box with the snapshot and be timed there as the CPU arm of bench.py
(`cpu_baseline.kind = "reference"`) and used to cross-check the oracle restatement.

    python oracle/build_ref.py          (also run by __graft_entry__.build() when /root/reference exists)

oracle/_ref/ is git-ignored (no reference source enters the history) but NOT gpurun-ignored.  The install is the
base contract's `pip install --no-index --no-build-isolation --no-deps --target ...` from a scratch copy of the tree
(/root/reference is read-only and setuptools writes an egg-info next to setup.py).  The reference imports
`fastprogress` (absent from the image, no network): our ten-line stand-in from tests/golden/_stubs is placed beside it.
TEST / BENCH INFRASTRUCTURE ONLY: nothing under littlemcmc_b200/ imports it.
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference"
DEST = os.path.join(HERE, "_ref")
STUB = os.path.join(os.path.dirname(HERE), "tests", "golden", "_stubs", "fastprogress")


def available():
    return os.path.exists(os.path.join(DEST, "littlemcmc", "__init__.py"))


def build(force=False):
    """-> path of oracle/_ref, or None when the reference tree is absent (GPU box: the prebuilt copy is used)."""
    if available() and not force:
        return DEST
    if not os.path.isdir(REF_SRC):
        return DEST if available() else None
    os.makedirs(DEST, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, "reference")
        shutil.copytree(REF_SRC, src, ignore=shutil.ignore_patterns(".git", "docs", "wheel"))
        cmd = [sys.executable, "-m", "pip", "install", "--quiet", "--no-index", "--no-build-isolation", "--no-deps",
               "--find-links", "/opt/wheelhouse", "--upgrade", "--target", DEST, src]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("pip install of the reference failed:\n%s\n%s" % (r.stdout[-2000:], r.stderr[-2000:]))
    stub_dst = os.path.join(DEST, "fastprogress")
    if os.path.isdir(stub_dst):
        shutil.rmtree(stub_dst)
    shutil.copytree(STUB, stub_dst)
    return DEST


def import_reference():
    """Import the reference package from oracle/_ref (never from the product tree).  -> module `littlemcmc`."""
    if not available():
        raise ImportError("oracle/_ref is not built (run python oracle/build_ref.py where /root/reference exists)")
    if DEST not in sys.path:
        sys.path.insert(0, DEST)
    import littlemcmc
    assert os.path.abspath(littlemcmc.__file__).startswith(DEST), littlemcmc.__file__
    return littlemcmc


if __name__ == "__main__":
    print("reference installed under", build(force="--force" in sys.argv))

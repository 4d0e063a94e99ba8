"""ORACLE (test infrastructure): snapshot of the reference's own hot-path files into `oracle/_ref/`.

The reference is pure Python, so "building" it means packing -- by this committed recipe, never by hand, and only
into the git-ignored `oracle/_ref/nf4ad_ref.tar` (it travels to the GPU box with the snapshot like a built `.so`; it
never enters the history, and it is only ever unpacked into a temporary directory outside the repo, `unpack()`).
What is taken, unmodified, from `/root/reference`:

    src/nf4ad/flows.py, transforms.py     the two in-tree files that hold hot-path arithmetic
    src/nf4ad/adbench_wrapper.py          the caller that is benchmarked (`ADBenchFlow.fit / predict_score`)
    src/nf4ad/vaeflow.py                  imported by adbench_wrapper.py at module level
    tests/conftest.py, test_flows.py, test_adbench_flow_wrapper.py, __init__.py, pytest.ini
                                          the reference's own tests for this path

Used by: `tests/test_reference_on_gpu.py` (the reference's unmodified classes and tests over the B200 drop-in),
`tests/test_oracle_golden.py` (the same over the CPU oracle shim) and `bench.py --impl reference` /
`cpu_baseline` (kind "reference": the reference's own `NonUSFlow` + `MaskedAffineCoupling` on the host cores).

    python oracle/make_ref.py            (also run by __graft_entry__.build() when /root/reference exists)
"""
import atexit
import hashlib
import io
import json
import os
import shutil
import sys
import tarfile
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("NF4AD_REFERENCE", "/root/reference")
DEST = os.path.join(HERE, "_ref")
FILES = [
    "src/nf4ad/flows.py",
    "src/nf4ad/transforms.py",
    "src/nf4ad/adbench_wrapper.py",
    "src/nf4ad/vaeflow.py",
    "tests/__init__.py",
    "tests/conftest.py",
    "tests/test_flows.py",
    "tests/test_adbench_flow_wrapper.py",
    "pytest.ini",
]


TAR = os.path.join(DEST, "nf4ad_ref.tar")
_unpacked = None


def available():
    return os.path.exists(TAR)


def make(force=False):
    """Packs the files listed above into `oracle/_ref/nf4ad_ref.tar`; returns the manifest (path -> sha256).
    No-op (returns None) without a reference checkout."""
    if not os.path.isdir(REF_ROOT):
        return None
    os.makedirs(DEST, exist_ok=True)
    manifest, blobs = {}, {}
    for rel in FILES:
        data = open(os.path.join(REF_ROOT, rel), "rb").read()
        manifest[rel] = hashlib.sha256(data).hexdigest()
        blobs[rel] = data
    mpath = os.path.join(DEST, "MANIFEST.json")
    if not force and available() and os.path.exists(mpath):
        try:
            if json.load(open(mpath)).get("sha256") == manifest:
                return manifest
        except ValueError:
            pass
    with tarfile.open(TAR, "w") as tar:
        for rel in FILES:
            info = tarfile.TarInfo(rel)
            info.size = len(blobs[rel])
            info.mtime = 0
            tar.addfile(info, io.BytesIO(blobs[rel]))
    with open(mpath, "w") as f:
        json.dump({"source": REF_ROOT, "sha256": manifest}, f, indent=1, sort_keys=True)
    return manifest


def unpack():
    """Extracts the snapshot into a fresh temporary directory (removed at interpreter exit) and returns its path:
    `<dir>/src` holds the `nf4ad` package files, `<dir>/tests` the reference's tests.  Every file is checked against
    the manifest."""
    global _unpacked
    if _unpacked is not None and os.path.isdir(_unpacked):
        return _unpacked
    if not available():
        raise FileNotFoundError(f"{TAR} is missing: run `python oracle/make_ref.py` where /root/reference exists")
    want = json.load(open(os.path.join(DEST, "MANIFEST.json")))["sha256"]
    d = tempfile.mkdtemp(prefix="nf4ad_ref_")
    atexit.register(shutil.rmtree, d, True)
    with tarfile.open(TAR) as tar:
        for m in tar.getmembers():
            if m.name not in want or not m.isfile():
                raise RuntimeError(f"unexpected member {m.name!r} in {TAR}")
            data = tar.extractfile(m).read()
            if hashlib.sha256(data).hexdigest() != want[m.name]:
                raise RuntimeError(f"{m.name} does not match the manifest")
            dst = os.path.join(d, m.name)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            with open(dst, "wb") as f:
                f.write(data)
    _unpacked = d
    return d


if __name__ == "__main__":
    m = make(force="--force" in sys.argv)
    print("no reference checkout at " + REF_ROOT if m is None else f"{len(m)} files -> {TAR}")

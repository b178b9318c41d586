"""TEST INFRASTRUCTURE -- never imported by the product path (hsimae_b200/, Models.py).

Brings the UNMODIFIED reference (Ryan21wy/HSIMAE) next to the oracle so that it can travel to the GPU box
(`/root/reference` does not exist there): copies, byte for byte,

    Models.py, Model_Pretraining.py, Model_Finetuning.py, Utils/*.py, LICENSE

from `/root/reference` into `oracle/_ref/`.  `oracle/_ref/` is git-ignored (it never enters history) and NOT
gpurun-ignored (it ships with the snapshot, like the built `.so`).  A manifest with the SHA-256 of every copied file is
written beside them; `verify()` re-checks it, so a test can assert that what ran on the GPU box is the reference as it
was mounted here.

Used by: `bench.py --impl reference` / `cpu_baseline` / `gpu_eager_baseline` (the unmodified `HSIMAE` training step as
the reported baseline), `tests/test_drivers_*.py` (the unchanged drivers against this repository's drop-in
`Models`), `tests/test_reference_gpu.py` (CUDA parity with the live reference).

    python oracle/fetch_ref.py            # copy (idempotent); exit 0 with a note when /root/reference is absent
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("HSIMAE_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ["Models.py", "Model_Pretraining.py", "Model_Finetuning.py", "LICENSE",
         "Utils/Early_Stop.py", "Utils/GroupWisePCA.py", "Utils/Label_to_Colormap.py", "Utils/Preprocessing.py",
         "Utils/Seed_Everything.py"]


def _sha(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def fetch(verbose: bool = True) -> bool:
    """copy the reference sources; returns False (and leaves any earlier copy in place) when the mount is absent"""
    if not os.path.exists(os.path.join(SRC, "Models.py")):
        if verbose:
            print(f"fetch_ref: {SRC} not present; keeping {'the existing copy' if available() else 'nothing'}")
        return False
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(SRC, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = _sha(dst)
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": "Ryan21wy/HSIMAE (unmodified copy made by oracle/fetch_ref.py)", "sha256": manifest}, f, indent=1)
    if verbose:
        print(f"fetch_ref: {len(FILES)} files -> {DST}")
    return True


def available() -> bool:
    return os.path.exists(os.path.join(DST, "Models.py")) and os.path.exists(os.path.join(DST, "MANIFEST.json"))


def verify() -> bool:
    """every copied file still has the hash recorded when it was copied (i.e. nobody edited the reference)"""
    if not available():
        return False
    want = json.load(open(os.path.join(DST, "MANIFEST.json")))["sha256"]
    return all(os.path.exists(os.path.join(DST, rel)) and _sha(os.path.join(DST, rel)) == h for rel, h in want.items())


def root() -> str | None:
    """directory to put on sys.path to import the reference: the live mount if present, else the travelling copy"""
    if os.path.exists(os.path.join(SRC, "Models.py")):
        return SRC
    return DST if available() else None


def import_models():
    """the reference's `Models` module under the name `hsimae_reference_models` (never shadows this repo's `Models`)"""
    import importlib.util
    r = root()
    if r is None:
        raise RuntimeError("the reference is neither mounted at /root/reference nor copied to oracle/_ref (run oracle/fetch_ref.py)")
    name = "hsimae_reference_models"
    if name in sys.modules:
        return sys.modules[name]
    sys.dont_write_bytecode = True
    import contextlib
    import io
    spec = importlib.util.spec_from_file_location(name, os.path.join(r, "Models.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    with contextlib.redirect_stdout(io.StringIO()):
        spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    fetch()
    sys.exit(0)

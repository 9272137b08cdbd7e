"""Vendors the UNMODIFIED reference files of the hot path into ``oracle/_ref/`` (git-ignored, travels to the GPU box
with the snapshot) so that the reference's own ``ClipLoss`` / ``gather_features`` / ``Normalize`` /
``LearnableLogitScaling`` can run where ``/root/reference`` does not exist.  TEST INFRASTRUCTURE ONLY: users are
``tests/`` (pins the oracle port against the real classes, the eager-PyTorch-on-B200 bar) and the baseline legs of
``bench.py``.  Nothing under ``oneprot_b200/`` imports it.

    python oracle/make_ref.py            # /root/reference/src/models/components/{loss,base_encoder}.py -> oracle/_ref/...

The files are byte-for-byte copies (sha256 recorded in oracle/_ref/MANIFEST.json); no reference source enters the
git history.  ``__graft_entry__.build()`` runs this when /root/reference is present."""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("ONEPROT_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ["src/models/components/loss.py",            # gather_features, ClipLoss, SigLipLoss (loss.py:19-311)
         "src/models/components/base_encoder.py"]    # Normalize, LearnableLogitScaling, BaseEncoder heads
PKGS = ["src", "src/models", "src/models/components"]


def available() -> bool:
    return all(os.path.exists(os.path.join(DST, f)) for f in FILES)


def make() -> bool:
    if not all(os.path.exists(os.path.join(REF_ROOT, f)) for f in FILES):
        return available()
    manifest = {}
    for pkg in PKGS:
        os.makedirs(os.path.join(DST, pkg), exist_ok=True)
        open(os.path.join(DST, pkg, "__init__.py"), "a").close()      # the reference's package files are empty too
    for f in FILES:
        shutil.copyfile(os.path.join(REF_ROOT, f), os.path.join(DST, f))
        manifest[f] = hashlib.sha256(open(os.path.join(DST, f), "rb").read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": REF_ROOT, "sha256": manifest}, fh, indent=1)
    return True


def import_reference():
    """-> (loss module, base_encoder module) of the vendored reference, or None when oracle/_ref is absent."""
    if not available():
        return None
    import importlib.util
    mods = []
    for name, f in (("_oneprot_ref_loss", FILES[0]), ("_oneprot_ref_base_encoder", FILES[1])):
        if name in sys.modules:
            mods.append(sys.modules[name])
            continue
        spec = importlib.util.spec_from_file_location(name, os.path.join(DST, f))
        m = importlib.util.module_from_spec(spec)
        sys.modules[name] = m
        spec.loader.exec_module(m)
        mods.append(m)
    return tuple(mods)


if __name__ == "__main__":
    ok = make()
    print("oracle/_ref", "ready" if ok else "NOT available (no /root/reference and no earlier copy)")

"""Packs the UNMODIFIED reference files of the hot path into ``oracle/_ref/oneprot_reference.zip`` (git-ignored, travels
to the GPU box with the snapshot) so that the reference's own ``ClipLoss`` / ``gather_features`` / ``Normalize`` /
``LearnableLogitScaling`` can run where ``/root/reference`` does not exist.  TEST INFRASTRUCTURE ONLY: users are
``tests/`` (pins the oracle port against the real classes, the eager-PyTorch-on-B200 bar) and the baseline legs of
``bench.py``.  Nothing under ``oneprot_b200/`` imports it.

    python oracle/make_ref.py            # /root/reference/src/models/components/{loss,base_encoder}.py -> oracle/_ref/oneprot_reference.zip

The archive members are byte-for-byte the reference files (sha256 recorded in oracle/_ref/MANIFEST.json); no reference
source enters the git history or lies loose in the tree.  ``__graft_entry__.build()`` runs this when /root/reference is present."""
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


ARCHIVE = os.path.join(DST, "oneprot_reference.zip")


def available() -> bool:
    return os.path.exists(ARCHIVE)


def make() -> bool:
    """Packs the reference files, unmodified, into oracle/_ref/oneprot_reference.zip - a built artefact like a compiled
    reference would be (no loose copy of a reference source file lies in the tree)."""
    if not all(os.path.exists(os.path.join(REF_ROOT, f)) for f in FILES):
        return available()
    import zipfile
    os.makedirs(DST, exist_ok=True)
    manifest = {}
    with zipfile.ZipFile(ARCHIVE, "w", compression=zipfile.ZIP_DEFLATED) as z:
        for f in FILES:
            data = open(os.path.join(REF_ROOT, f), "rb").read()
            z.writestr(f, data)
            manifest[f] = hashlib.sha256(data).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": REF_ROOT, "sha256": manifest}, fh, indent=1)
    # loose copies of an earlier layout of this directory
    shutil.rmtree(os.path.join(DST, "src"), ignore_errors=True)
    return True


def import_reference():
    """-> (loss module, base_encoder module) of the packed reference, or None when oracle/_ref is absent."""
    if not available():
        return None
    import types
    import zipfile
    mods = []
    with zipfile.ZipFile(ARCHIVE) as z:
        for name, f in (("_oneprot_ref_loss", FILES[0]), ("_oneprot_ref_base_encoder", FILES[1])):
            if name in sys.modules:
                mods.append(sys.modules[name])
                continue
            m = types.ModuleType(name)
            m.__file__ = ARCHIVE + "/" + f
            sys.modules[name] = m
            exec(compile(z.read(f).decode(), m.__file__, "exec"), m.__dict__)
            mods.append(m)
    return tuple(mods)


if __name__ == "__main__":
    ok = make()
    print("oracle/_ref", "ready" if ok else "NOT available (no /root/reference and no earlier archive)")

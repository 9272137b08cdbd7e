"""CPU: the C-ABI shared library builds/loads and exports every symbol include/oneprot_clip.h
declares (no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from oneprot_b200 import _lib, build
    build.build()
    return _lib.load()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "oneprot_clip.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(oneprot_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound(lib):
    from oneprot_b200 import _lib
    declared = _declared_symbols()
    assert len(declared) >= 20
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), f"{name} declared in the header but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in oneprot_b200/_lib.py"
    assert sorted(_lib.SIGNATURES) == declared


def test_ctypes_signatures_have_the_headers_argument_counts():
    """A wrong argtypes list is a silent stack mismatch on a GPU box: every prototype of the header and its ctypes
    signature must agree on the number of parameters (and on pointer vs. integer for each of them)."""
    from oneprot_b200 import _lib
    text = open(os.path.join(ROOT, "include", "oneprot_clip.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = re.findall(r"\b(oneprot_[a-z0-9_]+)\s*\(([^()]*)\)\s*;", text)
    assert len(protos) == len(_declared_symbols())
    for name, params in protos:
        params = [q.strip() for q in params.split(",")] if params.strip() not in ("", "void") else []
        argtypes = _lib.SIGNATURES[name][1]
        assert len(argtypes) == len(params), f"{name}: header has {len(params)} parameters, _lib.py binds {len(argtypes)}"
        for q, t in zip(params, argtypes):
            is_ptr_c = "*" in q
            is_ptr_py = t in (ctypes.c_void_p, ctypes.c_char_p) or hasattr(t, "contents") or getattr(t, "_type_", None) == "P"
            assert is_ptr_c == is_ptr_py, f"{name}: parameter '{q}' bound as {t}"


def test_abi_version_and_error_paths_without_gpu(lib):
    assert lib.oneprot_abi_version() == 1
    # argument validation happens before any CUDA call, so it is testable on CPU
    rc = lib.oneprot_clip_rowstats(None, None, 4, 4, 8, 0, None, None, None)
    assert rc == 1 and b"rowstats" in lib.oneprot_last_error()
    rc = lib.oneprot_gemm_bf16(None, 8, 0, None, 8, 0, 8, 8, 8, None, None, None, 8, None)
    assert rc == 1
    assert lib.oneprot_clip_fwd_scratch_bytes(32768, 32768) > 0


def test_built_for_sm100a_with_tcgen05():
    import shutil
    import subprocess
    from oneprot_b200 import _lib
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnemonic in sass, f"{mnemonic} missing: the tensor-core path is not tcgen05/TMA"
    assert "HMMA.16" not in sass   # no legacy mma.sync path

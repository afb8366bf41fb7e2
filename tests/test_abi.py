"""CPU-side checks of the C-ABI boundary: the library builds, loads, and exports exactly the
entry points include/b200vqa.h declares.  No compute is launched (no GPU here)."""
import ctypes
import os
import re

from relax_vqa_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "b200vqa.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200vqa_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(_lib.SIGNATURES) == names          # the ctypes table mirrors the header


def test_version_and_error_strings():
    lib = _lib.load()
    assert lib.b200vqa_version() >= 100
    assert lib.b200vqa_error_string(0) == b"ok"
    assert lib.b200vqa_error_string(-1) == b"invalid argument"


def test_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        return
    lib = _lib.load()
    h = ctypes.c_void_p()
    assert lib.b200vqa_create(0, ctypes.byref(h)) != 0        # no CPU fallback
    from relax_vqa_b200 import ops
    import pytest
    with pytest.raises(_lib.B200VQAError):
        ops.Context(0)

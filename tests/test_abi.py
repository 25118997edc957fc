"""CPU: libb200sat.so loads without a GPU and exports every symbol include/b200sat.h declares; the ctypes table matches."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "b200sat.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200sat_\w+)\s*\(", src)))


def test_library_exports_header_symbols():
    from b200sat import _lib
    assert os.path.exists(_lib.LIB_PATH), "run `python __graft_entry__.py` (build()) first"
    L = ctypes.CDLL(_lib.LIB_PATH)
    names = _header_functions()
    assert len(names) >= 10
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/b200sat.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes SIGNATURES and include/b200sat.h disagree"


def test_version_and_error_string_without_gpu():
    from b200sat import _lib
    L = _lib.lib()
    assert L.b200sat_version() >= 100
    assert isinstance(L.b200sat_last_error(), bytes)


def test_invalid_arguments_are_rejected_not_crashing():
    from b200sat import _lib
    L = _lib.lib()
    # null pointers -> B200SAT_EINVAL (-1), no launch attempted
    rc = L.b200sat_layernorm_fwd(None, 0, None, None, None, None, 0, 0, None, 0, 0, 0, 1e-5, None)
    assert rc == -1 and b"layernorm" in L.b200sat_last_error()
    rc = L.b200sat_attention_fwd(None, None, None, None, None, 1, 1, 1, 1, 1, *([0] * 12), 64, 0.125, None)
    assert rc == -1

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


def test_round2_entries_validate_before_touching_the_device():
    """The entries added in round 2 (batched / window weight gradients, spectrogram packing, concatenated DiT input) reject bad arguments and
    tap tables the window kernel cannot serve with the documented codes (-1 EINVAL, -2 EUNSUPPORTED) before any CUDA call is made."""
    from b200sat import _lib
    L = _lib.lib()
    buf = ctypes.create_string_buffer(64)          # a non-null host pointer: validation must fail first, nothing is ever dereferenced
    p = ctypes.cast(buf, ctypes.c_void_p)
    fp = ctypes.cast(buf, ctypes.POINTER(ctypes.c_float))

    def taps(v):
        return ctypes.cast((ctypes.c_int * len(v))(*v), ctypes.c_void_p)

    assert L.b200sat_conv_wgrad_taps_win(None, p, 128, taps([0, 1]), 2, fp, 1, None) == -1
    assert L.b200sat_conv_wgrad_taps_win(p, p, 128, taps([0, 1]), 0, fp, 1, None) == -1
    assert L.b200sat_conv_wgrad_taps_win(p, p, 128, taps([3, 1]), 2, fp, 1, None) == -2 and b"ascend" in L.b200sat_last_error()
    assert L.b200sat_conv_wgrad_taps_win(p, p, 128, taps([0, 100, 200, 300]), 4, fp, 1, None) == -2 and b"three bands" in L.b200sat_last_error()
    assert L.b200sat_conv_wgrad_taps_win(p, p, 128, taps(list(range(0, 290, 10))), 29, fp, 1, None) == -2
    assert L.b200sat_conv_wgrad_taps_win(p, p, 128, taps([5]), 1, fp, 1, None) == -2
    assert L.b200sat_conv_wgrad_taps_cat(p, p, 128, taps([0]), 33, fp, 1, None) == -1
    assert L.b200sat_disc_spec_pack(None, p, 1, 4, 17, 0, None) == -1
    assert L.b200sat_dit_concat(fp, fp, p, 1, 64, 65, 130, 16, 1, None, None, None) == -1 and b"multiple of 8" in L.b200sat_last_error()
    assert L.b200sat_dit_concat(fp, None, p, 1, 64, 65, 136, 16, 1, None, None, None) == -1

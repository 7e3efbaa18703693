"""CPU tests of the drop-in boundary: the C-ABI library builds, loads and exports every
symbol include/lens_b200.h declares (no compute calls without a GPU)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "lens_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lens_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_entry_points():
    syms = declared_symbols()
    for s in ["lens_bin_events", "lens_pool_frames", "lens_snn_create", "lens_snn_forward",
              "lens_snn_forward_float", "lens_seqmatch_topk", "lens_recall", "lens_last_error"]:
        assert s in syms


def test_library_builds_and_exports_every_declared_symbol():
    from lens_b200 import _lib, build
    path = build.build()
    assert os.path.exists(path)
    L = C.CDLL(path)
    for s in declared_symbols():
        assert hasattr(L, s), f"{s} declared in include/lens_b200.h but not exported"
    _lib.lib()
    assert sorted(_lib.EXPORTS) == declared_symbols()


def test_version_and_error_string():
    from lens_b200 import _lib
    L = _lib.lib()
    a, b = C.c_int(-1), C.c_int(-1)
    assert L.lens_version(C.byref(a), C.byref(b)) == 0
    assert (a.value, b.value) == (0, 1)
    assert isinstance(L.lens_last_error(), bytes)


def test_bad_arguments_are_rejected_without_a_gpu():
    from lens_b200 import _lib
    L = _lib.lib()
    # argument validation happens before any CUDA call
    rc = L.lens_seqmatch_topk(None, 1, 4, 4, 9, 5, None, None, None, None)
    assert rc == -1 and b"L" in L.lens_last_error()
    rc = L.lens_bin_events(None, None, None, 0, 0, 0, 0, 0, 8, 2, 1, 1, None, None, None, None, 1, None)
    assert rc == -1 and b"window_us" in L.lens_last_error()
    h = C.c_void_p()
    rc = L.lens_snn_create(0, 1, 1, 1, 1.0, -1.0, None, None, None, 1, C.byref(h), None, None)
    assert rc == -1 and not h.value


def test_product_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from lens_b200.network import B200Network
    from lens_b200._lib import LensError
    with pytest.raises(LensError):
        B200Network(torch.zeros(8, 4), torch.zeros(3, 8), roi=2, k=1, num_timesteps=5)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under lens_b200/ may import, link or load it."""
    pkg = os.path.join(ROOT, "lens_b200")
    pat = re.compile(r"(^\s*(from|import)\s+oracle\b)|liblens_oracle|oracle[/\\]|lens_oracle_", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not pat.search(txt), f"{f} references the oracle"

"""The C-ABI library: it loads without a GPU, exports every symbol that
include/*.h declares, answers its host-side layout queries, and the Python
binding fails loudly when the library or a CUDA device is missing (the product
has no CPU path)."""
import ctypes
import os
import re

import pytest
import torch

from bounded_lsq_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DECL = re.compile(r"^\s*(?:const\s+char\s*\*|int64_t|int)\s+(blsq_\w+)\s*\(", re.M)


def declared_symbols():
    names = []
    for h in ("blsq.h", "blsq_models.h"):
        names += DECL.findall(open(os.path.join(ROOT, "include", h)).read())
    return sorted(set(names))


@pytest.fixture(scope="module")
def dll():
    if not os.path.exists(L.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return ctypes.CDLL(L.LIB_PATH)


def test_every_declared_symbol_is_exported(dll):
    names = declared_symbols()
    assert len(names) >= 25, names
    missing = [n for n in names if not hasattr(dll, n)]
    assert not missing, missing


def test_binding_table_matches_header():
    """_lib.SIGNATURES (ctypes argtypes) covers the int-returning entry points."""
    names = set(declared_symbols())
    assert set(L.SIGNATURES) <= names, set(L.SIGNATURES) - names


def test_host_side_queries(dll):
    assert dll.blsq_version() >= 100
    lib = L.Lib()
    for method in (L.METHOD_TRF, L.METHOD_DOGBOX):
        for n in range(1, L.MAX_BATCHED_N + 1):
            lay = lib.state_layout(method, n)
            assert lay["size"] > 3 * n and lay["x"] == 0 and lay["x_new"] >= n
            if method == L.METHOD_TRF:
                # every block of the record is 32-byte aligned (256-bit moves)
                assert lay["size"] % 4 == 0 and lay["x_new"] % 4 == 0
            assert lib.lin_record_size(n) % 4 == 0
    with pytest.raises(L.BlsqError):
        lib.state_layout(L.METHOD_TRF, L.MAX_BATCHED_N + 1)
    for n in (10, 16, 64, 100, 256):
        t = lib.tall_layout(n)
        assert t["state_size"] > n * n and t["fac_size"] > 3 * n * n
        assert t["record"] >= n * n + 2 * n + 1 and t["record"] % 2 == 0
        assert t["rinvp"] % 2 == 0            # 16-byte aligned inside fac
    dll.blsq_error_string.restype = ctypes.c_char_p
    assert dll.blsq_error_string(-1)


def test_missing_library_is_loud(tmp_path):
    with pytest.raises(L.BlsqError, match="no CPU fallback"):
        L.Lib(str(tmp_path / "nope.so"))


def test_cpu_tensors_are_rejected():
    """No CPU path: host tensors never reach a kernel."""
    from bounded_lsq_b200 import least_squares, least_squares_batched
    if torch.cuda.is_available():
        pytest.skip("GPU box: covered by the parity suite")
    f = lambda x: x - 1.0                                      # noqa: E731
    with pytest.raises(L.BlsqError):
        least_squares(f, [2.0], bounds=(0.0, 3.0))
    with pytest.raises(L.BlsqError):
        least_squares_batched(f, torch.ones((4, 2), dtype=torch.float64))

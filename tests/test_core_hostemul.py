"""CPU suite: the kernels' per-problem C++ (bounded_lsq_b200/csrc/blsq_core.cuh)
compiled for the host by tests/host_emul and driven through the same Python
front end, against the golden vectors of the unmodified reference.

This checks the host logic and the branch structure of the device code in the
GPU-less container.  It is NOT the parity proof -- tests/test_gpu_parity.py
runs the same cases through the CUDA library on a B200."""
import pytest
import torch

import cases
import hostemul


@pytest.fixture(scope="module")
def lib():
    return hostemul.get()


DEV = torch.device("cpu")


def test_helpers_bit_exact(lib):
    assert cases.check_helpers_bit_exact(lib, DEV) > 50


def test_helpers_batched_rows(lib):
    cases.check_helpers_batched_rows(lib, DEV)


@pytest.mark.parametrize("name", ["c2_trf_exact", "c2_dogbox_exact"])
def test_golden_exact_jac(lib, name):
    cases.check_golden_exact_jac(lib, DEV, name)


@pytest.mark.parametrize("name", ["c3_dogbox_2point", "c3_trf_2point",
                                  "c3_trf_3point", "c3_dogbox_3point"])
def test_golden_fd_jac(lib, name):
    cases.check_golden_fd_jac(lib, DEV, name)


@pytest.mark.parametrize("name", ["rat_trf_2point", "rat_dogbox_2point",
                                  "rat_trf_3point", "rat_dogbox_3point"])
def test_golden_fd_exact(lib, name):
    print(cases.check_golden_fd_exact(lib, DEV, name))


def test_fd_linearise_bit_exact(lib):
    cases.check_fd_linearise_bit_exact(lib, DEV)


def test_edge_cases_vs_oracle(lib):
    print(cases.check_edge_cases_vs_oracle(lib, DEV))


def test_tall_edge_cases_vs_oracle(lib):
    print(cases.check_tall_edge_cases_vs_oracle(lib, DEV))


# the C5 family (n = 128 ... 256, half of the bounds active); the n = 256
# dogbox cases take minutes on the single-threaded host emulation and run on
# the GPU only
@pytest.mark.parametrize("tag,method", [("f", "trf"), ("f", "dogbox"), ("e", "trf")])
def test_c5_golden(lib, tag, method):
    print(cases.check_tall_golden(lib, DEV, tag, method, file="c5.npz"))


def test_compaction_invariance(lib):
    cases.check_compaction_invariance(lib, DEV)


def test_prologue_invariance(lib):
    cases.check_prologue_invariance(lib, DEV)


def test_per_problem_bounds(lib):
    cases.check_per_problem_bounds(lib, DEV)


def test_corpus_single(lib):
    st = cases.check_corpus_single(lib, DEV)
    print(st)
    assert st["total"] == 406 and st["exact_status"] >= 190


@pytest.mark.parametrize("tag,method", [("a", "trf"), ("a", "dogbox"),
                                        ("c", "trf"), ("c", "dogbox"),
                                        ("d", "trf"), ("d", "dogbox")])
def test_tall_golden(lib, tag, method):
    print(cases.check_tall_golden(lib, torch.device("cpu"), tag, method))


def test_tall_options_vs_oracle(lib):
    print(cases.check_tall_options_vs_oracle(lib, torch.device("cpu")))


def test_benchmark_table(lib, tmp_path):
    st = cases.check_benchmark_table(lib, DEV, tmp_path / "table.txt")
    print(st)
    assert st["instances"] == 58 and st["checked"] >= 130


def test_x_covariance(lib):
    cases.check_x_covariance(lib, DEV)


def test_mode_routing(lib):
    cases.check_mode_routing(lib, DEV)


def test_chunked_batch(lib):
    cases.check_chunked_batch(lib, DEV)


def test_compact_batched(lib):
    cases.check_compact_batched(lib, DEV)


def test_random_small_vs_oracle(lib):
    print(cases.check_random_small_vs_oracle(lib, DEV))


def test_random_tall_vs_oracle(lib):
    print(cases.check_random_tall_vs_oracle(lib, DEV))

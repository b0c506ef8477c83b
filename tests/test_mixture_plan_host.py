"""Host-side planner of the tensor-core hyper-prior kernels (csrc/mixture_tc.cu: pick_splits): the component chunks of a
256-query tile are split into work units so that the persistent grid's makespan `rounds x chunks_per_unit` is minimal, with at
most 4x the split count of the old "2 units per SM" rule.  The workspace-size entry points run without a device (148 SMs are
assumed), so the rule is checked here against a Python restatement and against its properties."""
import ctypes

import pytest

from ladder_latent_data_distribution_modelling_b200 import lib

SMS, BN, QROWS = 148, 128, 256


def ceil_div(a, b):
    return -(-a // b)


def plan(N, K):
    row_tiles, n_chunks = max(1, ceil_div(N, QROWS)), ceil_div(K, BN)
    s0 = min(max(ceil_div(2 * SMS, row_tiles), 1), n_chunks)
    best = None
    for sp in range(1, min(4 * s0, n_chunks) + 1):
        cps = ceil_div(n_chunks, sp)
        se = ceil_div(n_chunks, cps)
        cost = ceil_div(row_tiles * se, SMS) * cps * 64 + se
        if best is None or cost < best[0]:
            best = (cost, se, cps)
    return best[1], best[2], row_tiles, n_chunks


@pytest.fixture(scope='module')
def L():
    L = lib.load()
    L.ladder_mixture_tc_workspace_bytes.restype = ctypes.c_size_t
    L.ladder_mixture_tc_workspace_bytes.argtypes = [ctypes.c_longlong, ctypes.c_int]
    L.ladder_mixture_tc_grad_workspace_bytes.restype = ctypes.c_size_t
    L.ladder_mixture_tc_grad_workspace_bytes.argtypes = [ctypes.c_longlong, ctypes.c_int, ctypes.c_int]
    return L


@pytest.mark.parametrize('N,K', [(65536, 65536), (6400, 50), (1, 1), (768, 5120), (1000, 300), (256 * 148, 128), (100000, 4096)])
def test_workspace_follows_the_makespan_rule(L, N, K):
    splits, cps, row_tiles, n_chunks = plan(N, K)
    assert L.ladder_mixture_tc_workspace_bytes(N, K) == splits * N * 4 + 256
    for D in (32, 64):
        assert L.ladder_mixture_tc_grad_workspace_bytes(N, K, D) == splits * N * (1 + D) * 4 + 256
    assert 1 <= splits <= n_chunks and splits * cps >= n_chunks and (splits - 1) * cps < n_chunks


def test_bench_size_needs_seven_rounds_not_eight():
    """65 536 x 65 536: 256 query tiles x 512 chunks.  Two units per SM (512 units) run 4 rounds of 256 chunks = 1024 chunk-times;
    the chosen 4-way split runs 7 rounds of 128 = 896 (ideal 885.6)."""
    splits, cps, row_tiles, n_chunks = plan(65536, 65536)
    assert (splits, cps) == (4, 128)
    assert ceil_div(row_tiles * splits, SMS) * cps == 896
    assert ceil_div(row_tiles * 2, SMS) * 256 == 1024

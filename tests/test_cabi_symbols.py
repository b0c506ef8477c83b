"""CPU-only: the C-ABI library builds/loads and exports every symbol include/ladder_sm100.h declares
(no compute calls are made without a GPU)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT

PKG = os.path.join(ROOT, 'ladder_latent_data_distribution_modelling_b200')


@pytest.fixture(scope='module')
def libpath():
    from ladder_latent_data_distribution_modelling_b200.build import build
    return build(verbose=False)


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'ladder_sm100.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(ladder_[a-z0-9_]+)\s*\(', text)))


def test_header_symbols_exported(libpath):
    lib = ctypes.CDLL(libpath)
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), 'missing export: ' + n


def test_binding_covers_header(libpath):
    from ladder_latent_data_distribution_modelling_b200 import lib as L
    L.load()
    assert sorted(L.SIGNATURES) == declared_symbols()


def test_host_side_entry_points(libpath):
    """Pure host functions can run without a GPU: version, table stride, packing + error text."""
    import numpy as np
    from ladder_latent_data_distribution_modelling_b200 import lib as L
    lib = L.load()
    assert lib.ladder_version() >= 100
    assert lib.ladder_mixture_table_stride(2, 0) == 4
    assert lib.ladder_mixture_table_stride(2, 1) == 8
    assert lib.ladder_mixture_table_stride(2, 2) == 8
    mean = np.zeros((2, 2)); cov = np.tile(np.eye(2)[None], (2, 1, 1)); w = np.ones(2)
    cov[1] = [[1.0, 2.0], [2.0, 1.0]]           # not positive definite
    table = np.zeros((2, 8), np.float32); ref = ctypes.c_float()
    rc = lib.ladder_mixture_pack_full(mean.ctypes.data_as(L.c_double_p), cov.ctypes.data_as(L.c_double_p),
                                      w.ctypes.data_as(L.c_double_p), 2, 2, table.ctypes.data_as(L.c_float_p),
                                      ctypes.byref(ref))
    assert rc == -1 and b'positive definite' in lib.ladder_last_error()


def test_pack_full_matches_oracle_canonical_form(libpath, golden_dir):
    import numpy as np
    from ladder_latent_data_distribution_modelling_b200 import lib as L
    from oracle import mixture as OM
    lib = L.load()
    d = np.load(os.path.join(golden_dir, 'gm_prior_golden.npz'))
    m, K, w = d['m_full'], d['K_full'], d['w_full']
    table = np.zeros((50, 8), np.float32); ref = ctypes.c_float()
    rc = lib.ladder_mixture_pack_full(np.ascontiguousarray(m).ctypes.data_as(L.c_double_p),
                                      np.ascontiguousarray(K).ctypes.data_as(L.c_double_p),
                                      np.ascontiguousarray(w).ctypes.data_as(L.c_double_p), 50, 2,
                                      table.ctypes.data_as(L.c_float_p), ctypes.byref(ref))
    assert rc == 0
    mu, A, c = OM.canonical_from_full(m, K, w)
    sc = np.sqrt(0.5 * np.log2(np.e))
    np.testing.assert_allclose(table[:, 0], sc * A[:, 0, 0], rtol=1e-6)
    np.testing.assert_allclose(table[:, 1], sc * A[:, 1, 0], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(table[:, 2], sc * A[:, 1, 1], rtol=1e-6)
    np.testing.assert_allclose(table[:, 3:5], -sc * np.einsum('kij,kj->ki', A, mu), rtol=1e-5, atol=1e-6)
    c2 = c * np.log2(np.e)
    np.testing.assert_allclose(ref.value, c2.max(), rtol=1e-6)
    np.testing.assert_allclose(table[:, 5], c2 - c2.max(), rtol=1e-5, atol=1e-5)


def test_product_does_not_import_oracle():
    """The product path must never route through the oracle (tier rule 3)."""
    bad = []
    for base in (PKG, os.path.join(ROOT, 'codes')):
        for dp, _, files in os.walk(base):
            for f in files:
                if f.endswith('.py'):
                    src = open(os.path.join(dp, f)).read()
                    if re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M):
                        bad.append(os.path.join(dp, f))
    for f in ('train.py',):
        p = os.path.join(ROOT, f)
        if os.path.exists(p) and re.search(r'^\s*(from|import)\s+oracle\b', open(p).read(), flags=re.M):
            bad.append(p)
    assert not bad, bad

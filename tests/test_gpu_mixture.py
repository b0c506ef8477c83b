"""GPU parity of the fused mixture kernel (K9) against the oracle and the golden fixture,
called through the C ABI (ops -> ctypes -> libladder_sm100.so)."""
import os

import numpy as np
import pytest
import torch

from oracle import mixture as OM

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ops():
    from ladder_latent_data_distribution_modelling_b200 import ops
    return ops


@pytest.fixture(scope='module')
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, 'gm_prior_golden.npz'))


def _dev(a):
    return torch.tensor(np.asarray(a, dtype=np.float32), device='cuda')


@pytest.mark.parametrize('tag', ['full', 'active'])
def test_reference_fixture_full_cov(ops, gold, tag):
    """The reference's fitted 50-component hyper-prior incl. far outliers (rescue path)."""
    t = gold['t_' + tag]
    tab = ops.mixture_pack_full(gold['m_' + tag], gold['K_' + tag], gold['w_' + tag], 'cuda')
    lp, g = ops.mixture_logprob(_dev(t), tab, want_grad=True)
    mu, A, c = OM.canonical_from_full(gold['m_' + tag], gold['K_' + tag], gold['w_' + tag])
    t32 = t.astype(np.float32).astype(np.float64)
    ref, gref = OM.mixture_logprob(t32, mu, A, c, with_grad=True)
    lp, g = lp.cpu().numpy(), g.cpu().numpy()
    assert np.all(np.isfinite(lp))
    # tolerance: fp32 arithmetic + ex2.approx (2 ulp); |logp| reaches 6e3 for the far outliers
    np.testing.assert_allclose(lp, ref, rtol=2e-5, atol=2e-4)
    np.testing.assert_allclose(lp, gold['logp_sklearn_' + tag], rtol=5e-5, atol=5e-4)
    scale = np.abs(gref).max(axis=1, keepdims=True) + 1e-3      # per-row relative error
    err = np.abs(g - gref) / scale
    assert err.max() < 2e-3, (err.max(), np.argmax(err.max(axis=1)), t[np.argmax(err.max(axis=1))])


@pytest.mark.parametrize('D', [1, 2, 3, 4, 8, 16, 32, 64])
@pytest.mark.parametrize('mode', ['iso', 'diag'])
def test_diag_iso_all_dims(ops, D, mode):
    rng = np.random.default_rng(D)
    K, N = 77, 1000
    m = rng.normal(size=(K, D)); t = rng.normal(size=(N, D)).astype(np.float32)
    w = rng.uniform(0.1, 1.0, size=K)
    if mode == 'iso':
        std = 0.8
        tab = ops.mixture_pack_diag(m, std, w, 'cuda')
    else:
        std = rng.uniform(0.5, 1.5, size=(K, D))
        tab = ops.mixture_pack_diag(m, std, w, 'cuda')
    lp, g = ops.mixture_logprob(_dev(t), tab, want_grad=True, exact=True)   # the fp32 SIMT kernel (D = 32/64 iso would go to tcgen05)
    lp_only = ops.mixture_logprob(_dev(t), tab, exact=True)
    mu, A, c = OM.canonical_from_diag(m, std, w)
    ref, gref = OM.mixture_logprob(t.astype(np.float64), mu, A, c, with_grad=True)
    np.testing.assert_allclose(lp.cpu().numpy(), ref, rtol=1e-5, atol=1e-4 * max(1, D / 8))
    np.testing.assert_allclose(lp_only.cpu().numpy(), lp.cpu().numpy(), rtol=0, atol=0)
    np.testing.assert_allclose(g.cpu().numpy(), gref, rtol=2e-3, atol=2e-3)


def test_full_cov_other_dims(ops):
    rng = np.random.default_rng(0)
    for D in (1, 3, 4, 8, 16):        # 8 / 16: the z-space mixture of the "GMM" prior branch (D = code_size)
        K, N = 13, 500
        m = rng.normal(size=(K, D)); a = rng.normal(size=(K, D, D))
        cov = a @ a.transpose(0, 2, 1) + 0.2 * np.eye(D)
        w = rng.uniform(0.1, 1, size=K)
        t = rng.normal(size=(N, D)).astype(np.float32)
        tab = ops.mixture_pack_full(m, cov, w, 'cuda')
        lp, g = ops.mixture_logprob(_dev(t), tab, want_grad=True)
        mu, A, c = OM.canonical_from_full(m, cov, w)
        ref, gref = OM.mixture_logprob(t.astype(np.float64), mu, A, c, with_grad=True)
        np.testing.assert_allclose(lp.cpu().numpy(), ref, rtol=2e-5, atol=2e-4)
        np.testing.assert_allclose(g.cpu().numpy(), gref, rtol=2e-3, atol=2e-3)


def test_split_components_large_k_and_ragged_sizes(ops):
    """K large enough to be split over CTAs (deterministic partial reduce) and N not a tile multiple."""
    rng = np.random.default_rng(3)
    for N, K in ((1, 1), (129, 5000), (4097, 20011)):
        m = rng.normal(size=(K, 2)) * 2; t = (rng.normal(size=(N, 2)) * 2).astype(np.float32)
        tab = ops.mixture_pack_diag(m, 0.5, None, 'cuda')
        lp1, g1 = ops.mixture_logprob(_dev(t), tab, want_grad=True)
        lp2, g2 = ops.mixture_logprob(_dev(t), tab, want_grad=True)
        assert torch.equal(lp1, lp2) and torch.equal(g1, g2)          # bitwise reproducible
        mu, A, c = OM.canonical_from_diag(m, 0.5)
        ref, gref = OM.mixture_logprob(t.astype(np.float64), mu, A, c, with_grad=True)
        np.testing.assert_allclose(lp1.cpu().numpy(), ref, rtol=1e-5, atol=1e-4)
        np.testing.assert_allclose(g1.cpu().numpy(), gref, rtol=2e-3, atol=2e-3)


def test_empty_input(ops):
    tab = ops.mixture_pack_diag(np.zeros((3, 2)), 1.0, None, 'cuda')
    lp = ops.mixture_logprob(torch.empty(0, 2, device='cuda'), tab)
    assert lp.shape == (0,)


def test_far_queries_stay_finite(ops):
    """Queries far from every component underflow the fixed frame and must be rescued exactly."""
    rng = np.random.default_rng(5)
    m = rng.normal(size=(50, 2)); t = (rng.normal(size=(300, 2)) * 200).astype(np.float32)
    tab = ops.mixture_pack_diag(m, 0.3, None, 'cuda')
    lp, g = ops.mixture_logprob(_dev(t), tab, want_grad=True)
    mu, A, c = OM.canonical_from_diag(m, 0.3)
    ref, gref = OM.mixture_logprob(t.astype(np.float64), mu, A, c, with_grad=True)
    assert torch.isfinite(lp).all() and torch.isfinite(g).all()
    np.testing.assert_allclose(lp.cpu().numpy(), ref, rtol=1e-5)
    np.testing.assert_allclose(g.cpu().numpy(), gref, rtol=1e-3, atol=1e-2)


def test_component_sharded_partials_combine(ops, gold):
    """ranks hold K/P components; (m, s, g) partials combine to the full answer."""
    t = _dev(gold['t_full'])
    tab = ops.mixture_pack_full(gold['m_full'], gold['K_full'], gold['w_full'], 'cuda')
    full, gfull = ops.mixture_logprob(t, tab, want_grad=True)
    for P in (2, 4, 8):
        parts = [ops.mixture_logprob(t, tab.shard(r, P), want_grad=True, partial=True) for r in range(P)]
        lp, g = ops.mixture_combine(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]),
                                    torch.stack([p[2] for p in parts]))
        np.testing.assert_allclose(lp.cpu().numpy(), full.cpu().numpy(), rtol=1e-5, atol=1e-4)
        np.testing.assert_allclose(g.cpu().numpy(), gfull.cpu().numpy(), rtol=1e-3, atol=1e-3)


def test_full_size_properties(ops):
    """BASELINE microbench size (65 536 x 65 536, D=2): size-independent checks -- a subsample
    against the oracle, and the K copies of one component == that component identity."""
    rng = np.random.default_rng(1234)
    N = K = 65536
    t = rng.normal(size=(N, 2)).astype(np.float32); m = rng.normal(size=(K, 2))
    tab = ops.mixture_pack_diag(m, 1.0, None, 'cuda')
    lp, g = ops.mixture_logprob(_dev(t), tab, want_grad=True)
    idx = rng.choice(N, 256, replace=False)
    mu, A, c = OM.canonical_from_diag(m, 1.0)
    ref, gref = OM.mixture_logprob(t[idx].astype(np.float64), mu, A, c, with_grad=True)
    np.testing.assert_allclose(lp.cpu().numpy()[idx], ref, rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(g.cpu().numpy()[idx], gref, rtol=2e-3, atol=2e-3)
    same = ops.mixture_pack_diag(np.tile(m[:1], (K, 1)), 1.0, None, 'cuda')
    lp1 = ops.mixture_logprob(_dev(t), same).cpu().numpy()
    want = -0.5 * ((t - m[:1]) ** 2).sum(1) - np.log(2 * np.pi)
    # 65 536 identical addends: fp32 accumulation error is systematic here (~n*eps/4), hence 2e-3
    np.testing.assert_allclose(lp1, want, rtol=1e-5, atol=2e-3)


@pytest.mark.parametrize('D', [32, 64])
def test_tensor_core_path_iso(ops, D):
    """tcgen05 (tf32) forward path for isotropic D in {32, 64}: against the float64 oracle and the exact fp32
    SIMT kernel.  Stated tolerance: |d logp| <= 5e-2 for unit-scale data with
    |logp| ~ 1e2 (tf32 rounds the operands of the cross term to 11 bits)."""
    rng = np.random.default_rng(D)
    for N, K in ((1000, 777), (4096, 4096), (257, 129)):
        m = rng.normal(size=(K, D)); t = rng.normal(size=(N, D)).astype(np.float32)
        w = rng.uniform(0.1, 1.0, size=K)
        tab = ops.mixture_pack_diag(m, 0.9, w, 'cuda')
        assert tab.tc_image is not None
        lp_tc = ops.mixture_logprob(_dev(t), tab)
        lp_exact = ops.mixture_logprob(_dev(t), tab, exact=True)
        mu, A, c = OM.canonical_from_diag(m, 0.9, w)
        ref = OM.mixture_logprob(t.astype(np.float64), mu, A, c)
        assert torch.isfinite(lp_tc).all()
        np.testing.assert_allclose(lp_exact.cpu().numpy(), ref, rtol=1e-5, atol=2e-3)
        err = np.abs(lp_tc.cpu().numpy() - ref).max()
        assert err < 5e-2, err
        lp2 = ops.mixture_logprob(_dev(t), tab)
        assert torch.equal(lp_tc, lp2)                               # deterministic


@pytest.mark.parametrize('D', [32, 64])
def test_tensor_core_forward_and_gradient(ops, D):
    """tcgen05 forward + gradient for isotropic D in {32, 64}: scores t.mu^T on kind::tf32, the exponentials written back into
    TMEM in place of the scores, and the gradient's contraction W.mu as a second kind::tf32 MMA whose A operand is that
    TMEM-resident W and whose B operand is the same component tile read MN-major -- against the float64 oracle and the exact
    fp32 SIMT kernel.  Stated tolerance: |d logp| <= 5e-2 as for the forward kernel; the gradient (= -(t - sum_k p_k mu_k) /
    sigma^2, both operands of the second contraction rounded to tf32's 11 bits) within 5e-3 of its largest entry per call
    (measured 3.1e-3 / 3.9e-3 at D = 32 / 64), 2e-3 relative L2."""
    rng = np.random.default_rng(100 + D)
    for N, K in ((1000, 777), (4096, 4096), (257, 129), (300, 5000)):
        m = rng.normal(size=(K, D)); t = rng.normal(size=(N, D)).astype(np.float32)
        w = rng.uniform(0.1, 1.0, size=K)
        tab = ops.mixture_pack_diag(m, 0.9, w, 'cuda')
        lp_tc, g_tc = ops.mixture_logprob(_dev(t), tab, want_grad=True)
        lp_ex, g_ex = ops.mixture_logprob(_dev(t), tab, want_grad=True, exact=True)
        if N * K <= 2_000_000:
            mu, A, c = OM.canonical_from_diag(m, 0.9, w)
            ref, gref = OM.mixture_logprob(t.astype(np.float64), mu, A, c, with_grad=True, chunk=128)
            np.testing.assert_allclose(g_ex.cpu().numpy(), gref, rtol=2e-3, atol=2e-3)
        else:                        # the float64 oracle materialises [N, K, D]: at this size the exact fp32 kernel is the checker
            ref, gref = lp_ex.cpu().numpy().astype(np.float64), g_ex.cpu().numpy().astype(np.float64)
        assert torch.isfinite(lp_tc).all() and torch.isfinite(g_tc).all()
        assert np.abs(lp_tc.cpu().numpy() - ref).max() < 5e-2
        g = g_tc.cpu().numpy().astype(np.float64)
        assert np.abs(g - gref).max() <= 5e-3 * np.abs(gref).max(), np.abs(g - gref).max() / np.abs(gref).max()
        assert np.linalg.norm(g - gref) <= 2e-3 * np.linalg.norm(gref), np.linalg.norm(g - gref) / np.linalg.norm(gref)
        lp2, g2 = ops.mixture_logprob(_dev(t), tab, want_grad=True)
        assert torch.equal(lp_tc, lp2) and torch.equal(g_tc, g2)             # deterministic (split partials summed in order)
    # far queries: the fixed-frame sum underflows, the finalising kernel recomputes log p and the gradient exactly
    t = (rng.normal(size=(64, D)) * 30).astype(np.float32)
    m = rng.normal(size=(300, D))
    tab = ops.mixture_pack_diag(m, 0.5, None, 'cuda')
    lp, g = ops.mixture_logprob(_dev(t), tab, want_grad=True)
    mu, A, c = OM.canonical_from_diag(m, 0.5)
    ref, gref = OM.mixture_logprob(t.astype(np.float64), mu, A, c, with_grad=True)
    np.testing.assert_allclose(lp.cpu().numpy(), ref, rtol=2e-5)
    np.testing.assert_allclose(g.cpu().numpy(), gref, rtol=1e-3, atol=1e-3 * np.abs(gref).max())


def test_tensor_core_path_far_queries_rescued(ops):
    rng = np.random.default_rng(9)
    D, K, N = 32, 300, 512
    m = rng.normal(size=(K, D)); t = (rng.normal(size=(N, D)) * 30).astype(np.float32)
    tab = ops.mixture_pack_diag(m, 0.5, None, 'cuda')
    lp = ops.mixture_logprob(_dev(t), tab)
    mu, A, c = OM.canonical_from_diag(m, 0.5)
    ref = OM.mixture_logprob(t.astype(np.float64), mu, A, c)
    assert torch.isfinite(lp).all()
    np.testing.assert_allclose(lp.cpu().numpy(), ref, rtol=2e-5)


@pytest.mark.parametrize('D', [2, 8, 16])
def test_vamp_prior_device_pack_and_param_grads(ops, D):
    """VampPrior mixture (base.py:241-254): table packed on the device from device mean/std, log p and d/dt through
    the device frame, and d/d mean, d/d std of coef * sum_n log p(t_n) against the float64 oracle."""
    from oracle import tape as T
    rng = np.random.default_rng(D)
    K, N = 11, 700
    mean = rng.normal(size=(K, D)); std = rng.uniform(0.3, 1.5, size=(K, D))
    t = rng.normal(size=(N, D)) * 1.5
    tv, mv, sv = T.Var(t), T.Var(mean), T.Var(std)
    lp = OM.diag_mixture_logprob_var(tv, mv, sv)
    coef = -1.0 / N
    T.backward(T.reduce_sum(lp) * coef)
    md, sd, td = _dev(mean.astype(np.float32)), _dev(std.astype(np.float32)), _dev(t.astype(np.float32))
    tab = ops.mixture_pack_diag_device(md, sd)
    logp, g = ops.mixture_logprob(td, tab, want_grad=True)
    np.testing.assert_allclose(logp.cpu().numpy(), lp.v, rtol=2e-5, atol=2e-4)
    np.testing.assert_allclose(g.cpu().numpy() * coef, tv.g, rtol=2e-3, atol=2e-3 / N)
    host = ops.mixture_pack_diag(mean, std, None, 'cuda')           # same table as the host (double) packer
    np.testing.assert_allclose(tab.table.cpu().numpy(), host.table.cpu().numpy(), rtol=1e-5, atol=1e-5)
    dm, ds = torch.full((K, D), 7.0, device='cuda'), torch.full((K, D), 7.0, device='cuda')
    ops.mixture_diag_param_grad(td, md, sd, logp, coef, dm, ds)
    np.testing.assert_allclose(dm.cpu().numpy(), mv.g, rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(ds.cpu().numpy(), sv.g, rtol=1e-3, atol=1e-5)
    tab2 = ops.mixture_pack_diag_device(md * 0 + 1, sd, tab)         # repack in place reuses the table
    assert tab2 is tab


@pytest.mark.parametrize('D,mode', [(2, 'iso'), (2, 'full'), (16, 'diag'), (32, 'iso')])
@pytest.mark.parametrize('want_grad', [False, True])
def test_packed_shard_partials_combine_to_the_full_answer(D, mode, want_grad):
    """One-exchange form of the component-sharded evaluation with the exchange done in place on one GPU: the
    packed (m, s, g) partials of 3 component shards, stacked rank-major like the all-gather delivers them, combine to the
    unsharded answer; `ShardedMixture` (world 1, CUDA-graph replay) returns the same."""
    import torch
    from ladder_latent_data_distribution_modelling_b200 import ops, parallel
    rng = np.random.default_rng(17 + D)
    K, N = 301, 1000
    mean = rng.normal(size=(K, D)) * 1.5
    if mode == 'full':
        a = rng.normal(size=(K, D, D))
        tab = ops.mixture_pack_full(mean, a @ a.transpose(0, 2, 1) * 0.3 + 0.1 * np.eye(D), rng.uniform(0.1, 1, size=K), 'cuda')
    elif mode == 'diag':
        tab = ops.mixture_pack_diag(mean, rng.uniform(0.5, 1.5, size=(K, D)), rng.uniform(0.1, 1, size=K), 'cuda')
    else:
        tab = ops.mixture_pack_diag(mean, 0.8, None, 'cuda')
    t = torch.tensor(rng.normal(size=(N, D)).astype(np.float32) * 1.5, device='cuda')
    t[:5] += 200.0                                              # far queries: the exact rescue path of every shard
    full = ops.mixture_logprob(t, tab, want_grad=want_grad, exact=True)
    W = 2 + D if want_grad else 2
    parts = torch.stack([ops.mixture_logprob_packed(t, tab.shard(r, 3), torch.empty(N, W, device='cuda'), want_grad)
                         for r in range(3)])
    got = ops.mixture_combine_packed(parts, D, want_grad)
    lp, lp_full = (got[0], full[0]) if want_grad else (got, full)
    np.testing.assert_allclose(lp.cpu().numpy(), lp_full.cpu().numpy(), rtol=2e-5, atol=2e-4)
    if want_grad:
        np.testing.assert_allclose(got[1].cpu().numpy(), full[1].cpu().numpy(), rtol=2e-3, atol=2e-3)
    sm = parallel.ShardedMixture(tab, N, want_grad=want_grad)
    for _ in range(2):                                          # capture, then replay
        out = sm(t)
    lp2 = out[0] if want_grad else out
    # world 1 keeps the whole table, tensor-core images included: isotropic D = 32 evaluates on kind::tf32, whose cross-term
    # error grows with |t||mu| -- both are scaled by 1.5 here: |d logp| <= 0.1 (measured 0.07; 5e-2 for unit-scale data)
    tc = mode == 'iso' and D in (32, 64)
    np.testing.assert_allclose(lp2.cpu().numpy(), lp_full.cpu().numpy(), rtol=2e-5, atol=0.1 if tc else 2e-4)
    assert sm.use_graph and sm._graph is not None


@pytest.mark.parametrize('D', [32, 64])
@pytest.mark.parametrize('want_grad', [False, True])
def test_tensor_core_packed_shard_partials(D, want_grad):
    """Component shards of an isotropic D in {32, 64} mixture on the tcgen05 kernels: a shard that starts on a 128-component
    chunk keeps its slice of the operand images (MixtureTable.shard), its kernel emits the packed (m, s, g) row, and the
    combine of all shards equals the unsharded exact answer within the tensor-core kernels' stated tolerance (|d logp| <= 5e-2,
    gradient 5e-3 of its largest entry / 2e-3 relative L2).  K = 383 in 3 shards leaves a ragged last chunk; far queries take
    every shard's exact rescue path."""
    import torch
    from ladder_latent_data_distribution_modelling_b200 import ops
    rng = np.random.default_rng(300 + D)
    for K, world, N in ((383, 3, 1000), (4096, 4, 2048), (1024, 8, 300)):
        mean = rng.normal(size=(K, D))
        tab = ops.mixture_pack_diag(mean, 0.9, rng.uniform(0.1, 1.0, size=K), 'cuda')
        t = torch.tensor(rng.normal(size=(N, D)).astype(np.float32), device='cuda')
        t[:4] *= 40.0
        shards = [tab.shard(r, world) for r in range(world)]
        assert all((sh.tc_image_grad if want_grad else sh.tc_image) is not None for sh in shards)
        assert sum(sh.K for sh in shards) == K
        W = 2 + D if want_grad else 2
        parts = torch.stack([ops.mixture_logprob_packed(t, sh, torch.empty(N, W, device='cuda'), want_grad) for sh in shards])
        got = ops.mixture_combine_packed(parts, D, want_grad)
        full = ops.mixture_logprob(t, tab, want_grad=want_grad, exact=True)
        lp, lp_full = (got[0], full[0]) if want_grad else (got, full)
        assert torch.isfinite(lp).all()
        assert (lp - lp_full).abs().max().item() < 5e-2
        np.testing.assert_allclose(lp[:4].cpu().numpy(), lp_full[:4].cpu().numpy(), rtol=2e-5)      # rescued rows are exact
        if want_grad:
            g, gref = got[1].double().cpu().numpy(), full[1].double().cpu().numpy()
            assert np.isfinite(g).all()
            assert np.abs(g[4:] - gref[4:]).max() <= 5e-3 * np.abs(gref[4:]).max()
            assert np.linalg.norm(g - gref) <= 2e-3 * np.linalg.norm(gref)
        # the exact=True switch keeps the fp32 SIMT kernel
        ex = ops.mixture_combine_packed(torch.stack([ops.mixture_logprob_packed(t, sh, torch.empty(N, W, device='cuda'), want_grad,
                                                                                 exact=True) for sh in shards]), D, want_grad)
        np.testing.assert_allclose((ex[0] if want_grad else ex).cpu().numpy(), lp_full.cpu().numpy(), rtol=2e-5, atol=2e-4)


@pytest.mark.parametrize('D', [32, 64, 128, 256])
def test_full_cov_large_dims(ops, D):
    """Full-covariance mixture at the CelebA code sizes (prior "GMM", base.py:323-329 with code_size 128 / 256): the SGEMM-tiled
    kernel of csrc/mixture_bigd.cu against the float64 oracle.  N is not a multiple of the 64-row tile; one far query makes every
    responsibility but one vanish.  fp32 throughout: |log p| reaches ~1e3 here, hence rtol 2e-5 (+ 2e-3 absolute)."""
    rng = np.random.default_rng(D)
    K, N = 7, 333
    m = rng.normal(size=(K, D)); a = rng.normal(size=(K, D, D))
    cov = a @ a.transpose(0, 2, 1) / D + 0.3 * np.eye(D)
    w = rng.uniform(0.1, 1, size=K)
    t = (rng.normal(size=(N, D)) * 1.2).astype(np.float32)
    t[0] = m[3] + 25.0
    tab = ops.mixture_pack_full(m, cov, w, 'cuda')
    assert tab.mode == ops.MODE_FULL_BIGD
    lp, g = ops.mixture_logprob(_dev(t), tab, want_grad=True)
    lp_only = ops.mixture_logprob(_dev(t), tab)
    mu, A, c = OM.canonical_from_full(m, cov, w)
    ref, gref = OM.mixture_logprob(t.astype(np.float64), mu, A, c, with_grad=True, chunk=64)
    assert torch.equal(lp, lp_only)
    np.testing.assert_allclose(lp.cpu().numpy(), ref, rtol=2e-5, atol=2e-3)
    gg = g.cpu().numpy()
    assert np.abs(gg - gref).max() <= 2e-4 * np.abs(gref).max() + 2e-3, np.abs(gg - gref).max()
    lp2, g2 = ops.mixture_logprob(_dev(t), tab, want_grad=True)
    assert torch.equal(lp, lp2) and torch.equal(g, g2)                    # deterministic: no atomics
    with pytest.raises(RuntimeError):
        ops.mixture_logprob(_dev(t), tab, partial=True)


def test_full_cov_unsupported_dim_raises(ops):
    rng = np.random.default_rng(0)
    with pytest.raises(RuntimeError):
        ops.mixture_pack_full(rng.normal(size=(3, 24)), np.tile(np.eye(24), (3, 1, 1)), np.ones(3), 'cuda')


@pytest.mark.parametrize('D', [128, 256, 40])
def test_vamp_prior_large_dims(ops, D):
    """VampPrior mixture at the CelebA code sizes (base.py:241-254 with code_size 128 / 256; D = 40 checks a dimension that is
    not a lane multiple): log p, d/dt, and d/d mean, d/d std of coef * sum_n log p(t_n) from csrc/mixture_bigd.cu's diagonal
    kernels against the float64 oracle."""
    from oracle import tape as T
    rng = np.random.default_rng(D)
    K, N = 9, 301
    mean = rng.normal(size=(K, D)); std = rng.uniform(0.5, 1.5, size=(K, D))
    t = rng.normal(size=(N, D)) * 1.2
    tv, mv, sv = T.Var(t), T.Var(mean), T.Var(std)
    lp = OM.diag_mixture_logprob_var(tv, mv, sv)
    coef = -1.0 / N
    T.backward(T.reduce_sum(lp) * coef)
    md, sd, td = _dev(mean.astype(np.float32)), _dev(std.astype(np.float32)), _dev(t.astype(np.float32))
    logp, g, resp = torch.empty(N, device='cuda'), torch.empty(N, D, device='cuda'), torch.empty(N, K, device='cuda')
    ops.mixture_diag_bigd(td, md, sd, logp, g, resp)
    np.testing.assert_allclose(logp.cpu().numpy(), lp.v, rtol=2e-5, atol=2e-3)
    np.testing.assert_allclose(resp.sum(1).cpu().numpy(), np.ones(N), rtol=1e-4)
    np.testing.assert_allclose(g.cpu().numpy() * coef, tv.g, rtol=2e-3, atol=2e-3 / N)
    dm, ds = torch.full((K, D), 7.0, device='cuda'), torch.full((K, D), 7.0, device='cuda')
    ops.mixture_diag_bigd_param_grad(td, md, sd, resp, coef, dm, ds)
    np.testing.assert_allclose(dm.cpu().numpy(), mv.g, rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(ds.cpu().numpy(), sv.g, rtol=1e-3, atol=2e-5)

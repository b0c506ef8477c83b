"""K8: the in-kernel Philox noise (ladder_philox_normal) against its NumPy restatement, its data-parallel invariance (a rank
draws its rows of the GLOBAL noise tensor) and its device-resident draw counter (graph replays advance like eager launches)."""
import numpy as np
import pytest
import torch

import philox_ref

pytestmark = pytest.mark.gpu


def _draw(B, Bg, off, seed, ctr, C=8, R=2, L=5):
    from ladder_latent_data_distribution_modelling_b200 import ops
    z = torch.empty(B, C, device='cuda'); t = torch.empty(B, R, device='cuda'); mc = torch.empty(L, B, R, device='cuda')
    c = torch.tensor([ctr], dtype=torch.int32, device='cuda')
    ops.philox_normal([z, t, mc], B, Bg, off, seed, c)
    return z.cpu().numpy(), t.cpu().numpy(), mc.cpu().numpy()


def test_matches_numpy_restatement():
    seed = 0x1234567890ABCDEF
    z, t, mc = _draw(6, 6, 0, seed, 3)
    for got, seg in ((z, 0), (t, 1), (mc, 2)):
        want = philox_ref.normal(got.shape, 6, 0, seed, 3, seg)
        np.testing.assert_allclose(got, want, rtol=2e-5, atol=2e-6)


def test_rank_rows_are_rows_of_the_global_draw():
    seed = 77
    zg, tg, mg = _draw(8, 8, 0, seed, 1)
    for r in range(2):
        z, t, mc = _draw(4, 8, 4 * r, seed, 1)
        assert np.array_equal(z, zg[4 * r:4 * r + 4]) and np.array_equal(t, tg[4 * r:4 * r + 4])
        assert np.array_equal(mc, mg[:, 4 * r:4 * r + 4])


def test_counter_and_segments_decorrelate_and_moments():
    a = _draw(256, 256, 0, 5, 1, C=64, R=2, L=100)
    b = _draw(256, 256, 0, 5, 2, C=64, R=2, L=100)
    assert not np.array_equal(a[0], b[0])
    x = np.concatenate([a[0].ravel(), a[2].ravel(), b[0].ravel(), b[2].ravel()])
    assert abs(x.mean()) < 0.02 and abs(x.std() - 1) < 0.02 and abs((x ** 4).mean() - 3) < 0.15
    assert abs(np.corrcoef(a[0].ravel(), b[0].ravel())[0, 1]) < 0.03
    assert abs(np.corrcoef(a[2][:, :, 0].ravel(), a[2][:, :, 1].ravel())[0, 1]) < 0.03


def test_engine_noise_skips_unrequested_tensors():
    from conftest import load_config
    from ladder_latent_data_distribution_modelling_b200.engine import LadderEngine
    cfg = load_config('mnist_digit', batch_size=4, n_MC_samples=3, compute_dtype='fp32')
    eng = LadderEngine(cfg, 4, 'cuda', seed=9)
    eng.draw_noise()
    z1, t1, m1 = eng.eps_z.clone(), eng.eps_t.clone(), eng.eps_mc.clone()
    eng.draw_noise(z=True, t=False, mc=False)
    assert not torch.equal(eng.eps_z, z1) and torch.equal(eng.eps_t, t1) and torch.equal(eng.eps_mc, m1)
    assert int(eng.noise_ctr.item()) == 2
    want = philox_ref.normal((4, cfg['code_size']), 4, 0, 9, 2, 0)
    np.testing.assert_allclose(eng.eps_z.cpu().numpy(), want, rtol=2e-5, atol=2e-6)

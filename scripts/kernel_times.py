"""Per-kernel device time of the timed loop of bench.py (CUPTI through torch.profiler; concurrent, warm caches -- the
quick companion of the serialised ncu launch list).  python scripts/kernel_times.py [workload] [batch] [iterations]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from ladder_latent_data_distribution_modelling_b200.engine import LadderEngine  # noqa: E402


def main():
    workload = sys.argv[1] if len(sys.argv) > 1 else 'mnist_fashion'
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    bench.WORKLOAD = workload
    cfg = bench.load_config(B)
    cfg['seed'] = 1234
    if len(sys.argv) > 4:
        cfg['code_size'] = int(sys.argv[4])
    dev = torch.device('cuda', 0)
    eng = LadderEngine(cfg, B, dev, seed=1234)
    gm = bench.synthetic_mixture(cfg['n_mixtures'], cfg['representation_size'])
    eng.set_feeds(prior_mean=gm[0], prior_cov=gm[1], prior_weight=gm[2], use_standard_gaussian_prior=False, use_mask=False)
    eng.set_lrs(cfg['learning_rate_ae'], cfg['learning_rate_sigma'], cfg['learning_rate_prior'], cfg['learning_rate_inner_sigma'])
    x = torch.rand(B, *bench.image_shape(cfg), device=dev)

    def iteration():
        for name in ('ae', 'sigma', 'prior', 'inner_sigma'):
            eng.run_step(name, x)
    for _ in range(3):
        iteration()
    torch.cuda.synchronize()
    if os.environ.get('LADDER_NCU') == '1':      # under `ncu --profile-from-start off`: exactly one iteration is captured
        torch.cuda.profiler.start()
        iteration()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(iters):
            iteration()
        torch.cuda.synchronize()
    rows = []
    for e in prof.key_averages():
        t = getattr(e, 'device_time_total', None)
        if t is None:
            t = getattr(e, 'cuda_time_total', 0.0)
        if t > 0:
            rows.append((t / iters, e.count / iters, e.key[:90]))
    rows.sort(reverse=True)
    tot = sum(r[0] for r in rows)
    print('%s B=%d: %.1f us of kernel time per iteration, %d launches' % (workload, B, tot, sum(r[1] for r in rows)))
    for t, n, k in rows[:45]:
        print('%9.1f us %6.1f %5.1f%%  %s' % (t, n, 100 * t / tot, k))


if __name__ == '__main__':
    main()

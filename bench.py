"""Benchmark of the LaDDer ELBO hot path on B200 (contract: one JSON line on rank 0).

  python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
  python bench.py --impl reference ...                      # CPU restatement of the reference (oracle port)

A "step" is one reference training iteration (codes/trainers.py:33-40): train_step_ae,
train_step_sigma, train_step_prior, train_step_inner_sigma -- four forward passes with fresh
noise and two backward passes -- on one batch of synthetic 28x28x1 images.  Workload at N=1:
BASELINE.json configs[1], `codes/mnist_fashion_config.json` at batch 1024 (weak scaling:
batch 1024 PER GPU for N>1, gradients / batch sums all-reduced over NCCL).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'train imgs/sec (ELBO fwd+bwd, 4 sub-steps per image batch)'


def ncu_traffic(key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of a roofline kernel, read from THIS round's committed ncu
    summary (profiles/ncu_traffic.json: {key: {"bytes": ..., "source": "profiles/<capture summary>"}}, written by
    scripts/ncu_traffic.py from the `ncu --set full` report); None when the round has no capture of that kernel -- never a
    constant copied from an older build."""
    p = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
    try:
        with open(p) as f:
            return json.load(f).get(key, {}).get('bytes')
    except (OSError, ValueError):
        return None


WORKLOAD = 'mnist_fashion'       # set from --workload; the default is BASELINE.json configs[1]


def workload_string(batch, epoch):
    """config.workload of BOTH arms (the driver compares them verbatim)."""
    return ('codes/%s_config.json @ batch %d per GPU, epoch %d (all 4 sub-steps, 50-component hyper-prior, L=100 MC samples)'
            % (WORKLOAD, batch, epoch))


def load_config(batch):
    with open(os.path.join(ROOT, 'codes', WORKLOAD + '_config.json')) as f:
        cfg = json.load(f)
    cfg['batch_size'] = batch
    return cfg


def image_shape(cfg):
    return (int(cfg['dim_input_x']), int(cfg['dim_input_y']), int(cfg['dim_input_channel']))


def synthetic_mixture(K, R, seed=7):
    rng = np.random.default_rng(seed)
    a = rng.normal(size=(K, R, R))
    return (rng.normal(size=(K, R)) * 1.5, a @ a.transpose(0, 2, 1) * 0.2 + 0.05 * np.eye(R),
            rng.uniform(0.05, 1.0, size=K))


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return {'hbm_gbs': d['hbm_gbs'], 'bf16_tflops': d['bf16_tflops'],
                'bf16_tflops_sustained': d.get('bf16_tflops_sustained', d['bf16_tflops']), 'source': 'measured'}
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0, 'source': 'fallback'}


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons (the profiling recipe's query, one row every 20 ms, time-stamped) recorded for
    the whole run; `summary(windows)` keeps the rows whose timestamp falls DURING the timed regions."""
    Q = ('timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, device_index):
        self.idx = device_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix='.csv')
            os.close(fd)
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.idx), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '20'],
                                         stdout=open(self.path, 'w'), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.proc = None

    def summary(self, windows):
        """windows: [(t0, t1)] host wall-clock bounds (time.time()) of the timed regions."""
        import datetime
        self.stop()
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0,
               'windows': 'rows of `nvidia-smi -lms 20` time-stamped inside the %d timed region(s)' % len(windows)}
        sm, reasons = [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(',')]
                if len(f) < 10:
                    continue
                try:
                    ts = datetime.datetime.strptime(f[0], '%Y/%m/%d %H:%M:%S.%f').timestamp()
                    mhz, mx = float(f[2]), float(f[3])
                except ValueError:
                    continue
                if not any(t0 <= ts <= t1 for t0, t1 in windows):
                    continue
                sm.append(mhz)
                out['sm_max_mhz'] = mx
                for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[6:10]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out['sm_mhz'] = float(np.median(sm))
            out['samples'] = len(sm)
        out['reasons'] = sorted(reasons)
        return out


# ------------------------------------------------------------------------------------ CPU arm
def cpu_reference(cfg_full, sample_B, steps, warmup):
    """The oracle port of one reference iteration on the host cores, float32: the torch-CPU restatement
    (`oracle/torch_cpu.py`: oneDNN convolutions + autograd, all host threads -- a multi-threaded framework graph like the
    reference's TF1.15 one; pinned to the float64 NumPy oracle at 1e-8 by tests/test_oracle_torch_cpu.py)."""
    from oracle import params as oparams, steps as osteps
    cfg = dict(cfg_full)
    cfg['batch_size'] = sample_B
    rng = np.random.default_rng(0)
    spec = oparams.vae_param_specs(cfg) + oparams.prior_param_specs(cfg)
    P = oparams.glorot_init(spec, cfg, 1, dtype=np.float32)
    C, R, L, K = cfg['code_size'], cfg['representation_size'], cfg['n_MC_samples'], cfg['n_mixtures']
    x = rng.uniform(size=(sample_B,) + image_shape(cfg)).astype(np.float32)
    epoch = cfg['sg_pretraining'] + 1
    feeds = osteps.compute_feeds(cfg, epoch, synthetic_mixture(K, R))
    cores = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else os.cpu_count()
    import torch
    from oracle import torch_cpu
    torch.set_num_threads(cores)
    tr = torch_cpu.TorchTrainer(cfg, P)
    how = 'fp32 torch-CPU restatement of the reference graph (oneDNN + autograd, %d threads)' % torch.get_num_threads()

    def noise():
        return [dict(eps_z=rng.normal(size=(sample_B, C)), eps_t=rng.normal(size=(sample_B, R)),
                     eps_mc=rng.normal(size=(L, sample_B, R))) for _ in range(4)]
    for _ in range(warmup):
        tr.iteration(x, noise(), feeds, epoch)
    t0 = time.perf_counter()
    for _ in range(steps):
        tr.iteration(x, noise(), feeds, epoch)
    dt = time.perf_counter() - t0
    return {'value': sample_B * steps / dt, 'unit': 'imgs/s', 'cores': cores, 'kind': 'port',
            'sample': '%d iteration(s) of the 4-sub-step protocol on a %d-image batch of the same config, %s'
                      % (steps, sample_B, how), 'seconds': dt}


def run_reference(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    cfg = load_config(args.batch)
    # --steps / --warmup are honoured as given; each step is ONE iteration on a bounded sample of the workload (cpu_sample
    # images, default the full 1024-image batch: ~1.3 s per step on 16 cores), so the default run ends within a minute
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    base = cpu_reference(cfg, args.cpu_sample, steps, warmup)
    line = {'impl': 'reference', 'metric': METRIC, 'value': base['value'], 'unit': 'imgs/s', 'n_gpus': args.gpus,
            'steps': steps, 'warmup': warmup, 'ms_per_step': 1e3 * base['seconds'] / steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': workload_string(args.batch, cfg['sg_pretraining'] + 1)},
            'cpu_baseline': {k: base[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')},
            'e2e': {'value': base['value'], 'unit': 'imgs/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------ GPU arm
def celeba_leg(args, dev, rank, world, group, B, code_size=None, steps=None):
    """Same 4-sub-step iteration on codes/celeba_config.json (128x128x3 conv VAE + prior VAE + hyper-prior) through the
    reference-facing model / trainer classes: `value` with synthetic batches of B images per GPU resident in HBM, `e2e` through
    `CelebATrainer_joint_training.train_step_ae / train_step_prior` on pinned HOST batches (H2D copy + loss read-back inside the
    timed region), max over ranks of the CUDA-event time, its own clock window, and (rank 0) the roofline of its dominant GEMM
    (decoder/conv2d_7: fprop / dgrad / wgrad) and of its heaviest HBM-bound pass.  code_size overrides the JSON
    (BASELINE.json configs[4]: latent dim 128 at batch 512 per GPU)."""
    import contextlib
    import torch
    import torch.distributed as dist
    from ladder_latent_data_distribution_modelling_b200 import ops
    from ladder_latent_data_distribution_modelling_b200.host import models as hmodels, trainers as htrainers
    with open(os.path.join(ROOT, 'codes', 'celeba_config.json')) as f:
        cfg = json.load(f)
    cfg.update(batch_size=B, seed=1234, compute_dtype=args.dtype, synthetic=True, synthetic_pool=B)
    cfg['checkpoint_dir'] = cfg['result_dir'] = tempfile.mkdtemp() + '/'
    if code_size:
        cfg['code_size'] = int(code_size)
    if args.no_graphs:
        cfg['cuda_graphs'] = False
    with contextlib.redirect_stdout(sys.stderr):
        model = hmodels.CelebAModel_densenet(cfg, device=dev, dist_group=group)
    eng = model.engine
    if world > 1:
        for g in eng.groups.values():
            dist.broadcast(g.param, 0)
    gm = synthetic_mixture(cfg['n_mixtures'], cfg['representation_size'])
    eng.set_feeds(prior_mean=gm[0], prior_cov=gm[1], prior_weight=gm[2], use_standard_gaussian_prior=False, use_mask=False)
    eng.set_lrs(cfg['learning_rate_ae'], cfg['learning_rate_sigma'], cfg['learning_rate_prior'],
                cfg['learning_rate_inner_sigma'])
    gen = torch.Generator(device=dev)
    gen.manual_seed(7 + rank)
    pool = [torch.rand(B, 128, 128, 3, device=dev, generator=gen) for _ in range(4)]
    steps = steps or args.celeba_steps

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync()
        w0 = time.time()
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        sync()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), (w0, time.time())

    def iteration(i):
        for name in ('ae', 'sigma', 'prior', 'inner_sigma'):
            eng.run_step(name, pool[i % 4])
    sampler = ClockSampler(dev.index or 0)
    if rank == 0:
        sampler.start()
    for i in range(3):
        iteration(i)
    ms, win = timed(iteration, steps)
    windows = [win]
    fl = 3 * 10.04e9          # algorithmic fwd+bwd flop per image (one fused pass; the 4-sub-step protocol runs more)
    v = B * world * steps / (ms * 1e-3)
    out = {'workload': 'codes/celeba_config.json @ batch %d per GPU (128x128x3, H=512, C=%d), all 4 sub-steps' % (B, cfg['code_size']),
           'value': v, 'unit': 'imgs/s', 'ms_per_step': ms / steps, 'steps': steps, 'warmup': 3, 'global_batch': B * world,
           'algorithmic_tflops': v * fl / 1e12 / world, 'fused_norm_layers': bool(eng.outer.fused),
           'loss_ae': eng.fetch(['loss_ae'])['loss_ae']}
    del pool
    # ---- e2e: the reference-facing trainer on pinned host batches
    try:
        class _Data:
            n_train, n_val = 180000, 20000
        with contextlib.redirect_stdout(sys.stderr):
            trainer = htrainers.CelebATrainer_joint_training(None, model, _Data(), cfg)
        epoch = cfg['sg_pretraining'] + 1
        trainer.cur_epoch = epoch
        model.GM_prior_training.means_, model.GM_prior_training.covariances_, model.GM_prior_training.weights_ = gm
        trainer._gm_version += 1
        host_pool = [torch.rand(B, 128, 128, 3).pin_memory() for _ in range(2)]
        trainer.compute_cur_lr()

        def e2e_iteration(i):
            loss = trainer.train_step_ae(cur_lr=trainer.cur_lr, batch_data=host_pool[i % 2])
            trainer.train_step_prior(batch_data=host_pool[i % 2])
            return float(loss)                       # device -> host read of the step's loss
        for i in range(2):
            e2e_iteration(i)
        trainer._pending = []
        e_steps = max(2, steps // 2)
        e_ms, win = timed(e2e_iteration, e_steps)
        windows.append(win)
        out['e2e'] = {'value': B * world * e_steps / (e_ms * 1e-3), 'unit': 'imgs/s', 'ms_per_step': e_ms / e_steps,
                      'steps': e_steps, 'h2d_bytes_per_step': 2 * B * 128 * 128 * 3 * 4, 'd2h_bytes_per_step': 4,
                      'api': 'CelebATrainer_joint_training.train_step_ae + train_step_prior on pinned host batches'}
        del trainer, host_pool
    except Exception as e:                            # noqa: BLE001
        out['e2e'] = {'error': '%s: %s' % (type(e).__name__, str(e).splitlines()[0][:200] if str(e) else '')}
    if rank == 0:
        out['clocks'] = sampler.summary(windows)
    eng.release_graphs()
    H, bf = int(cfg['num_hidden_units']), torch.bfloat16
    del eng, model
    torch.cuda.empty_cache()
    # ---- roofline of this leg (rank 0): the GEMMs of decoder/conv2d_7 alone, and the heaviest HBM-bound pass
    if rank == 0 and args.dtype == 'bf16':
        try:
            peaks = measured_peaks()
            c = H // 4
            g = ops.ConvGeom(B, 128, 128, c, 3, 3, c, 1, 'same')
            xk = torch.randn(B, 128, 128, c, device=dev).to(bf)
            dyk = torch.randn(B, 128, 128, c, device=dev).to(bf)
            yk = torch.empty(B, 128, 128, c, device=dev, dtype=bf)
            wk = torch.randn(3, 3, c, c, device=dev) * 0.05
            bk = torch.zeros(c, device=dev)
            dwk = torch.empty(3, 3, c, c, device=dev)
            wf, wd = ops.tma_pack(wk, g, ops.FPROP), ops.tma_pack(wk, g, ops.DGRAD)
            flops = 2.0 * B * 128 * 128 * c * 9 * c
            act_bytes = 2.0 * B * 128 * 128 * c

            def kernel_ms(fn, reps=5):
                for _ in range(2):
                    fn()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                return e0.elapsed_time(e1) / reps
            rf = {}
            for name, fn, nbytes in (
                    ('fprop', lambda: ops.conv2d_fprop(xk, wk, bk, yk, g, 'leaky_relu', wimg=wf), 2 * act_bytes),
                    ('dgrad', lambda: ops.conv2d_dgrad(dyk, wk, yk, g, wimg=wd), 2 * act_bytes),
                    ('wgrad', lambda: ops.conv2d_wgrad(xk, dyk, dwk, None, g), 2 * act_bytes)):
                t = kernel_ms(fn)
                rf[name] = {'kernel': 'tma_kernel<%s> decoder/conv2d_7 [%d,128,128,%d] 3x3 -> %d' % (name.upper(), B, c, c),
                            'bound': 'tensor', 'achieved': flops / (t * 1e-3) / 1e12, 'peak': peaks['bf16_tflops_sustained'],
                            'unit': 'TFLOP/s', 'frac': flops / (t * 1e-3) / 1e12 / peaks['bf16_tflops_sustained'],
                            'ms_per_launch': t, 'algorithmic_flops_per_launch': flops, 'algorithmic_bytes_per_launch': nbytes,
                            'traffic': ncu_traffic('celeba_bf16_b%d_conv7_%s' % (B, name)),
                            'peak_source': peaks['source'] + ' bf16 sustained (kernel inside a long step)'}
            del xk, dyk, dwk
            # heaviest HBM-bound pass: instance norm + style + leaky + 64->128 resize in one pass (decoder block 6)
            cin = torch.randn(B, 64, 64, c, device=dev).to(bf)
            insum = torch.stack([cin.float().sum((1, 2)), (cin.float() ** 2).sum((1, 2))]).contiguous()
            sty = torch.randn(B, 2 * c, device=dev)
            t = kernel_ms(lambda: ops.in_style_resize16(cin, insum, sty, yk))
            nbytes = 2.0 * B * c * (64 * 64 + 128 * 128)
            rf['hbm_pass'] = {'kernel': 'in_style_resize16_kernel: instance norm + style + leaky + 64->128 legacy bilinear, '
                                        '[%d,64,64,%d] -> [%d,128,128,%d] bf16' % (B, c, B, c),
                              'bound': 'hbm', 'achieved': nbytes / (t * 1e-3) / 1e9, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                              'frac': nbytes / (t * 1e-3) / 1e9 / peaks['hbm_gbs'], 'ms_per_launch': t,
                              'algorithmic_bytes_per_launch': nbytes, 'traffic': ncu_traffic('celeba_bf16_b%d_in_style_resize' % B),
                              'peak_source': peaks['source'] + ' copy bandwidth'}
            out['roofline'] = rf
            del cin, yk, insum, sty
            torch.cuda.empty_cache()
        except Exception as e:                        # noqa: BLE001
            out['roofline'] = {'error': '%s: %s' % (type(e).__name__, str(e).splitlines()[0][:200] if str(e) else '')}
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    from ladder_latent_data_distribution_modelling_b200 import ops
    from ladder_latent_data_distribution_modelling_b200.engine import LadderEngine
    from ladder_latent_data_distribution_modelling_b200.host import models as hmodels, trainers as htrainers

    world = int(os.environ.get('WORLD_SIZE', 1))
    rank = int(os.environ.get('RANK', 0))
    local = int(os.environ.get('LOCAL_RANK', 0))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py needs a CUDA device; there is no CPU fallback for the product path')
    torch.cuda.set_device(local)
    group = None
    if world > 1:
        os.environ['NCCL_DEBUG'] = os.environ.get('LADDER_NCCL_DEBUG', 'WARN')     # NCCL's version banner goes to STDOUT otherwise
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
        group = dist.group.WORLD
    dev = torch.device('cuda', local)
    B = args.batch
    cfg = load_config(B)
    cfg['seed'] = 1234
    cfg['compute_dtype'] = args.dtype
    if args.no_graphs:
        cfg['cuda_graphs'] = False
    K, R = cfg['n_mixtures'], cfg['representation_size']
    epoch = cfg['sg_pretraining'] + 1            # past pretraining: all four sub-steps active, fitted mixture fed
    gm = synthetic_mixture(K, R)

    # ---- device-resident loop (value)
    eng = LadderEngine(cfg, B, dev, seed=1234 + rank, dist_group=group)
    if world > 1:                                 # identical initial weights on every rank
        for g in eng.groups.values():
            dist.broadcast(g.param, 0)
    eng.set_feeds(prior_mean=gm[0], prior_cov=gm[1], prior_weight=gm[2], use_standard_gaussian_prior=False,
                  use_mask=False)
    eng.set_lrs(cfg['learning_rate_ae'], cfg['learning_rate_sigma'], cfg['learning_rate_prior'],
                cfg['learning_rate_inner_sigma'])
    gen = torch.Generator(device=dev)
    gen.manual_seed(99 + rank)
    n_pool = 4
    shp = image_shape(cfg)
    pool = [torch.rand(B, *shp, device=dev, generator=gen) for _ in range(n_pool)]

    def iteration(x):
        for name in ('ae', 'sigma', 'prior', 'inner_sigma'):
            eng.run_step(name, x)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for i in range(args.warmup):
        iteration(pool[i % n_pool])
    barrier()
    windows = []
    w0 = time.time()
    launches0 = ops.launch_count() + eng.replayed_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    profile = os.environ.get('LADDER_BENCH_PROFILE') == '1'     # ncu --profile-from-start off: timed region only
    if profile:
        torch.cuda.profiler.start()
    e0.record()
    for i in range(args.steps):
        iteration(pool[i % n_pool])
    e1.record()
    barrier()
    if profile:
        torch.cuda.profiler.stop()
        if rank == 0:
            sampler.stop()
        print('profiled %d iterations (no bench line under a profiler)' % args.steps, file=sys.stderr)
        return
    ms = e0.elapsed_time(e1)
    launches = ops.launch_count() + eng.replayed_launches - launches0
    windows.append((w0, time.time()))
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = B * world * args.steps / (ms * 1e-3)
    loss_check = eng.fetch(['loss_prior'])['loss_prior']

    # ---- end-to-end through the reference-facing API with HOST buffers (e2e)
    cfg2 = dict(cfg)
    cfg2['checkpoint_dir'] = cfg2['result_dir'] = tempfile.mkdtemp() + '/'
    import contextlib
    with contextlib.redirect_stdout(sys.stderr):      # keep stdout to the single JSON line
        model_cls = {'mnist_fashion': hmodels.MNISTModel_fashion, 'mnist_digit': hmodels.MNISTModel_digit,
                     'celeba': hmodels.CelebAModel_densenet}[WORKLOAD]
        model = model_cls(cfg2, device=dev, dist_group=group)
    if world > 1:
        for g in model.engine.groups.values():
            dist.broadcast(g.param, 0)

    class _Data:
        n_train, n_val = 60000, 10000
        test_set = {'image': np.zeros((B,) + shp, np.float32)}
    if WORKLOAD == 'celeba':
        cfg2['synthetic_pool'] = B
        with contextlib.redirect_stdout(sys.stderr):
            trainer = htrainers.CelebATrainer_joint_training(None, model, _Data(), cfg2)
    else:
        trainer = htrainers.MNISTTrainer_joint_training(None, model, _Data(), cfg2)
    trainer.cur_epoch = epoch
    model.GM_prior_training.means_, model.GM_prior_training.covariances_, model.GM_prior_training.weights_ = gm
    host_pool = [torch.rand(B, *shp).pin_memory() for _ in range(n_pool)]
    lr = cfg['learning_rate_ae'] * 0.99 ** (epoch - 1)

    def e2e_iteration(hx):
        loss = trainer.train_step_ae(cur_lr=lr, batch_data=hx)
        trainer.train_step_prior(batch_data=hx)
        return float(loss)                       # device -> host read of the step's loss

    for i in range(args.warmup):
        e2e_iteration(host_pool[i % n_pool])
    trainer._pending = []
    barrier()
    w0 = time.time()
    e0.record()
    for i in range(args.steps):
        last = e2e_iteration(host_pool[i % n_pool])
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    e2e_value = B * world * args.steps / (e2e_ms * 1e-3)
    windows.append((w0, time.time()))               # second sampling window: the end-to-end timed region
    clocks = sampler.summary(windows) if rank == 0 else None

    # ---- component-sharded hyper-prior (BASELINE.json configs[2]): every rank evaluates K / world components
    # of the 65 536 x 65 536 problem, (m, s[, g]) partials are all-gathered over NCCL and combined; max over ranks
    sharded = None
    if world > 1:
        try:
            from ladder_latent_data_distribution_modelling_b200 import parallel
            sharded = {'N': 65536, 'K': 65536, 'components_per_rank': -(-65536 // world),
                       'exchange': 'one packed [N, 2+D] partial per rank, one all_gather_into_tensor, one combine kernel; the '
                                   'three launches replayed as one CUDA graph'}
            Ns = 65536
            for D in (2, 32, 64):                                    # BASELINE.json configs[2]: D = 2 / 32 / 64
                rng = np.random.default_rng(1234 + D)                # same queries and components on every rank
                sc = 1.0 if D == 2 else 1.0 / np.sqrt(D / 2.0)
                tq_s = torch.tensor((rng.normal(size=(Ns, D)) * sc).astype(np.float32), device=dev)
                tab_s = ops.mixture_pack_diag(rng.normal(size=(Ns, D)) * sc, 1.0, None, dev)
                for grad in (False, True):
                    sm = parallel.ShardedMixture(tab_s, Ns, group=group, want_grad=grad, graph=not args.no_graphs)
                    for _ in range(3):
                        sm(tq_s)
                    barrier()
                    reps = 5 if D == 2 else 2
                    e0.record()
                    for _ in range(reps):
                        sm(tq_s)
                    e1.record()
                    barrier()
                    ts = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
                    dist.all_reduce(ts, op=dist.ReduceOp.MAX)
                    key = 'pairs_per_s_fwd_grad' if grad else 'pairs_per_s_fwd'
                    sharded[key if D == 2 else '%s_D%d' % (key, D)] = Ns * Ns / (float(ts.item()) * 1e-3)
                    sharded['graph'] = bool(sm.use_graph)
                    del sm
            sharded['D'] = 2
            sharded['kernels'] = {'D2': 'mix_kernel (registers, fp32), packed (m, s, g) partial',
                                  'D32_D64': 'mix_tc_kernel / mix_tc_grad_kernel (tcgen05 kind::tf32) on the rank\'s slice of the '
                                             'operand images, packed partial from the finalising kernel'}
        except Exception as e:                                       # noqa: BLE001
            sharded = {'error': '%s: %s' % (type(e).__name__, str(e).splitlines()[0][:200] if str(e) else '')}

    line = None
    if rank == 0:
        peaks = measured_peaks()
        # ---- roofline of the dominant kernel: the implicit-GEMM of decoder/conv2d_3 (16x16, 64 -> 256, 3x3),
        # 68 % of the model's MACs; timed alone with CUDA events on the launch stream
        H = cfg['num_hidden_units']
        # (name, HW, Cin, Cout) of the layer that carries the most MACs of the model
        dom_name, hw, ci, co = {'mnist_fashion': ('decoder/conv2d_3', 16, H // 4, H), 'mnist_digit': ('decoder/conv2d', 4, H, H),
                                'celeba': ('decoder/conv2d_7', 128, H // 4, H // 4)}[WORKLOAD]
        g = ops.ConvGeom(B, hw, hw, ci, 3, 3, co, 1, 'same')
        tma = args.dtype == 'bf16' and ops.tma_supported(g, ops.FPROP)
        adt = torch.bfloat16 if tma else torch.float32     # the step keeps this layer's activations bf16-resident
        xk = torch.randn(B, hw, hw, ci, device=dev).to(adt)
        wk = torch.randn(3, 3, ci, co, device=dev) * 0.05
        bk = torch.zeros(co, device=dev)
        yk = torch.empty(B, hw, hw, co, device=dev, dtype=adt)
        wimg = ops.tma_pack(wk, g, ops.FPROP) if tma else None   # the step packs all weights once per sub-step
        d2s = 2 if tma and WORKLOAD != 'celeba' else 0     # the MNIST decoders store straight in depth_to_space layout
        for _ in range(3):
            ops.conv2d_fprop(xk, wk, bk, yk, g, 'leaky_relu', wimg=wimg, out_d2s=d2s)
        torch.cuda.synchronize()
        reps = 20
        e0.record()
        for _ in range(reps):
            ops.conv2d_fprop(xk, wk, bk, yk, g, 'leaky_relu', wimg=wimg, out_d2s=d2s)
        e1.record()
        torch.cuda.synchronize()
        k_ms = e0.elapsed_time(e1) / reps
        flops = 2.0 * B * hw * hw * co * 9 * ci
        ach = flops / (k_ms * 1e-3) / 1e12
        esz = 2 if tma else 4
        alg_bytes = esz * B * hw * hw * (ci + co) + 2 * 9 * ci * co
        kname = ('tma_kernel<FPROP> (TMA im2col + tcgen05, bf16 in/out)' if tma else
                 'tc_kernel<FPROP> (tcgen05, fp32 activations)' if args.dtype == 'bf16' else 'igemm_kernel<FPROP> (fp32 SIMT)')
        roofline = {'kernel': kname + ' %s [B,%d,%d,%d]->%d 3x3' % (dom_name, hw, hw, ci, co),
                    'bound': 'tensor', 'achieved': ach, 'peak': peaks['bf16_tflops'], 'unit': 'TFLOP/s',
                    'frac': ach / peaks['bf16_tflops'], 'traffic': ncu_traffic('%s_%s_b%d_fprop' % (WORKLOAD, args.dtype, B)),
                    'peak_source': peaks['source'] + ' bf16 burst (kernel timed alone)',
                    'ms_per_launch': k_ms, 'algorithmic_flops_per_launch': flops, 'algorithmic_bytes_per_launch': alg_bytes,
                    'hbm_floor_ms': alg_bytes / (peaks['hbm_gbs'] * 1e9) * 1e3,
                    'tensor_floor_ms': flops / (peaks['bf16_tflops'] * 1e12) * 1e3,
                    'timed': 'CUDA events on the launch stream around %d back-to-back launches; input + output = %d MB '
                             '(%s the 126 MB L2)' % (reps, alg_bytes >> 20, 'exceeds' if alg_bytes > 126 << 20 else 'FITS in'),
                    'note': 'bf16 activations and pre-packed bf16 weights in HBM, fp32 accumulation in TMEM, fused bias + leaky_relu%s, '
                            'exactly the launch the step makes' % (' + depth_to_space(2) store' if d2s else '')
                    if tma else 'denominator is the bf16 tensor peak'}
        if tma and ops.tma_supported(g, ops.DGRAD):
            # The largest kernel of the step (ncu launch list / CUPTI, profiles/): the decoder data gradients,
            # tma_kernel<DGRAD, BN=64, bf16>.  Its largest launch is the dgrad of this same layer, with the fused producer
            # leaky_relu' and (MNIST decoders) space_to_depth scatter -- it becomes `roofline`, the fprop above `roofline_fprop`.
            dyk = torch.randn(B, hw, hw, co, device=dev).to(adt)
            auxk = torch.randn(B, hw, hw, ci, device=dev).to(adt)
            dxk = torch.empty(B * hw * hw * ci, device=dev, dtype=adt)
            dxk = dxk.view(B, hw // d2s, hw // d2s, ci * d2s * d2s) if d2s else dxk.view(B, hw, hw, ci)
            wimg_d = ops.tma_pack(wk, g, ops.DGRAD)
            for _ in range(3):
                ops.conv2d_dgrad(dyk, wk, dxk, g, act_out=auxk, act='leaky_relu', out_s2d=d2s, wimg=wimg_d)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(reps):
                ops.conv2d_dgrad(dyk, wk, dxk, g, act_out=auxk, act='leaky_relu', out_s2d=d2s, wimg=wimg_d)
            e1.record()
            torch.cuda.synchronize()
            d_ms = e0.elapsed_time(e1) / reps
            d_bytes = 2 * B * hw * hw * (co + 2 * ci) + 2 * 9 * ci * co          # dy + saved activation + dx + weights, bf16
            d_ach = flops / (d_ms * 1e-3) / 1e12
            roofline_fprop = roofline
            roofline = {'kernel': 'tma_kernel<DGRAD,64> (TMA im2col + tcgen05, bf16 in/out) %s [B,%d,%d,%d]<-%d 3x3'
                                  % (dom_name, hw, hw, ci, co),
                        'bound': 'tensor', 'achieved': d_ach, 'peak': peaks['bf16_tflops'], 'unit': 'TFLOP/s',
                        'frac': d_ach / peaks['bf16_tflops'], 'traffic': ncu_traffic('%s_%s_b%d_dgrad' % (WORKLOAD, args.dtype, B)),
                        'peak_source': peaks['source'] + ' bf16 burst (kernel timed alone)',
                        'ms_per_launch': d_ms, 'algorithmic_flops_per_launch': flops, 'algorithmic_bytes_per_launch': d_bytes,
                        'hbm_floor_ms': d_bytes / (peaks['hbm_gbs'] * 1e9) * 1e3,
                        'tensor_floor_ms': flops / (peaks['bf16_tflops'] * 1e12) * 1e3,
                        'timed': 'CUDA events on the launch stream around %d back-to-back launches; dy + aux + dx = %d MB (%s the '
                                 '126 MB L2)' % (reps, d_bytes >> 20, 'exceeds' if d_bytes > 126 << 20 else 'FITS in'),
                        'note': 'largest launch of the largest kernel of the step (11-12 %% of the step, profiles/r2ac_kernel_times_fashion.txt); '
                                'fused producer leaky_relu\'%s; issue warps run converged with one elected lane since r2w '
                                '(profiles/r2w_issue_warp_analysis.md); what remains is the L2->SM operand stream at N = 64'
                                % (' + space_to_depth(2) scatter' if d2s else ''),
                        'roofline_fprop': roofline_fprop}
        # ---- hyper-prior micro-benchmark (second half of the metric, BASELINE.json configs[2]): 65 536 x 65 536 pairs at
        # D = 2 (register / SIMT kernel) and D = 32, 64 (tcgen05 kernels: scores on kind::tf32, exponentials out of TMEM; the
        # gradient's second contraction W . mu with W kept in TMEM), forward and forward + gradient
        n_ex2 = ops.pipe_peak(1, 148 * 8, 512)
        torch.cuda.synchronize()
        e0.record()
        ops.pipe_peak(1, 148 * 8, 512)
        e1.record()
        torch.cuda.synchronize()
        ex2_peak = n_ex2 / (e0.elapsed_time(e1) * 1e-3)
        N = 65536
        hyper = {'N': N, 'K': N, 'bound': 'sfu (1 MUFU.EX2 per pair)', 'ex2_peak_per_s_measured': ex2_peak}
        for D in (2, 32, 64):
            rng = np.random.default_rng(1234 + D)
            sc = 1.0 if D == 2 else 1.0 / np.sqrt(D / 2.0)           # keep |t - mu|^2 ~ O(1) so the sums stay in fp32 range
            tq = torch.tensor((rng.normal(size=(N, D)) * sc).astype(np.float32), device=dev)
            tab = ops.mixture_pack_diag(rng.normal(size=(N, D)) * sc, 1.0, None, dev)
            for grad in (False, True):
                for _ in range(3):
                    ops.mixture_logprob(tq, tab, want_grad=grad)
                torch.cuda.synchronize()
                e0.record()
                for _ in range(5):
                    ops.mixture_logprob(tq, tab, want_grad=grad)
                e1.record()
                torch.cuda.synchronize()
                rate = N * N / (e0.elapsed_time(e1) / 5 * 1e-3)
                key = ('fwd_grad' if grad else 'fwd') + ('' if D == 2 else '_D%d' % D)
                hyper['pairs_per_s_' + key] = rate
                hyper['frac_' + key] = rate / ex2_peak
            del tq, tab
        hyper['D'] = 2
        hyper['kernels'] = {'D2': 'mix_kernel (registers, fp32)', 'D32_D64': 'mix_tc_kernel / mix_tc_grad_kernel (tcgen05 kind::tf32, TMEM)'}
        # reported baseline, rank 0 at N = 1 only (at N > 1 the other ranks' processes would share the host cores with it)
        cpu = cpu_reference(cfg, args.cpu_sample, 2, 1) if world == 1 else None
        line = {'metric': METRIC, 'value': value, 'unit': 'imgs/s', 'n_gpus': world, 'steps': args.steps,
                'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': 'bf16' if args.dtype == 'bf16' else 'f32', 'data': 'synthetic',
                'config': {'workload': workload_string(B, epoch),
                           'global_batch': B * world, 'parallelism': 'dp%d' % world, 'cuda_graphs': bool(eng.use_graphs),
                           'l2': 'per-step activation working set (>1 GB) exceeds the 126 MB L2; 4 input batches rotate'},
                'clocks': clocks, 'gpu_launches': launches,
                'e2e': {'value': e2e_value, 'unit': 'imgs/s', 'h2d_bytes_per_step': 2 * B * int(np.prod(shp)) * 4,
                        'd2h_bytes_per_step': 4, 'ms_per_step': e2e_ms / args.steps,
                        'api': '*Trainer_joint_training.train_step_ae + train_step_prior on pinned host batches'},
                'roofline': roofline,
                'cpu_baseline': {k: cpu[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')} if cpu else None,
                'hyper_prior': hyper, 'hyper_prior_component_sharded': sharded, 'celeba_shape': None, 'loss_prior_last': loss_check, 'loss_ae_last_e2e': last}
    # ---- secondary workload: CelebA-shape (128x128x3) training step, BASELINE.json configs[3] per-GPU batch
    celeba = None
    if WORKLOAD != 'celeba' and args.celeba_batch > 0:
        del model, trainer
        eng.release_graphs()
        torch.cuda.empty_cache()
        celeba = {}
        for key, b, cs, st in (('reference_batch', args.celeba_batch, None, args.celeba_steps),
                               ('weak_scaling_config', args.celeba_big_batch, 128, max(3, args.celeba_steps // 2))):
            if b <= 0:
                continue
            try:                                      # a secondary leg must never take the headline line down with it
                celeba[key] = celeba_leg(args, dev, rank, world, group, b, cs, st)
            except Exception as e:                    # noqa: BLE001
                celeba[key] = {'error': '%s: %s' % (type(e).__name__, str(e).splitlines()[0][:200] if str(e) else '')}
                try:
                    torch.cuda.synchronize()
                except Exception:                     # noqa: BLE001
                    break

    if line is not None:
        line['celeba_shape'] = celeba
    try:
        if world > 1:
            eng.release_graphs()             # captured graphs hold NCCL kernels: drop them before the group goes away
            try:
                model.engine.release_graphs()
            except NameError:                # already deleted in front of the CelebA legs
                pass
            torch.cuda.synchronize()
            dist.barrier()
            dist.destroy_process_group()
    except Exception:                                 # noqa: BLE001
        pass
    if line is not None:
        print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=None,
                    help='timed iterations (default: 320 for the GPU arm = a timed region of about 1 s at 3.2 ms per iteration; 20 for --impl reference)')
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=1024, help='images per GPU')
    ap.add_argument('--no-graphs', action='store_true', help='launch every kernel eagerly instead of CUDA-graph replay')
    ap.add_argument('--dtype', default='bf16', choices=['bf16', 'fp32'], help='GEMM math: bf16 tcgen05 or fp32 SIMT')
    ap.add_argument('--cpu-sample', type=int, default=0,
                    help='batch of the bounded CPU-baseline sample (0 = the full batch, at most 1024 images, for the MNIST workloads; 16 for celeba: ~10-20 s)')
    ap.add_argument('--celeba-batch', type=int, default=64, help='per-GPU batch of the secondary CelebA-shape leg (0 = skip)')
    ap.add_argument('--celeba-big-batch', type=int, default=512,
                    help='per-GPU batch of the CelebA-shape weak-scaling leg (BASELINE.json configs[4]: 4096 over 8 GPUs, code_size 128)')
    ap.add_argument('--celeba-steps', type=int, default=10)
    ap.add_argument('--workload', default='mnist_fashion', choices=['mnist_fashion', 'mnist_digit', 'celeba'],
                    help='config file under codes/ (default: BASELINE.json configs[1])')
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 320 if args.impl == 'ours' and args.workload != 'celeba' else 20
    global WORKLOAD
    WORKLOAD = args.workload
    if args.cpu_sample <= 0:
        # bounded sample: the full batch when the run is short, fewer images per step when --steps is large, so that the
        # reference arm's whole run stays within a few minutes (~700 img/s on 16 host cores)
        cap = max(32, int(90e3 / max(1, args.steps + args.warmup))) if args.impl == 'reference' else 1024
        args.cpu_sample = 16 if WORKLOAD == 'celeba' else min(args.batch, 1024, cap)
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()

"""Halo-mode probe: stride-1 3x3 fprop / dgrad through the halo kernel (both base-offset modes) against the plain TMA kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ladder_latent_data_distribution_modelling_b200 import ops

ops.set_math_mode('bf16')
bf = torch.bfloat16
torch.manual_seed(0)
for (B, HW, Cin, Cout) in [(2, 16, 64, 64), (3, 16, 256, 64), (2, 32, 128, 128), (4, 16, 64, 256)]:
    g = ops.ConvGeom(B, HW, HW, Cin, 3, 3, Cout, 1, 'same')
    x = torch.randn(B, HW, HW, Cin, device='cuda').to(bf)
    w = torch.randn(3, 3, Cin, Cout, device='cuda') * 0.05
    b = torch.randn(Cout, device='cuda')
    dy = torch.randn(B, HW, HW, Cout, device='cuda').to(bf)
    aux = torch.randn(B, HW, HW, Cin, device='cuda').to(bf)
    res = {}
    for name, (en, bm) in (("plain", (0, 0)), ("halo_base1", (2, 1)), ("halo_base0", (2, 0))):
        ops.set_halo(en, bm)
        y = torch.empty(B, HW, HW, Cout, device='cuda', dtype=bf)
        ops.conv2d_fprop(x, w, b, y, g, 'leaky_relu')
        dx = torch.empty(B, HW, HW, Cin, device='cuda', dtype=bf)
        ops.conv2d_dgrad(dy, w, dx, g, act_out=aux, act='leaky_relu')
        torch.cuda.synchronize()
        res[name] = (y.float(), dx.float())
    for name in ('halo_base1', 'halo_base0'):
        ey = (res[name][0] - res['plain'][0]).abs().max().item() / res['plain'][0].abs().max().item()
        ed = (res[name][1] - res['plain'][1]).abs().max().item() / res['plain'][1].abs().max().item()
        print('B%d HW%d %d->%d  %-11s fprop relerr %.3e  dgrad relerr %.3e' % (B, HW, Cin, Cout, name, ey, ed))
# timing on the two bench-relevant layers
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for (B, HW, Cin, Cout, tag) in [(1024, 16, 64, 256, 'fashion conv2d_3'), (128, 128, 128, 128, 'celeba conv2d_7 (B=128)'), (256, 64, 256, 128, 'celeba conv2d_6 (B=256)')]:
    g = ops.ConvGeom(B, HW, HW, Cin, 3, 3, Cout, 1, 'same')
    x = torch.randn(B, HW, HW, Cin, device='cuda').to(bf)
    w = torch.randn(3, 3, Cin, Cout, device='cuda') * 0.05
    b = torch.zeros(Cout, device='cuda')
    y = torch.empty(B, HW, HW, Cout, device='cuda', dtype=bf)
    dy = torch.randn(B, HW, HW, Cout, device='cuda').to(bf)
    dx = torch.empty(B, HW, HW, Cin, device='cuda', dtype=bf)
    wf, wd = ops.tma_pack(w, g, ops.FPROP), ops.tma_pack(w, g, ops.DGRAD)
    fl = 2.0 * B * HW * HW * Cin * Cout * 9
    for en in (0, 2):
        ops.set_halo(en, int(os.environ.get('HALO_BASE', '0')))
        tf = t(lambda: ops.conv2d_fprop(x, w, b, y, g, 'leaky_relu', wimg=wf))
        td = t(lambda: ops.conv2d_dgrad(dy, w, dx, g, act_out=x, act='leaky_relu', wimg=wd))
        print('%-26s halo=%d  fprop %.3f ms %.0f TF/s   dgrad %.3f ms %.0f TF/s' % (tag, en, tf, fl / tf / 1e9, td, fl / td / 1e9))

"""The four launches that lead the mnist_fashion step after r1f, each alone, for one `ncu --set full` capture:
  (a) dgrad of decoder/conv2d_3 (TMA kernel, N = 64, bf16 out, fused leaky' + space_to_depth)
  (b) first encoder conv on the 1-channel image (thin_k_fprop)
  (c) shifted copy DYS of the last decoder layer (tap_scatter_bf16)   (d) its tap-GEMM dgrad   (e) tap_sum of its fprop
  ncu --set full --clock-control none --import-source on -k regex:"tma_kernel|thin_k|tap_" -c 12 -o gpurun_out/x python scripts/ncu_targets.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ladder_latent_data_distribution_modelling_b200 import ops  # noqa: E402


def main():
    ops.set_math_mode('bf16')
    B = 1024
    bf = torch.bfloat16
    # (a)
    g = ops.ConvGeom(B, 16, 16, 64, 3, 3, 256, 1, 'same')
    dy = torch.randn(B, 16, 16, 256, device='cuda').to(bf)
    w = torch.randn(3, 3, 64, 256, device='cuda') * 0.05
    aux = torch.randn(B, 16, 16, 64, device='cuda').to(bf)
    dx = torch.empty(B, 8, 8, 256, device='cuda', dtype=bf)
    wimg = ops.tma_pack(w, g, ops.DGRAD)
    # (b)
    g0 = ops.ConvGeom(B, 32, 32, 1, 3, 3, 64, 2, 'same')
    x0 = torch.rand(B, 32, 32, 1, device='cuda')
    w0 = torch.randn(3, 3, 1, 64, device='cuda') * 0.3
    b0 = torch.zeros(64, device='cuda')
    y0 = torch.empty(B, 16, 16, 64, device='cuda')
    # (c, d, e)
    gl = ops.ConvGeom(B, 32, 32, 64, 5, 5, 1, 1, 'valid')
    xl = torch.randn(B, 32, 32, 64, device='cuda').to(bf)
    wl = torch.randn(5, 5, 64, 1, device='cuda') * 0.05
    bl = torch.zeros(1, device='cuda')
    yl = torch.empty(B, 28, 28, 1, device='cuda')
    dyl = torch.randn(B, 28, 28, 1, device='cuda')
    auxl = torch.randn(B, 32, 32, 64, device='cuda').to(bf)
    dxl = torch.empty(B, 16, 16, 256, device='cuda', dtype=bf)
    for _ in range(2):
        ops.conv2d_dgrad(dy, w, dx, g, act_out=aux, act='leaky_relu', out_s2d=2, wimg=wimg)
        ops.conv2d_fprop(x0, w0, b0, y0, g0, 'leaky_relu')
        ops.conv2d_fprop(xl, wl, bl, yl, gl, 'relu')
        ops.conv2d_dgrad(dyl, wl, dxl, gl, act_out=auxl, act='leaky_relu', out_s2d=2)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    ev[0].record()
    for _ in range(10):
        ops.conv2d_dgrad(dy, w, dx, g, act_out=aux, act='leaky_relu', out_s2d=2, wimg=wimg)
    ev[1].record()
    for _ in range(10):
        ops.conv2d_fprop(x0, w0, b0, y0, g0, 'leaky_relu')
    ev[2].record()
    for _ in range(10):
        ops.conv2d_fprop(xl, wl, bl, yl, gl, 'relu')
    ev[3].record()
    for _ in range(10):
        ops.conv2d_dgrad(dyl, wl, dxl, gl, act_out=auxl, act='leaky_relu', out_s2d=2)
    ev[4].record()
    torch.cuda.synchronize()
    for name, i in (('dgrad conv2d_3 (77.3 GFLOP)', 0), ('conv0 thin_k_fprop', 1), ('last conv fprop (tap-GEMM + tap_sum)', 2),
                    ('last conv dgrad (tap_scatter + tap-GEMM)', 3)):
        print('%-45s %.4f ms' % (name, ev[i].elapsed_time(ev[i + 1]) / 10))


if __name__ == '__main__':
    main()

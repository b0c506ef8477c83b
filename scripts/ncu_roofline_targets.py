"""The launches whose `roofline.traffic` bench.py reports, each alone and in a fixed order, for ONE `ncu --set full` capture:
  0 fashion decoder/conv2d_3 fprop  [1024,16,16,64] -> 256, bf16, fused leaky + depth_to_space(2)      key mnist_fashion_bf16_b1024_fprop
  1 fashion decoder/conv2d_3 dgrad  (N = 64, fused leaky' + space_to_depth(2))                          key mnist_fashion_bf16_b1024_dgrad
  2-4 celeba decoder/conv2d_7 fprop / dgrad / wgrad at batch B (default 64)                             key celeba_bf16_b<B>_conv7_*
  5 in_style_up2 (instance norm + style + leaky + 64->128 resize) at batch B                            key celeba_bf16_b<B>_in_style_resize
usage (GPU box):  ncu --set full --clock-control none -k regex:"tma_kernel|in_style_up2" -o gpurun_out/<tag>_roofline python scripts/ncu_roofline_targets.py [B]
then (here):      python scripts/ncu_traffic.py gpurun_out/<tag>_roofline.ncu-rep <B> profiles/<tag>_ncu_roofline.md"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ladder_latent_data_distribution_modelling_b200 import ops  # noqa: E402


def main():
    ops.set_math_mode('bf16')
    bf = torch.bfloat16
    Bc = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    B = 1024
    g = ops.ConvGeom(B, 16, 16, 64, 3, 3, 256, 1, 'same')
    x = torch.randn(B, 16, 16, 64, device='cuda').to(bf)
    w = torch.randn(3, 3, 64, 256, device='cuda') * 0.05
    b = torch.zeros(256, device='cuda')
    y = torch.empty(B, 32, 32, 64, device='cuda', dtype=bf)
    dy = torch.randn(B, 16, 16, 256, device='cuda').to(bf)
    dx = torch.empty(B, 8, 8, 256, device='cuda', dtype=bf)
    wf, wd = ops.tma_pack(w, g, ops.FPROP), ops.tma_pack(w, g, ops.DGRAD)
    ops.conv2d_fprop(x, w, b, y, g, 'leaky_relu', wimg=wf, out_d2s=2)
    ops.conv2d_dgrad(dy, w, dx, g, act_out=x, act='leaky_relu', out_s2d=2, wimg=wd)
    c = 128
    g7 = ops.ConvGeom(Bc, 128, 128, c, 3, 3, c, 1, 'same')
    x7 = torch.randn(Bc, 128, 128, c, device='cuda').to(bf)
    d7 = torch.randn(Bc, 128, 128, c, device='cuda').to(bf)
    y7 = torch.empty(Bc, 128, 128, c, device='cuda', dtype=bf)
    w7 = torch.randn(3, 3, c, c, device='cuda') * 0.05
    b7 = torch.zeros(c, device='cuda')
    dw7 = torch.empty(3, 3, c, c, device='cuda')
    ops.conv2d_fprop(x7, w7, b7, y7, g7, 'leaky_relu', wimg=ops.tma_pack(w7, g7, ops.FPROP))
    ops.conv2d_dgrad(d7, w7, y7, g7, wimg=ops.tma_pack(w7, g7, ops.DGRAD))
    ops.conv2d_wgrad(x7, d7, dw7, None, g7)
    cin = torch.randn(Bc, 64, 64, c, device='cuda').to(bf)
    insum = torch.stack([cin.float().sum((1, 2)), (cin.float() ** 2).sum((1, 2))]).contiguous()
    sty = torch.randn(Bc, 2 * c, device='cuda')
    ops.in_style_resize16(cin, insum, sty, y7)
    torch.cuda.synchronize()


if __name__ == '__main__':
    main()

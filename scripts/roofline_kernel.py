"""Launch the bench workload's dominant kernel alone (the TMA-fed tcgen05 fprop of decoder/conv2d_3 of mnist_fashion at
batch 1024: [1024,16,16,64] -> 256, 3x3, bf16 in/out, pre-packed weights) so that ONE `ncu --set full` capture gives its
DRAM traffic, tensor-pipe activity and stall reasons:

  ncu --set full --clock-control none --import-source on -k regex:tma_kernel -s 3 -c 1 -o gpurun_out/prof python scripts/roofline_kernel.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ladder_latent_data_distribution_modelling_b200 import ops  # noqa: E402


D2S = int(os.environ.get('LADDER_D2S', '2'))      # the step stores this layer's output in depth_to_space(2) layout


def main():
    ops.set_math_mode('bf16')
    B, hw, ci, co = 1024, 16, 64, 256
    g = ops.ConvGeom(B, hw, hw, ci, 3, 3, co, 1, 'same')
    x = torch.randn(B, hw, hw, ci, device='cuda').to(torch.bfloat16)
    w = torch.randn(3, 3, ci, co, device='cuda') * 0.05
    b = torch.zeros(co, device='cuda')
    y = torch.empty(B, hw, hw, co, device='cuda', dtype=torch.bfloat16)
    wimg = ops.tma_pack(w, g, ops.FPROP)
    for _ in range(5):
        ops.conv2d_fprop(x, w, b, y, g, 'leaky_relu', wimg=wimg, out_d2s=D2S)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.conv2d_fprop(x, w, b, y, g, 'leaky_relu', wimg=wimg, out_d2s=D2S)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print('%.4f ms  %.1f TFLOP/s' % (ms, 2.0 * B * hw * hw * co * 9 * ci / ms / 1e9))


if __name__ == '__main__':
    main()

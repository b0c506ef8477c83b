"""conv2d_7-shaped fprop / dgrad launches for one `ncu --set full` capture: plain TMA kernel, then halo mode."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ladder_latent_data_distribution_modelling_b200 import ops
ops.set_math_mode('bf16')
bf = torch.bfloat16
B, HW, C = 64, 128, 128
g = ops.ConvGeom(B, HW, HW, C, 3, 3, C, 1, 'same')
x = torch.randn(B, HW, HW, C, device='cuda').to(bf)
w = torch.randn(3, 3, C, C, device='cuda') * 0.05
b = torch.zeros(C, device='cuda')
y = torch.empty(B, HW, HW, C, device='cuda', dtype=bf)
dx = torch.empty(B, HW, HW, C, device='cuda', dtype=bf)
wf, wd = ops.tma_pack(w, g, ops.FPROP), ops.tma_pack(w, g, ops.DGRAD)
for en in (0, 1):
    ops.set_halo(en, 0)
    for _ in range(2):
        ops.conv2d_fprop(x, w, b, y, g, 'leaky_relu', wimg=wf)
        ops.conv2d_dgrad(x, w, dx, g, wimg=wd)
torch.cuda.synchronize()

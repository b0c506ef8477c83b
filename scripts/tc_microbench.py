"""Isolated timing of the tensor-core implicit-GEMM kernels on the bench workload's dominant layer
(mnist_fashion decoder/conv2d_3: [B,16,16,64] -> 256, 3x3 same) and a CelebA-sized layer."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ladder_latent_data_distribution_modelling_b200 import ops  # noqa: E402


def time_ms(fn, warm=3, reps=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else 'bf16'
    ops.set_math_mode(mode)
    out = []
    for name, (B, HW, Cin, Cout) in {'fashion_dec3': (1024, 16, 64, 256), 'celeba_conv7_b32': (32, 128, 128, 128),
                                     'dense_512': (1024, 1, 512, 512)}.items():
        k = 1 if HW == 1 else 3
        g = ops.ConvGeom(B, HW, HW, Cin, k, k, Cout, 1, 'same')
        x = torch.randn(B, HW, HW, Cin, device='cuda')
        w = torch.randn(k, k, Cin, Cout, device='cuda') * 0.05
        b = torch.zeros(Cout, device='cuda')
        y = torch.empty(B, g.OH, g.OW, Cout, device='cuda')
        dy = torch.randn_like(y)
        dx = torch.empty_like(x)
        dw = torch.empty_like(w)
        flops = 2.0 * B * g.OH * g.OW * Cout * k * k * Cin
        for what, fn in (('fprop', lambda: ops.conv2d_fprop(x, w, b, y, g, 'leaky_relu')),
                         ('dgrad', lambda: ops.conv2d_dgrad(dy, w, dx, g)),
                         ('wgrad', lambda: ops.conv2d_wgrad(x, dy, dw, None, g))):
            ms = time_ms(fn)
            out.append({'layer': name, 'op': what, 'io': 'fp32', 'ms': ms, 'tflops': flops / ms / 1e9})
        if mode == 'bf16' and ops.tma_supported(g, ops.FPROP):
            # bf16-resident tensors: no conversion launches, only the weight repack + the TMA-fed kernel
            x16, y16, dy16, dx16 = (t.to(torch.bfloat16) for t in (x, y, dy, dx))
            for what, fn in (('fprop', lambda: ops.conv2d_fprop(x16, w, b, y16, g, 'leaky_relu')),
                             ('dgrad', lambda: ops.conv2d_dgrad(dy16, w, dx16, g)),
                             ('wgrad', lambda: ops.conv2d_wgrad(x16, dy16, dw, None, g))):
                ms = time_ms(fn)
                out.append({'layer': name, 'op': what, 'io': 'bf16', 'ms': ms, 'tflops': flops / ms / 1e9})
    for r in out:
        print(json.dumps(r))


if __name__ == '__main__':
    main()

"""One layer, one op, a few launches -- the target of `ncu --set full` captures (see profiles/)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ladder_latent_data_distribution_modelling_b200 import ops  # noqa: E402

LAYERS = {'fashion_dec3': (1024, 16, 64, 256), 'celeba_conv7': (32, 128, 128, 128), 'celeba_conv5': (64, 32, 256, 256)}


def main():
    layer, op = sys.argv[1], sys.argv[2]
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    ops.set_math_mode('bf16')
    B, HW, Cin, Cout = LAYERS[layer]
    g = ops.ConvGeom(B, HW, HW, Cin, 3, 3, Cout, 1, 'same')
    bf = torch.bfloat16
    x = torch.randn(B, HW, HW, Cin, device='cuda').to(bf)
    w = torch.randn(3, 3, Cin, Cout, device='cuda') * 0.05
    b = torch.zeros(Cout, device='cuda')
    y = torch.empty(B, g.OH, g.OW, Cout, device='cuda', dtype=bf)
    dy = torch.randn(B, g.OH, g.OW, Cout, device='cuda').to(bf)
    dx = torch.empty_like(x)
    dw = torch.empty_like(w)
    fn = {'fprop': lambda: ops.conv2d_fprop(x, w, b, y, g, 'leaky_relu'),
          'dgrad': lambda: ops.conv2d_dgrad(dy, w, dx, g, act_out=x, act='leaky_relu'),
          'wgrad': lambda: ops.conv2d_wgrad(x, dy, dw, None, g)}[op]
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()


if __name__ == '__main__':
    main()

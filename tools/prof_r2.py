"""Small driver for the round-2 ncu captures: one ConvLSTM recurrence (persistent cluster kernels) forward + backward at
the benchmark shape, one level-1 ConvBlock forward + backward through the 4-kernel path and one through the fused
whole-block forward kernel (148 images: one per SM).  Run plain first, then under ncu with -k regex filters."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pcm_b200  # noqa: E402,F401
from pcm_b200 import ops  # noqa: E402
from pcm_b200._lib import lib  # noqa: E402
from pcm_b200.src.unet import ConvBlock  # noqa: E402

torch.manual_seed(0)
T, B, Cin, Ch, H, W = 6, 64, 128, 64, 6, 9
w = (torch.randn(4 * Ch, Cin + Ch, 3, 3, device="cuda") * 0.02).requires_grad_(True)
b = torch.zeros(4 * Ch, device="cuda", requires_grad=True)
x = torch.randn(T * B, H, W, Cin, device="cuda").bfloat16().requires_grad_(True)
gy = torch.randn(B, H, W, Ch, device="cuda").bfloat16()
for _ in range(2):
    h = ops.ConvLSTMFn.apply(x, w, b, T, B, B, 1, True)
    h.backward(gy)
torch.cuda.synchronize()

for fused in ("0", "1"):
    os.environ["PCM_BLOCK_FWD_TC"] = fused
    m = ConvBlock(7, 16).cuda()
    xs = torch.randn(148, 48, 72, 16, device="cuda").bfloat16().requires_grad_(True)
    for _ in range(2):
        out = m.forward_nhwc(xs)
        out.backward(torch.randn_like(out))
    torch.cuda.synchronize()
print("tc errors", lib()._fn["pcm_tc_error_count"](), "launches", lib().launches)

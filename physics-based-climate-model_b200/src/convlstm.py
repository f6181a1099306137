"""B200-native ConvLSTM — interface of reference src/convlstm.py (ConvLSTMCell :5-19, ConvLSTM :21-35).

Same constructors, attribute names (`.conv`, `.cell`) and forward signatures; the bodies launch the
pcm_b200 kernels (gate convolution with W.cat(x,h) split into Wx.x + Wh.h, fused sigmoid/tanh cell
update) through ops.ConvLSTMFn.  nn.Conv2d is kept purely as the parameter container so the
state_dict keys and the default initialisation match the reference bit for bit."""
import torch
import torch.nn as nn

from .. import ops
from ..config import compute_dtype


class ConvLSTMCell(nn.Module):
    def __init__(self, c_in, c_hid, kernel_size=3):
        super().__init__()
        pad = kernel_size // 2
        self.conv = nn.Conv2d(c_in + c_hid, 4 * c_hid, kernel_size, padding=pad)

    def forward(self, x, h_c):
        """x (B,c_in,H,W), (h, c) each (B,c_hid,H,W) -> (h_next, c_next)   [reference :11-19].
        Gate convolution over cat([x, h]) followed by the fused sigmoid/tanh cell update."""
        h, c = h_c
        return _cell_step(self, x, h, c)


class ConvLSTM(nn.Module):
    """Temporal depth T is handled inside: x_seq is (T, B, C, H, W)."""

    def __init__(self, c_in, c_hid, kernel_size=3):
        super().__init__()
        self.cell = ConvLSTMCell(c_in, c_hid, kernel_size)

    def forward(self, x_seq):
        T, B, C, H, W = x_seq.shape
        dt = compute_dtype()
        x = ops.StageIn.apply(x_seq.reshape(T * B, C, H, W), dt)
        h = self.forward_nhwc(x, T, B, st_t=B, st_b=1, last_only=False)        # (T,B,H,W,Ch)
        Ch = h.shape[-1]
        out = ops.StageOut.apply(h.reshape(T * B, H, W, Ch), Ch)
        return out.reshape(T, B, Ch, H, W)

    def forward_nhwc(self, x, T, B, st_t, st_b, last_only):
        conv = self.cell.conv
        return ops.ConvLSTMFn.apply(x, conv.weight, conv.bias, T, B, st_t, st_b, last_only)


def _cell_step(cell: ConvLSTMCell, x, h, c):
    """Single cell step with explicit state: gate conv over cat([x, h]) -> fused cell update.
    Differentiable w.r.t. x, h, c and the cell parameters (ops.CellStepFn)."""
    conv = cell.conv
    Ch = conv.out_channels // 4
    dt = compute_dtype()
    xh = ops.StageIn.apply(torch.cat([x, h], dim=1), dt)        # cat is layout plumbing, no arithmetic
    cs = ops.StageIn.apply(c, torch.float32, 8)
    h2, c2 = ops.CellStepFn.apply(xh, cs, conv.weight, conv.bias)
    return ops.StageOut.apply(h2, Ch), ops.StageOut.apply(c2, Ch)

"""B200-native U-Net building blocks — interface of reference src/unet.py
(SEBlock :6-17, SpatialGate :19-29, ConvBlock :32-49, Down :51-58, Up :60-69, UNet :72-109).

The torch.nn layers are kept as parameter containers only (identical state_dict keys and default
init); forward bodies run the fused pcm_b200 kernels on NHWC activations.  `forward` keeps the
reference's NCHW fp32 signature; `forward_nhwc` is the internal channels-last entry the parent
networks chain without re-staging."""
import torch
import torch.nn as nn

from .. import ops
from ..config import compute_dtype


def _pointwise(c_in, c_out):
    return nn.Conv2d(c_in, c_out, 1, bias=False)


class SEBlock(nn.Module):
    """Squeeze-and-excitation with reduction r: children `avg` (global average pool) and `fc`
    (1x1 conv C -> C/r, ReLU, 1x1 conv C/r -> C, Sigmoid; no biases)."""

    def __init__(self, c: int, r: int = 8):
        super().__init__()
        hidden = c // r
        self.avg = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Sequential(_pointwise(c, hidden), nn.ReLU(inplace=True), _pointwise(hidden, c), nn.Sigmoid())

    def forward(self, x):
        N, C, H, W = x.shape
        if C % 8 != 0:
            raise RuntimeError("pcm_b200 SEBlock needs a channel count that is a multiple of 8")
        a = ops.StageIn.apply(x, compute_dtype(), 8)
        y = ops.SEFn.apply(a, self.fc[0].weight, self.fc[2].weight)
        return ops.StageOut.apply(y, C)


class SpatialGate(nn.Module):
    """CBAM-style spatial attention: child `conv` = 7x7, 2 -> 1 channels, no bias, over [mean_C, max_C]."""

    def __init__(self):
        super().__init__()
        self.conv = nn.Conv2d(2, 1, kernel_size=7, padding=3, bias=False)

    def forward(self, x):
        N, C, H, W = x.shape
        if C % 8 != 0:
            raise RuntimeError("pcm_b200 SpatialGate needs a channel count that is a multiple of 8")
        a = ops.StageIn.apply(x, compute_dtype(), 8)
        y = ops.SpatialGateFn.apply(a, self.conv.weight)
        return ops.StageOut.apply(y, C)


def _conv_gn_silu(c_in, c_out):
    return [nn.Conv2d(c_in, c_out, 3, padding=1, bias=False), nn.GroupNorm(8, c_out), nn.SiLU(inplace=True)]


class ConvBlock(nn.Module):
    """Children: `body` = Sequential[conv3x3, GroupNorm(8), SiLU, conv3x3, GroupNorm(8), SiLU], `se`, `spat`."""

    def __init__(self, c_in: int, c_out: int):
        super().__init__()
        self.body = nn.Sequential(*_conv_gn_silu(c_in, c_out), *_conv_gn_silu(c_out, c_out))
        self.se, self.spat = SEBlock(c_out), SpatialGate()
        self.c_out = c_out

    def forward_nhwc(self, x):
        b = self.body
        return ops.ConvBlockFn.apply(x, b[0].weight, b[1].weight, b[1].bias, b[3].weight, b[4].weight, b[4].bias,
                                     self.se.fc[0].weight, self.se.fc[2].weight, self.spat.conv.weight)

    def forward(self, x):
        y = self.forward_nhwc(ops.StageIn.apply(x, compute_dtype()))
        return ops.StageOut.apply(y, self.c_out)


class Down(nn.Module):
    """MaxPool2d(2) then ConvBlock (children `pool`, `conv`)."""

    def __init__(self, c_in, c_out):
        super().__init__()
        self.pool, self.conv = nn.MaxPool2d(2), ConvBlock(c_in, c_out)

    def forward_nhwc(self, x):
        return self.conv.forward_nhwc(ops.MaxPoolFn.apply(x))

    def forward(self, x):
        y = self.forward_nhwc(ops.StageIn.apply(x, compute_dtype()))
        return ops.StageOut.apply(y, self.conv.c_out)


class Up(nn.Module):
    """ConvTranspose2d(k2, s2) of x, concatenated in front of the skip, then ConvBlock (children `up`, `conv`)."""

    def __init__(self, c_in, c_skip, c_out):
        super().__init__()
        self.up = nn.ConvTranspose2d(c_in, c_out, 2, stride=2)
        self.conv = ConvBlock(c_out + c_skip, c_out)

    def forward_nhwc(self, x, skip):
        if skip.shape[1] != 2 * x.shape[1] or skip.shape[2] != 2 * x.shape[2]:
            # the reference fails at torch.cat for grids not divisible by 8
            raise RuntimeError(f"Up: skip {tuple(skip.shape[1:3])} does not match upsampled "
                               f"{(2 * x.shape[1], 2 * x.shape[2])}")
        return self.conv.forward_nhwc(ops.UpCatFn.apply(x, skip, self.up.weight, self.up.bias))

    def forward(self, x, skip):
        dt = compute_dtype()
        y = self.forward_nhwc(ops.StageIn.apply(x, dt), ops.StageIn.apply(skip, dt))
        return ops.StageOut.apply(y, self.conv.c_out)


class UNet(nn.Module):
    """Single-frame depth-4 U-Net with attention blocks.  Children in registration order: enc1, enc2, enc3, enc4,
    bott, up3, up2, up1, head (widths base, 2base, 4base, 8base)."""

    def __init__(self, in_ch: int = 5, out_ch: int = 2, base: int = 16):
        super().__init__()
        w1, w2, w3, w4 = (base << k for k in range(4))
        self.enc1 = ConvBlock(in_ch, w1)
        for name, (a, b) in zip(("enc2", "enc3", "enc4"), ((w1, w2), (w2, w3), (w3, w4))):
            setattr(self, name, Down(a, b))
        self.bott = ConvBlock(w4, w4)
        for name, (a, b) in zip(("up3", "up2", "up1"), ((w4, w3), (w3, w2), (w2, w1))):
            setattr(self, name, Up(a, b, b))
        self.head = nn.Conv2d(w1, out_ch, kernel_size=1)

    def forward_loss(self, x, target):
        """nn.MSELoss()(self(x), target) with the head and the loss fused (the training step's form)."""
        return self.forward(x, target)

    def forward(self, x, target=None):
        a = ops.StageIn.apply(x, compute_dtype())
        s1 = self.enc1.forward_nhwc(a)
        p1, k1 = ops.PoolSkipFn.apply(s1, 1, self.up1.up.out_channels)
        s2 = self.enc2.conv.forward_nhwc(p1)
        p2, k2 = ops.PoolSkipFn.apply(s2, 1, self.up2.up.out_channels)
        s3 = self.enc3.conv.forward_nhwc(p2)
        p3, k3 = ops.PoolSkipFn.apply(s3, 1, self.up3.up.out_channels)
        s4 = self.enc4.conv.forward_nhwc(p3)
        y = self.bott.forward_nhwc(s4)
        y = self.up3.forward_nhwc(y, k3)
        y = self.up2.forward_nhwc(y, k2)
        y = self.up1.forward_nhwc(y, k1)
        return ops.head_or_loss(y, self.head.weight, self.head.bias, target)

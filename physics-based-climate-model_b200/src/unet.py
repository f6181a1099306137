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


class SEBlock(nn.Module):
    """Channel-wise squeeze-and-excitation (ratio = 8)."""

    def __init__(self, c: int, r: int = 8):
        super().__init__()
        self.avg = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Sequential(
            nn.Conv2d(c, c // r, 1, bias=False), nn.ReLU(inplace=True),
            nn.Conv2d(c // r, c, 1, bias=False), nn.Sigmoid()
        )

    def forward(self, x):
        N, C, H, W = x.shape
        if C % 8 != 0:
            raise RuntimeError("pcm_b200 SEBlock needs a channel count that is a multiple of 8")
        a = ops.StageIn.apply(x, compute_dtype(), 8)
        y = ops.SEFn.apply(a, self.fc[0].weight, self.fc[2].weight)
        return ops.StageOut.apply(y, C)


class SpatialGate(nn.Module):
    """7x7 conv on concatenated mean- & max-over-channel maps (CBAM style)."""

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


class ConvBlock(nn.Module):
    def __init__(self, c_in: int, c_out: int):
        super().__init__()
        self.body = nn.Sequential(
            nn.Conv2d(c_in, c_out, 3, padding=1, bias=False),
            nn.GroupNorm(8, c_out), nn.SiLU(inplace=True),
            nn.Conv2d(c_out, c_out, 3, padding=1, bias=False),
            nn.GroupNorm(8, c_out), nn.SiLU(inplace=True),
        )
        self.se = SEBlock(c_out)
        self.spat = SpatialGate()
        self.c_out = c_out

    def forward_nhwc(self, x):
        b = self.body
        return ops.ConvBlockFn.apply(x, b[0].weight, b[1].weight, b[1].bias, b[3].weight, b[4].weight, b[4].bias,
                                     self.se.fc[0].weight, self.se.fc[2].weight, self.spat.conv.weight)

    def forward(self, x):
        y = self.forward_nhwc(ops.StageIn.apply(x, compute_dtype()))
        return ops.StageOut.apply(y, self.c_out)


class Down(nn.Module):
    def __init__(self, c_in, c_out):
        super().__init__()
        self.pool = nn.MaxPool2d(2)
        self.conv = ConvBlock(c_in, c_out)

    def forward_nhwc(self, x):
        return self.conv.forward_nhwc(ops.MaxPoolFn.apply(x))

    def forward(self, x):
        y = self.forward_nhwc(ops.StageIn.apply(x, compute_dtype()))
        return ops.StageOut.apply(y, self.conv.c_out)


class Up(nn.Module):
    def __init__(self, c_in, c_skip, c_out):
        super().__init__()
        self.up = nn.ConvTranspose2d(c_in, c_out, 2, stride=2)
        self.conv = ConvBlock(c_out + c_skip, c_out)

    def forward_nhwc(self, x, skip):
        if skip.shape[1] != 2 * x.shape[1] or skip.shape[2] != 2 * x.shape[2]:
            # the reference fails at torch.cat (src/unet.py:68) for grids not divisible by 8
            raise RuntimeError(f"Up: skip {tuple(skip.shape[1:3])} does not match upsampled "
                               f"{(2 * x.shape[1], 2 * x.shape[2])}")
        return self.conv.forward_nhwc(ops.UpCatFn.apply(x, skip, self.up.weight, self.up.bias))

    def forward(self, x, skip):
        dt = compute_dtype()
        y = self.forward_nhwc(ops.StageIn.apply(x, dt), ops.StageIn.apply(skip, dt))
        return ops.StageOut.apply(y, self.conv.c_out)


class UNet(nn.Module):
    """Depth-4 UNet with attention (single frame)."""

    def __init__(self, in_ch: int = 5, out_ch: int = 2, base: int = 16):
        super().__init__()
        self.enc1 = ConvBlock(in_ch, base)
        self.enc2 = Down(base, base * 2)
        self.enc3 = Down(base * 2, base * 4)
        self.enc4 = Down(base * 4, base * 8)
        self.bott = ConvBlock(base * 8, base * 8)
        self.up3 = Up(base * 8, base * 4, base * 4)
        self.up2 = Up(base * 4, base * 2, base * 2)
        self.up1 = Up(base * 2, base, base)
        self.head = nn.Conv2d(base, out_ch, kernel_size=1)

    def forward(self, x):
        a = ops.StageIn.apply(x, compute_dtype())
        s1 = self.enc1.forward_nhwc(a)
        p1, k1 = ops.PoolSkipFn.apply(s1, 1)
        s2 = self.enc2.conv.forward_nhwc(p1)
        p2, k2 = ops.PoolSkipFn.apply(s2, 1)
        s3 = self.enc3.conv.forward_nhwc(p2)
        p3, k3 = ops.PoolSkipFn.apply(s3, 1)
        s4 = self.enc4.conv.forward_nhwc(p3)
        y = self.bott.forward_nhwc(s4)
        y = self.up3.forward_nhwc(y, k3)
        y = self.up2.forward_nhwc(y, k2)
        y = self.up1.forward_nhwc(y, k1)
        return ops.HeadFn.apply(y, self.head.weight, self.head.bias)

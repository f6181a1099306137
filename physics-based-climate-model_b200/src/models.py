"""Model factory and the SimpleCNN family — interface of reference src/models.py.

`get_model(cfg)` (:7-38): unchanged semantics — keyed on cfg.model.type, `in_ch=7` hard-coded for
unet_convlstm_attention (:26), ValueError on unknown types (:37).
`ResidualBlock` (:44-73) / `SimpleCNN` (:76-123): same constructors and registered modules (state_dict keys, BN
buffers and default init match); bodies run the pcm_b200 kernels on NHWC activations — conv(+bias) on the tcgen05
path, BatchNorm with batch statistics (+ReLU, + the residual add of :70-71) in two HBM passes, Dropout2d with a
counter-based channel mask."""
import torch.nn as nn

from .. import ops, ops_nn
from ..config import compute_dtype
from .cnn_transformer import CNNTransformer
from .unet import UNet
from .unet_convlstm_attention import AttUNetConvLSTM


def _get(cfg, name):
    return cfg[name] if isinstance(cfg, dict) else getattr(cfg, name)


def get_model(cfg):
    model_cfg, data_cfg = _get(cfg, "model"), _get(cfg, "data")
    mtype = _get(model_cfg, "type")
    n_in, n_out = len(_get(data_cfg, "input_vars")), len(_get(data_cfg, "output_vars"))
    if mtype == "SimpleCNN":
        kwargs = {k: v for k, v in dict(model_cfg).items() if k != "type"}
        return SimpleCNN(n_input_channels=n_in, n_output_channels=n_out, **kwargs)
    elif mtype == "cnn_transformer":
        return CNNTransformer(in_channels=n_in, out_channels=n_out, embed_dim=_get(model_cfg, "embed_dim"),
                              depth=_get(model_cfg, "depth"), n_heads=_get(model_cfg, "n_heads"),
                              mlp_dim=_get(model_cfg, "mlp_dim"), dropout=_get(model_cfg, "dropout"))
    elif mtype == "unet_convlstm_attention":
        return AttUNetConvLSTM(in_ch=7, out_ch=n_out, base=_get(model_cfg, "base_channels"))
    elif mtype == "unet":
        return UNet(in_ch=n_in, out_ch=n_out, base=_get(model_cfg, "base_channels"))
    else:
        raise ValueError(f"Unknown model type: {mtype}")


def _conv_bn(x, conv: nn.Conv2d, bn: nn.BatchNorm2d, relu: bool, res=None):
    stride = conv.stride[0]
    y = ops_nn.Conv2dFn.apply(x, conv.weight, conv.bias, stride, conv.padding[0], False)
    return ops_nn.batch_norm(y, bn, res=res, relu=relu)


class ResidualBlock(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1):
        super().__init__()
        self.conv1 = nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride, padding=kernel_size // 2)
        self.bn1 = nn.BatchNorm2d(out_channels)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(out_channels, out_channels, kernel_size, padding=kernel_size // 2)
        self.bn2 = nn.BatchNorm2d(out_channels)
        self.skip = nn.Sequential()
        if stride != 1 or in_channels != out_channels:
            self.skip = nn.Sequential(
                nn.Conv2d(in_channels, out_channels, kernel_size=1, stride=stride), nn.BatchNorm2d(out_channels)
            )
        self.out_channels = out_channels

    def forward_nhwc(self, x):
        x, xs = ops_nn.fork(x)
        out = _conv_bn(x, self.conv1, self.bn1, relu=True)
        identity = _conv_bn(xs, self.skip[0], self.skip[1], relu=False) if len(self.skip) else xs
        # bn2 + residual add + ReLU in one pass (:68-71)
        return _conv_bn(out, self.conv2, self.bn2, relu=True, res=identity)

    def forward(self, x):
        y = self.forward_nhwc(ops.StageIn.apply(x, compute_dtype()))
        return ops.StageOut.apply(y, self.out_channels)


class SimpleCNN(nn.Module):
    def __init__(self, n_input_channels, n_output_channels, kernel_size=3, init_dim=64, depth=4, dropout_rate=0.2):
        super().__init__()
        self.initial = nn.Sequential(
            nn.Conv2d(n_input_channels, init_dim, kernel_size=kernel_size, padding=kernel_size // 2),
            nn.BatchNorm2d(init_dim),
            nn.ReLU(inplace=True),
        )
        self.res_blocks = nn.ModuleList()
        current_dim = init_dim
        for i in range(depth):
            out_dim = current_dim * 2 if i < depth - 1 else current_dim
            self.res_blocks.append(ResidualBlock(current_dim, out_dim))
            if i < depth - 1:
                current_dim *= 2
        self.dropout = nn.Dropout2d(dropout_rate)
        self.final = nn.Sequential(
            nn.Conv2d(current_dim, current_dim // 2, kernel_size=kernel_size, padding=kernel_size // 2),
            nn.BatchNorm2d(current_dim // 2),
            nn.ReLU(inplace=True),
            nn.Conv2d(current_dim // 2, n_output_channels, kernel_size=1),
        )

    def forward(self, x):
        a = ops.StageIn.apply(x, compute_dtype())
        a = _conv_bn(a, self.initial[0], self.initial[1], relu=True)
        for blk in self.res_blocks:
            a = blk.forward_nhwc(a)
        if self.training and self.dropout.p > 0.0:
            a = ops_nn.Dropout2dFn.apply(a, float(self.dropout.p), ops_nn.next_seed())
        a = _conv_bn(a, self.final[0], self.final[1], relu=True)
        return ops.HeadFn.apply(a, self.final[3].weight, self.final[3].bias)

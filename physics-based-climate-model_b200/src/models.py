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


def _build_simple_cnn(model_cfg, n_in, n_out):
    extra = {k: v for k, v in dict(model_cfg).items() if k != "type"}
    return SimpleCNN(n_input_channels=n_in, n_output_channels=n_out, **extra)


def _build_cnn_transformer(model_cfg, n_in, n_out):
    keys = ("embed_dim", "depth", "n_heads", "mlp_dim", "dropout")
    return CNNTransformer(in_channels=n_in, out_channels=n_out, **{k: _get(model_cfg, k) for k in keys})


def _build_att_unet(model_cfg, n_in, n_out):
    # the reference hard-codes 7 input channels here (5 forcings + sin/cos month), whatever cfg.data lists
    return AttUNetConvLSTM(in_ch=7, out_ch=n_out, base=_get(model_cfg, "base_channels"))


def _build_unet(model_cfg, n_in, n_out):
    return UNet(in_ch=n_in, out_ch=n_out, base=_get(model_cfg, "base_channels"))


_BUILDERS = {"SimpleCNN": _build_simple_cnn, "cnn_transformer": _build_cnn_transformer,
             "unet_convlstm_attention": _build_att_unet, "unet": _build_unet}


def get_model(cfg):
    """cfg.model.type selects the architecture; channel counts come from cfg.data.{input_vars,output_vars}."""
    model_cfg, data_cfg = _get(cfg, "model"), _get(cfg, "data")
    kind = _get(model_cfg, "type")
    if kind not in _BUILDERS:
        raise ValueError(f"Unknown model type: {kind}")
    return _BUILDERS[kind](model_cfg, len(_get(data_cfg, "input_vars")), len(_get(data_cfg, "output_vars")))


def _conv_bn(x, conv: nn.Conv2d, bn: nn.BatchNorm2d, relu: bool, res=None):
    y = ops_nn.Conv2dFn.apply(x, conv.weight, conv.bias, conv.stride[0], conv.padding[0], False)
    return ops_nn.batch_norm(y, bn, res=res, relu=relu)


def _same_conv(c_in, c_out, k, stride=1):
    return nn.Conv2d(c_in, c_out, k, stride=stride, padding=k // 2)


class ResidualBlock(nn.Module):
    """conv-BN-ReLU-conv-BN + (identity | 1x1 conv-BN) -> ReLU.  Registered children, in order: conv1, bn1, relu,
    conv2, bn2, skip — the parameter containers of the reference block."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1):
        super().__init__()
        k, projected = kernel_size, (stride != 1 or in_channels != out_channels)
        self.conv1, self.bn1 = _same_conv(in_channels, out_channels, k, stride), nn.BatchNorm2d(out_channels)
        self.relu = nn.ReLU(inplace=True)
        self.conv2, self.bn2 = _same_conv(out_channels, out_channels, k), nn.BatchNorm2d(out_channels)
        # (built only when needed: constructing it unconditionally would consume RNG draws the reference does not make)
        self.skip = nn.Sequential(*([nn.Conv2d(in_channels, out_channels, kernel_size=1, stride=stride),
                                     nn.BatchNorm2d(out_channels)] if projected else []))
        self.out_channels = out_channels

    def forward_nhwc(self, x):
        x, xs = ops_nn.fork(x)
        out = _conv_bn(x, self.conv1, self.bn1, relu=True)
        identity = _conv_bn(xs, self.skip[0], self.skip[1], relu=False) if len(self.skip) else xs
        # bn2 + residual add + ReLU in one pass
        return _conv_bn(out, self.conv2, self.bn2, relu=True, res=identity)

    def forward(self, x):
        y = self.forward_nhwc(ops.StageIn.apply(x, compute_dtype()))
        return ops.StageOut.apply(y, self.out_channels)


class SimpleCNN(nn.Module):
    """Stem (conv-BN-ReLU) -> `depth` residual blocks (width doubles in all but the last) -> Dropout2d -> conv-BN-ReLU
    -> 1x1 conv.  Children: initial, res_blocks, dropout, final."""

    def __init__(self, n_input_channels, n_output_channels, kernel_size=3, init_dim=64, depth=4, dropout_rate=0.2):
        super().__init__()
        k = kernel_size
        self.initial = nn.Sequential(_same_conv(n_input_channels, init_dim, k), nn.BatchNorm2d(init_dim), nn.ReLU(inplace=True))
        widths = [init_dim * 2 ** min(i, depth - 1) for i in range(depth + 1)]       # 64, 128, 256, 512, 512
        self.res_blocks = nn.ModuleList(ResidualBlock(a, b) for a, b in zip(widths[:-1], widths[1:]))
        top = widths[-1] if depth > 0 else init_dim
        self.dropout = nn.Dropout2d(dropout_rate)
        self.final = nn.Sequential(_same_conv(top, top // 2, k), nn.BatchNorm2d(top // 2), nn.ReLU(inplace=True),
                                   nn.Conv2d(top // 2, n_output_channels, kernel_size=1))

    def forward_loss(self, x, target):
        """nn.MSELoss()(self(x), target) with the final 1x1 convolution and the loss fused (the training step's form)."""
        return self.forward(x, target)

    def forward(self, x, target=None):
        a = ops.StageIn.apply(x, compute_dtype())
        a = _conv_bn(a, self.initial[0], self.initial[1], relu=True)
        for blk in self.res_blocks:
            a = blk.forward_nhwc(a)
        if self.training and self.dropout.p > 0.0:
            a = ops_nn.Dropout2dFn.apply(a, float(self.dropout.p), ops_nn.next_seed())
        a = _conv_bn(a, self.final[0], self.final[1], relu=True)
        return ops.head_or_loss(a, self.final[3].weight, self.final[3].bias, target)

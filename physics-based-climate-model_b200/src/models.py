"""Model factory — interface of reference src/models.py:7-38 (`get_model(cfg)`), unchanged
semantics: keyed on cfg.model.type, `in_ch=7` hard-coded for unet_convlstm_attention (:26),
ValueError on unknown types (:37)."""
import torch.nn as nn

from .unet import UNet
from .unet_convlstm_attention import AttUNetConvLSTM


def _get(cfg, name):
    return cfg[name] if isinstance(cfg, dict) else getattr(cfg, name)


def get_model(cfg):
    model_cfg, data_cfg = _get(cfg, "model"), _get(cfg, "data")
    mtype = _get(model_cfg, "type")
    n_in, n_out = len(_get(data_cfg, "input_vars")), len(_get(data_cfg, "output_vars"))
    if mtype == "SimpleCNN":
        from .simple_cnn import SimpleCNN
        kwargs = {k: v for k, v in dict(model_cfg).items() if k != "type"}
        return SimpleCNN(n_input_channels=n_in, n_output_channels=n_out, **kwargs)
    elif mtype == "cnn_transformer":
        from .cnn_transformer import CNNTransformer
        return CNNTransformer(in_channels=n_in, out_channels=n_out, embed_dim=_get(model_cfg, "embed_dim"),
                              depth=_get(model_cfg, "depth"), n_heads=_get(model_cfg, "n_heads"),
                              mlp_dim=_get(model_cfg, "mlp_dim"), dropout=_get(model_cfg, "dropout"))
    elif mtype == "unet_convlstm_attention":
        return AttUNetConvLSTM(in_ch=7, out_ch=n_out, base=_get(model_cfg, "base_channels"))
    elif mtype == "unet":
        return UNet(in_ch=n_in, out_ch=n_out, base=_get(model_cfg, "base_channels"))
    else:
        raise ValueError(f"Unknown model type: {mtype}")

"""B200-native CNNTransformer — interface of reference src/cnn_transformer.py (:5-54).

Same constructor (in_channels=5, out_channels=2, embed_dim=128, depth=4, n_heads=4, mlp_dim=256, dropout=0.1),
same registered modules (`encoder`, `pos_embedding`, `transformer`, `decoder`) so state_dict keys and the default
initialisation match the reference; nn.TransformerEncoder is kept purely as the parameter container.  The forward
body runs the pcm_b200 kernels on channels-last activations: the (B, E, 12, 18) feature map in NHWC *is* the
(B, 216, E) token matrix, so the reference's flatten/transpose/view round trips (:47,:52) cost nothing.

Per encoder layer (post-norm, batch_first — torch's nn.TransformerEncoderLayer, :25-31):
    x = LN1(x + drop(out_proj(MHA(in_proj(x)))));   x = LN2(x + drop(linear2(drop(relu(linear1(x))))))
Linear layers -> tcgen05 1x1 path; attention / LayerNorm -> csrc/transformer.cu; dropout masks are counter-based
(torch's RNG stream is not reproduced: with dropout > 0 results match the reference in distribution only)."""
import torch
import torch.nn as nn

from .. import ops, ops_nn
from ..config import compute_dtype


class CNNTransformer(nn.Module):
    def __init__(self, in_channels=5, out_channels=2, embed_dim=128, depth=4, n_heads=4, mlp_dim=256, dropout=0.1):
        super().__init__()
        # CNN Encoder: (48, 72) -> (12, 18)
        self.encoder = nn.Sequential(
            nn.Conv2d(in_channels, embed_dim // 2, kernel_size=3, stride=2, padding=1),
            nn.ReLU(),
            nn.Conv2d(embed_dim // 2, embed_dim, kernel_size=3, stride=2, padding=1),
            nn.ReLU(),
        )
        self.height = 12
        self.width = 18
        self.num_tokens = self.height * self.width
        self.embed_dim = embed_dim
        self.pos_embedding = nn.Parameter(torch.randn(1, self.num_tokens, embed_dim))
        encoder_layer = nn.TransformerEncoderLayer(d_model=embed_dim, nhead=n_heads, dim_feedforward=mlp_dim,
                                                   dropout=dropout, batch_first=True)
        self.transformer = nn.TransformerEncoder(encoder_layer, num_layers=depth)
        # CNN Decoder: back to 48x72
        self.decoder = nn.Sequential(
            nn.ConvTranspose2d(embed_dim, embed_dim // 2, kernel_size=2, stride=2),
            nn.ReLU(),
            nn.ConvTranspose2d(embed_dim // 2, embed_dim // 4, kernel_size=2, stride=2),
            nn.ReLU(),
            nn.Conv2d(embed_dim // 4, out_channels, kernel_size=1),
        )
        self.n_heads = n_heads
        self.p_drop = dropout
        if embed_dim % (8 * n_heads) != 0 or (embed_dim // 4) % 8 != 0:
            raise RuntimeError("pcm_b200 CNNTransformer needs embed_dim divisible by 8*n_heads and by 32")

    def _layer(self, x, lyr, from_norm=False):
        """from_norm: x is the output of the previous layer's norm2 — the residual fork then hands its two gradients to that
        LayerNorm's backward kernel instead of adding them in a launch of its own (ops_nn.ForkFn lazy)."""
        tr, p = self.training, self.p_drop
        attn = lyr.self_attn
        x, xr = ops_nn.fork(x, lazy=from_norm)
        qkv = ops_nn.LinearFn.apply(x, attn.in_proj_weight, attn.in_proj_bias, False)
        a = ops_nn.MHAFn.apply(qkv, self.n_heads, float(p) if tr else 0.0, ops_nn.next_seed())
        a = ops_nn.LinearFn.apply(a, attn.out_proj.weight, attn.out_proj.bias, False)
        pd = float(p) if tr else 0.0            # sub-layer dropouts are fused into the residual-add + LayerNorm kernels
        x = ops_nn.AddLayerNormFn.apply(xr, a, lyr.norm1.weight, lyr.norm1.bias, pd, ops_nn.next_seed() if pd > 0 else 0)
        x, xr = ops_nn.fork(x, lazy=True)                 # x = norm1's output
        pf = float(p) if tr else 0.0            # linear1 -> ReLU -> dropout in one launch (GEMM store epilogue)
        f = ops_nn.LinearFn.apply(x, lyr.linear1.weight, lyr.linear1.bias, True, pf, ops_nn.next_seed() if pf > 0 else 0)
        f = ops_nn.LinearFn.apply(f, lyr.linear2.weight, lyr.linear2.bias, False)
        return ops_nn.AddLayerNormFn.apply(xr, f, lyr.norm2.weight, lyr.norm2.bias, pd, ops_nn.next_seed() if pd > 0 else 0)

    def forward_loss(self, x, target):
        """nn.MSELoss()(self(x), target) with the final 1x1 convolution and the loss fused (the training step's form)."""
        return self.forward(x, target)

    def forward(self, x, target=None):
        B = x.size(0)
        e = self.encoder
        a = ops.StageIn.apply(x, compute_dtype())
        a = ops_nn.Conv2dFn.apply(a, e[0].weight, e[0].bias, 2, 1, True)
        a = ops_nn.Conv2dFn.apply(a, e[2].weight, e[2].bias, 2, 1, True)          # (B, Hh, Ww, E)
        Hh, Ww, E = a.shape[1], a.shape[2], a.shape[3]
        if Hh * Ww != self.num_tokens:
            raise RuntimeError(f"CNNTransformer: expected a {4 * self.height}x{4 * self.width} grid "
                               f"({self.num_tokens} tokens), got {Hh * Ww}")
        t = ops_nn.AddPosFn.apply(a.reshape(B, Hh * Ww, E), self.pos_embedding)
        for i, lyr in enumerate(self.transformer.layers):
            t = self._layer(t, lyr, from_norm=i > 0)
        d = self.decoder
        y = t.reshape(B, Hh, Ww, E)
        y = ops_nn.ConvT2x2Fn.apply(y, d[0].weight, d[0].bias, True)
        y = ops_nn.ConvT2x2Fn.apply(y, d[2].weight, d[2].bias, True)
        return ops.head_or_loss(y, d[4].weight, d[4].bias, target)

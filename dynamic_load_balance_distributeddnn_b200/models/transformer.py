"""Word-level Transformer language model (wikitext-2 configuration).

Capabilities of reference ``Net/Transformer.py:8-95`` as instantiated by ``dbs.py:337-343,360-362``:
``Embedding(33278,200)·√d → sinusoidal PE → dropout → 2 × post-norm encoder layer (2 heads ⇒
head_dim 100, ReLU FFN 200→200→200, causal mask) → Linear(200→33278) → log_softmax``;
13 828 478 parameters in 27 tensors with ``torch.nn.Transformer*``-compatible names
(``transformer_encoder.layers.N.self_attn.in_proj_weight`` …) so checkpoints interchange.

The encoder layer is implemented directly (not ``nn.TransformerEncoderLayer``) so each sub-block is
one fused op: QKV projection, causal attention over the whole S=35 tile, out-proj + residual +
LayerNorm, FFN + residual + LayerNorm (``ops/transformer_ops.py``).  Training uses
``forward_loss`` — decoder projection + log-softmax + NLL fused and chunked over the vocabulary so
the [S·B, 33278] logits are never materialised (K16) — while ``forward`` returns the full
log-probabilities like the reference.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from .layers import Linear


class PositionalEncoding(nn.Module):
    def __init__(self, d_model, dropout=0.1, max_len=5000):
        super().__init__()
        self.p = dropout
        position = torch.arange(max_len, dtype=torch.float32).unsqueeze(1)
        div = torch.exp(torch.arange(0, d_model, 2, dtype=torch.float32) * (-math.log(10000.0) / d_model))
        pe = torch.zeros(max_len, 1, d_model)
        pe[:, 0, 0::2] = torch.sin(position * div)
        pe[:, 0, 1::2] = torch.cos(position * div)
        self.register_buffer("pe", pe)

    def forward(self, x):
        return F.dropout(x + self.pe[:x.size(0)].to(x.dtype), self.p, self.training)


class SelfAttention(nn.Module):
    """Parameters named like ``nn.MultiheadAttention``."""

    def __init__(self, d_model, nhead, dropout):
        super().__init__()
        self.d_model, self.nhead, self.p = d_model, nhead, dropout
        self.in_proj_weight = nn.Parameter(torch.empty(3 * d_model, d_model))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * d_model))
        self.out_proj = Linear(d_model, d_model)
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.zeros_(self.out_proj.bias)

    def forward(self, x):                                   # x: [S, B, D]
        s, b, d = x.shape
        h, hd = self.nhead, d // self.nhead
        qkv = ops.linear(x, self.in_proj_weight, self.in_proj_bias)          # [S,B,3D]
        from ..ops import attention as fused_attn
        if fused_attn.supported(qkv, h):
            # whole-sequence fused kernel: scores, causal softmax, dropout and PV stay on chip
            return self.out_proj(fused_attn.causal_attention_packed(qkv, h, self.p if self.training else 0.0))
        q, k, v = qkv.view(s, b, 3, h, hd).permute(2, 1, 3, 0, 4)            # each [B,H,S,hd]
        o = ops.causal_attention(q, k, v, self.p if self.training else 0.0)  # [B,H,S,hd]
        o = o.permute(2, 0, 1, 3).reshape(s, b, d)
        return self.out_proj(o)


class EncoderLayer(nn.Module):
    def __init__(self, d_model, nhead, dim_ff, dropout):
        super().__init__()
        self.self_attn = SelfAttention(d_model, nhead, dropout)
        self.linear1 = Linear(d_model, dim_ff)
        self.linear2 = Linear(dim_ff, d_model)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        for p in list(self.norm1.parameters()) + list(self.norm2.parameters()):
            p._dlb_keep_fp32 = True
        self.p = dropout

    def forward(self, x):
        a = F.dropout(self.self_attn(x), self.p, self.training)
        x = ops.add_layer_norm(x, a, self.norm1.weight, self.norm1.bias, self.norm1.eps)
        f = self.linear2(F.dropout(F.relu(self.linear1(x)), self.p, self.training))
        f = F.dropout(f, self.p, self.training)
        return ops.add_layer_norm(x, f, self.norm2.weight, self.norm2.bias, self.norm2.eps)


class _Encoder(nn.Module):
    def __init__(self, make_layer, n):
        super().__init__()
        self.layers = nn.ModuleList([make_layer() for _ in range(n)])

    def forward(self, x):
        for layer in self.layers:
            x = layer(x)
        return x


class TransformerModel(nn.Module):
    def __init__(self, ntoken=33278, ninp=200, nhead=2, nhid=200, nlayers=2, dropout=0.5):
        super().__init__()
        self.model_type = "Transformer"
        self.ninp, self.ntoken = ninp, ntoken
        self.pos_encoder = PositionalEncoding(ninp, dropout)
        self.transformer_encoder = _Encoder(lambda: EncoderLayer(ninp, nhead, nhid, dropout), nlayers)
        self.encoder = nn.Embedding(ntoken, ninp)
        self.decoder = Linear(ninp, ntoken)
        nn.init.uniform_(self.encoder.weight, -0.1, 0.1)
        nn.init.uniform_(self.decoder.weight, -0.1, 0.1)

    def features(self, src):                                # src: int64 [S, B]
        # gather * sqrt(d) + positional table + dropout: one fused kernel on CUDA (ops/embedding.py), the plain ops elsewhere
        from ..ops import embedding as fused_embed
        x = fused_embed.embed_pe_dropout(src, self.encoder.weight, self.pos_encoder.pe, self.pos_encoder.p, self.training)
        return self.transformer_encoder(x)

    def forward(self, src, has_mask=True):
        """→ log-probabilities [S, B, ntoken] (reference semantics)."""
        out = self.decoder(self.features(src))
        return F.log_softmax(out.float(), dim=-1)

    def forward_loss(self, src, target):
        """Mean NLL over all S·B positions without materialising the logits."""
        feats = self.features(src)
        return ops.linear_cross_entropy(feats.reshape(-1, self.ninp), self.decoder.weight, self.decoder.bias,
                                        target.reshape(-1))

"""Transformer building blocks: causal attention over a short sequence, residual-add + LayerNorm,
and the fused decoder projection + log-softmax + NLL loss.

Each op has a PyTorch reference path; CUDA fast paths are registered by ``ops/lm_native.py`` once
the corresponding kernels are built.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from . import _native as nat


def causal_attention_reference(q, k, v, dropout_p: float = 0.0):
    """q,k,v: [B,H,S,hd] → [B,H,S,hd]; additive causal mask like reference Net/Transformer.py:71-74."""
    s = q.shape[-2]
    scores = (q.float() @ k.float().transpose(-1, -2)) / math.sqrt(q.shape[-1])
    mask = torch.ones(s, s, dtype=torch.bool, device=q.device).tril()
    scores = scores.masked_fill(~mask, float("-inf"))
    p = torch.softmax(scores, dim=-1)
    if dropout_p > 0:
        p = F.dropout(p, dropout_p, True)
    return (p @ v.float()).to(q.dtype)


def causal_attention(q, k, v, dropout_p: float = 0.0):
    if q.is_cuda:
        return F.scaled_dot_product_attention(q, k, v, dropout_p=dropout_p, is_causal=True)
    return causal_attention_reference(q, k, v, dropout_p)


def add_layer_norm_reference(x, residual, weight, bias, eps: float = 1e-5):
    y = F.layer_norm((x.float() + residual.float()), (x.shape[-1],), weight.float(), bias.float(), eps)
    return y.to(x.dtype)


def add_layer_norm(x, residual, weight, bias, eps: float = 1e-5):
    """LayerNorm(x + residual) (post-norm encoder layer, reference nn.TransformerEncoderLayer)."""
    if x.is_cuda and nat.available():
        from . import lm_native
        if lm_native.has_add_layer_norm():
            return lm_native.add_layer_norm(x, residual, weight, bias, eps)
    return add_layer_norm_reference(x, residual, weight, bias, eps)


def linear_cross_entropy_reference(feats, weight, bias, target):
    logits = F.linear(feats.float(), weight.float(), bias.float() if bias is not None else None)
    return F.cross_entropy(logits, target)


class _ChunkedLinearCE(torch.autograd.Function):
    """Vocabulary projection + log-softmax + NLL, chunked over ROWS so at most
    ``chunk × V`` logits exist at a time; the backward recomputes nothing: d(logits) is formed in
    the forward per chunk and immediately contracted into d(feats) and d(weight)."""

    @staticmethod
    def forward(ctx, feats, weight, bias, target, chunk):
        t, d = feats.shape
        v = weight.shape[0]
        w = weight if weight.dtype == feats.dtype else weight.to(feats.dtype)
        dfeats = torch.empty_like(feats)
        dweight = torch.zeros(v, d, dtype=torch.float32, device=feats.device)
        dbias = torch.zeros(v, dtype=torch.float32, device=feats.device) if bias is not None else None
        loss = torch.zeros((), dtype=torch.float32, device=feats.device)
        inv_t = 1.0 / t
        for s in range(0, t, chunk):
            e = min(t, s + chunk)
            logits = (feats[s:e] @ w.t()).float()
            if bias is not None:
                logits += bias.float()
            lse = torch.logsumexp(logits, dim=-1, keepdim=True)
            tgt = target[s:e].unsqueeze(1)
            loss += (lse - logits.gather(1, tgt)).sum() * inv_t
            p = torch.exp(logits - lse)                      # softmax
            p.scatter_add_(1, tgt, torch.full_like(tgt, -1.0, dtype=p.dtype))
            p *= inv_t                                        # d loss / d logits
            pl = p.to(feats.dtype)
            dfeats[s:e] = pl @ w
            dweight.addmm_(pl.t().float() if feats.dtype == torch.float32 else pl.t().float(), feats[s:e].float())
            if dbias is not None:
                dbias += p.sum(0)
        ctx.save_for_backward(dfeats, dweight, dbias)
        ctx.dtypes = (weight.dtype, bias.dtype if bias is not None else None)
        return loss

    @staticmethod
    def backward(ctx, g):
        dfeats, dweight, dbias = ctx.saved_tensors
        wdt, bdt = ctx.dtypes
        return (dfeats * g.to(dfeats.dtype), (dweight * g).to(wdt),
                (dbias * g).to(bdt) if dbias is not None else None, None, None)


def linear_cross_entropy(feats, weight, bias, target, chunk: int = 1024):
    if feats.is_cuda:
        if nat.available():
            from . import lm_native
            if lm_native.has_linear_ce():
                return lm_native.linear_cross_entropy(feats, weight, bias, target)
        return _ChunkedLinearCE.apply(feats, weight, bias, target, chunk)
    return linear_cross_entropy_reference(feats, weight, bias, target)

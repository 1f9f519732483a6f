"""Oracle: TransformerLM as a beam-search scorer (LM shallow fusion, SURVEY.md 8f-3 -- groundwork for a later round).  TEST INFRASTRUCTURE.

Reference: espnet2/lm/transformer_lm.py:13-133 (embed -> legacy Encoder(input_layer="linear") -> Linear; batch_score with a per-layer
cache), legacy/nets/pytorch_backend/transformer/encoder.py:132-139,366-392 (Linear + nn.LayerNorm(eps 1e-5) + ReLU + PositionalEncoding;
forward_one_step), encoder_layer.py:65-126, mask.py:9-22.  The cache only avoids recomputation: with a causal mask the layer outputs of
earlier positions do not change, so scoring the whole prefix each step gives the same last-position distribution.
Weights: the ESPnetLanguageModel state_dict (keys "lm.embed.weight", "lm.encoder.*", "lm.decoder.*").
"""
import math

import torch
import torch.nn.functional as F

from .encoder import _lin, _ln
from .search import _pos_enc


class OracleLM:
    def __init__(self, w, heads, num_blocks):
        self.w = {k: v.detach().float().cpu() for k, v in w.items()}
        self.heads, self.n = heads, num_blocks

    def _attn(self, x, mask, pfx):
        """MultiHeadedAttention default branch with the target mask: masked_fill(min) -> softmax -> masked_fill(0) (attention.py:121-151)."""
        n, L, d = x.shape
        dk = d // self.heads
        w = self.w
        q = _lin(x, w, pfx + ".linear_q").view(n, L, self.heads, dk).transpose(1, 2)
        k = _lin(x, w, pfx + ".linear_k").view(n, L, self.heads, dk).transpose(1, 2)
        v = _lin(x, w, pfx + ".linear_v").view(n, L, self.heads, dk).transpose(1, 2)
        sc = q @ k.transpose(-2, -1) / math.sqrt(dk)
        m = ~mask.unsqueeze(1)
        att = torch.softmax(sc.masked_fill(m, torch.finfo(sc.dtype).min), dim=-1).masked_fill(m, 0.0)
        return _lin((att @ v).transpose(1, 2).contiguous().view(n, L, d), w, pfx + ".linear_out")

    @torch.no_grad()
    def batch_score(self, ys):
        """ys (n, len) int64 prefixes (leading sos) -> log-probabilities of the next token (n, V)."""
        w = self.w
        x = F.embedding(ys, w["lm.embed.weight"])
        x = _lin(x, w, "lm.encoder.embed.0")
        x = F.layer_norm(x, (x.shape[-1],), w["lm.encoder.embed.1.weight"], w["lm.encoder.embed.1.bias"], 1e-5)   # torch.nn.LayerNorm default
        x = torch.relu(x)
        d = x.shape[-1]
        x = x * math.sqrt(d) + _pos_enc(ys.shape[1], d)
        L = ys.shape[1]
        mask = (ys != 0).unsqueeze(-2) & torch.tril(torch.ones(L, L, dtype=torch.bool)).unsqueeze(0)   # _target_mask (:59-62)
        for i in range(self.n):
            p = f"lm.encoder.encoders.{i}"
            x = x + self._attn(_ln(x, w, p + ".norm1"), mask, p + ".self_attn")
            x = x + _lin(torch.relu(_lin(_ln(x, w, p + ".norm2"), w, p + ".feed_forward.w_1")), w, p + ".feed_forward.w_2")
        x = _ln(x, w, "lm.encoder.after_norm")
        return torch.log_softmax(_lin(x[:, -1], w, "lm.decoder"), dim=-1)

"""Oracle: TransformerEncoder (abs-pos, conv2d input layer, pre-LN, ReLU feed-forward) -- the encoder of the "next" scope row
(BASELINE configs[4]); one utterance at a time.  TEST INFRASTRUCTURE.

Reference: espnet2/asr/encoder/transformer_encoder.py:216-299 (forward), legacy/nets/pytorch_backend/transformer/encoder_layer.py:65-126
(pre-LN block), attention.py:77-151,262-265 (default branch: matmul / sqrt(d_k), softmax, no flash / sdpa),
embedding.py:38-95 (PositionalEncoding: x * sqrt(d) + pe[:T]), subsampling.py:397-474 (Conv2dSubsampling),
positionwise_feed_forward.py:30-32 (w_2(relu(w_1 x))), layer_norm.py:12-42 (eps 1e-12).
Weights: flat dict with the reference's state_dict names.
"""
import math

import torch

from .encoder import _lin, _ln, conv2d_subsampling


def positional_encoding(T, d):
    """PositionalEncoding.extend_pe (embedding.py:62-83): pe[t, 2i] = sin(t / 10000^(2i/d)), pe[t, 2i+1] = cos(...)."""
    pos = torch.arange(0, T, dtype=torch.float32).unsqueeze(1)
    div = torch.exp(torch.arange(0, d, 2, dtype=torch.float32) * -(math.log(10000.0) / d))
    pe = torch.zeros(T, d)
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe


def self_attention(x, w, pfx, heads):
    """MultiHeadedAttention.forward, default branch (attention.py:262-265 + forward_attention :121-151); single utterance, so the
    key mask is all-true."""
    T, d = x.shape
    dk = d // heads
    q = _lin(x, w, pfx + ".linear_q").view(T, heads, dk).transpose(0, 1)
    k = _lin(x, w, pfx + ".linear_k").view(T, heads, dk).transpose(0, 1)
    v = _lin(x, w, pfx + ".linear_v").view(T, heads, dk).transpose(0, 1)
    scores = q @ k.transpose(-2, -1) / math.sqrt(dk)
    attn = torch.softmax(scores, dim=-1)
    ctx = (attn @ v).transpose(0, 1).contiguous().view(T, d)
    return _lin(ctx, w, pfx + ".linear_out")


def encoder_layer(x, w, pfx, heads):
    """EncoderLayer.forward with normalize_before=True, concat_after=False (encoder_layer.py:91-124)."""
    x = x + self_attention(_ln(x, w, pfx + ".norm1"), w, pfx + ".self_attn", heads)
    h = torch.relu(_lin(_ln(x, w, pfx + ".norm2"), w, pfx + ".feed_forward.w_1"))
    return x + _lin(h, w, pfx + ".feed_forward.w_2")


def transformer_encode(feats, w, heads, num_blocks, return_layers=False):
    """feats (T_f, 80) normalised log-mel -> (T, d).  Conv2dSubsampling's own PositionalEncoding scales by sqrt(d) and adds pe
    (conv2d_subsampling already applies the scaling that both positional encodings share)."""
    x = conv2d_subsampling(feats, w)
    x = x + positional_encoding(x.shape[0], x.shape[1])
    layers = [x]
    for i in range(num_blocks):
        x = encoder_layer(x, w, f"encoder.encoders.{i}", heads)
        layers.append(x)
    x = _ln(x, w, "encoder.after_norm")
    return (x, layers) if return_layers else x

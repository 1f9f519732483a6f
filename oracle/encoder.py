"""Oracle: ConformerEncoder (conv2d input layer, rel_pos "latest", macaron, cnn module) and
the CTC head, one utterance at a time (Speech2Text is batch-1).  TEST INFRASTRUCTURE.

Weights come as a flat dict with the reference's state_dict names.
"""
import math

import torch
import torch.nn.functional as F

LN_EPS = 1e-12  # espnet2/legacy/nets/pytorch_backend/transformer/layer_norm.py:22


class TooShortUttError(Exception):
    """Mirror of subsampling.py:14-28 (message, actual_size, limit)."""

    def __init__(self, message, actual_size, limit):
        super().__init__(message)
        self.actual_size, self.limit = actual_size, limit


def _ln(x, w, pfx):
    return F.layer_norm(x, (x.shape[-1],), w[pfx + ".weight"], w[pfx + ".bias"], LN_EPS)


def _lin(x, w, pfx, bias=True):
    return F.linear(x, w[pfx + ".weight"], w[pfx + ".bias"] if bias else None)


def rel_positional_encoding(T, d):
    """RelPositionalEncoding.extend_pe/forward: row k of the (2T-1, d) table is the sinusoid of
    relative position T-1-k (even=sin, odd=cos).  embedding.py:286-334."""
    pos = torch.arange(T - 1, -T, -1, dtype=torch.float32).unsqueeze(1)
    div = torch.exp(torch.arange(0, d, 2, dtype=torch.float32) * -(math.log(10000.0) / d))
    pe = torch.zeros(2 * T - 1, d)
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe


def conv2d_subsampling(feats, w, pfx="encoder.embed"):
    """Conv2dSubsampling.forward (subsampling.py:432-474): two 3x3/stride-2 convs + ReLU,
    flatten as feature index c*F'+f (:450-451), Linear, then x*sqrt(d) (embedding.py:329)."""
    if feats.shape[0] < 7:  # check_short_utt, subsampling.py:43-44
        raise TooShortUttError(
            f"has {feats.shape[0]} frames and is too short for subsampling "
            + "(it needs more than 7 frames), return empty results", feats.shape[0], 7)
    x = feats.unsqueeze(0).unsqueeze(0)
    x = F.relu(F.conv2d(x, w[pfx + ".conv.0.weight"], w[pfx + ".conv.0.bias"], stride=2))
    x = F.relu(F.conv2d(x, w[pfx + ".conv.2.weight"], w[pfx + ".conv.2.bias"], stride=2))
    _, c, t, f = x.shape
    x = x.transpose(1, 2).contiguous().view(t, c * f)
    x = _lin(x, w, pfx + ".out")
    return x * math.sqrt(x.shape[-1])


def rel_self_attention(x, pos_emb, w, pfx, heads):
    """RelPositionMultiHeadedAttention.forward + rel_shift + forward_attention
    (attention.py:416-459, 391-414, 121-151); single utterance so the key mask is all-true."""
    T, d = x.shape
    dk = d // heads
    q = _lin(x, w, pfx + ".linear_q").view(T, heads, dk)
    k = _lin(x, w, pfx + ".linear_k").view(T, heads, dk).transpose(0, 1)
    v = _lin(x, w, pfx + ".linear_v").view(T, heads, dk).transpose(0, 1)
    p = _lin(pos_emb, w, pfx + ".linear_pos", bias=False).view(-1, heads, dk).transpose(0, 1)
    qu = (q + w[pfx + ".pos_bias_u"]).transpose(0, 1)
    qv = (q + w[pfx + ".pos_bias_v"]).transpose(0, 1)
    ac = qu @ k.transpose(-2, -1)  # (h, T, T)
    bd = qv @ p.transpose(-2, -1)  # (h, T, 2T-1)
    idx = (T - 1) - torch.arange(T).unsqueeze(1) + torch.arange(T).unsqueeze(0)  # rel_shift
    bd = torch.gather(bd, 2, idx.unsqueeze(0).expand(heads, T, T))
    attn = torch.softmax((ac + bd) / math.sqrt(dk), dim=-1)
    ctx = (attn @ v).transpose(0, 1).contiguous().view(T, d)
    return _lin(ctx, w, pfx + ".linear_out")


def conv_module(x, w, pfx):
    """ConvolutionModule.forward (conformer/convolution.py:56-79), BatchNorm1d in eval mode."""
    y = _lin(x, {pfx + ".pointwise_conv1.weight": w[pfx + ".pointwise_conv1.weight"].squeeze(-1),
                 pfx + ".pointwise_conv1.bias": w[pfx + ".pointwise_conv1.bias"]}, pfx + ".pointwise_conv1")
    y = F.glu(y, dim=-1)
    dw = w[pfx + ".depthwise_conv.weight"]
    y = F.conv1d(y.t().unsqueeze(0), dw, w[pfx + ".depthwise_conv.bias"], padding=(dw.shape[-1] - 1) // 2,
                 groups=dw.shape[0])
    y = F.batch_norm(y, w[pfx + ".norm.running_mean"], w[pfx + ".norm.running_var"], w[pfx + ".norm.weight"],
                     w[pfx + ".norm.bias"], training=False, eps=1e-5)
    y = y * torch.sigmoid(y)  # Swish, conformer/swish.py:15-18
    y = y.squeeze(0).t()
    return F.linear(y, w[pfx + ".pointwise_conv2.weight"].squeeze(-1), w[pfx + ".pointwise_conv2.bias"])


def _ffn_swish(x, w, pfx):
    h = _lin(x, w, pfx + ".w_1")
    return _lin(h * torch.sigmoid(h), w, pfx + ".w_2")  # positionwise_feed_forward.py:30-32


def encoder_layer(x, pos_emb, w, pfx, heads):
    """EncoderLayer.forward, pre-LN macaron block (conformer/encoder_layer.py:79-179)."""
    x = x + 0.5 * _ffn_swish(_ln(x, w, pfx + ".norm_ff_macaron"), w, pfx + ".feed_forward_macaron")
    x = x + rel_self_attention(_ln(x, w, pfx + ".norm_mha"), pos_emb, w, pfx + ".self_attn", heads)
    x = x + conv_module(_ln(x, w, pfx + ".norm_conv"), w, pfx + ".conv_module")
    x = x + 0.5 * _ffn_swish(_ln(x, w, pfx + ".norm_ff"), w, pfx + ".feed_forward")
    return _ln(x, w, pfx + ".norm_final")


def conformer_encode(feats, w, heads, num_blocks, return_layers=False):
    """ConformerEncoder.forward for one utterance (asr/encoder/conformer_encoder.py:327-429).
    feats (T_f, 80) normalised log-mel -> (T, d)."""
    x = conv2d_subsampling(feats, w)
    pos_emb = rel_positional_encoding(x.shape[0], x.shape[1])
    layers = [x]
    for i in range(num_blocks):
        x = encoder_layer(x, pos_emb, w, f"encoder.encoders.{i}", heads)
        layers.append(x)
    x = _ln(x, w, "encoder.after_norm")
    return (x, layers) if return_layers else x


def ctc_logits(enc, w):
    """ctc_lo Linear (espnet2/asr/ctc.py:39)."""
    return _lin(enc, w, "ctc.ctc_lo")


def ctc_greedy(enc, w, blank=0):
    """CTC.argmax (asr/ctc.py:207-215) + unique_consecutive + drop blank
    (bin/s2t_inference_ctc.py:630-632, asr_inference.py:574-575). Returns (frame_argmax, token_ids)."""
    am = torch.argmax(ctc_logits(enc, w), dim=-1)
    ids = torch.unique_consecutive(am)
    return am, ids[ids != blank]

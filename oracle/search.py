"""Oracle: TransformerDecoder.batch_score, CTCPrefixScoreTH / CTCPrefixScorer and
BatchBeamSearch for one utterance.  TEST INFRASTRUCTURE.

Restated with the reference's per-step work kept as it is (self/cross K,V re-projected every
step from the cached layer inputs, per-frame loop in the CTC prefix recursion) so that timing
this oracle is a fair stand-in for the reference's CPU path (bench.py --impl reference).
"""
import math
from typing import Dict, List, NamedTuple

import torch
import torch.nn.functional as F

from .encoder import _lin, _ln

LOGZERO = -10000000000.0  # ctc_prefix_score.py:34


class Hyp(NamedTuple):
    """Hypothesis (beam_search.py:15-31): yseq with leading sos (and trailing eos when ended)."""

    yseq: torch.Tensor
    score: float
    scores: Dict[str, float]


def _pos_enc(length, d):
    """PositionalEncoding table (embedding.py:62-83)."""
    pos = torch.arange(0, length, dtype=torch.float32).unsqueeze(1)
    div = torch.exp(torch.arange(0, d, 2, dtype=torch.float32) * -(math.log(10000.0) / d))
    pe = torch.zeros(length, d)
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe


def _mha(q_in, kv_in, w, pfx, heads):
    """MultiHeadedAttention.forward default branch, no mask (attention.py:262-265,121-151)."""
    n, tq, d = q_in.shape
    dk = d // heads
    q = _lin(q_in, w, pfx + ".linear_q").view(n, tq, heads, dk).transpose(1, 2)
    k = _lin(kv_in, w, pfx + ".linear_k").view(n, -1, heads, dk).transpose(1, 2)
    v = _lin(kv_in, w, pfx + ".linear_v").view(n, -1, heads, dk).transpose(1, 2)
    att = torch.softmax(q @ k.transpose(-2, -1) / math.sqrt(dk), dim=-1)
    ctx = (att @ v).transpose(1, 2).contiguous().view(n, tq, d)
    return _lin(ctx, w, pfx + ".linear_out")


class OracleDecoder:
    """TransformerDecoder.batch_score / forward_one_step (asr/decoder/transformer_decoder.py:
    262-311, 191-238) with DecoderLayer.forward's cache branch (decoder_layer.py:73-179)."""

    def __init__(self, w, heads, num_blocks):
        self.w, self.heads, self.n = w, heads, num_blocks
        self.d = w["decoder.embed.0.weight"].shape[1]

    def batch_score(self, ys, cache, memory):
        """ys (n, len) int64; cache: list[layer] of (n, len-1, d) or None; memory (n, T, d).
        The last query attends to the whole prefix, so the causal mask row is all-true."""
        w = self.w
        x = F.embedding(ys, w["decoder.embed.0.weight"]) * math.sqrt(self.d) + _pos_enc(ys.shape[1], self.d)
        new_cache = []
        for i in range(self.n):
            p = f"decoder.decoders.{i}"
            tgt = _ln(x, w, p + ".norm1")
            y = x[:, -1:, :] + _mha(tgt[:, -1:, :], tgt, w, p + ".self_attn", self.heads)
            y = y + _mha(_ln(y, w, p + ".norm2"), memory, w, p + ".src_attn", self.heads)
            h = F.relu(_lin(_ln(y, w, p + ".norm3"), w, p + ".feed_forward.w_1"))
            y = y + _lin(h, w, p + ".feed_forward.w_2")
            x = y if cache is None else torch.cat([cache[i], y], dim=1)
            if cache is None and ys.shape[1] > 1:  # full forward of a primer longer than 1 (not used)
                raise NotImplementedError
            new_cache.append(x)
        y = _ln(x[:, -1], w, "decoder.after_norm")
        return torch.log_softmax(_lin(y, w, "decoder.output_layer"), dim=-1), new_cache


class CTCPrefixScorerTH:
    """CTCPrefixScorer.batch_init_state/batch_score_partial/select_state (scorers/ctc.py:87-126,
    40-63) over CTCPrefixScoreTH.__call__ (ctc_prefix_score.py:71-191), batch of one utterance."""

    def __init__(self, logp, blank, eos):
        self.x = logp  # (T, V) log-softmax of ctc_lo(enc)
        self.T, self.V = logp.shape
        self.blank, self.eos = blank, eos

    def score(self, ys, r_prev, s_prev, ids):
        """ys (n, len); r_prev (T, 2, n) or None; s_prev (n,) ; ids (n, k) int64 or None.
        Returns (scores (n, V) = log_psi - s_prev, r (T, 2, n, k), log_psi (n, V), idmap)."""
        n = ys.shape[0]
        out_len = ys.shape[1] - 1
        last = ys[:, -1]
        T, V = self.T, self.V
        if r_prev is None:
            r_prev = torch.full((T, 2, n), LOGZERO)
            r_prev[:, 1] = torch.cumsum(self.x[:, self.blank], 0).unsqueeze(1)
            s_prev = torch.zeros(n)
        if ids is not None:
            k = ids.shape[1]
            idmap = torch.full((n, V), -1, dtype=torch.long)
            idmap[torch.arange(n).unsqueeze(1), ids] = torch.arange(k)
            xn = self.x[:, ids.reshape(-1)].view(T, n, k)
        else:
            k, idmap = V, None
            xn = self.x.unsqueeze(1).expand(T, n, V)
        xb = self.x[:, self.blank].view(T, 1, 1).expand(T, n, k)
        r = torch.full((T, 2, n, k), LOGZERO)
        if out_len == 0:
            r[0, 0] = xn[0]
        r_sum = torch.logsumexp(r_prev, 1)  # (T, n)
        log_phi = r_sum.unsqueeze(2).repeat(1, 1, k)
        for i in range(n):  # same-label case: only the blank-ending path may precede (:138-144)
            pos = int(idmap[i, last[i]]) if idmap is not None else int(last[i])
            if pos >= 0:
                log_phi[:, i, pos] = r_prev[:, 1, i]
        start, end = max(out_len, 1), T
        for t in range(start, end):  # forward recursion (:159-164)
            rn = torch.logsumexp(torch.stack([r[t - 1, 0], log_phi[t - 1]]), 0) + xn[t]
            rb = torch.logsumexp(torch.stack([r[t - 1, 0], r[t - 1, 1]]), 0) + xb[t]
            r[t, 0], r[t, 1] = rn, rb
        log_phi_x = torch.cat((log_phi[0].unsqueeze(0), log_phi[:-1]), dim=0) + xn
        psi = torch.logsumexp(torch.cat((log_phi_x[start:end], r[start - 1, 0].unsqueeze(0)), dim=0), dim=0)
        if ids is not None:
            log_psi = torch.full((n, V), LOGZERO)
            log_psi[torch.arange(n).unsqueeze(1), ids] = psi
        else:
            log_psi = psi.clone()
        log_psi[:, self.eos] = r_sum[T - 1]  # (:184-185), end_frame = T-1
        if self.eos != self.blank:
            log_psi[:, self.blank] = LOGZERO  # (:187-189)
        return log_psi - s_prev.unsqueeze(1), r, log_psi, idmap


def _end_detect(ended: List[Hyp], i, M=3, D_end=math.log(1 * math.exp(-10))):
    """end_detect (e2e_asr_common.py:14-44)."""
    if not ended:
        return False
    best = max(h.score for h in ended)
    count = 0
    for m in range(M):
        same = [h.score for h in ended if len(h.yseq) == i - m]
        if same and max(same) - best < D_end:
            count += 1
    return count == M


def batch_beam_search(enc, decoder, ctc_logp, *, beam_size, ctc_weight, vocab, sos, eos, blank=0,
                      maxlenratio=0.0, minlenratio=0.0, penalty=0.0, normalize_length=False, trace=None, lm=None, lm_weight=0.0):
    """BeamSearch.forward + BatchBeamSearch.search/post_process for one utterance
    (beam_search.py:385-498; batch_beam_search.py:253-357, 359-423).  Scorers as Speech2Text wires
    them (asr_inference.py:168-176, 310-316): decoder weight 1-ctc_weight, ctc weight ctc_weight,
    length_bonus weight `penalty` (constant 1.0 score, scorers/length_bonus.py:38-60); zero-weight
    scorers are dropped (beam_search.py:83-85).  `trace`, if a list, receives per-step dicts."""
    T = enc.shape[0]
    w_dec, w_ctc = 1.0 - ctc_weight, ctc_weight
    use_dec = w_dec != 0 and decoder is not None
    use_ctc = w_ctc != 0
    use_lm = lm is not None and lm_weight != 0   # asr_inference.py:178-191: scorers["lm"] = lm.lm, a full scorer with weight lm_weight
    pre_beam = int(1.5 * beam_size)
    do_pre_beam = use_ctc and (use_dec or use_lm) and ctc_weight != 1.0 and pre_beam < vocab  # pre_beam_score_key "full" unless ctc_weight==1
    if maxlenratio == 0:
        maxlen = T
    elif maxlenratio < 0:
        maxlen = -int(maxlenratio)
    else:
        maxlen = max(1, int(maxlenratio * T))
    minlen = -int(minlenratio) if minlenratio < 0 else int(minlenratio * T)
    scorer = CTCPrefixScorerTH(ctc_logp, blank, eos) if use_ctc else None

    yseq = torch.full((1, 1), sos, dtype=torch.long)
    score = torch.zeros(1)
    sc_dec, sc_ctc, sc_lb, sc_lm = torch.zeros(1), torch.zeros(1), torch.zeros(1), torch.zeros(1)
    cache, r_state, s_state = None, None, None
    ended: List[Hyp] = []
    for i in range(maxlen):
        n = yseq.shape[0]
        weighted = torch.zeros(n, vocab)
        if use_dec:
            logp, new_cache = decoder.batch_score(yseq, cache, enc.unsqueeze(0).expand(n, T, -1))
            weighted += w_dec * logp
        if penalty != 0:
            weighted += penalty * 1.0
        if use_lm:
            lm_logp = lm.batch_score(yseq)
            weighted += lm_weight * lm_logp
        part_ids = torch.topk(weighted, pre_beam, dim=-1)[1] if do_pre_beam else None
        if use_ctc:
            part, r_new, log_psi, idmap = scorer.score(yseq, r_state, s_state, part_ids)
            weighted += w_ctc * part
        weighted += score.unsqueeze(1)
        top = weighted.view(-1).topk(beam_size)[1]  # batch_beam, batch_beam_search.py:98-122
        prev = torch.div(top, vocab, rounding_mode="trunc")
        tok = top % vocab
        if trace is not None:
            trace.append(dict(step=i, prev=prev.clone(), tok=tok.clone(), score=weighted[prev, tok].clone(),
                              part_ids=None if part_ids is None else part_ids.clone(),
                              dec_logp=logp.clone() if use_dec else None,
                              ctc_part=part.clone() if use_ctc else None))
        yseq = torch.cat([yseq[prev], tok.unsqueeze(1)], dim=1)
        score = weighted[prev, tok]
        if use_dec:
            sc_dec = sc_dec[prev] + logp[prev, tok]
            cache = [c[prev] for c in new_cache]
        if penalty != 0:
            sc_lb = sc_lb[prev] + 1.0
        if use_lm:
            sc_lm = sc_lm[prev] + lm_logp[prev, tok]
        if use_ctc:
            sc_ctc = sc_ctc[prev] + part[prev, tok]
            col = idmap[prev, tok] if idmap is not None else tok  # select_state, scorers/ctc.py:40-63
            r_state = r_new[:, :, prev, col]
            s_state = log_psi[prev, tok]
        # post_process
        if i == maxlen - 1:
            yseq = torch.cat([yseq, torch.full((yseq.shape[0], 1), eos, dtype=torch.long)], dim=1)
        is_eos = yseq[:, -1] == eos
        for b in torch.nonzero(is_eos).view(-1).tolist():
            if i >= minlen:
                scores = {}
                if use_dec:
                    scores["decoder"] = float(sc_dec[b])
                if use_ctc:
                    scores["ctc"] = float(sc_ctc[b])
                if penalty != 0:
                    scores["length_bonus"] = float(sc_lb[b])
                if use_lm:
                    scores["lm"] = float(sc_lm[b])
                ended.append(Hyp(yseq=yseq[b].clone(), score=float(score[b]), scores=scores))
        keep = torch.nonzero(~is_eos).view(-1)
        yseq, score = yseq[keep], score[keep]
        sc_dec, sc_ctc, sc_lb = sc_dec[keep] if use_dec else sc_dec, sc_ctc[keep] if use_ctc else sc_ctc, \
            sc_lb[keep] if penalty != 0 else sc_lb
        sc_lm = sc_lm[keep] if use_lm else sc_lm
        if use_dec:
            cache = [c[keep] for c in cache]
        if use_ctc:
            r_state, s_state = r_state[:, :, keep], s_state[keep]
        if maxlenratio == 0.0 and _end_detect(ended, i):
            break
        if yseq.shape[0] == 0:
            break
    key = (lambda h: h.score / (len(h.yseq) - 1)) if normalize_length else (lambda h: h.score)
    nbest = sorted(ended, key=key, reverse=True)
    if not nbest and minlenratio >= 0.1:  # beam_search.py:462-471
        return batch_beam_search(enc, decoder, ctc_logp, beam_size=beam_size, ctc_weight=ctc_weight, vocab=vocab,
                                 sos=sos, eos=eos, blank=blank, maxlenratio=maxlenratio,
                                 minlenratio=max(0.0, minlenratio - 0.1), penalty=penalty,
                                 normalize_length=normalize_length, lm=lm, lm_weight=lm_weight)
    return nbest

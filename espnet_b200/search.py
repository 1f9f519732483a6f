"""Utterance-batched, device-resident BatchBeamSearch (joint CTC/attention, attention-only, CTC-only).

Per-utterance results are those of the reference's BatchBeamSearch run on each utterance alone
(espnet2/legacy/nets/batch_beam_search.py, beam_search.py:385-498): same scorer weights
(decoder 1-ctc_weight, ctc ctc_weight, length_bonus `penalty`), pre-beam int(1.5*beam) on the full
scorers, eos handling, maxlen/minlen rules, end detection and final ordering.  What differs is where
the state lives: hypotheses never become Python objects during the search.
"""
from typing import Any, Dict, List, NamedTuple, Union

import torch

from . import ops
from .lib import call, ptr
from .ops import _count


class Hypothesis(NamedTuple):
    """Same fields as espnet2.legacy.nets.beam_search.Hypothesis (beam_search.py:15-31)."""

    yseq: torch.Tensor
    score: Union[float, torch.Tensor] = 0
    scores: Dict[str, Union[float, torch.Tensor]] = dict()
    states: Dict[str, Any] = dict()
    hs: List[torch.Tensor] = []

    def asdict(self) -> dict:
        return self._replace(yseq=self.yseq.tolist(), score=float(self.score),
                             scores={k: float(v) for k, v in self.scores.items()})._asdict()


class BatchBeamSearch(torch.nn.Module):
    """scorers: {"decoder": TransformerDecoder | None, "ctc": CTC | None}; weights as Speech2Text builds them
    (asr_inference.py:310-316).  Zero-weight scorers are dropped (beam_search.py:83-85)."""

    def __init__(self, scorers, weights, beam_size, vocab_size, sos, eos, token_list=None, pre_beam_ratio=1.5,
                 pre_beam_score_key=None, normalize_length=False):
        super().__init__()
        self.weights = dict(weights)
        self.w_dec = float(weights.get("decoder", 0.0)) if scorers.get("decoder") is not None else 0.0
        self.w_ctc = float(weights.get("ctc", 0.0)) if scorers.get("ctc") is not None else 0.0
        self.penalty = float(weights.get("length_bonus", 0.0))
        self.decoder = scorers.get("decoder") if self.w_dec != 0 else None
        self.ctc = scorers.get("ctc") if self.w_ctc != 0 else None
        if self.decoder is None and self.ctc is None:
            raise ValueError("no scorer with non-zero weight")
        if self.decoder is not None and self.w_dec < 0 or self.w_ctc < 0:
            raise NotImplementedError("negative scorer weights")
        self.nn_dict = torch.nn.ModuleDict({k: v for k, v in (("decoder", self.decoder), ("ctc", self.ctc)) if v is not None})
        self.sos, self.eos, self.n_vocab, self.beam_size = sos, eos, vocab_size, beam_size
        self.token_list = token_list
        self.pre_beam_size = int(pre_beam_ratio * beam_size)
        self.pre_beam_score_key = pre_beam_score_key
        self.do_pre_beam = (pre_beam_score_key is not None and self.pre_beam_size < vocab_size and self.ctc is not None
                            and self.decoder is not None)
        if self.decoder is not None and self.ctc is not None and not self.do_pre_beam:
            raise NotImplementedError("joint decoding without pre-beam (vocab <= 1.5*beam) is not implemented")
        self.normalize_length = normalize_length
        self.full_scorers = {k: v for k, v in (("decoder", self.decoder),) if v is not None}
        self.part_scorers = {k: v for k, v in (("ctc", self.ctc),) if v is not None}
        if beam_size > 16 and self.decoder is not None:
            raise NotImplementedError("beam_size > 16 with an attention decoder")

    @torch.no_grad()
    def forward_batch(self, enc, enc_lens, enc_split=None, maxlenratio=0.0, minlenratio=0.0, check_every=8):
        """enc (U, Tmax, D) CUDA, enc_lens (U,) -> list (per utterance) of n-best Hypothesis lists, sorted."""
        dev = enc.device
        U, Tmax, D = enc.shape
        W, V = self.beam_size, self.n_vocab
        n = U * W
        lens_cpu = enc_lens.detach().cpu().to(torch.int64)
        lens32 = lens_cpu.to(device=dev, dtype=torch.int32)
        if maxlenratio == 0:
            maxlen = lens_cpu.clone()
        elif maxlenratio < 0:
            maxlen = torch.full_like(lens_cpu, -int(maxlenratio))
        else:
            maxlen = torch.clamp((maxlenratio * lens_cpu.double()).long(), min=1)
        minlen = torch.full_like(lens_cpu, -int(minlenratio)) if minlenratio < 0 else (minlenratio * lens_cpu.double()).long()
        cap = int(maxlen.max())
        maxlen_d, minlen_d = maxlen.to(dev, torch.int32), minlen.to(dev, torch.int32)
        if enc_split is None:
            enc_split = ops.split_from(enc.contiguous().view(U * Tmax, D))
        use_dec, use_ctc = self.decoder is not None, self.ctc is not None
        mode = 1 if (use_dec and use_ctc) else (0 if use_dec else 2)
        P = self.pre_beam_size if mode == 1 else W
        PC = P + 1 if mode == 1 else P
        i32 = lambda *s, fill=0: torch.full(s, fill, dtype=torch.int32, device=dev)  # noqa: E731
        f32 = lambda *s, fill=0.0: torch.full(s, fill, dtype=torch.float32, device=dev)  # noqa: E731

        # ---- state (double-buffered)
        score, sc_dec, sc_ctc = [f32(n), f32(n)], [f32(n), f32(n)], [f32(n), f32(n)]
        active = [i32(n), i32(n)]
        active[0].view(U, W)[:, 0] = 1           # one initial hypothesis [sos] per utterance (batch_beam_search.py:124-153)
        last_tok = [i32(n, fill=self.sos), i32(n, fill=self.sos)]
        parent = i32(n)
        anc = [i32(n, cap + 1), i32(n, cap + 1)]
        bp_parent, bp_token = i32(cap, n, fill=-1), i32(cap, n, fill=self.eos)
        ended_cap = W * cap
        e_count, e_step, e_slot = i32(U), i32(U, ended_cap), i32(U, ended_cap)
        e_score, e_dec, e_ctc = f32(U, ended_cap), f32(U, ended_cap), f32(U, ended_cap)
        best_at, best_all, done = f32(U, cap, fill=float("-inf")), f32(U, fill=float("-inf")), i32(U)
        cand_ids, cand_val = i32(n, P), f32(n, P)
        n_active = i32(1)
        dst = None
        if use_dec:
            dst = self.decoder.init_memory(enc_split, U, Tmax, lens32, n, cap)
        logp_ctc = part = psi = valid = r = s_prev = None
        if use_ctc:
            logp_ctc = self.ctc.log_softmax(enc, enc_split)           # (U, Tmax, V), scorers/ctc.py:96-99
            r = [f32(n, Tmax, 4), f32(n, Tmax, 4)]   # per frame (r^n, r^b, r_sum, pad)
            s_prev = [f32(n), f32(n)]
            call("espb_ctc_init_state_f32", ptr(logp_ctc), U, Tmax, V, ptr(lens32), 0, W, ptr(r[0]), ptr(s_prev[0]))
            _count()
            if mode == 1:
                part, psi, valid = f32(n, PC), f32(n, PC), i32(n, PC)
            else:
                part = f32(n, V)
        end_detect = 1 if maxlenratio == 0.0 else 0
        # the CTC state update of step i is only consumed by the CTC scoring of step i+1 (after the next decoder pass): run it on a
        # side stream so that its T-step sequential recursion overlaps the decoder
        main = torch.cuda.current_stream()
        side = self._side_stream(dev) if (use_ctc and use_dec) else None
        ev_sel, ev_adv = torch.cuda.Event(), torch.cuda.Event()
        adv_pending = False
        cur = 0
        steps_run = 0
        for i in range(cap):
            nxt = cur ^ 1
            logp_dec = None
            if use_dec:
                logp_dec = self.decoder.step(dst, i, last_tok[cur], anc[cur], W)
            if adv_pending:
                main.wait_event(ev_adv)
                adv_pending = False
            if mode == 1:
                ops.rows_topk(logp_dec, self.w_dec, P, cand_ids, cand_val)
                call("espb_ctc_score_cands_f32", ptr(logp_ctc), U, Tmax, V, ptr(lens32), 0, self.eos, W, ptr(r[cur]), ptr(s_prev[cur]),
                     ptr(last_tok[cur]), i, ptr(cand_ids), P, ptr(part), ptr(psi), ptr(valid))
                _count()
            elif mode == 0:
                ops.rows_topk(logp_dec, self.w_dec, P, cand_ids, cand_val)
            else:
                call("espb_ctc_score_dense_f32", ptr(logp_ctc), U, Tmax, V, ptr(lens32), 0, self.eos, W, ptr(r[cur]), ptr(s_prev[cur]),
                     ptr(last_tok[cur]), i, ptr(part))
                _count()
                ops.rows_topk(part, self.w_ctc, P, cand_ids, cand_val)
            call("espb_beam_select", ptr(score[cur]), ptr(sc_dec[cur]), ptr(sc_ctc[cur]), ptr(active[cur]), ptr(score[nxt]),
                 ptr(sc_dec[nxt]), ptr(sc_ctc[nxt]), ptr(active[nxt]), ptr(last_tok[nxt]), ptr(parent), ptr(bp_parent), ptr(bp_token),
                 ptr(e_count), ptr(e_step), ptr(e_slot), ptr(e_score), ptr(e_dec), ptr(e_ctc), ended_cap, ptr(best_at), ptr(best_all),
                 ptr(done), U, W, P, V, i, ptr(maxlen_d), ptr(minlen_d), self.eos, self.w_dec, self.w_ctc, self.penalty, mode,
                 ptr(cand_ids), ptr(cand_val), ptr(logp_dec), ptr(part), ptr(valid), end_detect, cap)
            _count()
            if use_dec:
                call("espb_anc_update_i32", ptr(anc[cur]), ptr(anc[nxt]), cap + 1, ptr(parent), i, n)
                _count()
            if use_ctc:
                if side is not None:
                    ev_sel.record(main)
                    with torch.cuda.stream(side):
                        side.wait_event(ev_sel)
                        call("espb_ctc_advance_f32", ptr(logp_ctc), U, Tmax, V, ptr(lens32), 0, self.eos, W, ptr(r[cur]), ptr(parent),
                             ptr(last_tok[cur]), ptr(last_tok[nxt]), ptr(active[nxt]), i, ptr(r[nxt]), ptr(s_prev[nxt]))
                        ev_adv.record(side)
                    adv_pending = True
                else:
                    call("espb_ctc_advance_f32", ptr(logp_ctc), U, Tmax, V, ptr(lens32), 0, self.eos, W, ptr(r[cur]), ptr(parent),
                         ptr(last_tok[cur]), ptr(last_tok[nxt]), ptr(active[nxt]), i, ptr(r[nxt]), ptr(s_prev[nxt]))
                _count()
            cur = nxt
            steps_run = i + 1
            if (i + 1) % check_every == 0 or end_detect:
                call("espb_count_active_i32", ptr(active[cur]), n, ptr(n_active))
                _count()
                if int(n_active.item()) == 0:
                    break
        if side is not None:
            main.wait_stream(side)
        return self._collect(U, W, steps_run, maxlen, bp_parent, bp_token, e_count, e_step, e_slot, e_score, e_dec, e_ctc)

    def _side_stream(self, dev):
        key = str(dev)
        if not hasattr(self, "_streams"):
            self._streams = {}
        if key not in self._streams:
            self._streams[key] = torch.cuda.Stream(device=dev)
        return self._streams[key]

    def _collect(self, U, W, steps, maxlen, bp_parent, bp_token, e_count, e_step, e_slot, e_score, e_dec, e_ctc):
        """Host post-processing: rebuild token sequences from back-pointers and sort (beam_search.py:452-459)."""
        bpp, bpt = bp_parent[:steps].cpu().numpy(), bp_token[:steps].cpu().numpy()
        cnt = e_count.cpu().numpy()
        es, el = e_step.cpu().numpy(), e_slot.cpu().numpy()
        sc, sd, sct = e_score.cpu().numpy(), e_dec.cpu().numpy(), e_ctc.cpu().numpy()
        results = []
        for u in range(U):
            hyps = []
            for e in range(int(cnt[u])):
                step, slot = int(es[u, e]), int(el[u, e])
                toks = []
                s = slot
                for j in range(step, -1, -1):
                    toks.append(int(bpt[j, s]))
                    s = int(bpp[j, s])
                toks.reverse()
                yseq = [self.sos] + toks
                if step == int(maxlen[u]) - 1:
                    yseq.append(self.eos)  # "adding <eos> in the last position in the loop" (batch_beam_search.py:392-407)
                scores = {}
                if self.decoder is not None:
                    scores["decoder"] = float(sd[u, e])
                if self.ctc is not None:
                    scores["ctc"] = float(sct[u, e])
                if self.penalty != 0:
                    scores["length_bonus"] = float(step + 1)
                hyps.append(Hypothesis(yseq=torch.tensor(yseq, dtype=torch.int64), score=float(sc[u, e]), scores=scores))
            if self.normalize_length:
                hyps.sort(key=lambda h: h.score / (len(h.yseq) - 1), reverse=True)
            else:
                hyps.sort(key=lambda h: h.score, reverse=True)
            results.append(hyps)
        return results

    def forward(self, x, maxlenratio=0.0, minlenratio=0.0):
        """Reference signature: one utterance, x (T, D) -> n-best list (beam_search.py:385-498)."""
        lens = torch.tensor([x.shape[0]], dtype=torch.int64)
        res = self.forward_batch(x.unsqueeze(0).contiguous(), lens, None, maxlenratio, minlenratio)[0]
        if not res and minlenratio >= 0.1:  # beam_search.py:462-471
            return self.forward(x, maxlenratio, max(0.0, minlenratio - 0.1))
        return res

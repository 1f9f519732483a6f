"""Oracle: Speech2Text.__call__ (espnet2/bin/asr_inference.py:490-562, 583-677).  TEST INFRASTRUCTURE."""
import numpy as np
import torch

from . import encoder as E
from . import frontend as Fr
from . import transformer_encoder as TE
from .search import OracleDecoder, batch_beam_search


class OracleSpeech2Text:
    """cfg keys: d_model, heads, ff, enc_layers, dec_layers, vocab (+ encoder: "conformer" | "transformer").  `weights` = the reference
    ESPnetASRModel.state_dict() (float32 CPU tensors).  blank=0, sos=eos=vocab-1
    (espnet_model.py:76-87)."""

    def __init__(self, cfg, weights, beam_size=20, ctc_weight=0.5, penalty=0.0, nbest=1, maxlenratio=0.0,
                 minlenratio=0.0, normalize_length=False):
        self.cfg = cfg
        self.w = {k: v.detach().float().cpu() for k, v in weights.items()}
        self.melmat = self.w.get("frontend.logmel.melmat", None)
        if self.melmat is None:
            self.melmat = Fr.slaney_mel_matrix()
        self.beam_size, self.ctc_weight, self.penalty, self.nbest = beam_size, ctc_weight, penalty, nbest
        self.maxlenratio, self.minlenratio, self.normalize_length = maxlenratio, minlenratio, normalize_length
        self.vocab = cfg["vocab"]
        self.sos = self.eos = self.vocab - 1
        has_dec = any(k.startswith("decoder.") for k in self.w) and ctc_weight != 1.0
        self.decoder = OracleDecoder(self.w, cfg["heads"], cfg["dec_layers"]) if has_dec else None

    @torch.no_grad()
    def encode(self, speech):
        """ESPnetASRModel.encode for one utterance (espnet_model.py:380-448)."""
        if isinstance(speech, np.ndarray):
            speech = torch.tensor(speech)
        feats = Fr.log_mel(Fr.stft_power(speech.float()), self.melmat)
        feats = Fr.utterance_mvn(feats)
        if self.cfg.get("encoder", "conformer") == "transformer":   # abs-pos TransformerEncoder (next scope row)
            return TE.transformer_encode(feats, self.w, self.cfg["heads"], self.cfg["enc_layers"])
        return E.conformer_encode(feats, self.w, self.cfg["heads"], self.cfg["enc_layers"])

    @torch.no_grad()
    def ctc_greedy(self, speech):
        return E.ctc_greedy(self.encode(speech), self.w)

    @torch.no_grad()
    def __call__(self, speech, trace=None):
        enc = self.encode(speech)
        logp = torch.log_softmax(E.ctc_logits(enc, self.w), dim=-1)
        hyps = batch_beam_search(enc, self.decoder, logp, beam_size=self.beam_size, ctc_weight=self.ctc_weight,
                                 vocab=self.vocab, sos=self.sos, eos=self.eos, maxlenratio=self.maxlenratio,
                                 minlenratio=self.minlenratio, penalty=self.penalty,
                                 normalize_length=self.normalize_length, trace=trace)
        results = []
        for h in hyps[: self.nbest]:
            token_int = [t for t in h.yseq[1:-1].tolist() if t != 0]  # asr_inference.py:659-666
            results.append((None, [f"<{t}>" for t in token_int], token_int, h))
        return results

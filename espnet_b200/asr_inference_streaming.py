"""Speech2TextStreaming for N live streams in lock step: chunked frontend -> contextual-block encoder -> incremental CTC-greedy output
(BASELINE configs[3]).

Reference: espnet2/bin/asr_inference_streaming.py:205-335.  ``apply_frontend`` is the reference's algorithm (waveform overlap buffer of
(2*ceil(ceil(win/hop)/2) - 1) hops + residual, STFT of the buffered chunk, trimming of the ceil(ceil(win/hop)/2) edge frames) with a leading
stream dimension: every stream pushes the same number of samples per call, so one (N, n) tensor replaces N Python objects.  The encoder is
``ContextualBlockConformerEncoder.forward_infer`` (streaming_encoder.py).

NOT the reference's decoder: the reference scores the block-synchronous ``BatchBeamSearchOnline`` (legacy/nets/batch_beam_search_online.py) on every
push; here each push emits the CTC-greedy tokens of the new encoder frames (argmax -> collapse repeats across pushes -> drop blank, as
asr/ctc.py:207-215 does for whole utterances).  Joint / attention streaming decoding is not implemented and is refused.
"""
import math
from typing import List, Optional

import torch

from . import ops
from .lib import call, ptr
from .ops import _count


class Speech2TextStreaming:
    def __init__(self, asr_model=None, n_streams: int = 1, device: str = "cuda", ctc_weight: float = 1.0, asr_train_config=None,
                 asr_model_file=None, **unused):
        """``asr_model``: a built ESPnetASRModel, or -- as the reference's constructor (asr_inference_streaming.py:46-75) -- ``asr_train_config`` (+
        ``asr_model_file``) to build it from the reference's own config.yaml / checkpoint."""
        if asr_model is None or isinstance(asr_model, (str, bytes)) or hasattr(asr_model, "__fspath__"):
            from .asr_inference import build_model_from_file

            cfg = asr_train_config if asr_model is None else asr_model
            if cfg is None:
                raise ValueError("Speech2TextStreaming needs asr_model or asr_train_config")
            asr_model, _ = build_model_from_file(cfg, asr_model_file, device)
        if ctc_weight != 1.0:
            raise NotImplementedError("espnet_b200.Speech2TextStreaming emits CTC-greedy tokens (ctc_weight=1.0); the block-synchronous beam search "
                                      "of the reference (BatchBeamSearchOnline) is not implemented")
        self.asr_model = asr_model.to(device).eval()
        self.device, self.n = device, n_streams
        fe = asr_model.frontend
        self.hop_length, self.win_length = fe.hop_length, fe.win_length
        self.edge = math.ceil(math.ceil(self.win_length / self.hop_length) / 2)      # frames trimmed at a chunk edge
        self.blank = asr_model.blank_id
        self.reset()

    def reset(self):
        self.frontend_states, self.encoder_states = None, None
        self.last_tok = torch.full((self.n,), -1, dtype=torch.int32, device=self.device)    # previous frame's argmax per stream (collapse across pushes)
        self.tokens: List[List[int]] = [[] for _ in range(self.n)]

    @torch.no_grad()
    def apply_frontend(self, speech: torch.Tensor, prev_states=None, is_final: bool = False):
        """speech (N, n) float32 -> (feats (N, T, n_mels) | None, next_states)   [asr_inference_streaming.py:205-294]"""
        speech = speech.to(self.device, non_blocking=True).float()
        if prev_states is not None:
            speech = torch.cat([prev_states["waveform_buffer"], speech], dim=1)
        n = speech.shape[1]
        if n <= self.win_length:
            if not is_final:
                return None, {"waveform_buffer": speech.clone()}
            speech = torch.cat([speech, speech.new_zeros(speech.shape[0], self.win_length - n)], dim=1)
            n = speech.shape[1]
        if is_final:
            to_process, buf = speech, None
        else:
            n_frames, n_res = n // self.hop_length, n % self.hop_length
            keep = (2 * self.edge - 1) * self.hop_length + n_res
            to_process = speech[:, : n_frames * self.hop_length]
            buf = speech[:, n - keep:].clone()
        to_process = to_process.contiguous()
        lens = torch.full((to_process.shape[0],), to_process.shape[1], dtype=torch.long)
        feats, flens = self.asr_model.frontend(to_process, lens)
        feats._espb_partial = None          # chunk statistics are not utterance statistics
        if self.asr_model.normalize is not None:
            feats, flens = self.asr_model.normalize(feats, flens)
        T = feats.shape[1]
        if is_final:
            if prev_states is not None:
                feats = feats[:, self.edge:]
        elif prev_states is None:
            feats = feats[:, : T - self.edge]
        else:
            feats = feats[:, self.edge: T - self.edge]
        return feats.contiguous(), (None if is_final else {"waveform_buffer": buf})

    @torch.no_grad()
    def __call__(self, speech: torch.Tensor, is_final: bool = False) -> List[List[int]]:
        """speech (N, n): the next n samples of every stream.  Returns, per stream, the token ids this push added."""
        assert speech.dim() == 2 and speech.shape[0] == self.n
        feats, self.frontend_states = self.apply_frontend(speech, self.frontend_states, is_final)
        new = [[] for _ in range(self.n)]
        if feats is not None and feats.shape[1] > 0:
            lens = torch.full((self.n,), feats.shape[1], dtype=torch.long)
            enc, _, self.encoder_states = self.asr_model.encoder(feats, lens, self.encoder_states, is_final=is_final, infer_mode=True)
            if enc.shape[1] > 0:
                ctc = self.asr_model.ctc
                am = ctc.argmax(enc).to(torch.int32)                        # (N, T)
                prev = torch.cat([self.last_tok.view(-1, 1), am[:, :-1]], dim=1)
                keep = (am != prev) & (am != self.blank)
                self.last_tok = am[:, -1].contiguous()
                am_c, keep_c = am.cpu(), keep.cpu()
                for s in range(self.n):
                    new[s] = am_c[s][keep_c[s]].tolist()
                    self.tokens[s].extend(new[s])
        if is_final:
            done = [list(t) for t in self.tokens]
            self.reset()
            self.final_tokens = done
        return new

"""CTC head (espnet2/asr/ctc.py:13-215): ``ctc_lo`` Linear + log_softmax / argmax, plus on-device greedy collapse."""
import torch

from . import ops
from .lib import call, ptr
from .ops import _count, linear, split_from


class CTC(torch.nn.Module):
    def __init__(self, odim: int, encoder_output_size: int, dropout_rate: float = 0.0, ctc_type: str = "builtin",
                 reduce: bool = True, ignore_nan_grad=None, zero_infinity: bool = True, brctc_risk_strategy: str = "exp",
                 brctc_group_strategy: str = "end", brctc_risk_factor: float = 0.0):
        super().__init__()
        self.ctc_lo = torch.nn.Linear(encoder_output_size, odim)
        self.odim, self.eprojs = odim, encoder_output_size
        self._packed = None

    def _load_from_state_dict(self, *args, **kwargs):
        self._packed = None
        return super()._load_from_state_dict(*args, **kwargs)

    def _pack(self):
        w = self.ctc_lo.weight.detach().float().contiguous()
        self._packed = (split_from(w), self.ctc_lo.bias.detach().float().contiguous())
        return self._packed

    def _split_input(self, hs_pad, hs_split):
        if hs_split is not None:
            return hs_split
        return split_from(hs_pad.contiguous().float().view(-1, hs_pad.shape[-1]))

    @torch.no_grad()
    def logits(self, hs_pad, hs_split=None, out=None):
        """(B, T, D) -> (B, T, V) = ctc_lo(hs_pad); `out` optionally provides the (B*T, V) result buffer."""
        w, b = self._packed or self._pack()
        B, T, D = hs_pad.shape
        xs = self._split_input(hs_pad, hs_split)
        if out is None:
            out = torch.empty(B * T, self.odim, dtype=torch.float32, device=hs_pad.device)
        out = out.view(B * T, self.odim)
        linear(xs, w, out, bias=b)
        return out.view(B, T, self.odim)

    @torch.no_grad()
    def log_softmax(self, hs_pad, hs_split=None, out=None):
        lg = self.logits(hs_pad, hs_split, out)
        ops.log_softmax_rows_(lg.view(-1, self.odim))
        return lg

    @torch.no_grad()
    def argmax(self, hs_pad, hs_split=None):
        lg = self.logits(hs_pad, hs_split)
        out = torch.empty(lg.shape[0] * lg.shape[1], dtype=torch.int32, device=lg.device)
        ops.argmax_rows(lg.view(-1, self.odim), out)
        return out.view(lg.shape[0], lg.shape[1]).long()

    @torch.no_grad()
    def greedy(self, hs_pad, hlens, hs_split=None, blank=0):
        """argmax -> unique_consecutive -> drop blank, on device.  Returns (ids (B, T) int32, counts (B,) int32, argmax (B,T) int32)."""
        lg = self.logits(hs_pad, hs_split)
        B, T, V = lg.shape
        am = torch.empty(B, T, dtype=torch.int32, device=lg.device)
        ops.argmax_rows(lg.view(-1, V), am)
        ids = torch.zeros(B, T, dtype=torch.int32, device=lg.device)
        cnt = torch.zeros(B, dtype=torch.int32, device=lg.device)
        lens32 = hlens.to(device=lg.device, dtype=torch.int32).contiguous()
        call("espb_ctc_collapse_i32", ptr(am), B, T, ptr(lens32), blank, ptr(ids), ptr(cnt))
        _count()
        return ids, cnt, am

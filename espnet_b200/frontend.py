"""DefaultFrontend / LogMel / UtteranceMVN with the reference's class surface, executed by the fused
CUDA frontend kernel (espnet_b200/csrc/frontend.cu).

Reference: espnet2/asr/frontend/default.py:24-131, espnet2/layers/stft.py, espnet2/layers/log_mel.py,
espnet2/layers/utterance_mvn.py.  Same constructor arguments, same ``state_dict`` key
(``logmel.melmat``), same ``forward(input, input_lengths) -> (feats, feats_lens)``.
"""
import math
from typing import Optional, Tuple, Union

import numpy as np
import torch

from . import lib
from .lib import call, ptr
from .ops import _count


def slaney_mel_matrix(sr=16000, n_fft=512, n_mels=80, fmin=0.0, fmax=None):
    """(n_fft/2+1, n_mels) Slaney-scale, area-normalised triangular filterbank = librosa.filters.mel(htk=False).T,
    which is what the reference registers as ``melmat`` (espnet2/layers/log_mel.py:50-52)."""
    fmax = sr / 2.0 if fmax is None else float(fmax)
    f_sp, min_log_hz = 200.0 / 3.0, 1000.0
    min_log_mel, logstep = min_log_hz / f_sp, math.log(6.4) / 27.0
    to_mel = lambda f: min_log_mel + math.log(f / min_log_hz) / logstep if f >= min_log_hz else f / f_sp  # noqa: E731
    mels = np.linspace(to_mel(float(fmin)), to_mel(fmax), n_mels + 2)
    hz = np.where(mels >= min_log_mel, min_log_hz * np.exp(logstep * (mels - min_log_mel)), f_sp * mels)
    bins = np.linspace(0.0, sr / 2.0, 1 + n_fft // 2)
    ramps = hz[:, None] - bins[None, :]
    fd = np.diff(hz)
    w = np.maximum(0.0, np.minimum(-ramps[:-2] / fd[:-1, None], ramps[2:] / fd[1:, None]))
    w *= (2.0 / (hz[2:] - hz[:-2]))[:, None]
    return torch.from_numpy(np.ascontiguousarray(w.T).astype(np.float32))


class LogMel(torch.nn.Module):
    """Holds the mel matrix buffer (state_dict key ``melmat``); the matmul+log runs inside the fused kernel."""

    def __init__(self, fs=16000, n_fft=512, n_mels=80, fmin=None, fmax=None, htk=False, log_base=None):
        super().__init__()
        if htk or log_base is not None:
            raise NotImplementedError("espnet_b200 LogMel: only htk=False, natural log (the reference defaults)")
        fmin = 0 if fmin is None else fmin
        fmax = fs / 2 if fmax is None else fmax
        self.mel_options = dict(sr=fs, n_fft=n_fft, n_mels=n_mels, fmin=fmin, fmax=fmax, htk=htk)
        self.register_buffer("melmat", slaney_mel_matrix(fs, n_fft, n_mels, fmin, fmax))


# The fused kernel also emits per-block column sums; they travel with the feats tensor object itself (attribute ``_espb_partial``), so that
# UtteranceMVN can skip its reduction pass when it is handed exactly that tensor (no module-global state: re-entrant).


class DefaultFrontend(torch.nn.Module):
    """Stft -> power -> log-mel, one fused kernel.  Single-channel only (the WPE/beamformer branch of the
    reference is an identity for 3-D input, legacy/.../frontends/frontend.py:104)."""

    def __init__(self, fs: Union[int, str] = 16000, n_fft: int = 512, win_length: Optional[int] = None, hop_length: int = 128,
                 window: Optional[str] = "hann", center: bool = True, normalized: bool = False, onesided: bool = True,
                 n_mels: int = 80, fmin: Optional[int] = None, fmax: Optional[int] = None, htk: bool = False,
                 frontend_conf: Optional[dict] = None, apply_stft: bool = True):
        super().__init__()
        if isinstance(fs, str):
            s = fs.strip().lower()
            fs = int(float(s[:-1]) * {"k": 1000, "m": 1000000}[s[-1]]) if s[-1] in "km" else int(s)
        win_length = n_fft if win_length is None else int(win_length)
        if (n_fft, center, normalized, onesided, apply_stft) != (512, True, False, True, True) or not (0 < win_length <= n_fft) or hop_length < 1:
            raise NotImplementedError("espnet_b200 DefaultFrontend: n_fft=512, center=True, onesided, not normalized (any hop_length, "
                                      "win_length <= 512, any torch window function or None)")
        if window is not None and not hasattr(torch, f"{window}_window"):
            raise ValueError(f"{window} window is not implemented")       # stft.py:41-42
        self.hop_length, self.n_fft, self.n_mels, self.win_length, self.window = int(hop_length), n_fft, n_mels, win_length, window
        self.logmel = LogMel(fs=fs, n_fft=n_fft, n_mels=n_mels, fmin=fmin, fmax=fmax, htk=htk)
        self.frontend_type = "default"
        self._dev = None  # cached device-side constants

    def output_size(self) -> int:
        return self.n_mels

    def _constants(self, device):
        mm = self.logmel.melmat
        key = (str(device), mm._version, mm.data_ptr())
        if self._dev is not None and self._dev["key"] == key:
            return self._dev
        m = mm.detach().float().cpu().numpy()  # (257, n_mels)
        starts, counts, offsets, weights = [], [], [], []
        for j in range(m.shape[1]):
            nz = np.nonzero(m[:, j])[0]
            lo, hi = (int(nz[0]), int(nz[-1]) + 1) if nz.size else (0, 0)
            starts.append(lo); counts.append(hi - lo); offsets.append(len(weights)); weights.extend(m[lo:hi, j].tolist())
        k = np.arange(256, dtype=np.float64)
        tw = np.stack([np.cos(2 * np.pi * k / 512), -np.sin(2 * np.pi * k / 512)], axis=1).astype(np.float32)
        nk = np.outer(np.arange(16), np.arange(16)).astype(np.float64)          # [k1][n2] -> W256^{n2 k1} (four-step FFT twiddles)
        twt = np.stack([np.cos(2 * np.pi * nk / 256), -np.sin(2 * np.pi * nk / 256)], axis=-1).reshape(256, 2).astype(np.float32)
        # torch.stft: the window (win_length taps, periodic) is zero-padded on both sides to n_fft (stft.py:84-93)
        if self.window is not None:
            wwin = getattr(torch, f"{self.window}_window")(self.win_length, dtype=torch.float32)
        else:
            wwin = torch.ones(self.win_length, dtype=torch.float32)
        left = (self.n_fft - self.win_length) // 2
        wfull = torch.zeros(self.n_fft, dtype=torch.float32)
        wfull[left:left + self.win_length] = wwin
        i32 = lambda a: torch.tensor(a, dtype=torch.int32, device=device)  # noqa: E731
        self._dev = dict(key=key, start=i32(starts), count=i32(counts), offset=i32(offsets), nnz=len(weights),
                         weight=torch.tensor(weights if weights else [0.0], dtype=torch.float32, device=device),
                         tw=torch.from_numpy(tw).to(device), twt=torch.from_numpy(twt).to(device), window=wfull.to(device))
        return self._dev

    @torch.no_grad()
    def forward(self, input: torch.Tensor, input_lengths: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """input (B, L) float32 CUDA, input_lengths (B,) int64 -> feats (B, 1 + Lmax//hop_length, n_mels), feats_lens."""
        if input.dim() != 2:
            raise NotImplementedError("espnet_b200 DefaultFrontend: single-channel (B, L) input only")
        lib.load()
        input = input.contiguous().float()
        B, L = input.shape
        lens_dev = input_lengths.to(device=input.device, dtype=torch.int64).contiguous()
        lens_cpu = input_lengths.detach().cpu()
        if int(lens_cpu.min()) <= self.n_fft // 2:
            raise RuntimeError("reflect padding needs utterances longer than n_fft/2 samples (torch.stft raises too)")
        Tf = 1 + int(lens_cpu.max()) // self.hop_length
        c = self._constants(input.device)
        feats = torch.empty(B, Tf, self.n_mels, dtype=torch.float32, device=input.device)
        nblk = lib.load().espb_frontend_blocks(Tf)
        partial = torch.empty(B, nblk, self.n_mels, dtype=torch.float32, device=input.device)
        call("espb_stft_logmel_f32", ptr(input), ptr(lens_dev), B, L, self.hop_length, ptr(c["window"]), ptr(c["tw"]), ptr(c["twt"]), ptr(c["start"]),
             ptr(c["count"]), ptr(c["offset"]), ptr(c["weight"]), c["nnz"], self.n_mels, ptr(feats), Tf, ptr(partial))
        _count()
        feats._espb_partial = (partial, lens_dev, self.hop_length)
        feats_lens = torch.div(input_lengths, self.hop_length, rounding_mode="trunc") + 1
        return feats, feats_lens


class UtteranceMVN(torch.nn.Module):
    """espnet2/layers/utterance_mvn.py:9-88 (norm_means=True, norm_vars=False is the reference default and the
    only mode implemented).  In place, like the reference (``x -= mean``)."""

    def __init__(self, norm_means: bool = True, norm_vars: bool = False, eps: float = 1.0e-20):
        super().__init__()
        if not norm_means or norm_vars:
            raise NotImplementedError("espnet_b200 UtteranceMVN: norm_means=True, norm_vars=False only")
        self.norm_means, self.norm_vars, self.eps = norm_means, norm_vars, eps

    @torch.no_grad()
    def forward(self, x: torch.Tensor, ilens: torch.Tensor = None) -> Tuple[torch.Tensor, torch.Tensor]:
        B, T, D = x.shape
        if ilens is None:
            ilens = torch.full((B,), T, dtype=torch.int64, device=x.device)
        stash = getattr(x, "_espb_partial", None)
        if stash is not None and x.is_contiguous():
            partial, wave_lens, hop = stash
            x._espb_partial = None                       # the sums describe the un-normalised features only
            call("espb_utt_mvn_from_partial_f32", ptr(x), ptr(wave_lens), B, T, D, hop, ptr(partial))
            _count()
        else:
            x = x.contiguous()
            lens_dev = ilens.to(device=x.device, dtype=torch.int64).contiguous()
            ws = torch.empty(B, (T + 31) // 32, D, dtype=torch.float32, device=x.device)
            call("espb_utt_mvn_f32", ptr(x), ptr(lens_dev), B, T, D, ptr(ws))
            _count(2)
        return x, ilens


class GlobalMVN(torch.nn.Module):
    """espnet2/layers/global_mvn.py:12-103: mean/variance normalisation with statistics from an npy/npz file (same buffers
    ``mean`` / ``std``).  In place like the reference."""

    def __init__(self, stats_file, norm_means: bool = True, norm_vars: bool = True, eps: float = 1.0e-20):
        super().__init__()
        self.norm_means, self.norm_vars, self.eps, self.stats_file = norm_means, norm_vars, eps, stats_file
        stats = np.load(stats_file)
        if isinstance(stats, np.ndarray):   # Kaldi-like stats
            count = stats[0].flatten()[-1]
            mean = stats[0, :-1] / count
            var = stats[1, :-1] / count - mean * mean
        else:
            count, sum_v, sum_square_v = stats["count"], stats["sum"], stats["sum_square"]
            mean = sum_v / count
            var = sum_square_v / count - mean * mean
        std = np.sqrt(np.maximum(var, eps))
        self.register_buffer("mean", torch.from_numpy(np.asarray(mean)))
        self.register_buffer("std", torch.from_numpy(np.asarray(std)))

    @torch.no_grad()
    def forward(self, x: torch.Tensor, ilens: torch.Tensor = None) -> Tuple[torch.Tensor, torch.Tensor]:
        B, T, D = x.shape
        if ilens is None:
            ilens = torch.full((B,), T, dtype=torch.int64, device=x.device)
        assert x.is_contiguous()
        lens_dev = ilens.to(device=x.device, dtype=torch.int64).contiguous()
        mean = self.mean.to(device=x.device, dtype=torch.float32).contiguous()
        std = self.std.to(device=x.device, dtype=torch.float32).contiguous()
        call("espb_global_mvn_f32", ptr(x), ptr(lens_dev), B, T, D, ptr(mean), ptr(std), int(self.norm_means), int(self.norm_vars))
        _count()
        return x, ilens

"""Oracle: DefaultFrontend (Stft + power + LogMel) and UtteranceMVN.  TEST INFRASTRUCTURE."""
import math

import numpy as np
import torch


def slaney_mel_matrix(sr=16000, n_fft=512, n_mels=80, fmin=0.0, fmax=None):
    """librosa.filters.mel(htk=False) restated; call site espnet2/layers/log_mel.py:50-52.

    Third-party arithmetic (librosa>=0.10.2, pyproject.toml:40): Slaney mel scale (linear below
    1 kHz, log above, logstep = ln(6.4)/27), triangular filters on fftfreqs = linspace(0, sr/2,
    1+n_fft/2), Slaney area normalisation 2/(f[m+2]-f[m]); float64 then cast to float32.
    Returns melmat transposed as the reference stores it: (n_fft/2+1, n_mels).
    """
    fmax = sr / 2.0 if fmax is None else float(fmax)
    f_sp = 200.0 / 3.0
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = math.log(6.4) / 27.0

    def hz_to_mel(f):
        return min_log_mel + math.log(f / min_log_hz) / logstep if f >= min_log_hz else f / f_sp

    mels = np.linspace(hz_to_mel(float(fmin)), hz_to_mel(fmax), n_mels + 2)
    freqs = np.where(mels >= min_log_mel, min_log_hz * np.exp(logstep * (mels - min_log_mel)), f_sp * mels)
    fft_f = np.linspace(0.0, sr / 2.0, 1 + n_fft // 2)
    fdiff = np.diff(freqs)
    ramps = freqs[:, None] - fft_f[None, :]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    w = np.maximum(0.0, np.minimum(lower, upper))
    w *= (2.0 / (freqs[2:] - freqs[:-2]))[:, None]
    return torch.from_numpy(w.T.astype(np.float32).copy())


def stft_power(wave, n_fft=512, hop=128, win_length=None, window="hann"):
    """Stft.forward + power, one utterance.  espnet2/layers/stft.py:75-120 (torch.stft with
    center=True -> reflect pad n_fft//2, periodic hann(win=n_fft), onesided, normalized=False);
    power = re^2+im^2, espnet2/asr/frontend/default.py:110.   wave (L,) -> (T_f, n_fft/2+1),
    T_f = 1 + L // hop (stft.py:107-115)."""
    L = wave.numel()
    pad = n_fft // 2
    x = torch.nn.functional.pad(wave.view(1, 1, L), (pad, pad), mode="reflect").view(-1)
    n_frames = 1 + L // hop
    frames = x.unfold(0, n_fft, hop)[:n_frames]
    # window_func(win_length), periodic; torch.stft zero-pads it on both sides to n_fft (stft.py:84-93); window None = rectangular
    win_length = n_fft if win_length is None else win_length
    w = getattr(torch, f"{window}_window")(win_length, dtype=wave.dtype) if window is not None else torch.ones(win_length, dtype=wave.dtype)
    left = (n_fft - win_length) // 2
    win = torch.zeros(n_fft, dtype=wave.dtype)
    win[left:left + win_length] = w
    spec = torch.fft.rfft(frames * win, dim=-1)
    return spec.real**2 + spec.imag**2


def log_mel(power, melmat):
    """LogMel.forward: matmul, clamp(1e-10), natural log.  espnet2/layers/log_mel.py:57-84."""
    return torch.clamp(power @ melmat, min=1e-10).log()


def utterance_mvn(feats):
    """UtteranceMVN defaults (norm_means=True, norm_vars=False): subtract the per-utterance
    per-bin mean over frames.  espnet2/layers/utterance_mvn.py:45-88."""
    return feats - feats.sum(dim=0, keepdim=True) / feats.shape[0]


def global_mvn_stats(stats, eps=1.0e-20):
    """mean / std from accumulated statistics (espnet2/layers/global_mvn.py:41-57): either a dict with
    count / sum / sum_square, or a Kaldi-like [2][D+1] array whose last column holds the count."""
    if isinstance(stats, np.ndarray):
        count = stats[0].flatten()[-1]
        mean = stats[0, :-1] / count
        var = stats[1, :-1] / count - mean * mean
    else:
        count = stats["count"]
        mean = stats["sum"] / count
        var = stats["sum_square"] / count - mean * mean
    return torch.from_numpy(np.asarray(mean)), torch.from_numpy(np.asarray(np.sqrt(np.maximum(var, eps))))


def global_mvn(x, ilens, mean, std, norm_means=True, norm_vars=True):
    """GlobalMVN.forward (espnet2/layers/global_mvn.py:74-103) on a padded batch [B][T][D]: subtract the
    mean, zero the padded frames, divide by std (the float64 statistics are cast to x.dtype first, :86-87)."""
    x = x.clone()
    mean, std = mean.to(x.dtype), std.to(x.dtype)
    pad = torch.arange(x.shape[1])[None, :] >= ilens[:, None]
    if norm_means:
        x -= mean.to(x.device)
    x = x.masked_fill(pad[:, :, None], 0.0)
    if norm_vars:
        x /= std.to(x.device)
    return x


def frontend_forward(wave, melmat=None):
    """DefaultFrontend.forward for one single-channel utterance (frontend/default.py:82-117)."""
    melmat = slaney_mel_matrix() if melmat is None else melmat
    return log_mel(stft_power(wave), melmat)

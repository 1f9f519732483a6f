"""Import scaffolding for the UNMODIFIED reference at /root/reference (test infrastructure only).

Only used by ``tests/golden/make_golden.py`` (fixture generation) and by the optional
``tests/test_oracle_vs_reference.py`` (skipped when /root/reference is absent, e.g. on the GPU box).
Contains no reference code: three functional shims for absent third-party packages the path
touches (torch_complex.ComplexTensor container, humanfriendly.parse_size, librosa.filters.mel)
and an import finder that fabricates empty modules for absent packages the import chain names
but this path never calls.
"""
import importlib.abc
import importlib.machinery
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("ESPNET_REFERENCE_ROOT", "/root/reference")

_STUBS = (
    "fairscale g2p_en hydra jaconv jamo kaldiio loralib miditoolkit music21 omegaconf opt_einsum "
    "s3prl soundfile tacotron_cleaner torch_optimizer vietnamese_cleaner wandb whisper "
    "espnet_model_zoo lhotse k2 pyworld pypinyin nltk configargparse editdistance "
    "pytorch_lightning lightning transformers_stream_generator sacrebleu"
).split()


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "espnet2"))


class _ComplexTensor:
    """Container with the attributes espnet2/asr/frontend/default.py:98-110,130 touches."""

    def __init__(self, real, imag):
        self.real, self.imag = real, imag

    def dim(self):
        return self.real.dim()

    def size(self, *a):
        return self.real.size(*a)

    @property
    def shape(self):
        return self.real.shape

    def __len__(self):
        return len(self.real)

    def __getitem__(self, idx):
        return _ComplexTensor(self.real[idx], self.imag[idx])


def _slaney_mel(sr=16000, n_fft=512, n_mels=80, fmin=0.0, fmax=None, htk=False, **_):
    """librosa.filters.mel restated (librosa>=0.10.2, Slaney scale + Slaney area norm)."""
    assert not htk
    fmax = sr / 2 if fmax is None else fmax

    def hz2mel(f):
        f = np.asarray(f, dtype=np.float64)
        m = f / (200.0 / 3)
        min_log_hz, min_log_mel, logstep = 1000.0, 1000.0 / (200.0 / 3), np.log(6.4) / 27.0
        return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, m)

    def mel2hz(m):
        m = np.asarray(m, dtype=np.float64)
        f = m * (200.0 / 3)
        min_log_hz, min_log_mel, logstep = 1000.0, 1000.0 / (200.0 / 3), np.log(6.4) / 27.0
        return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f)

    fftfreqs = np.linspace(0, sr / 2, 1 + n_fft // 2)
    mel_f = mel2hz(np.linspace(hz2mel(fmin), hz2mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    w = np.zeros((n_mels, 1 + n_fft // 2))
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        w[i] = np.maximum(0, np.minimum(lower, upper))
    w *= (2.0 / (mel_f[2 : n_mels + 2] - mel_f[:n_mels]))[:, None]
    return w.astype(np.float32)


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, name, path=None, target=None):
        if name.split(".")[0] in _STUBS:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        m = types.ModuleType(spec.name)
        m.__path__ = []
        m.__version__ = "9.9.9"
        m.__getattr__ = lambda attr: _stub_attr(spec.name, attr)
        return m

    def exec_module(self, module):
        pass


def _stub_attr(mod, attr):
    if attr.startswith("__"):
        raise AttributeError(attr)
    return type(attr, (), {"__init__": lambda self, *a, **k: None})


def _mod(name, pkg=False):
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None, is_package=pkg)
    if pkg:
        m.__path__ = []
    return m


def install():
    """Make `import espnet2...` resolve to the reference with the shims in place."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    if "torch_complex" not in sys.modules:
        tc = _mod("torch_complex", True)
        tct = _mod("torch_complex.tensor"); tct.ComplexTensor = _ComplexTensor
        tcf = _mod("torch_complex.functional")
        tc.tensor, tc.functional, tc.ComplexTensor = tct, tcf, _ComplexTensor
        sys.modules.update({"torch_complex": tc, "torch_complex.tensor": tct, "torch_complex.functional": tcf})
    if "humanfriendly" not in sys.modules:
        hf = _mod("humanfriendly")

        def parse_size(s):
            s = str(s).strip().lower()
            mult = {"k": 1000, "m": 1000**2, "g": 1000**3}
            return int(float(s[:-1]) * mult[s[-1]]) if s[-1] in mult else int(s)

        hf.parse_size = parse_size
        sys.modules["humanfriendly"] = hf
    try:
        import librosa  # noqa: F401
    except Exception:
        lb = _mod("librosa", True)
        lf = _mod("librosa.filters"); lf.mel = _slaney_mel
        lb.filters = lf; lb.__version__ = "0.10.2"
        sys.modules.update({"librosa": lb, "librosa.filters": lf})
    if not any(isinstance(f, _StubFinder) for f in sys.meta_path):
        sys.meta_path.append(_StubFinder())

"""Generate tests/golden/streaming_enc.npz from the UNMODIFIED reference ContextualBlockConformerEncoder.forward_infer
(espnet2/asr/encoder/contextual_block_conformer_encoder.py:386-600) -- the streaming row.  Build container only.

    python tests/golden/make_golden_streaming.py

Three scenarios on seeded features: a stream pushed in uneven chunks (outputs of every push), one whole-utterance call (is_final on the first
call, several blocks), and a short segment (is_final on the first call, fewer frames than a block)."""
import logging
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refshim  # noqa: E402

logging.disable(logging.WARNING)
refshim.install()
from espnet2.asr.encoder.contextual_block_conformer_encoder import ContextualBlockConformerEncoder  # noqa: E402

CFG = dict(output_size=64, attention_heads=4, linear_units=96, num_blocks=2, cnn_module_kernel=15, block_size=40, hop_size=16, look_ahead=16)
torch.manual_seed(0)
enc = ContextualBlockConformerEncoder(80, input_layer="conv2d", macaron_style=True, use_cnn_module=True,
                                      activation_type="swish", normalize_before=True, **CFG).eval()
with torch.no_grad():
    for m in enc.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            m.running_mean.normal_(0, 0.1)
            m.running_var.uniform_(0.5, 1.5)
out = {"cfg_keys": np.array(list(CFG.keys())), "cfg_vals": np.array(list(CFG.values()), dtype=np.int64)}
for k, v in enc.state_dict().items():
    out["w:" + k] = v.numpy()
g = torch.Generator().manual_seed(1)
feats = torch.randn(1, 700, 80, generator=g)
out["feats"] = feats[0].numpy()
pushes = [64, 64, 37, 100, 64, 150, 221]
assert sum(pushes) == 700
out["pushes"] = np.array(pushes)
with torch.no_grad():
    states, pos = None, 0
    for i, n in enumerate(pushes):
        final = i == len(pushes) - 1
        y, ylen, states = enc(feats[:, pos:pos + n], torch.tensor([n]), states, is_final=final, infer_mode=True)
        out[f"stream:{i}:y"] = y[0].numpy()
        pos += n
    y, _, _ = enc(feats[:, :500], torch.tensor([500]), None, is_final=True, infer_mode=True)
    out["whole:y"] = y[0].numpy()
    y, _, _ = enc(feats[:, :131], torch.tensor([131]), None, is_final=True, infer_mode=True)     # 31 frames after subsampling < block_size
    out["short:y"] = y.reshape(-1, y.shape[-1]).numpy()
# ---- Speech2TextStreaming.apply_frontend (espnet2/bin/asr_inference_streaming.py:205-294): waveform chunking, overlap buffer, edge-frame trimming
import types  # noqa: E402

from espnet2.asr.frontend.default import DefaultFrontend  # noqa: E402
from espnet2.bin.asr_inference_streaming import Speech2TextStreaming  # noqa: E402

fe = DefaultFrontend(fs=16000, n_fft=512, hop_length=128, n_mels=80).eval()
fake = types.SimpleNamespace(win_length=512, hop_length=128, dtype="float32", device="cpu",
                             asr_model=types.SimpleNamespace(_extract_feats=lambda speech, speech_lengths: fe(speech, speech_lengths), normalize=None))
wave = 0.1 * torch.randn(12000, generator=g)
wpush = [640, 640, 300, 2000, 640, 640, 7140]
assert sum(wpush) == 12000
out["wave"], out["wave_pushes"] = wave.numpy(), np.array(wpush)
with torch.no_grad():
    st, pos = None, 0
    for i, n in enumerate(wpush):
        feats, fl, st = Speech2TextStreaming.apply_frontend(fake, wave[pos:pos + n], st, is_final=(i == len(wpush) - 1))
        out[f"fe:{i}:feats"] = np.zeros((0, 80), np.float32) if feats is None else feats[0].numpy()
        pos += n
path = os.path.join(HERE, "streaming_enc.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path) // 1024, "KiB", {k: v.shape for k, v in out.items() if ":y" in k})

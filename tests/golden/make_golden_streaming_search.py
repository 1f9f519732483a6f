"""Generate tests/golden/streaming_search.npz from the UNMODIFIED reference Speech2TextStreaming (espnet2/bin/asr_inference_streaming.py:35-357:
apply_frontend -> ContextualBlockConformerEncoder.forward_infer -> BatchBeamSearchOnline with decoder + CTC prefix scorer + length bonus) -- the
streaming row.  Build container only.

    python tests/golden/make_golden_streaming_search.py

A tiny random-init model (seed 0); one waveform pushed in uneven chunks; for several decoding settings the n-best of EVERY push (token ids, total
score, per-scorer scores) as the reference returned them."""
import json
import logging
import os
import sys
import tempfile

import numpy as np
import torch
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refbuild  # noqa: E402
import refshim  # noqa: E402

logging.disable(logging.WARNING)
refshim.install()
from espnet2.bin.asr_inference_streaming import Speech2TextStreaming  # noqa: E402

CFG = dict(d_model=64, heads=4, ff=96, enc_layers=2, dec_layers=1, vocab=30, kernel=15)
y = refbuild.model_yaml(CFG)
y.update(encoder="contextual_block_conformer", normalize=None, normalize_conf={},
         encoder_conf=dict(output_size=64, attention_heads=4, linear_units=96, num_blocks=2, macaron_style=True, use_cnn_module=True,
                           cnn_module_kernel=15, block_size=40, hop_size=16, look_ahead=16, input_layer="conv2d", activation_type="swish",
                           normalize_before=True, dropout_rate=0.0, positional_dropout_rate=0.0, attention_dropout_rate=0.0))
tmp = tempfile.mkdtemp(prefix="espref_stream_")
cfg_path = os.path.join(tmp, "config.yaml")
with open(cfg_path, "w") as f:
    yaml.safe_dump(y, f)

LM = dict(embed_unit=32, att_unit=64, head=4, unit=96, layer=2)
lm_path = os.path.join(tmp, "lm.yaml")
with open(lm_path, "w") as f:
    yaml.safe_dump(dict(token_list=refbuild.token_list(CFG["vocab"]), lm="transformer",
                        lm_conf=dict(pos_enc="sinusoidal", dropout_rate=0.1, positional_dropout_rate=0.1, attention_dropout_rate=0.1, **LM),
                        model_conf={}, init=None, use_preprocessor=False), f)

SETTINGS = {
    "joint": dict(beam_size=3, ctc_weight=0.3),
    "joint_pen_norep": dict(beam_size=4, ctc_weight=0.5, penalty=0.4, disable_repetition_detection=True, nbest=3),
    "ctc_only": dict(beam_size=3, ctc_weight=1.0, nbest=2),
    "att_heavy_maxlen": dict(beam_size=2, ctc_weight=0.1, maxlenratio=0.2, nbest=2),
    "joint_lm": dict(beam_size=3, ctc_weight=0.3, lm_weight=0.5, nbest=2),       # + TransformerLM shallow fusion (lm_train_config)
    "joint_normlen_minlen": dict(beam_size=3, ctc_weight=0.4, normalize_length=True, minlenratio=0.3, maxlenratio=0.6, penalty=0.2, nbest=3),
}
wave = refbuild.waveform(7, 52000)
pushes = [8000, 640, 640, 9000, 12000, 3000, 18720]
assert sum(pushes) == wave.numel()
out = {"lm_conf": np.array(json.dumps(dict(pos_enc="sinusoidal", **LM))), "yaml": np.array(json.dumps(y)), "wave": wave.numpy(), "pushes": np.array(pushes), "settings": np.array(json.dumps(SETTINGS))}
weights = None
for name, kw in SETTINGS.items():
    torch.manual_seed(0)
    extra = dict(lm_train_config=lm_path, lm_file=None) if "lm_weight" in kw else {}
    s2t = Speech2TextStreaming(asr_train_config=cfg_path, asr_model_file=None, device="cpu", **kw, **extra)
    if extra:
        for k, v in s2t.beam_search.full_scorers["lm"].state_dict().items():
            out["lm:" + k] = v.numpy()
    with torch.no_grad():
        for m in s2t.asr_model.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.running_mean.fill_(0.05)
                m.running_var.fill_(1.2)
    if weights is None:
        weights = {k: v.detach().clone() for k, v in s2t.asr_model.state_dict().items()}
        for k, v in weights.items():
            out["w:" + k] = v.numpy()
    else:
        for k, v in s2t.asr_model.state_dict().items():
            assert torch.equal(v, weights[k]), k
    pos = 0
    n_nonempty = 0
    for i, n in enumerate(pushes):
        res = s2t(wave[pos:pos + n], is_final=(i == len(pushes) - 1))
        pos += n
        out[f"{name}:{i}:n"] = np.array(len(res))
        n_nonempty += len(res) > 0
        for j, (_, _, token_int, hyp) in enumerate(res):
            out[f"{name}:{i}:{j}:yseq"] = np.array(hyp.yseq.tolist(), dtype=np.int64)
            out[f"{name}:{i}:{j}:token_int"] = np.array(token_int, dtype=np.int64)
            out[f"{name}:{i}:{j}:score"] = np.array(float(hyp.score), dtype=np.float64)
            for k, v in hyp.scores.items():
                out[f"{name}:{i}:{j}:score:{k}"] = np.array(float(v), dtype=np.float64)
    print(name, "pushes with output:", n_nonempty, "final:", [out[f"{name}:{len(pushes) - 1}:{j}:yseq"].tolist() for j in range(int(out[f"{name}:{len(pushes) - 1}:n"]))])
path = os.path.join(HERE, "streaming_search.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path) // 1024, "KiB")

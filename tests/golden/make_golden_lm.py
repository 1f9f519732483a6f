"""Generate tests/golden/tiny_lm.npz from the UNMODIFIED reference Speech2Text with a TransformerLM scorer (LM shallow fusion,
espnet2/bin/asr_inference.py:178-191, espnet2/lm/transformer_lm.py) -- groundwork for the LM row.  Build container only.

    python tests/golden/make_golden_lm.py
"""
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
from espnet2.bin.asr_inference import Speech2Text  # noqa: E402

CFG = dict(d_model=64, heads=4, ff=128, enc_layers=2, dec_layers=2, vocab=50, kernel=15)
LM = dict(embed_unit=32, att_unit=64, head=4, unit=96, layer=2)
DECODES = [("joint_lm", 4, 0.3, 0.6, -8.0), ("att_lm", 3, 0.0, 0.4, -6.0), ("ctc_lm", 3, 1.0, 0.5, -5.0)]   # name, beam, ctc_weight, lm_weight, maxlenratio

tmp = tempfile.mkdtemp(prefix="espref_lm_")
asr_yaml, lm_yaml = os.path.join(tmp, "asr.yaml"), os.path.join(tmp, "lm.yaml")
yaml.safe_dump(refbuild.model_yaml(CFG), open(asr_yaml, "w"))
yaml.safe_dump(dict(token_list=refbuild.token_list(CFG["vocab"]), lm="transformer",
                    lm_conf=dict(pos_enc="sinusoidal", dropout_rate=0.1, positional_dropout_rate=0.1, attention_dropout_rate=0.1, **LM),
                    model_conf={}, init=None, use_preprocessor=False), open(lm_yaml, "w"))
wave = refbuild.waveform(0, 12000)
out = {"cfg_keys": np.array(list(CFG.keys())), "cfg_vals": np.array(list(CFG.values()), dtype=np.int64),
       "lm_keys": np.array(list(LM.keys())), "lm_vals": np.array(list(LM.values()), dtype=np.int64), "wave": wave.numpy()}
for j, (dn, beam, cw, lw, mlr) in enumerate(DECODES):
    torch.manual_seed(0)
    s2t = Speech2Text(asr_train_config=asr_yaml, asr_model_file=None, lm_train_config=lm_yaml, lm_file=None, device="cpu", dtype="float32",
                      beam_size=beam, ctc_weight=cw, lm_weight=lw, maxlenratio=mlr, nbest=10)
    if j == 0:
        for k, v in s2t.asr_model.state_dict().items():
            out["w:" + k] = v.numpy()
        lm = s2t.beam_search.full_scorers["lm"]
        for k, v in lm.state_dict().items():
            out["w:lm." + k] = v.numpy()
    res = s2t(wave)
    out[f"dec:{dn}:params"] = np.array([beam, cw, lw, mlr], dtype=np.float64)
    out[f"dec:{dn}:n"] = np.array(len(res))
    for i, (_, _, ids, hyp) in enumerate(res):
        out[f"dec:{dn}:{i}:yseq"] = hyp.yseq.numpy()
        out[f"dec:{dn}:{i}:score"] = np.array(float(hyp.score))
        out[f"dec:{dn}:{i}:scores"] = np.array([float(hyp.scores.get(k, np.nan)) for k in ("decoder", "ctc", "lm")])
path = os.path.join(HERE, "tiny_lm.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path) // 1024, "KiB", [(dn, int(out[f'dec:{dn}:n'])) for dn, *_ in DECODES])

"""Generate tests/golden/transformer_enc.npz from the UNMODIFIED reference TransformerEncoder
(espnet2/asr/encoder/transformer_encoder.py) -- the encoder of the next scope row.  Build container only.

    python tests/golden/make_golden_transformer.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refshim  # noqa: E402

refshim.install()
from espnet2.asr.encoder.transformer_encoder import TransformerEncoder  # noqa: E402

CFG = dict(d_model=64, heads=4, ff=96, layers=3)
torch.manual_seed(0)
enc = TransformerEncoder(80, output_size=CFG["d_model"], attention_heads=CFG["heads"], linear_units=CFG["ff"], num_blocks=CFG["layers"],
                         input_layer="conv2d", normalize_before=True, use_flash_attn=False).eval()
g = torch.Generator().manual_seed(5)
feats = torch.randn(1, 83, 80, generator=g)
with torch.no_grad():
    (out, inter), olens, _ = enc(feats, torch.tensor([83]), return_all_hs=True)
z = {"cfg_keys": np.array(list(CFG.keys())), "cfg_vals": np.array(list(CFG.values()), dtype=np.int64), "feats": feats[0].numpy(),
     "out": out[0].numpy(), "olens": olens.numpy()}
for i, h in enumerate(inter):
    z[f"layer{i + 1}"] = h[0].numpy()
for k, v in enc.state_dict().items():
    z["w:encoder." + k] = v.numpy()
np.savez_compressed(os.path.join(HERE, "transformer_enc.npz"), **z)
print("wrote transformer_enc.npz", out.shape, olens, sorted(k for k in z if not k.startswith("w:")))

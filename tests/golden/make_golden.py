"""Generate tests/golden/*.npz from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py

Imports the real espnet2.bin.asr_inference.Speech2Text from /root/reference (via refshim),
builds random-init models with torch.manual_seed(0), runs them on seeded synthetic waveforms and
stores inputs, weights and stage-boundary outputs.  These are the golden vectors the oracle is
pinned to (the reference holds none for this path).
"""
import logging
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refbuild  # noqa: E402

logging.disable(logging.WARNING)

CASES = {
    "tiny": dict(cfg=dict(d_model=64, heads=4, ff=128, enc_layers=2, dec_layers=2, vocab=50, kernel=15),
                 nsamples=12000, wave_id=0),
    "small": dict(cfg=dict(d_model=96, heads=2, ff=192, enc_layers=3, dec_layers=1, vocab=97, kernel=31),
                  nsamples=20000, wave_id=3),
    # next scope row: abs-pos TransformerEncoder + TransformerDecoder
    "tfm": dict(cfg=dict(d_model=64, heads=4, ff=128, enc_layers=3, dec_layers=2, vocab=50), encoder="transformer",
                nsamples=16000, wave_id=5),
}
DECODES = [  # (name, beam, ctc_weight, maxlenratio, minlenratio, penalty, normalize_length)
    ("joint", 4, 0.3, -8.0, 0.0, 0.0, False),
    ("joint_auto", 5, 0.5, 0.0, 0.0, 0.0, False),
    ("att", 3, 0.0, -6.0, 0.0, 0.0, False),
    ("ctc", 3, 1.0, -5.0, 0.0, 0.0, False),
    ("joint_pen", 4, 0.3, -7.0, -2.0, 0.5, True),
]


def run_case(name, spec):
    cfg = spec["cfg"]
    out = {"cfg_keys": np.array(list(cfg.keys())), "cfg_vals": np.array(list(cfg.values()), dtype=np.int64)}
    if "encoder" in spec:
        cfg = dict(cfg, encoder=spec["encoder"])
        out["encoder_type"] = np.array(spec["encoder"])
    wave = refbuild.waveform(spec["wave_id"], spec["nsamples"])
    out["wave"] = wave.numpy()
    s2t = refbuild.build_reference(cfg, seed=0, beam_size=4, ctc_weight=0.3, maxlenratio=-8.0, nbest=10)
    model = s2t.asr_model
    for k, v in model.state_dict().items():
        out["w:" + k] = v.numpy()
    with torch.no_grad():
        lens = torch.tensor([wave.numel()])
        feats, flens = model.frontend(wave[None], lens)
        out["feats"] = feats[0].numpy().copy()
        norm, _ = model.normalize(feats.clone(), flens)
        out["feats_norm"] = norm[0].numpy()
        layers = []
        def first(o):   # conformer modules return ((x, pos_emb), mask), transformer modules (x, mask)
            x = o[0]
            return (x[0] if isinstance(x, tuple) else x)[0].clone()

        hooks = [model.encoder.embed.register_forward_hook(lambda m, i, o: layers.append(first(o)))]
        for lyr in model.encoder.encoders:
            hooks.append(lyr.register_forward_hook(lambda m, i, o: layers.append(first(o))))
        enc, olens = model.encode(wave[None], lens)
        for h in hooks:
            h.remove()
        for i, t in enumerate(layers):
            out[f"layer{i}"] = t.numpy()
        out["enc"] = enc[0].numpy()
        out["ctc_logits"] = model.ctc.ctc_lo(enc)[0].numpy()
        out["ctc_logp"] = model.ctc.log_softmax(enc)[0].numpy()
        am = model.ctc.argmax(enc)[0]
        out["ctc_argmax"] = am.numpy()
        ids = torch.unique_consecutive(am)
        out["ctc_greedy"] = ids[ids != 0].numpy()
    for (dn, beam, cw, mlr, minr, pen, nl) in DECODES:
        s = refbuild.build_reference(cfg, seed=0, beam_size=beam, ctc_weight=cw, maxlenratio=mlr, minlenratio=minr,
                                     penalty=pen, normalize_length=nl, nbest=10)
        res = s(wave)
        out[f"dec:{dn}:params"] = np.array([beam, cw, mlr, minr, pen, float(nl)], dtype=np.float64)
        out[f"dec:{dn}:n"] = np.array(len(res))
        for j, (_, _, ids, hyp) in enumerate(res):
            out[f"dec:{dn}:{j}:yseq"] = hyp.yseq.numpy()
            out[f"dec:{dn}:{j}:score"] = np.array(float(hyp.score))
            out[f"dec:{dn}:{j}:scores"] = np.array([float(hyp.scores.get(k, np.nan)) for k in ("decoder", "ctc", "length_bonus")])
    path = os.path.join(HERE, f"{name}.npz")
    np.savez_compressed(path, **out)
    print(name, "->", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    for name in (sys.argv[1:] or list(CASES)):   # optional case names: `make_golden.py tfm`
        run_case(name, CASES[name])

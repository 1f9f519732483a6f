"""Next scope row: TransformerEncoder (abs-pos) on the CUDA path vs the reference fixture and the oracle.

The oracle side is pinned to the reference on CPU (tests/test_oracle_golden.py).  Tolerance: encoder outputs atol 1e-4 (default GEMM)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _load():
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "transformer_enc.npz"))
    cfg = dict(zip(z["cfg_keys"].tolist(), z["cfg_vals"].tolist()))
    w = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w:")}
    return z, cfg, w


def _build(cfg, w):
    import espnet_b200

    enc = espnet_b200.TransformerEncoder(80, output_size=cfg["d_model"], attention_heads=cfg["heads"], linear_units=cfg["ff"],
                                         num_blocks=cfg["layers"])
    enc.load_state_dict({k[len("encoder."):]: v for k, v in w.items()}, strict=True)
    return enc.cuda().eval()


def test_transformer_encoder_vs_reference_fixture():
    z, cfg, w = _load()
    enc = _build(cfg, w)
    enc.trace = []
    feats = torch.from_numpy(z["feats"])[None].cuda()
    out, olens, _ = enc(feats, torch.tensor([feats.shape[1]]).cuda())
    assert int(olens[0]) == int(z["olens"][0]) == out.shape[1]
    for i in range(cfg["layers"]):
        err = float((enc.trace[i + 1][0].cpu() - torch.from_numpy(z[f"layer{i + 1}"])).abs().max())
        print(f"layer {i + 1} max abs err {err:.3e}")
        assert err < TOL
    assert float((out[0].cpu() - torch.from_numpy(z["out"])).abs().max()) < TOL


def test_transformer_encoder_ragged_batch_vs_oracle():
    from oracle import transformer_encoder as TE

    z, cfg, w = _load()
    enc = _build(cfg, w)
    g = torch.Generator().manual_seed(3)
    lens = [700, 233, 480]            # T = 174, 57, 119: masked keys and rows longer than 128
    feats = torch.zeros(len(lens), max(lens), 80)
    for i, n in enumerate(lens):
        feats[i, :n] = torch.randn(n, 80, generator=g)
    out, olens, _ = enc(feats.cuda(), torch.tensor(lens).cuda())
    for i, n in enumerate(lens):
        ref = TE.transformer_encode(feats[i, :n], w, cfg["heads"], cfg["layers"])
        assert int(olens[i]) == ref.shape[0]
        e = float((out[i, : ref.shape[0]].cpu() - ref).abs().max())
        print(f"utt{i} (T={ref.shape[0]}) max abs err {e:.3e}")
        assert e < TOL


def test_transformer_enc_dec_speech2text_vs_reference_fixture():
    """Whole path with the TransformerEncoder (tests/golden/tfm.npz, made from the reference Speech2Text): encoder output, greedy ids,
    and the n-best lists of the attention-only / joint / CTC-only decode settings."""
    from golden_util import DEC_NAMES, decode_params, decode_results, load
    from gpu_util import speech2text

    z, cfg, w = load("tfm")
    s2t = speech2text(cfg, w, beam_size=2, ctc_weight=0.3)
    wave = torch.from_numpy(z["wave"])
    speech, sl = s2t._to_batch([wave])
    enc, _ = s2t.asr_model.encode(speech, sl)
    assert float((enc[0].cpu() - torch.from_numpy(z["enc"])).abs().max()) < TOL
    assert s2t.ctc_greedy([wave])[0] == z["ctc_greedy"].tolist()
    for dn in DEC_NAMES:
        res = speech2text(cfg, w, nbest=10, **decode_params(z, dn))(z["wave"])
        gold = decode_results(z, dn)
        assert len(res) == len(gold), dn
        for (_, _, _, h), (yseq, score, _) in zip(res, gold):
            assert h.yseq.tolist() == yseq, dn
            assert abs(h.score - score) <= 2e-4 * max(1.0, abs(score))

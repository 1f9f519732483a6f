"""The oracle is pinned to outputs of the reference itself (tests/golden/*.npz)."""
import numpy as np
import pytest
import torch

import oracle
from oracle import encoder as E
from oracle import frontend as Fr
from golden_util import DEC_NAMES, decode_params, decode_results, load

CASES = ["tiny", "small"]


@pytest.mark.parametrize("case", CASES)
def test_mel_matrix_matches_reference_buffer(case):
    z, cfg, w = load(case)
    # reference melmat came from the librosa restatement in refshim; cross-check independent formula
    import torchaudio

    ta = torchaudio.functional.melscale_fbanks(257, 0.0, 8000.0, 80, 16000, norm="slaney", mel_scale="slaney")
    m = Fr.slaney_mel_matrix()
    assert (m - ta).abs().max() < 1e-7
    assert (m - w["frontend.logmel.melmat"]).abs().max() < 1e-7


@pytest.mark.parametrize("case", CASES)
def test_frontend_and_mvn(case):
    z, cfg, w = load(case)
    wave = torch.from_numpy(z["wave"])
    feats = Fr.frontend_forward(wave, w["frontend.logmel.melmat"])
    assert feats.shape == z["feats"].shape
    # reference's own STFT self-consistency tolerance is atol 7e-6 (test/espnet2/layers/test_stft.py:43-55)
    np.testing.assert_allclose(feats.numpy(), z["feats"], atol=2e-4, rtol=1e-5)
    np.testing.assert_allclose(Fr.utterance_mvn(feats).numpy(), z["feats_norm"], atol=2e-4, rtol=1e-5)


@pytest.mark.parametrize("case", CASES)
def test_encoder_layers_and_ctc(case):
    z, cfg, w = load(case)
    feats = torch.from_numpy(z["feats_norm"])
    enc, layers = E.conformer_encode(feats, w, cfg["heads"], cfg["enc_layers"], return_layers=True)
    for i, t in enumerate(layers):
        np.testing.assert_allclose(t.numpy(), z[f"layer{i}"], atol=1e-5, rtol=1e-5)
    np.testing.assert_allclose(enc.numpy(), z["enc"], atol=1e-5, rtol=1e-5)
    np.testing.assert_allclose(E.ctc_logits(enc, w).numpy(), z["ctc_logits"], atol=1e-5, rtol=1e-5)
    am, ids = E.ctc_greedy(enc, w)
    assert am.tolist() == z["ctc_argmax"].tolist()
    assert ids.tolist() == z["ctc_greedy"].tolist()


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("dn", DEC_NAMES)
def test_beam_search_nbest(case, dn):
    z, cfg, w = load(case)
    o = oracle.OracleSpeech2Text(cfg, w, nbest=10, **decode_params(z, dn))
    res = o(torch.from_numpy(z["wave"]))
    gold = decode_results(z, dn)
    assert len(res) == len(gold)
    for (_, _, _, h), (yseq, score, scores) in zip(res, gold):
        assert h.yseq.tolist() == yseq
        assert abs(h.score - score) <= 1e-4 * max(1.0, abs(score))  # rtol 1e-4: test_transformer_decode.py:9
        for k, ref in zip(("decoder", "ctc", "length_bonus"), scores):
            if not np.isnan(ref):
                assert abs(h.scores[k] - ref) <= 1e-4 * max(1.0, abs(ref))


def test_too_short_utterance_raises():
    z, cfg, w = load("tiny")
    o = oracle.OracleSpeech2Text(cfg, w, beam_size=2, ctc_weight=0.3)
    with pytest.raises(E.TooShortUttError):
        o(torch.zeros(700))  # 6 frames < 7 (subsampling.py:43-44)


def test_global_mvn_vs_reference_fixture():
    import os

    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gmvn.npz"))
    stats = {k: z["stats_" + k] for k in ("count", "sum", "sum_square")}
    mean, std = Fr.global_mvn_stats(stats)
    np.testing.assert_array_equal(mean.float().numpy(), z["mean"])   # the reference casts its buffers to x.dtype on first use
    np.testing.assert_array_equal(std.float().numpy(), z["std"])
    kaldi = np.zeros((2, 81))
    kaldi[0, :80], kaldi[1, :80], kaldi[0, 80] = stats["sum"], stats["sum_square"], stats["count"]
    mk, sk = Fr.global_mvn_stats(kaldi)
    np.testing.assert_array_equal(mk.float().numpy(), z["mean"])
    np.testing.assert_array_equal(sk.float().numpy(), z["std"])
    for nm in (1, 0):
        for nv in (1, 0):
            y = Fr.global_mvn(torch.from_numpy(z["x"]), torch.from_numpy(z["ilens"]), mean, std, bool(nm), bool(nv))
            np.testing.assert_array_equal(y.numpy(), z[f"y_m{nm}_v{nv}"])


def test_transformer_encoder_oracle_vs_reference_fixture():
    """Next scope row: the abs-pos TransformerEncoder restatement is pinned to the reference's own outputs
    (tests/golden/transformer_enc.npz from tests/golden/make_golden_transformer.py), layer by layer."""
    import os

    from oracle import transformer_encoder as TE

    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "transformer_enc.npz"))
    cfg = dict(zip(z["cfg_keys"].tolist(), z["cfg_vals"].tolist()))
    w = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w:")}
    out, layers = TE.transformer_encode(torch.from_numpy(z["feats"]), w, cfg["heads"], cfg["layers"], return_layers=True)
    assert out.shape[0] == int(z["olens"][0])
    for i in range(cfg["layers"]):
        np.testing.assert_allclose(layers[i + 1].numpy(), z[f"layer{i + 1}"], atol=2e-5, rtol=1e-5)
    np.testing.assert_allclose(out.numpy(), z["out"], atol=2e-5, rtol=1e-5)


def test_transformer_enc_dec_pipeline_vs_reference_fixture():
    """Next scope row (BASELINE configs[4] shape family): the reference Speech2Text with a TransformerEncoder
    (tests/golden/tfm.npz): encoder layers, CTC head and the n-best lists of all five decode settings."""
    from oracle import transformer_encoder as TE

    z, cfg, w = load("tfm")
    assert cfg["encoder"] == "transformer"
    enc, layers = TE.transformer_encode(torch.from_numpy(z["feats_norm"]), w, cfg["heads"], cfg["enc_layers"], return_layers=True)
    for i, t in enumerate(layers):
        np.testing.assert_allclose(t.numpy(), z[f"layer{i}"], atol=2e-5, rtol=1e-5)
    np.testing.assert_allclose(enc.numpy(), z["enc"], atol=2e-5, rtol=1e-5)
    am, ids = E.ctc_greedy(enc, w)
    assert am.tolist() == z["ctc_argmax"].tolist() and ids.tolist() == z["ctc_greedy"].tolist()
    for dn in DEC_NAMES:
        o = oracle.OracleSpeech2Text(cfg, w, nbest=10, **decode_params(z, dn))
        res = o(torch.from_numpy(z["wave"]))
        gold = decode_results(z, dn)
        assert len(res) == len(gold), dn
        for (_, _, _, h), (yseq, score, scores) in zip(res, gold):
            assert h.yseq.tolist() == yseq, dn
            assert abs(h.score - score) <= 1e-4 * max(1.0, abs(score))


@pytest.mark.parametrize("dn", ["joint_lm", "att_lm", "ctc_lm"])
def test_lm_fusion_oracle_vs_reference_fixture(dn):
    """Groundwork for the LM row: TransformerLM as a full scorer (asr_inference.py:178-191) -- n-best lists and the
    per-scorer score split of the reference Speech2Text with an LM (tests/golden/tiny_lm.npz)."""
    import os

    from oracle import encoder as OE
    from oracle.lm import OracleLM
    from oracle.search import OracleDecoder, batch_beam_search

    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tiny_lm.npz"))
    cfg = dict(zip(z["cfg_keys"].tolist(), z["cfg_vals"].tolist()))
    lmc = dict(zip(z["lm_keys"].tolist(), z["lm_vals"].tolist()))
    w = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w:")}
    beam, cw, lw, mlr = z[f"dec:{dn}:params"].tolist()
    o = oracle.OracleSpeech2Text(cfg, {k: v for k, v in w.items() if not k.startswith("lm.")}, beam_size=int(beam), ctc_weight=cw, maxlenratio=mlr)
    enc = o.encode(torch.from_numpy(z["wave"]))
    logp = torch.log_softmax(OE.ctc_logits(enc, o.w), dim=-1)
    dec = OracleDecoder(o.w, cfg["heads"], cfg["dec_layers"]) if cw != 1.0 else None
    lm = OracleLM({k: v for k, v in w.items() if k.startswith("lm.")}, lmc["head"], lmc["layer"])
    res = batch_beam_search(enc, dec, logp, beam_size=int(beam), ctc_weight=cw, vocab=cfg["vocab"], sos=cfg["vocab"] - 1, eos=cfg["vocab"] - 1,
                            maxlenratio=mlr, lm=lm, lm_weight=lw)[:10]
    n = int(z[f"dec:{dn}:n"])
    assert len(res) == n
    for i, h in enumerate(res):
        assert h.yseq.tolist() == z[f"dec:{dn}:{i}:yseq"].tolist()
        score = float(z[f"dec:{dn}:{i}:score"])
        assert abs(h.score - score) <= 1e-4 * max(1.0, abs(score))
        for k, ref in zip(("decoder", "ctc", "lm"), z[f"dec:{dn}:{i}:scores"]):
            if not np.isnan(ref):
                assert abs(h.scores[k] - ref) <= 1e-4 * max(1.0, abs(ref)), k


@pytest.mark.parametrize("hop,win_length,window", [(160, None, "hann"), (160, 400, "hann"), (100, 320, "hamming"), (75, 512, None)])
def test_oracle_stft_generalised_vs_torch_stft(hop, win_length, window):
    """The oracle's framing for other hop / window settings (recipes such as train_asr_conformer10_hop_length160.yaml:37-38) is torch.stft's, the
    call the reference makes (espnet2/layers/stft.py:94-105)."""
    import torch

    from oracle import frontend as OF

    g = torch.Generator().manual_seed(hop)
    wave = torch.randn(7000, generator=g)
    wl = 512 if win_length is None else win_length
    w = getattr(torch, f"{window}_window")(wl) if window is not None else None
    spec = torch.stft(wave, 512, hop_length=hop, win_length=wl, window=w, center=True, pad_mode="reflect", normalized=False, onesided=True,
                      return_complex=True)
    ref = (spec.real ** 2 + spec.imag ** 2).t()
    got = OF.stft_power(wave, hop=hop, win_length=win_length, window=window)
    assert got.shape == ref.shape == (1 + 7000 // hop, 257)
    assert float((got - ref).abs().max()) <= 2e-3 * float(ref.abs().max())

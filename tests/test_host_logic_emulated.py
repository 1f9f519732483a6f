"""Host logic of the encoder classes on a CPU-only box: the C-ABI entry points are replaced by their torch restatements
(tests/emu_backend.py), everything else -- weight packing, GEMM descriptors (strides, offsets, batch dims, implicit-GEMM conv2 layout),
buffer pitches, kernel order -- is the product code, checked against the reference fixtures."""
import numpy as np
import pytest
import torch

import emu_backend
from golden_util import load


def _conformer(cfg, w):
    import espnet_b200

    enc = espnet_b200.ConformerEncoder(80, output_size=cfg["d_model"], attention_heads=cfg["heads"], linear_units=cfg["ff"],
                                       num_blocks=cfg["enc_layers"], input_layer="conv2d", normalize_before=True, macaron_style=True,
                                       rel_pos_type="latest", pos_enc_layer_type="rel_pos", selfattention_layer_type="rel_selfattn",
                                       activation_type="swish", use_cnn_module=True, cnn_module_kernel=cfg.get("kernel", 31))
    enc.load_state_dict({k[len("encoder."):]: v for k, v in w.items() if k.startswith("encoder.")}, strict=True)
    return enc.eval()


@pytest.mark.parametrize("case", ["tiny", "small"])
def test_conformer_encoder_host_logic_vs_reference_fixture(case, monkeypatch):
    emu_backend.install(monkeypatch)
    z, cfg, w = load(case)
    enc = _conformer(cfg, w)
    enc.trace = []
    feats = torch.from_numpy(z["feats_norm"])[None]
    out, olens, _ = enc(feats, torch.tensor([feats.shape[1]]))
    assert int(olens[0]) == z["enc"].shape[0]
    for i, t in enumerate(enc.trace):
        np.testing.assert_allclose(t[0].numpy(), z[f"layer{i}"], atol=5e-5, rtol=1e-5)
    np.testing.assert_allclose(out[0].numpy(), z["enc"], atol=5e-5, rtol=1e-5)
    assert emu_backend.calls.count("espb_relpos_softmax_f32") == cfg["enc_layers"]


def test_conformer_encoder_ragged_batch_host_logic(monkeypatch):
    """Per-utterance semantics of a ragged batch: own conv boundaries, own attention keys."""
    from oracle import encoder as OE

    emu_backend.install(monkeypatch)
    z, cfg, w = load("tiny")
    enc = _conformer(cfg, w)
    g = torch.Generator().manual_seed(1)
    lens = [90, 61, 75]
    feats = torch.zeros(3, 90, 80)
    for i, n in enumerate(lens):
        feats[i, :n] = torch.randn(n, 80, generator=g)
    out, olens, _ = enc(feats, torch.tensor(lens))
    for i, n in enumerate(lens):
        ref = OE.conformer_encode(feats[i, :n], w, cfg["heads"], cfg["enc_layers"])
        assert int(olens[i]) == ref.shape[0]
        np.testing.assert_allclose(out[i, : ref.shape[0]].numpy(), ref.numpy(), atol=5e-5, rtol=1e-5)


def test_transformer_encoder_host_logic_vs_reference_fixture(monkeypatch):
    """Next scope row: the new class's orchestration (abs-pos table as a broadcast residual of the embed GEMM, q / k
    read in place from the fused qkv buffer, masked softmax) against the reference's layer outputs."""
    import espnet_b200

    emu_backend.install(monkeypatch)
    z, cfg, w = load("tfm")
    enc = espnet_b200.TransformerEncoder(80, output_size=cfg["d_model"], attention_heads=cfg["heads"], linear_units=cfg["ff"],
                                         num_blocks=cfg["enc_layers"])
    enc.load_state_dict({k[len("encoder."):]: v for k, v in w.items() if k.startswith("encoder.")}, strict=True)
    enc.eval()
    enc.trace = []
    feats = torch.from_numpy(z["feats_norm"])[None]
    out, olens, _ = enc(feats, torch.tensor([feats.shape[1]]))
    for i, t in enumerate(enc.trace):
        np.testing.assert_allclose(t[0].numpy(), z[f"layer{i}"], atol=5e-5, rtol=1e-5)
    np.testing.assert_allclose(out[0].numpy(), z["enc"], atol=5e-5, rtol=1e-5)
    assert emu_backend.calls.count("espb_masked_softmax_f32") == cfg["enc_layers"]


def test_transformer_encoder_ragged_batch_host_logic(monkeypatch):
    import espnet_b200
    from oracle import transformer_encoder as TE

    emu_backend.install(monkeypatch)
    z, cfg, w = load("tfm")
    enc = espnet_b200.TransformerEncoder(80, output_size=cfg["d_model"], attention_heads=cfg["heads"], linear_units=cfg["ff"],
                                         num_blocks=cfg["enc_layers"])
    enc.load_state_dict({k[len("encoder."):]: v for k, v in w.items() if k.startswith("encoder.")}, strict=True)
    g = torch.Generator().manual_seed(2)
    lens = [83, 50, 64]
    feats = torch.zeros(3, 83, 80)
    for i, n in enumerate(lens):
        feats[i, :n] = torch.randn(n, 80, generator=g)
    out, olens, _ = enc.eval()(feats, torch.tensor(lens))
    for i, n in enumerate(lens):
        ref = TE.transformer_encode(feats[i, :n], w, cfg["heads"], cfg["enc_layers"])
        assert int(olens[i]) == ref.shape[0]
        np.testing.assert_allclose(out[i, : ref.shape[0]].numpy(), ref.numpy(), atol=5e-5, rtol=1e-5)


def _search_setup(case, dn):
    """The scorer / weight wiring of Speech2Text.__init__ (asr_inference.py:310-381) around a CPU-built model."""
    import argparse

    import espnet_b200
    from espnet_b200.search import BatchBeamSearch
    from gpu_util import refbuild   # model_yaml only (does not import the reference)
    from golden_util import decode_params

    z, cfg, w = load(case)
    model = espnet_b200.build_model(argparse.Namespace(**refbuild.model_yaml(cfg)))
    model.load_state_dict(w, strict=True)
    model.eval()
    kw = decode_params(z, dn)
    cw = kw["ctc_weight"]
    scorers = dict(decoder=model.decoder if cw != 1.0 else None, ctc=model.ctc)
    weights = dict(decoder=1.0 - cw, ctc=cw, lm=1.0, ngram=0.9, length_bonus=kw["penalty"])
    bs = BatchBeamSearch(scorers, weights, kw["beam_size"], len(model.token_list), model.sos, model.eos, token_list=model.token_list,
                         pre_beam_score_key=None if cw == 1.0 else "full", normalize_length=kw["normalize_length"])
    return z, model, bs, kw


@pytest.mark.parametrize("case,dn", [("tiny", "joint"), ("tiny", "att"), ("tiny", "ctc"), ("tiny", "joint_pen"), ("tiny", "joint_auto"),
                                     ("tfm", "att"), ("tfm", "joint")])
def test_search_and_decoder_host_logic_vs_reference_fixture(case, dn, monkeypatch):
    """Encoder -> CTC head -> decoder memory / incremental decoder -> device-resident beam search, all host code of the product with the
    kernels emulated on CPU: n-best lists of the reference fixtures (attention-only, joint, CTC-only, penalty / minlen /
    normalize_length, end detection)."""
    from golden_util import decode_results

    emu_backend.install_search(monkeypatch)
    z, model, bs, kw = _search_setup(case, dn)
    feats = torch.from_numpy(z["feats_norm"])[None]
    enc, enc_lens, _ = model.encoder(feats, torch.tensor([feats.shape[1]]))
    hyps = bs.forward_batch(enc, enc_lens, model.enc_split(enc), kw["maxlenratio"], kw["minlenratio"])[0]
    gold = decode_results(z, dn)
    got = hyps[:10]
    assert len(got) == len(gold)
    for h, (yseq, score, scores) in zip(got, gold):
        assert h.yseq.tolist() == yseq
        assert abs(h.score - score) <= 2e-4 * max(1.0, abs(score))
    assert "espb_beam_select" in emu_backend.calls
    assert ("espb_anc_update_i32" in emu_backend.calls) == (dn != "ctc")          # the ancestor table belongs to the decoder cache
    assert ("espb_transpose_tv_f32" in emu_backend.calls) == (dn not in ("att", "ctc"))   # token-major posteriors: joint mode only


def test_ctc_greedy_host_logic_vs_reference_fixture(monkeypatch):
    emu_backend.install_search(monkeypatch)
    z, model, bs, kw = _search_setup("small", "joint")
    feats = torch.from_numpy(z["feats_norm"])[None]
    enc, enc_lens, _ = model.encoder(feats, torch.tensor([feats.shape[1]]))
    ids, cnt, am = model.ctc.greedy(enc, enc_lens, model.enc_split(enc), blank=model.blank_id)
    assert am[0, : int(enc_lens[0])].tolist() == z["ctc_argmax"].tolist()
    assert ids[0, : int(cnt[0])].tolist() == z["ctc_greedy"].tolist()


def test_ragged_batch_and_utterance_groups_host_logic(monkeypatch):
    """A ragged batch searched (a) as one group and (b) as two groups on separate (here: stand-in) streams gives, per utterance, the
    oracle's batch-1 n-best list."""
    import oracle

    emu_backend.install_search(monkeypatch)
    z, model, bs, kw = _search_setup("tiny", "joint")
    cfgz, cfg, w = load("tiny")
    g = torch.Generator().manual_seed(4)
    lens = [90, 61, 75]
    feats = torch.zeros(3, 90, 80)
    for i, n in enumerate(lens):
        feats[i, :n] = torch.randn(n, 80, generator=g)
    enc, enc_lens, _ = model.encoder(feats, torch.tensor(lens))
    single = bs.forward_batch(enc, enc_lens, model.enc_split(enc), -6.0, 0.0)
    bs.group_min_utts, bs.n_groups = 2, 2
    assert bs._group_bounds(3) == [(0, 3)] and bs._group_bounds(4) == [(0, 2), (2, 4)]
    enc4 = torch.cat([enc, enc[:1]], 0).contiguous()          # 4 utterances -> two groups of two
    lens4 = torch.cat([enc_lens, enc_lens[:1]])
    grouped = bs.forward_batch(enc4, lens4, None, -6.0, 0.0)
    o = oracle.OracleSpeech2Text(cfg, w, beam_size=kw["beam_size"], ctc_weight=kw["ctc_weight"], maxlenratio=-6.0, nbest=10)
    for i, n in enumerate(lens):
        from oracle import encoder as OE
        from oracle.search import batch_beam_search

        ref_enc = OE.conformer_encode(feats[i, :n], w, cfg["heads"], cfg["enc_layers"])
        logp = torch.log_softmax(OE.ctc_logits(ref_enc, o.w), dim=-1)
        ref = batch_beam_search(ref_enc, o.decoder, logp, beam_size=o.beam_size, ctc_weight=o.ctc_weight, vocab=o.vocab, sos=o.sos, eos=o.eos,
                                maxlenratio=-6.0, minlenratio=0.0, penalty=0.0, normalize_length=False)
        for res in (single[i], grouped[i]):
            assert [h.yseq.tolist() for h in res] == [h.yseq.tolist() for h in ref]
            for a, b in zip(res, ref):
                assert abs(a.score - float(b.score)) <= 2e-4 * max(1.0, abs(float(b.score)))
    assert [h.yseq.tolist() for h in grouped[3]] == [h.yseq.tolist() for h in grouped[0]]


@pytest.mark.parametrize("case", ["tiny", "small"])
def test_frontend_host_logic_vs_reference_fixture(case, monkeypatch):
    """DefaultFrontend's host side (sparse mel tables built from the melmat buffer, window, frame counts) and UtteranceMVN (fused partial-sum
    path and standalone path) against the reference's features."""
    import espnet_b200

    emu_backend.install_frontend(monkeypatch)
    z, cfg, w = load(case)
    fe = espnet_b200.DefaultFrontend()
    fe.load_state_dict({"logmel.melmat": w["frontend.logmel.melmat"]}, strict=True)
    wave = torch.from_numpy(z["wave"])
    feats, flens = fe(wave[None], torch.tensor([wave.numel()]))
    assert int(flens[0]) == z["feats"].shape[0] == feats.shape[1]
    np.testing.assert_allclose(feats[0].numpy(), z["feats"], atol=2e-4, rtol=1e-5)
    norm, _ = espnet_b200.UtteranceMVN()(feats, flens)                       # consumes the stashed partial sums, in place
    assert norm.data_ptr() == feats.data_ptr()
    np.testing.assert_allclose(norm[0].numpy(), z["feats_norm"], atol=2e-4, rtol=1e-5)
    alone, _ = espnet_b200.UtteranceMVN()(torch.from_numpy(z["feats"])[None].clone(), flens)   # no stash: standalone kernels
    np.testing.assert_allclose(alone[0].numpy(), z["feats_norm"], atol=2e-4, rtol=1e-5)


def test_frontend_ragged_and_global_mvn_host_logic(monkeypatch, tmp_path):
    import espnet_b200
    from oracle import frontend as OF

    emu_backend.install_frontend(monkeypatch)
    fe = espnet_b200.DefaultFrontend()
    g = torch.Generator().manual_seed(9)
    lens = [9000, 4000, 6789]
    wave = torch.zeros(3, 9000)
    for i, n in enumerate(lens):
        wave[i, :n] = 0.1 * torch.randn(n, generator=g)
    feats, flens = fe(wave, torch.tensor(lens))
    for i, n in enumerate(lens):
        ref = OF.frontend_forward(wave[i, :n], fe.logmel.melmat)
        assert int(flens[i]) == ref.shape[0]
        np.testing.assert_allclose(feats[i, : ref.shape[0]].numpy(), ref.numpy(), atol=2e-4, rtol=1e-5)
        assert float(feats[i, ref.shape[0]:].abs().max() if ref.shape[0] < feats.shape[1] else 0.0) == 0.0
    zg = np.load(__import__("os").path.join(__import__("os").path.dirname(__import__("os").path.abspath(__file__)), "golden", "gmvn.npz"))
    f = tmp_path / "stats.npz"
    np.savez(f, count=zg["stats_count"], sum=zg["stats_sum"], sum_square=zg["stats_sum_square"])
    for nm in (1, 0):
        for nv in (1, 0):
            m = espnet_b200.GlobalMVN(str(f), norm_means=bool(nm), norm_vars=bool(nv))
            y, _ = m(torch.from_numpy(zg["x"]).clone(), torch.from_numpy(zg["ilens"]))
            np.testing.assert_array_equal(y.numpy(), zg[f"y_m{nm}_v{nv}"])
    with pytest.raises(RuntimeError):
        fe(torch.zeros(1, 200), torch.tensor([200]))          # shorter than n_fft/2: reflect padding impossible (torch.stft raises too)


@pytest.mark.parametrize("case,dn", [("tiny", "joint"), ("small", "joint_pen"), ("tfm", "att")])
def test_waveform_to_nbest_host_logic(case, dn, monkeypatch):
    """The whole path from the waveform: ESPnetASRModel.encode (frontend -> normalize -> encoder) and the search, every kernel emulated,
    against the reference Speech2Text's n-best list for the same waveform."""
    from golden_util import decode_results

    emu_backend.install_search(monkeypatch)
    emu_backend.install_frontend(monkeypatch)
    z, model, bs, kw = _search_setup(case, dn)
    wave = torch.from_numpy(z["wave"])
    enc, enc_lens = model.encode(wave[None], torch.tensor([wave.numel()]))
    np.testing.assert_allclose(enc[0].numpy(), z["enc"], atol=3e-4, rtol=1e-4)
    hyps = bs.forward_batch(enc, enc_lens, model.enc_split(enc), kw["maxlenratio"], kw["minlenratio"])[0][:10]
    gold = decode_results(z, dn)
    assert len(hyps) == len(gold)
    for h, (yseq, score, _) in zip(hyps, gold):
        assert h.yseq.tolist() == yseq
        assert abs(h.score - score) <= 3e-4 * max(1.0, abs(score))


def _random_model(cfg, seed):
    import argparse

    import espnet_b200
    from gpu_util import random_weights, refbuild

    w = random_weights(cfg, seed=seed)
    model = espnet_b200.build_model(argparse.Namespace(**refbuild.model_yaml(cfg)))
    model.load_state_dict(w, strict=True)
    return model.eval(), {k: v.float() for k, v in w.items()}


@pytest.mark.parametrize("cfg,lens", [
    (dict(d_model=32, heads=1, ff=48, enc_layers=1, dec_layers=1, vocab=20, kernel=3), [7]),               # the shortest legal input: T = 1
    (dict(d_model=32, heads=2, ff=48, enc_layers=2, dec_layers=1, vocab=20, kernel=7), [7, 30, 11]),       # T = 1, 6, 2 in one batch
    (dict(d_model=64, heads=4, ff=64, enc_layers=1, dec_layers=1, vocab=20, kernel=15), [140, 8, 67, 139]),
    (dict(d_model=32, heads=2, ff=32, enc_layers=1, dec_layers=1, vocab=20, kernel=31), [16, 15]),         # conv kernel wider than the sequence
])
def test_encoder_edge_shapes_host_logic(cfg, lens, monkeypatch):
    from oracle import encoder as OE

    emu_backend.install(monkeypatch)
    model, w = _random_model(cfg, seed=1)
    g = torch.Generator().manual_seed(0)
    feats = torch.zeros(len(lens), max(lens), 80)
    for i, n in enumerate(lens):
        feats[i, :n] = torch.randn(n, 80, generator=g)
    out, olens, _ = model.encoder(feats, torch.tensor(lens))
    for i, n in enumerate(lens):
        ref = OE.conformer_encode(feats[i, :n], w, cfg["heads"], cfg["enc_layers"])
        assert int(olens[i]) == ref.shape[0]
        np.testing.assert_allclose(out[i, : ref.shape[0]].numpy(), ref.numpy(), atol=1e-4, rtol=1e-4)


@pytest.mark.parametrize("lens", [[7], [11], [7, 23, 12]])
@pytest.mark.parametrize("beam,cw,mlr,minr,pen,nl", [(2, 0.3, 0.0, 0.0, 0.0, False), (2, 0.0, 0.0, 0.0, 0.0, False), (2, 1.0, 0.0, 0.0, 0.0, False),
                                                    (3, 0.3, 1.0, 0.5, 0.3, True)])
def test_search_on_very_short_utterances_host_logic(lens, beam, cw, mlr, minr, pen, nl, monkeypatch):
    """T = 1 .. 5 encoder frames (maxlen = T, minlen > 0, length bonus, normalised ranking): per-utterance results equal the oracle's.
    Hypotheses CTC cannot align (score <= -1e9: ties of -7e9 that absorb every other term in fp32) are excluded from the comparison."""
    from espnet_b200.search import BatchBeamSearch
    from oracle import encoder as OE
    from oracle.search import OracleDecoder, batch_beam_search

    emu_backend.install_search(monkeypatch)
    cfg = dict(d_model=32, heads=2, ff=48, enc_layers=1, dec_layers=1, vocab=12, kernel=7)
    model, w = _random_model(cfg, seed=3)
    g = torch.Generator().manual_seed(0)
    feats = torch.zeros(len(lens), max(lens), 80)
    for i, n in enumerate(lens):
        feats[i, :n] = torch.randn(n, 80, generator=g)
    bs = BatchBeamSearch(dict(decoder=model.decoder if cw != 1.0 else None, ctc=model.ctc), dict(decoder=1.0 - cw, ctc=cw, length_bonus=pen), beam,
                         cfg["vocab"], model.sos, model.eos, token_list=model.token_list, pre_beam_score_key=None if cw == 1.0 else "full",
                         normalize_length=nl)
    enc, el, _ = model.encoder(feats, torch.tensor(lens))
    res = bs.forward_batch(enc, el, model.enc_split(enc), mlr, minr)
    for i, n in enumerate(lens):
        renc = OE.conformer_encode(feats[i, :n], w, cfg["heads"], cfg["enc_layers"])
        logp = torch.log_softmax(OE.ctc_logits(renc, w), -1)
        dec = OracleDecoder(w, cfg["heads"], cfg["dec_layers"]) if cw != 1.0 else None
        ref = batch_beam_search(renc, dec, logp, beam_size=beam, ctc_weight=cw, vocab=cfg["vocab"], sos=model.sos, eos=model.eos, maxlenratio=mlr,
                                minlenratio=minr, penalty=pen, normalize_length=nl)
        got = [h for h in res[i] if h.score > -1e9]
        ref = [h for h in ref if h.score > -1e9]
        assert [h.yseq.tolist() for h in got] == [h.yseq.tolist() for h in ref]
        for a, b in zip(got, ref):
            assert abs(a.score - b.score) <= 2e-4 * max(1.0, abs(b.score))


def test_ctc_scoring_beyond_encoder_length_is_refused(monkeypatch):
    """maxlen > T + 1 with a CTC scorer: the reference dies with an IndexError inside ctc_prefix_score.py once the prefix outgrows the
    encoder output (only if a hypothesis is still alive then); here the kernels keep an all-logzero state instead of indexing out of bounds and
    the search reports the utterance with the same exception type once it has really happened."""
    from espnet_b200.search import BatchBeamSearch

    emu_backend.install_search(monkeypatch)
    cfg = dict(d_model=32, heads=2, ff=48, enc_layers=1, dec_layers=1, vocab=12, kernel=7)
    model, _ = _random_model(cfg, seed=3)
    feats = torch.randn(1, 11, 80, generator=torch.Generator().manual_seed(0))      # T = 2
    enc, el, _ = model.encoder(feats, torch.tensor([11]))
    mk = lambda cw: BatchBeamSearch(dict(decoder=model.decoder if cw != 1.0 else None, ctc=model.ctc), dict(decoder=1.0 - cw, ctc=cw), 2,  # noqa: E731
                                    cfg["vocab"], model.sos, model.eos, pre_beam_score_key=None if cw == 1.0 else "full")
    with pytest.raises(IndexError):
        mk(0.3).forward_batch(enc, el, model.enc_split(enc), -4.0, 0.0)             # maxlen 4 > T + 1 = 3
    assert len(mk(0.3).forward_batch(enc, el, model.enc_split(enc), -3.0, 0.0)) == 1   # maxlen = T + 1 is the last legal value
    assert len(mk(0.0).forward_batch(enc, el, model.enc_split(enc), -6.0, 0.0)) == 1   # attention-only: no such limit


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_randomised_search_settings_host_logic(seed, monkeypatch):
    """Seeded sweep over model shapes, ragged batches and decode settings (beam 1-5, ctc_weight 0 .. 1, maxlenratio 0 / > 0 / < 0, minlenratio
    incl. the "decode again with a smaller minlenratio" retry of beam_search.py:462-471, length bonus, normalised ranking, different
    termination-poll intervals): every utterance's n-best list equals the oracle's."""
    import random

    from espnet_b200.search import BatchBeamSearch
    from oracle import encoder as OE
    from oracle.search import OracleDecoder, batch_beam_search

    emu_backend.install_search(monkeypatch)
    rnd = random.Random(seed)
    checked = 0
    for trial in range(12):
        V, heads = rnd.choice([9, 12, 17]), rnd.choice([1, 2, 4])
        cfg = dict(d_model=32, heads=heads, ff=rnd.choice([32, 48]), enc_layers=1, dec_layers=rnd.choice([1, 2]), vocab=V, kernel=rnd.choice([3, 7, 15]))
        model, w = _random_model(cfg, seed=rnd.randrange(1000))
        lens = [rnd.randrange(7, 60) for _ in range(rnd.choice([1, 2, 3]))]
        g = torch.Generator().manual_seed(100 * seed + trial)
        feats = torch.zeros(len(lens), max(lens), 80)
        for i, n in enumerate(lens):
            feats[i, :n] = torch.randn(n, 80, generator=g)
        Ts = [((x - 1) // 2 - 1) // 2 for x in lens]
        beam, cw = rnd.choice([1, 2, 3, 4, 5]), rnd.choice([0.0, 0.2, 0.5, 0.9, 1.0])
        if cw not in (0.0, 1.0) and int(1.5 * beam) >= V:
            beam = 2
        mlr, minr = rnd.choice([0.0, 0.0, 0.5, 1.0, -2.0, -5.0]), rnd.choice([0.0, 0.0, 0.3, -1.0, -2.0])
        pen, nl = rnd.choice([0.0, 0.0, 0.4, 1.5]), rnd.choice([False, True])
        if cw != 0.0 and mlr < 0 and -mlr > min(Ts) + 1:
            mlr = 0.0
        bs = BatchBeamSearch(dict(decoder=model.decoder if cw != 1.0 else None, ctc=model.ctc), dict(decoder=1.0 - cw, ctc=cw, length_bonus=pen), beam,
                             V, model.sos, model.eos, token_list=model.token_list, pre_beam_score_key=None if cw == 1.0 else "full", normalize_length=nl)
        enc, el, _ = model.encoder(feats, torch.tensor(lens))
        res = bs.forward_batch(enc, el, model.enc_split(enc), mlr, minr, check_every=rnd.choice([1, 3, 8]))
        for i, n in enumerate(lens):
            renc = OE.conformer_encode(feats[i, :n], w, heads, 1)
            logp = torch.log_softmax(OE.ctc_logits(renc, w), -1)
            dec = OracleDecoder(w, heads, cfg["dec_layers"]) if cw != 1.0 else None
            ref = batch_beam_search(renc, dec, logp, beam_size=beam, ctc_weight=cw, vocab=V, sos=model.sos, eos=model.eos, maxlenratio=mlr,
                                    minlenratio=minr, penalty=pen, normalize_length=nl)
            got, ref = [h for h in res[i] if h.score > -1e9], [h for h in ref if h.score > -1e9]
            assert [h.yseq.tolist() for h in got] == [h.yseq.tolist() for h in ref], (trial, cfg, lens, beam, cw, mlr, minr, pen, nl)
            for a, b in zip(got, ref):
                assert abs(a.score - b.score) <= 3e-4 * max(1.0, abs(b.score))
            checked += 1
    assert checked >= 12


def test_minlenratio_retry_is_per_utterance(monkeypatch):
    """maxlen 2 but minlen = int(0.3 T) > 2: nothing may end -> the reference decodes again with minlenratio lowered by 0.1 until something ends
    (beam_search.py:462-471); in a batch only the utterances without a result are searched again."""
    from espnet_b200.search import BatchBeamSearch

    emu_backend.install_search(monkeypatch)
    cfg = dict(d_model=32, heads=2, ff=48, enc_layers=1, dec_layers=1, vocab=17, kernel=7)
    model, w = _random_model(cfg, seed=5)
    g = torch.Generator().manual_seed(0)
    lens = [59, 19]                 # T = 14 (minlen 4 > maxlen 2 -> retries) and T = 4 (minlen 1 -> ends in the first pass)
    feats = torch.zeros(2, 59, 80)
    for i, n in enumerate(lens):
        feats[i, :n] = torch.randn(n, 80, generator=g)
    enc, el, _ = model.encoder(feats, torch.tensor(lens))
    bs = BatchBeamSearch(dict(decoder=model.decoder, ctc=None), dict(decoder=1.0, ctc=0.0), 2, cfg["vocab"], model.sos, model.eos)
    first = bs._search_once(enc, el, model.enc_split(enc), -2.0, 0.3, 8)
    assert first[0] == [] and len(first[1]) > 0
    full = bs.forward_batch(enc, el, model.enc_split(enc), -2.0, 0.3)
    assert len(full[0]) > 0 and [h.yseq.tolist() for h in full[1]] == [h.yseq.tolist() for h in first[1]]
    assert all(len(h.yseq) == 4 for h in full[0])          # sos + 2 tokens + the eos appended at maxlen


@pytest.mark.parametrize("kind", ["conformer", "transformer"])
def test_fused_attention_call_host_logic(kind, monkeypatch):
    """d_k = 64 takes the fused attention entry point (espb_flash_attn_f32: pointer / offset / stride arguments of q, k, V^T, bd and the
    context buffer); ragged batch against the oracle, and the same numbers as the materialised sequence."""
    import espnet_b200
    from espnet_b200 import ops
    from oracle import encoder as OE
    from oracle import transformer_encoder as TE

    emu_backend.install(monkeypatch)
    torch.manual_seed(4)
    if kind == "conformer":
        enc = espnet_b200.ConformerEncoder(80, output_size=128, attention_heads=2, linear_units=64, num_blocks=2, input_layer="conv2d",
                                           macaron_style=True, rel_pos_type="latest", pos_enc_layer_type="rel_pos",
                                           selfattention_layer_type="rel_selfattn", cnn_module_kernel=7).eval()
    else:
        enc = espnet_b200.TransformerEncoder(80, output_size=128, attention_heads=2, linear_units=64, num_blocks=2).eval()
    w = {"encoder." + k: v.detach().clone() for k, v in enc.state_dict().items()}
    g = torch.Generator().manual_seed(2)
    lens = [70, 41]
    feats = torch.zeros(2, 70, 80)
    for i, n in enumerate(lens):
        feats[i, :n] = torch.randn(n, 80, generator=g)
    monkeypatch.setattr(ops, "_ATTN_MODE", "fused")
    out, olens, _ = enc(feats, torch.tensor(lens))
    assert emu_backend.calls.count("espb_flash_attn_f32") == 2 and "espb_relpos_softmax_f32" not in emu_backend.calls
    monkeypatch.setattr(ops, "_ATTN_MODE", "materialized")
    out2, _, _ = enc(feats, torch.tensor(lens))
    for i, n in enumerate(lens):
        ref = (OE.conformer_encode if kind == "conformer" else TE.transformer_encode)(feats[i, :n], w, 2, 2)
        T = ref.shape[0]
        assert int(olens[i]) == T
        np.testing.assert_allclose(out[i, :T].numpy(), ref.numpy(), atol=5e-5, rtol=1e-5)
        np.testing.assert_allclose(out[i, :T].numpy(), out2[i, :T].numpy(), atol=2e-5, rtol=1e-5)

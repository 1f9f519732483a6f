"""The reference's scorer protocol (espnet2/legacy/nets/scorer_interface.py:85-188) on the espnet_b200 classes, driven by the REFERENCE's own
BatchBeamSearch (unmodified files under oracle/_ref): TransformerDecoder.batch_score / select_state and CTCPrefixScorer.batch_init_state /
batch_score_partial / select_state must reproduce the n-best lists the reference produced with its own scorers (tests/golden/*.npz).
CPU: C-ABI entry points emulated (tests/emu_backend.py) -> host logic of the protocol; GPU (-m gpu): the CUDA kernels, plus the registry path
(espnet_b200.integration.register + the reference's own Speech2Text on device="cuda")."""
import os
import sys

import numpy as np
import pytest
import torch

from golden_util import decode_params, decode_results, load

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _reference():
    """Import the reference (oracle/_ref, or a fresh copy when /root/reference is mounted); skip if neither exists."""
    sys.path.insert(0, ROOT)
    from oracle import install_ref

    if not install_ref.available():
        try:
            install_ref.install(verbose=False)
        except Exception:
            pass
    if not install_ref.available():
        pytest.skip("oracle/_ref is absent (python oracle/install_ref.py needs /root/reference)")
    install_ref.activate()


def _ref_search(model, z, dn, ctc_scorer, decoder):
    """espnet2.bin.asr_inference.Speech2Text.__init__'s BatchBeamSearch wiring (asr_inference.py:168-176, 310-381) with our scorers."""
    from espnet2.legacy.nets.batch_beam_search import BatchBeamSearch
    from espnet2.legacy.nets.scorers.length_bonus import LengthBonus

    kw = decode_params(z, dn)
    cw = kw["ctc_weight"]
    V = model.vocab_size
    scorers = dict(decoder=decoder, ctc=ctc_scorer, length_bonus=LengthBonus(V))
    weights = dict(decoder=1.0 - cw, ctc=cw, lm=1.0, ngram=0.9, length_bonus=kw.get("penalty", 0.0))
    bs = BatchBeamSearch(beam_size=kw["beam_size"], weights=weights, scorers=scorers, sos=model.sos, eos=model.eos, vocab_size=V,
                         token_list=model.token_list, pre_beam_score_key=None if cw == 1.0 else "full",
                         normalize_length=kw.get("normalize_length", False))
    return bs, kw


def _check(hyps, gold):
    assert len(hyps) >= len(gold)
    for h, (yseq, score, _) in zip(hyps, gold):
        assert h.yseq.tolist() == yseq
        assert abs(float(h.score) - score) <= 2e-4 * max(1.0, abs(score))


def _drive(model, z, dn, device):
    from espnet_b200 import integration

    enc = torch.from_numpy(z["enc"]).to(device)
    dec = integration.register()["decoder"]["b200_transformer"]
    decoder = model.decoder
    decoder.__class__ = dec      # same weights, now an instance of AbsDecoder + BatchScorerInterface (a subclass of its former class)
    bs, kw = _ref_search(model, z, dn, integration.ctc_prefix_scorer(model.ctc, model.eos), decoder)
    hyps = bs(x=enc, maxlenratio=kw.get("maxlenratio", 0.0), minlenratio=kw.get("minlenratio", 0.0))
    _check(hyps[:10], decode_results(z, dn))


@pytest.mark.parametrize("dn", ["joint", "att", "ctc", "joint_pen"])
def test_reference_beam_search_drives_our_scorers_host_logic(dn, monkeypatch):
    import argparse

    import emu_backend
    import espnet_b200
    from gpu_util import refbuild

    _reference()
    emu_backend.install_search(monkeypatch)
    z, cfg, w = load("tiny")
    model = espnet_b200.build_model(argparse.Namespace(**refbuild.model_yaml(cfg)))
    model.load_state_dict(w, strict=True)
    _drive(model.eval(), z, dn, "cpu")


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["tiny", "small"])
@pytest.mark.parametrize("dn", ["joint", "att", "ctc", "joint_pen"])
def test_reference_beam_search_drives_our_scorers_cuda(case, dn):
    from gpu_util import build_cuda_model

    _reference()
    z, cfg, w = load(case)
    model, _ = build_cuda_model(cfg, w)
    _drive(model, z, dn, "cuda")


@pytest.mark.gpu
@pytest.mark.parametrize("dn", ["joint", "att"])
def test_reference_speech2text_with_registered_b200_classes(dn, tmp_path):
    """config.yaml names the b200_ classes -> the reference's ASRTask.build_model builds them, the reference's Speech2Text (device cuda)
    encodes through them and its BatchBeamSearch scores through TransformerDecoder.batch_score; result == the reference-only fixture."""
    import yaml

    from gpu_util import refbuild

    _reference()
    from espnet_b200 import integration

    integration.register()
    from espnet2.bin.asr_inference import Speech2Text

    z, cfg, w = load("tiny")
    y = refbuild.model_yaml(cfg)
    y.update(frontend="b200_default", normalize="b200_utterance_mvn", encoder="b200_conformer", decoder="b200_transformer")
    y["encoder_conf"].pop("use_flash_attn", None); y["decoder_conf"].pop("use_flash_attn", None)
    path = str(tmp_path / "config.yaml")
    with open(path, "w") as f:
        yaml.safe_dump(y, f)
    kw = decode_params(z, dn)
    s2t = Speech2Text(asr_train_config=path, asr_model_file=None, device="cuda", dtype="float32", nbest=10, **kw)
    missing = s2t.asr_model.load_state_dict(w, strict=False)
    assert not [k for k in missing.unexpected_keys], missing
    s2t.asr_model.eval()
    res = s2t(z["wave"])
    gold = decode_results(z, dn)
    assert len(res) == len(gold)
    for (_, _, _, h), (yseq, score, _) in zip(res, gold):
        assert h.yseq.tolist() == yseq
        assert abs(float(h.score) - score) <= 2e-4 * max(1.0, abs(score))

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

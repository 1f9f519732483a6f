"""Streaming row: espnet_b200.ContextualBlockConformerEncoder.forward_infer against outputs of the UNMODIFIED reference
(tests/golden/streaming_enc.npz, made by tests/golden/make_golden_streaming.py): a stream pushed in uneven chunks (every push), one whole-utterance
call, a short segment; and N streams in lock step == the single stream.  CPU: host logic with the kernels emulated; -m gpu: the CUDA kernels.
Tolerance: atol 1e-4 on O(1) encoder outputs."""
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "streaming_enc.npz")


def _build(device):
    import espnet_b200

    z = np.load(GOLD)
    cfg = {k: int(v) for k, v in zip(z["cfg_keys"].tolist(), z["cfg_vals"].tolist())}
    enc = espnet_b200.ContextualBlockConformerEncoder(80, input_layer="conv2d", macaron_style=True, use_cnn_module=True, **cfg)
    enc.load_state_dict({k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w:")}, strict=True)
    return z, enc.to(device).eval()


def _run(z, enc, device, n_streams=1):
    feats = torch.from_numpy(z["feats"]).to(device)[None].repeat(n_streams, 1, 1)
    pushes = z["pushes"].tolist()
    states, pos = None, 0
    for i, n in enumerate(pushes):
        y, ylen, states = enc(feats[:, pos:pos + n], torch.full((n_streams,), n), states, is_final=(i == len(pushes) - 1), infer_mode=True)
        ref = z[f"stream:{i}:y"]
        assert tuple(y.shape) == (n_streams,) + ref.shape, (i, y.shape, ref.shape)
        for s in range(n_streams):
            if ref.size:
                e = float((y[s].cpu() - torch.from_numpy(ref)).abs().max())
                assert e < 1e-4, (i, s, e)
        pos += n
    assert states is None
    y, _, _ = enc(feats[:1, :500], torch.tensor([500]), None, is_final=True, infer_mode=True)
    assert float((y[0].cpu() - torch.from_numpy(z["whole:y"])).abs().max()) < 1e-4
    y, _, _ = enc(feats[:1, :131], torch.tensor([131]), None, is_final=True, infer_mode=True)
    assert float((y[0].cpu() - torch.from_numpy(z["short:y"])).abs().max()) < 1e-4


def test_streaming_encoder_host_logic_vs_reference_fixture(monkeypatch):
    import emu_backend

    emu_backend.install(monkeypatch)
    z, enc = _build("cpu")
    _run(z, enc, "cpu")
    assert "espb_cbe_ctx_propagate_f32" in emu_backend.calls and "espb_gather_rows_f32" in emu_backend.calls


def test_streaming_encoder_state_dict_names_match_reference():
    z, enc = _build("cpu")
    assert sorted(enc.state_dict()) == sorted(k[2:] for k in z.files if k.startswith("w:"))


@pytest.mark.gpu
@pytest.mark.parametrize("n_streams", [1, 5])
def test_streaming_encoder_cuda_vs_reference_fixture(n_streams):
    z, enc = _build("cuda")
    _run(z, enc, "cuda", n_streams)


def _frontend_model(device, n_streams):
    import types

    import espnet_b200

    fe = espnet_b200.DefaultFrontend(fs=16000, n_fft=512, hop_length=128, n_mels=80).to(device)
    model = types.SimpleNamespace(frontend=fe, normalize=None, blank_id=0, encoder=None, ctc=None)
    model.to = lambda d: model
    model.eval = lambda: model
    return espnet_b200.Speech2TextStreaming(model, n_streams=n_streams, device=device, greedy=True)


def _run_frontend(s2t, z, device, n_streams):
    wave = torch.from_numpy(z["wave"]).to(device)[None].repeat(n_streams, 1)
    st, pos = None, 0
    pushes = z["wave_pushes"].tolist()
    for i, n in enumerate(pushes):
        feats, st = s2t.apply_frontend(wave[:, pos:pos + n], st, is_final=(i == len(pushes) - 1))
        ref = z[f"fe:{i}:feats"]
        got = torch.zeros(n_streams, 0, 80) if feats is None else feats.cpu()
        assert tuple(got.shape) == (n_streams,) + ref.shape, (i, got.shape, ref.shape)
        if ref.size:
            assert float((got - torch.from_numpy(ref)).abs().max()) < 5e-4, i
        pos += n
    assert st is None


def test_streaming_frontend_chunking_host_logic_vs_reference_fixture(monkeypatch):
    """Speech2TextStreaming.apply_frontend == the reference's (asr_inference_streaming.py:205-294) push by push."""
    import emu_backend

    emu_backend.install(monkeypatch)
    emu_backend.install_frontend(monkeypatch)
    z = np.load(GOLD)
    _run_frontend(_frontend_model("cpu", 1), z, "cpu", 1)


@pytest.mark.gpu
def test_streaming_frontend_chunking_cuda_vs_reference_fixture():
    z = np.load(GOLD)
    _run_frontend(_frontend_model("cuda", 3), z, "cuda", 3)


@pytest.mark.gpu
def test_streaming_speech2text_ctc_greedy_equals_whole_utterance_collapse():
    """40-ms pushes through frontend -> block encoder -> CTC head: the tokens emitted push by push equal argmax/collapse of the concatenated
    encoder frames (and all streams of the lock-step batch agree)."""
    import argparse

    import espnet_b200
    from gpu_util import refbuild

    cfg = dict(d_model=64, heads=4, ff=96, enc_layers=2, dec_layers=1, vocab=40, kernel=15)
    y = refbuild.model_yaml(cfg)
    y.update(encoder="contextual_block_conformer", normalize=None, normalize_conf={},
             encoder_conf=dict(output_size=64, attention_heads=4, linear_units=96, num_blocks=2, macaron_style=True, cnn_module_kernel=15,
                               block_size=40, hop_size=16, look_ahead=16))
    torch.manual_seed(3)
    model = espnet_b200.build_model(argparse.Namespace(**y)).cuda().eval()
    n_streams, wave = 4, refbuild.waveform(5, 48000)
    s2t = espnet_b200.Speech2TextStreaming(model, n_streams=n_streams, device="cuda", greedy=True)
    frames = []
    orig = model.encoder.forward

    def spy(*a, **k):
        out = orig(*a, **k)
        frames.append(out[0][0].clone())
        return out
    model.encoder.forward = spy
    emitted = [[] for _ in range(n_streams)]
    for p in range(0, 48000, 640):
        new = s2t(wave[None, p:p + 640].repeat(n_streams, 1), is_final=(p + 640 >= 48000))
        for s in range(n_streams):
            emitted[s] += new[s]
    enc = torch.cat(frames, 0)
    am = model.ctc.argmax(enc[None])[0].cpu().tolist()
    ref = [t for i, t in enumerate(am) if t != 0 and (i == 0 or am[i - 1] != t)]
    assert emitted[0] == ref and len(ref) > 0
    assert all(e == emitted[0] for e in emitted)
    assert s2t.final_tokens[0] == ref

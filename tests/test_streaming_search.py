"""Streaming row: espnet_b200.Speech2TextStreaming with the block-synchronous beam search (search_online.BatchBeamSearchOnline over
TransformerDecoder.batch_score + CTCPrefixScorer.batch_score_partial / extend_prob / extend_state + LengthBonus) against the n-best the UNMODIFIED
reference Speech2TextStreaming returned for EVERY push of a waveform (tests/golden/streaming_search.npz, made by
tests/golden/make_golden_streaming_search.py): joint CTC/attention, joint with length bonus and repetition detection off, CTC only, attention-heavy
with a maxlenratio, joint + TransformerLM shallow fusion.  CPU: host logic with the kernels emulated; -m gpu: the CUDA kernels.
Tolerance: identical token sequences; total and per-scorer scores rtol 2e-4 + atol 2e-3 (scores are O(10..100) sums of log-probabilities)."""
import argparse
import json
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "streaming_search.npz")


def _model(z, device):
    import espnet_b200

    y = json.loads(str(z["yaml"]))
    model = espnet_b200.build_model(argparse.Namespace(**y))
    model.load_state_dict({k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w:")}, strict=True)
    return model.to(device).eval(), argparse.Namespace(**y)


def _run(device, names=None, n_streams=1):
    import espnet_b200

    z = np.load(GOLD)
    model, _ = _model(z, device)
    settings = json.loads(str(z["settings"]))
    wave, pushes = torch.from_numpy(z["wave"]), z["pushes"].tolist()
    for name, kw in settings.items():
        if names is not None and name not in names:
            continue
        extra = {}
        if "lm_weight" in kw:      # TransformerLM shallow fusion: the fixture carries the reference LM's weights (lm:*) and its lm_conf
            lm = espnet_b200.TransformerLM(len(json.loads(str(z["yaml"]))["token_list"]), **json.loads(str(z["lm_conf"])))
            lm.load_state_dict({k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("lm:")}, strict=True)
            extra["lm"] = lm.to(device).eval()
        s2t = espnet_b200.Speech2TextStreaming(model, n_streams=n_streams, device=device, **kw, **extra)
        pos = 0
        for i, n in enumerate(pushes):
            chunk = wave[pos:pos + n]
            res = s2t(chunk if n_streams == 1 else chunk[None].repeat(n_streams, 1), is_final=(i == len(pushes) - 1))
            pos += n
            for per_stream in ([res] if n_streams == 1 else res):
                assert len(per_stream) == int(z[f"{name}:{i}:n"]), (name, i, len(per_stream))
                for j, (_, _, token_int, hyp) in enumerate(per_stream):
                    key = f"{name}:{i}:{j}"
                    assert hyp.yseq.tolist() == z[key + ":yseq"].tolist(), (key, hyp.yseq.tolist(), z[key + ":yseq"].tolist())
                    assert token_int == z[key + ":token_int"].tolist(), key
                    ref = float(z[key + ":score"])
                    assert abs(float(hyp.score) - ref) <= 2e-4 * abs(ref) + 2e-3, (key, float(hyp.score), ref)
                    for k, v in hyp.scores.items():
                        r = float(z[f"{key}:score:{k}"])
                        assert abs(float(v) - r) <= 2e-4 * abs(r) + 2e-3, (key, k, float(v), r)


@pytest.mark.parametrize("name", ["joint", "joint_pen_norep", "ctc_only", "att_heavy_maxlen", "joint_lm", "joint_normlen_minlen"])
def test_streaming_beam_search_host_logic_vs_reference_fixture(monkeypatch, name):
    import emu_backend

    emu_backend.install_search(monkeypatch)
    emu_backend.install_frontend(monkeypatch)
    _run("cpu", [name])
    assert "espb_ctc_extend_state_f32" in emu_backend.calls


def test_streaming_speech2text_from_config_and_checkpoint_files(monkeypatch, tmp_path):
    """The reference's constructor form (asr_inference_streaming.py:46-75): asr_train_config + asr_model_file; block sizes come from encoder_conf."""
    import yaml

    import emu_backend
    import espnet_b200

    emu_backend.install_search(monkeypatch)
    emu_backend.install_frontend(monkeypatch)
    z = np.load(GOLD)
    cfg_path, pth = tmp_path / "config.yaml", tmp_path / "model.pth"
    cfg_path.write_text(yaml.safe_dump(json.loads(str(z["yaml"]))))
    torch.save({k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w:")}, pth)
    kw = json.loads(str(z["settings"]))["joint"]
    s2t = espnet_b200.Speech2TextStreaming(asr_train_config=str(cfg_path), asr_model_file=str(pth), device="cpu", **kw)
    assert (s2t.searches[0].block_size, s2t.searches[0].hop_size, s2t.searches[0].look_ahead) == (40, 16, 16)
    wave, pushes = torch.from_numpy(z["wave"]), z["pushes"].tolist()
    pos = 0
    for i, n in enumerate(pushes):
        res = s2t(wave[pos:pos + n].numpy(), is_final=(i == len(pushes) - 1))      # numpy input like the reference accepts
        pos += n
        assert len(res) == int(z[f"joint:{i}:n"])
        for j, r in enumerate(res):
            assert r[2] == z[f"joint:{i}:{j}:token_int"].tolist() and r[1] == [f"t{t - 2}" for t in r[2]]
    with pytest.raises(NotImplementedError):
        espnet_b200.BatchBeamSearchOnline({}, {}, 2, 10, 9, 9, time_sync=True)


def test_streaming_cli_sim_chunk_length_writes_result_dir(monkeypatch, tmp_path):
    """bin_asr_inference_streaming.inference (asr_inference_streaming.py:360-487): scp in, simulated chunks of --sim_chunk_length samples, the final
    push's n-best out as {n}best_recog/{token,token_int,score}.  The block-synchronous search only depends on the encoder blocks, not on how the
    samples were chunked, so the result equals the reference fixture's final n-best (pushed in different, uneven chunks there)."""
    import emu_backend
    import espnet_b200
    from espnet_b200.bin_asr_inference_streaming import get_parser, inference

    emu_backend.install_search(monkeypatch)
    emu_backend.install_frontend(monkeypatch)
    z = np.load(GOLD)
    model, _ = _model(z, "cpu")
    kw = json.loads(str(z["settings"]))["joint_pen_norep"]
    s2t = espnet_b200.Speech2TextStreaming(model, device="cpu", **kw)
    np.save(tmp_path / "utt1.npy", z["wave"])
    (tmp_path / "wav.scp").write_text(f"utt1 {tmp_path / 'utt1.npy'}\n")
    out = inference(str(tmp_path / "dec"), [(str(tmp_path / "wav.scp"), "speech", "sound")], nbest=kw["nbest"], sim_chunk_length=4000, speech2text=s2t)
    last = len(z["pushes"]) - 1
    n_ref = int(z[f"joint_pen_norep:{last}:n"])
    assert len(out["utt1"]) == n_ref
    lines = dict(ln.split(maxsplit=1) for ln in (tmp_path / "dec/1best_recog/token_int").read_text().splitlines())
    assert lines["utt1"].split() == [str(t) for t in z[f"joint_pen_norep:{last}:0:token_int"].tolist()]
    for j in range(n_ref):
        assert out["utt1"][j][3].yseq.tolist() == z[f"joint_pen_norep:{last}:{j}:yseq"].tolist()
        assert abs(float(out["utt1"][j][3].score) - float(z[f"joint_pen_norep:{last}:{j}:score"])) < 2e-3 + 2e-4 * abs(float(z[f"joint_pen_norep:{last}:{j}:score"]))
    assert (tmp_path / f"dec/{n_ref}best_recog/score").exists()
    a = get_parser().parse_args(["--output_dir", "o", "--data_path_and_name_and_type", "wav.scp,speech,sound", "--asr_train_config", "c", "--asr_model_file", "m",
                                 "--sim_chunk_length", "640", "--disable_repetition_detection", "true"])
    assert a.sim_chunk_length == 640 and a.disable_repetition_detection is True and a.ctc_weight == 0.5 and a.beam_size == 20


@pytest.mark.gpu
def test_streaming_beam_search_cuda_vs_reference_fixture():
    _run("cuda")


@pytest.mark.gpu
def test_streaming_beam_search_cuda_three_streams_in_lock_step():
    _run("cuda", ["joint"], n_streams=3)

"""Build the reference's own Speech2Text (random init, fixed seed) from a small config dict.

Test infrastructure: used to generate tests/golden/*.npz and for oracle-vs-reference checks.
The yaml holds exactly the keys ASRTask.build_model reads (espnet2/tasks/asr.py:512-651).
"""
import os
import tempfile

import torch
import yaml


def token_list(vocab):
    return ["<blank>", "<unk>"] + [f"t{i}" for i in range(vocab - 3)] + ["<sos/eos>"]


def model_yaml(cfg):
    d, h = cfg["d_model"], cfg["heads"]
    y = dict(
        token_list=token_list(cfg["vocab"]),
        input_size=None,
        frontend="default",
        frontend_conf=dict(fs=16000, n_fft=512, win_length=None, hop_length=128, n_mels=80),
        specaug=None,
        normalize="utterance_mvn",
        normalize_conf={},
        preencoder=None,
        encoder="conformer",
        encoder_conf=dict(
            output_size=d, attention_heads=h, linear_units=cfg["ff"], num_blocks=cfg["enc_layers"],
            dropout_rate=0.1, positional_dropout_rate=0.1, attention_dropout_rate=0.0,
            input_layer="conv2d", normalize_before=True, macaron_style=True,
            rel_pos_type="latest", pos_enc_layer_type="rel_pos", selfattention_layer_type="rel_selfattn",
            activation_type="swish", use_cnn_module=True, cnn_module_kernel=cfg.get("kernel", 31),
            use_flash_attn=False,
        ),
        postencoder=None,
        decoder="transformer",
        decoder_conf=dict(
            attention_heads=h, linear_units=cfg["ff"], num_blocks=cfg["dec_layers"],
            dropout_rate=0.1, positional_dropout_rate=0.1, self_attention_dropout_rate=0.0,
            src_attention_dropout_rate=0.0, use_flash_attn=False,
        ),
        ctc_conf={},
        joint_net_conf=None,
        model="espnet",
        model_conf=dict(ctc_weight=cfg.get("train_ctc_weight", 0.3), lsm_weight=0.1, length_normalized_loss=False),
        init=None,
        token_type=None,
        bpemodel=None,
        use_preprocessor=False,
    )
    if cfg.get("encoder", "conformer") == "transformer":   # abs-pos TransformerEncoder (SURVEY.md 8f-1 / BASELINE configs[4])
        y["encoder"] = "transformer"
        y["encoder_conf"] = dict(output_size=d, attention_heads=h, linear_units=cfg["ff"], num_blocks=cfg["enc_layers"], dropout_rate=0.1,
                                 positional_dropout_rate=0.1, attention_dropout_rate=0.0, input_layer="conv2d", normalize_before=True,
                                 use_flash_attn=False)
    return y


def build_reference(cfg, seed=0, **s2t_kwargs):
    """Returns the reference Speech2Text instance (CPU, float32)."""
    import refshim

    refshim.install()
    from espnet2.bin.asr_inference import Speech2Text

    tmp = tempfile.mkdtemp(prefix="espref_")
    path = os.path.join(tmp, "config.yaml")
    with open(path, "w") as f:
        yaml.safe_dump(model_yaml(cfg), f)
    torch.manual_seed(seed)
    s2t = Speech2Text(asr_train_config=path, asr_model_file=None, device="cpu", dtype="float32", **s2t_kwargs)
    return s2t


def waveform(i, nsamples):
    g = torch.Generator().manual_seed(1234 + i)
    return 0.1 * torch.randn(nsamples, generator=g)

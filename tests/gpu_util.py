"""Shared helpers for the -m gpu parity tests (CUDA path vs oracle / golden fixtures)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import refbuild  # noqa: E402  (model_yaml / waveform helpers only; does not import the reference)


def build_cuda_model(cfg, weights):
    import espnet_b200

    args = argparse.Namespace(**refbuild.model_yaml(cfg))
    model = espnet_b200.build_model(args)
    missing = model.load_state_dict(weights, strict=True)
    return model.cuda().eval(), args


def speech2text(cfg, weights, **kw):
    import espnet_b200

    model, args = build_cuda_model(cfg, weights)
    return espnet_b200.Speech2Text(asr_model=model, asr_train_args=args, device="cuda", **kw)


def random_weights(cfg, seed=0):
    """Random-init weights with the reference's parameter names/shapes (our containers mirror them), PyTorch default init."""
    import espnet_b200

    torch.manual_seed(seed)
    args = argparse.Namespace(**refbuild.model_yaml(cfg))
    model = espnet_b200.build_model(args)
    with torch.no_grad():  # BatchNorm running stats away from (0,1) so the folded affine is exercised
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.running_mean.normal_(0, 0.1)
                m.running_var.uniform_(0.5, 1.5)
    return {k: v.detach().clone() for k, v in model.state_dict().items()}

"""Latency of the streaming beam-search mode (Speech2TextStreaming with BatchBeamSearchOnline) on ONE live stream.

    python scripts/streaming_search_microbench.py [seconds=20] [push_ms=500] [beam=10]

Model of the streaming bench workload (contextual-block Conformer 12L/512d/8h, block 40 / hop 16 / look-ahead 16) + 6L Transformer decoder, V = 5000,
joint CTC/attention (ctc_weight 0.3), random-init weights, synthetic waveform.  Reports wall time per push and the real-time factor."""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import espnet_b200  # noqa: E402
from gpu_util import refbuild  # noqa: E402

secs = float(sys.argv[1]) if len(sys.argv) > 1 else 20.0
push_ms = float(sys.argv[2]) if len(sys.argv) > 2 else 500.0
beam = int(sys.argv[3]) if len(sys.argv) > 3 else 10
y = refbuild.model_yaml(dict(d_model=512, heads=8, ff=2048, enc_layers=12, dec_layers=6, vocab=5000, kernel=31))
y.update(encoder="contextual_block_conformer", normalize=None, normalize_conf={},
         encoder_conf=dict(output_size=512, attention_heads=8, linear_units=2048, num_blocks=12, macaron_style=True, cnn_module_kernel=31,
                           block_size=40, hop_size=16, look_ahead=16))
torch.manual_seed(0)
model = espnet_b200.build_model(argparse.Namespace(**y)).cuda().eval()
s2t = espnet_b200.Speech2TextStreaming(model, n_streams=1, device="cuda", beam_size=beam, ctc_weight=0.3)
n, push = int(secs * 16000), int(push_ms * 16)
wave = refbuild.waveform(11, n)
for rep in range(2):      # first pass warms up (allocations, kernel attribute set-up)
    torch.cuda.synchronize()
    per_push, steps = [], 0
    t_all = time.perf_counter()
    for p in range(0, n, push):
        t0 = time.perf_counter()
        res = s2t(wave[p:p + push], is_final=(p + push >= n))
        torch.cuda.synchronize()
        per_push.append(time.perf_counter() - t0)
    wall = time.perf_counter() - t_all
per_push.sort()
print(f"[streaming_search] {secs:.0f} s of audio in {push_ms:.0f}-ms pushes, beam {beam}: wall {wall:.2f} s -> RTF {wall / secs:.3f}; per push median "
      f"{1e3 * per_push[len(per_push) // 2]:.1f} ms, p90 {1e3 * per_push[int(0.9 * len(per_push))]:.1f} ms, max {1e3 * per_push[-1]:.1f} ms; "
      f"final hypothesis {len(res[0][2]) if res else 0} tokens")

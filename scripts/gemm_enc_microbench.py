"""Encoder-shaped GEMMs of the benchmark (M = 64 x 937 frames) through ops.linear: CUDA events, L2 flushed between launches.

    python scripts/gemm_enc_microbench.py            # current dispatch
    ESPB_GEMM_EW8=1 python scripts/gemm_enc_microbench.py   # 8 epilogue warps in the 256-column CTA-pair kernel
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from espnet_b200 import ops  # noqa: E402

M = 64 * 937
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
tag = "ew8" if os.environ.get("ESPB_GEMM_EW8") else "ew16"
# (name, N, K, act, split_out, residual)
CASES = [("ffn_w1 swish split", 2048, 512, ops.ACT_SWISH, True, False), ("ffn_w2 +res", 512, 2048, ops.ACT_NONE, False, True),
         ("qkv", 1536, 512, ops.ACT_NONE, False, False), ("out +res", 512, 512, ops.ACT_NONE, False, True),
         ("pw1 (glu in)", 1024, 512, ops.ACT_NONE, False, False), ("plain split", 512, 512, ops.ACT_NONE, True, False)]
for name, N, K, act, split, res in CASES:
    a = ops.split_from(torch.randn(M, K, device="cuda"))
    w = ops.split_from(torch.randn(N, K, device="cuda") / K ** 0.5)
    bias = torch.randn(N, device="cuda")
    out = torch.zeros((2, M, N) if split else (M, N), device="cuda")
    r = torch.randn(M, N, device="cuda") if res else None
    fn = lambda: ops.linear(a, w, out, bias=bias, residual=r, act=act, split_out=split)  # noqa: E731
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(7):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    us = ts[len(ts) // 2]
    print(f"[gemm_enc {tag}] {name:20s} M{M} N{N} K{K}: median {us:8.1f} us  min {ts[0]:8.1f}  ({2.0 * M * N * K / us / 1e6:6.1f} TFLOP/s algorithmic)")
    del a, w, out, r

"""Fused attention kernel vs the materialised GEMM + softmax + GEMM sequence at the benchmarked shape (B 64, H 8, T 937, d_k 64):
CUDA-event time per call (L2 flushed between calls) and the algorithmic figures.  usage: python scripts/attn_microbench.py [B] [T] [reps]"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from espnet_b200 import ops  # noqa: E402
from espnet_b200.lib import call, ptr  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 937
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
H, dk = 8, 64
D, M = H * dk, B * T
Tp, Rp, R = (T + 31) // 32 * 32, (2 * T - 1 + 31) // 32 * 32, 2 * T - 1
g = torch.Generator().manual_seed(0)
q = torch.randn(M, D, generator=g).cuda()
kv = torch.randn(M, 3 * D, generator=g).cuda()
qu, qkv = ops.split_from(q), ops.split_from(kv)
pos = ops.split_from(torch.randn(R, D, generator=g).cuda())
lens32 = torch.full((B,), T, dtype=torch.int32, device="cuda")
vt = torch.empty(2, B, H, dk, Tp, device="cuda")
call("espb_v_transpose_f32", ptr(qkv), M * 3 * D, B, T, D, H, ptr(lens32), ptr(vt), B * H * dk * Tp, Tp)
bd = torch.zeros(B, H, T, Rp, device="cuda")
ctx = torch.empty(2, M, D, device="cuda")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")


def bd_gemm():
    ops.gemm(T, R, dk, qu, M * D, D, pos, R * D, D, bd, Rp, nbx=H, nby=B, sa=(dk, T * D), sb=(dk, 0), sc=(T * Rp, H * T * Rp), band_t=T)


def fused():
    ops.flash_attn(qu, 0, D, qkv, D, 3 * D, vt, Tp, bd, Rp, lens32, B, H, T, dk, ctx)


def fused_plain():
    ops.flash_attn(qu, 0, D, qkv, D, 3 * D, vt, Tp, None, 0, lens32, B, H, T, dk, ctx)


ac = probs = None


def materialised():
    global ac, probs
    if ac is None:
        ac = torch.empty(B, H, T, Tp, device="cuda")
        probs = torch.empty(2, B, H, T, Tp, device="cuda")
    ops.gemm(T, T, dk, qu, M * D, D, qkv, M * 3 * D, 3 * D, ac, Tp, nbx=H, nby=B, sa=(dk, T * D), sb=(dk, T * 3 * D), sc=(T * Tp, H * T * Tp), b_off=D)
    call("espb_relpos_softmax_f32", ptr(ac), ptr(bd), B, H, T, Tp, Rp, ptr(lens32), math.sqrt(dk), ptr(probs), B * H * T * Tp)
    ops.gemm(T, dk, T, probs, B * H * T * Tp, Tp, vt, B * H * dk * Tp, Tp, ctx, D, c_plane=M * D, split_out=True, nbx=H, nby=B,
             sa=(T * Tp, H * T * Tp), sb=(dk * Tp, H * dk * Tp), sc=(dk, T * D))


def timeit(fn, name, flops):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    ms = ts[len(ts) // 2]
    print(f"{name:28s} {ms * 1e3:9.1f} us   {flops / ms / 1e9:8.1f} TFLOP/s algorithmic (median of {reps})", flush=True)
    return ms


fl_qk = 2.0 * B * H * T * T * dk
only = os.environ.get("ATTN_BENCH_ONLY", "")
if only in ("", "fused"):
    timeit(bd_gemm, "bd = (q+v) p^T band GEMM", 2.0 * B * H * T * T * dk)
    timeit(fused, "fused rel-pos attention", 2 * fl_qk)
    timeit(fused_plain, "fused plain attention", 2 * fl_qk)
if only in ("", "materialised"):
    timeit(materialised, "q k^T + softmax + p v", 2 * fl_qk)

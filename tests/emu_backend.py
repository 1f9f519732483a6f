"""TEST INFRASTRUCTURE: a torch-CPU emulation of the C-ABI entry points the encoder classes call, so that the HOST logic of
espnet_b200 (GEMM descriptors: strides / offsets / batch dims, buffer pitches, layouts, the order of kernels) can be tested on a box
without a GPU.  Each function restates the contract of the CUDA kernel of the same name (espnet_b200/csrc/*.cu; the GEMM follows
gemm_simt_kernel, which is itself the validator of the tensor-core kernels).  Nothing in the product imports this file; it is
installed by monkey-patching `call` / `ptr` / `gemm` inside the espnet_b200 modules for the duration of a test.
"""
import math

import torch

MASK = -8192  # 0xFFFFE000 as int32


def tf32_hi(x):
    return (x.contiguous().view(torch.int32) & MASK).view(torch.float32)


def tf32_lo(x, hi):
    return ((x - hi).contiguous().view(torch.int32) & MASK).view(torch.float32)


def _flat(t):
    assert t.is_contiguous()
    return t.view(-1)


def _store(flat, off, v, split, plane):
    if split:
        hi = tf32_hi(v)
        flat[off] = hi
        flat[off + plane] = tf32_lo(v, hi)
    else:
        flat[off] = v


def gemm(M, N, K, A, a_plane, lda, B, b_plane, ldb, C, ldc, *, c_plane=0, split_out=False, bias=None, R=None, ldr=0, alpha=1.0, act=0,
         nbx=1, nby=1, sa=(0, 0), sb=(0, 0), sc=(0, 0), sr=(0, 0), kob=0, a_mode=0, conv=(0, 0, 0), a_off=0, b_off=0, c_off=0, r_off=0,
         sbias_x=0, bias_off=0, band_t=0, force=None):
    """EspbGemmDesc semantics (gemm.cu: load_a / load_b / gemm_simt_kernel epilogue).  band_t only lets the kernels skip tiles nobody
    reads, so computing everything is a valid superset."""
    Af, Bf, Cf = _flat(A), _flat(B), _flat(C)
    m = torch.arange(M).view(M, 1)
    n = torch.arange(N).view(N, 1)
    k = torch.arange(K).view(1, K)
    rows, cols = torch.arange(M).view(M, 1), torch.arange(N).view(1, N)
    for by in range(nby):
        for bx in range(nbx):
            if a_mode == 0:
                ko, ki = (k // (kob * 32), k % (kob * 32)) if kob > 0 else (0, k)
                off = a_off + by * sa[1] + (bx + ko) * sa[0] + m * lda + ki
                Am = Af[off] + Af[off + a_plane]
            else:  # conv2 over [b][plane*4 + pt*2 + pf][F1h][T1h][C]
                t1h, f1h, cin = conv
                tap, c = k // cin, k % cin
                kt, kf = tap // 3, tap % 3
                par = (kt & 1) * 2 + (kf & 1)
                tt, ff = m + (kt >> 1), bx + (kf >> 1)
                ok = (tt < t1h) & (ff < f1h)
                sub = f1h * t1h * cin
                off = a_off + by * 8 * sub + (torch.clamp(ff, max=f1h - 1) * t1h + torch.clamp(tt, max=t1h - 1)) * cin + c
                Am = torch.where(ok, Af[off + par * sub] + Af[off + (4 + par) * sub], torch.zeros(()))
            offb = b_off + by * sb[1] + bx * sb[0] + n * ldb + k
            Bm = Bf[offb] + Bf[offb + b_plane]
            v = (Am.double() @ Bm.double().t()).float()
            if bias is not None:
                v = v + _flat(bias)[bias_off + bx * sbias_x + cols]
            if act == 1:
                v = torch.relu(v)
            elif act == 2:
                v = v / (1.0 + torch.exp(-v))
            v = v * alpha
            if R is not None:
                v = v + _flat(R)[r_off + by * sr[1] + bx * sr[0] + rows * ldr + cols]
            _store(Cf, c_off + by * sc[1] + bx * sc[0] + rows * ldc + cols, v, split_out, c_plane)
    return True


def ptr(t):
    return t


def _split_tf32(x, n, out, plane):
    _store(_flat(out), torch.arange(n), _flat(x)[:n], True, plane)


def _layernorm(x, rows, D, gamma, beta, eps, out_plain, out_split, split_plane):
    xr = _flat(x)[: rows * D].view(rows, D)
    mean = xr.sum(-1, keepdim=True) / D
    var = ((xr - mean) ** 2).sum(-1, keepdim=True) / D
    y = ((xr - mean) * (1.0 / torch.sqrt(var + eps)) * gamma + beta).reshape(-1)   # clone: out_plain may alias x
    if out_split is not None:
        _store(_flat(out_split), torch.arange(rows * D), y, True, split_plane)
    if out_plain is not None:
        _flat(out_plain)[: rows * D] = y


def _conv1_relu(feats, B, Tf, F, w, bias, C, out, T1, F1, T1h, F1h):
    x = feats.view(B, 1, Tf, F)
    y = torch.relu(torch.nn.functional.conv2d(x, w.view(C, 1, 3, 3), bias, stride=2))   # [B][C][T1][F1]
    sub = F1h * T1h * C
    of = _flat(out)
    b, c, t1, f1 = torch.meshgrid(torch.arange(B), torch.arange(C), torch.arange(T1), torch.arange(F1), indexing="ij")
    par = (t1 & 1) * 2 + (f1 & 1)
    off = b * 8 * sub + par * sub + ((f1 >> 1) * T1h + (t1 >> 1)) * C + c
    _store(of, off, y, True, 4 * sub)


def _qu_qv(qkv, qkv_plane, M, D, pos_u, pos_v, qu, qv, out_plane):
    qf = _flat(qkv)
    off = torch.arange(M).view(M, 1) * 3 * D + torch.arange(D).view(1, D)
    q = qf[off] + qf[off + qkv_plane]
    o = torch.arange(M * D).view(M, D)
    _store(_flat(qu), o, q + pos_u.view(1, D), True, out_plane)
    _store(_flat(qv), o, q + pos_v.view(1, D), True, out_plane)


def _v_transpose(qkv, qkv_plane, B, Tmax, D, H, lens, vt, vt_plane, Tp):
    qf, dk = _flat(qkv), D // H
    b, h, d, t = torch.meshgrid(torch.arange(B), torch.arange(H), torch.arange(dk), torch.arange(Tp), indexing="ij")
    ok = t < lens.view(B, 1, 1, 1).long()
    src = (b * Tmax + torch.clamp(t, max=Tmax - 1)) * 3 * D + 2 * D + h * dk + d
    v = torch.where(ok, qf[src] + qf[src + qkv_plane], torch.zeros(()))
    _store(_flat(vt), ((b * H + h) * dk + d) * Tp + t, v, True, vt_plane)


def _softmax_rows(scores, lens_rows, Tp, probs, plane):
    """scores [R][Tp] already scaled; keys >= len -> probability 0; split store."""
    Rn = scores.shape[0]
    j = torch.arange(Tp).view(1, Tp)
    ok = j < lens_rows.view(Rn, 1)
    s = torch.where(ok, scores, torch.full((), float("-inf")))
    p = torch.where(ok, torch.softmax(s, dim=-1), torch.zeros(()))
    _store(_flat(probs), torch.arange(Rn * Tp).view(Rn, Tp), p, True, plane)


def _relpos_softmax(ac, bd, B, H, T, Tp, Rp, lens, sqrt_dk, probs, plane):
    a = _flat(ac)[: B * H * T * Tp].view(B * H * T, Tp)
    bdv = _flat(bd)[: B * H * T * Rp].view(B * H * T, Rp)
    i = (torch.arange(B * H * T) % T).view(-1, 1)
    j = torch.arange(Tp).view(1, Tp)
    col = torch.clamp(T - 1 - i + j, max=Rp - 1)                      # rel_shift: bd[i][T-1-i+j]
    s = (a + torch.gather(bdv, 1, col)) / sqrt_dk
    _softmax_rows(s, lens.long().repeat_interleave(H * T), Tp, probs, plane)


def _masked_softmax(sc, B, H, T, Tp, lens, sqrt_dk, probs, plane):
    s = _flat(sc)[: B * H * T * Tp].view(B * H * T, Tp) / sqrt_dk
    _softmax_rows(s, lens.long().repeat_interleave(H * T), Tp, probs, plane)


def _glu_dwconv_bn_swish(y, B, Tmax, C, lens, dw_w, dw_b, K, bn_a, bn_b, out, out_plane):
    yv = _flat(y)[: B * Tmax * 2 * C].view(B, Tmax, 2 * C)
    t = torch.arange(Tmax).view(1, Tmax, 1)
    valid = t < lens.view(B, 1, 1).long()
    g = torch.where(valid, yv[..., :C] * (1.0 / (1.0 + torch.exp(-yv[..., C:]))), torch.zeros(()))
    z = torch.nn.functional.conv1d(g.transpose(1, 2), dw_w.view(C, 1, K), dw_b, padding=(K - 1) // 2, groups=C).transpose(1, 2)
    z = z * bn_a + bn_b
    z = z / (1.0 + torch.exp(-z))
    z = torch.where(valid, z, torch.zeros(()))
    _store(_flat(out), torch.arange(B * Tmax * C).view(B, Tmax, C), z, True, out_plane)


_TABLE = {"espb_split_tf32_f32": _split_tf32, "espb_layernorm_f32": _layernorm, "espb_conv1_relu_f32": _conv1_relu, "espb_qu_qv_f32": _qu_qv,
          "espb_v_transpose_f32": _v_transpose, "espb_relpos_softmax_f32": _relpos_softmax, "espb_masked_softmax_f32": _masked_softmax,
          "espb_glu_dwconv_bn_swish_f32": _glu_dwconv_bn_swish}
calls = []   # names of the emulated entry points, in call order (tests can assert on the sequence)


def call(name, *args):
    if name not in _TABLE:
        raise NotImplementedError(f"emu_backend: {name} is not emulated")
    calls.append(name)
    _TABLE[name](*args)


def install(monkeypatch):
    """Route the encoder-side modules of espnet_b200 through the emulation (CPU tensors)."""
    import espnet_b200.encoder as enc
    import espnet_b200.ops as ops
    import espnet_b200.transformer_encoder as tenc

    del calls[:]
    for mod in (ops, enc, tenc):
        monkeypatch.setattr(mod, "call", call, raising=True)
        monkeypatch.setattr(mod, "ptr", ptr, raising=True)
    monkeypatch.setattr(ops, "gemm", gemm, raising=True)
    monkeypatch.setattr(enc, "gemm", gemm, raising=True)
    monkeypatch.setattr(tenc, "gemm", gemm, raising=True)
    monkeypatch.setattr(ops, "new_split", lambda *shape, device="cpu": torch.zeros((2,) + tuple(shape), dtype=torch.float32), raising=True)
    monkeypatch.setattr(enc, "new_split", ops.new_split, raising=True)

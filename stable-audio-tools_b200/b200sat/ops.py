"""Tensor-level wrappers over the C ABI.  Each wrapper validates shapes/dtypes, allocates outputs with torch (the
caching allocator owns all memory) and launches on torch's current stream."""
import torch
from ._lib import lib, check

GEMM_BIAS, GEMM_RESIDUAL, GEMM_SILU, GEMM_SWIGLU, GEMM_ROPE, GEMM_OUT_F32, GEMM_ROW_REMAP, GEMM_GATE = 1, 2, 4, 8, 16, 32, 64, 128
GEMM_A_MN, GEMM_B_MN, GEMM_ACCUM, GEMM_SWIGLU_BWD, GEMM_LN_A, GEMM_ROWSTATS = 256, 512, 1024, 2048, 8192, 16384


LAUNCHES = [0]  # kernel launches issued through this module (bench.py's gpu_launches claim)


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return 0 if t is None else t.data_ptr()


def _chk_bf16(t, name):
    if t.dtype != torch.bfloat16 or not t.is_cuda:
        raise TypeError(f"{name}: expected a CUDA bfloat16 tensor, got {t.dtype} on {t.device}")
    if t.stride(-1) != 1:
        raise ValueError(f"{name}: innermost dimension must be contiguous")


def linear(x, w, bias=None, residual=None, out=None, silu=False, out_f32=False, swiglu=False, rope=None,
           row_remap=None, gate=None, force_bn=0, save_pre=None, ln=None, out_stats=None):
    """out = epilogue(x @ w.T).  x [M,K] bf16, w [N,K] bf16 (nn.Linear layout), bias fp32 [N].

    swiglu: w is the GLU projection [2*Nh, K]; out [M, Nh] = (u[:, :Nh]) * silu(u[:, Nh:]).
    rope: (cos [S,16] fp32, sin [S,16] fp32, seq_len S, d_model, dim_heads) applied to the q,k thirds.
    row_remap: (seg_in, seg_out, seg_off) -> out_row = r // seg_in * seg_out + seg_off + r % seg_in.
    gate: fp32 [B, N] multiplied into the branch output before the residual add (adaLN), needs row_remap[0]=rows/batch.
    """
    _chk_bf16(x, "x"); _chk_bf16(w, "w")
    assert x.dim() == 2 and w.dim() == 2 and x.shape[1] == w.shape[1]
    M, K = x.shape
    flags = 0
    n_half = 0
    if swiglu:
        n_half = w.shape[0] // 2
        N = n_half
        flags |= GEMM_SWIGLU
    else:
        N = w.shape[0]
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.is_contiguous()
        flags |= GEMM_BIAS
    if silu:
        flags |= GEMM_SILU
    if out_f32:
        flags |= GEMM_OUT_F32
    seg_in = seg_out = seg_off = 0
    out_rows = M
    if row_remap is not None:
        seg_in, seg_out, seg_off = row_remap
        if seg_out:
            flags |= GEMM_ROW_REMAP
    if out is None:
        assert not (flags & GEMM_ROW_REMAP), "row_remap needs an explicit output buffer"
        out = torch.empty((out_rows, N), device=x.device, dtype=torch.float32 if out_f32 else torch.bfloat16)
    ldr = 0
    if residual is not None:
        _chk_bf16(residual, "residual")
        flags |= GEMM_RESIDUAL
        ldr = residual.stride(0)
    rc_, rs_, rseq, rdm, rdh = None, None, 0, 0, 0
    if rope is not None:
        rc_, rs_, rseq, rdm, rdh = rope
        flags |= GEMM_ROPE
    if gate is not None:
        assert gate.dtype == torch.float32 and gate.is_contiguous()
        flags |= GEMM_GATE
    ln_stats = ln_colsum = None
    ln_eps = 0.0
    if ln is not None:      # (row stats [M,2] fp32, colsum [N] fp32, eps): LayerNorm folded into this GEMM
        ln_stats, ln_colsum, ln_eps = ln
        flags |= GEMM_LN_A
    if out_stats is not None:
        flags |= GEMM_ROWSTATS
    rc = lib().b200sat_gemm_bf16(x.data_ptr(), x.stride(0), w.data_ptr(), w.stride(0), out.data_ptr(), out.stride(0),
                                 M, N, K, flags, _p(bias), _p(residual), ldr, _p(rc_), _p(rs_), rseq, rdm, rdh, n_half,
                                 seg_in, seg_out, seg_off, _p(gate), _p(save_pre), save_pre.stride(0) if save_pre is not None else 0,
                                 _p(ln_stats), _p(ln_colsum), float(ln_eps), _p(out_stats), force_bn, _stream())
    LAUNCHES[0] += 1
    check(rc, "gemm_bf16")
    return out


def gemm(a, b, out, M, N, K, a_mn=False, b_mn=False, accumulate=False, swiglu_bwd_aux=None, residual=None, force_bn=0):
    """General D[M,N] (=|+=) A . B^T for the backward pass.  a: [M,K] (or [K,M] if a_mn), b: [N,K] (or [K,N] if b_mn), bf16.
    out: bf16 [M,N], or fp32 when accumulate (gradient accumulation), or bf16 [M,2N] with swiglu_bwd_aux = saved u [M,2N]."""
    _chk_bf16(a, "a"); _chk_bf16(b, "b")
    flags = (GEMM_A_MN if a_mn else 0) | (GEMM_B_MN if b_mn else 0)
    n_half = 0
    if out.dtype == torch.float32:
        flags |= GEMM_OUT_F32 | (GEMM_ACCUM if accumulate else 0)
    if swiglu_bwd_aux is not None:
        flags |= GEMM_SWIGLU_BWD
        n_half = N
    if residual is not None:
        flags |= GEMM_RESIDUAL
    rc = lib().b200sat_gemm_bf16(a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), out.data_ptr(), out.stride(0), M, N, K, flags,
                                 0, _p(residual), residual.stride(0) if residual is not None else 0, 0, 0, 0, 0, 0, n_half, 0, 0, 0, 0,
                                 _p(swiglu_bwd_aux),
                                 swiglu_bwd_aux.stride(0) if swiglu_bwd_aux is not None else 0, 0, 0, 0.0, 0, force_bn, _stream())
    LAUNCHES[0] += 1
    check(rc, "gemm_bf16")
    return out


def attention(q, k, v, out=None, lse=None, scale=None):
    """q [B,Nq,Hq,64], k/v [B,Nk,Hkv,64] bf16 views (any batch/seq/head strides, unit stride on the last dim).
    Returns out [B,Nq,Hq,64] (contiguous unless `out` is given)."""
    for n, t in (("q", q), ("k", k), ("v", v)):
        _chk_bf16(t, n)
        assert t.dim() == 4
    B, Nq, Hq, D = q.shape
    _, Nk, Hkv, _ = k.shape
    if out is None:
        out = torch.empty((B, Nq, Hq, D), device=q.device, dtype=torch.bfloat16)
    if scale is None:
        scale = D ** -0.5
    rc = lib().b200sat_attention_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), _p(lse), B, Hq, Hkv, Nq, Nk,
                                     q.stride(0), q.stride(1), q.stride(2), k.stride(0), k.stride(1), k.stride(2),
                                     v.stride(0), v.stride(1), v.stride(2), out.stride(0), out.stride(1), out.stride(2),
                                     D, float(scale), _stream())
    LAUNCHES[0] += 1
    check(rc, "attention_fwd")
    return out


def layernorm(x, gamma, beta=None, scale=None, shift=None, rows_per_batch=0, out=None, eps=1e-5):
    """x [rows, D] bf16; gamma/beta fp32 [D]; adaLN scale/shift fp32 [B, >=D] (row stride = stride(0))."""
    _chk_bf16(x, "x")
    rows, D = x.shape
    if out is None:
        out = torch.empty((rows, D), device=x.device, dtype=torch.bfloat16)
    ld_mod = scale.stride(0) if scale is not None else 0
    rc = lib().b200sat_layernorm_fwd(x.data_ptr(), x.stride(0), gamma.data_ptr(), _p(beta), _p(scale), _p(shift), ld_mod,
                                     rows_per_batch, out.data_ptr(), out.stride(0), rows, D, float(eps), _stream())
    LAUNCHES[0] += 1
    check(rc, "layernorm_fwd")
    return out


def small_linear(x, w, bias=None, add=None, out=None, silu=False, out_f32=False, sigmoid_1m=False, stats=None, stats_stride=0):
    _chk_bf16(x, "x"); _chk_bf16(w, "w")
    M, K = x.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty((M, N), device=x.device, dtype=torch.float32 if out_f32 else torch.bfloat16)
    rc = lib().b200sat_small_linear(x.data_ptr(), x.stride(0), w.data_ptr(), w.stride(0), _p(bias), _p(add),
                                    add.stride(0) if add is not None else 0, out.data_ptr(), out.stride(0), M, N, K,
                                    int(silu), int(out_f32), int(sigmoid_1m), _p(stats), stats_stride, _stream())
    LAUNCHES[0] += 1
    check(rc, "small_linear")
    return out


def fourier_features(t, w, out=None, step=None, t_stride=0):
    B = out.shape[0] if out is not None else t.shape[0]
    half = w.shape[0]
    if out is None:
        out = torch.empty((B, 2 * half), device=w.device, dtype=torch.bfloat16)
    rc = lib().b200sat_fourier_features(t.data_ptr(), w.data_ptr(), out.data_ptr(), B, half, _p(step), t_stride, _stream())
    LAUNCHES[0] += 1
    check(rc, "fourier_features")
    return out


def dit_pre(x, wconv, out, reps=1, cin_table=None, step=None):
    B, C, T = x.shape
    assert x.dtype == torch.float32 and x.is_contiguous()
    rc = lib().b200sat_dit_pre(x.data_ptr(), wconv.data_ptr(), out.data_ptr(), B, C, T, reps, _p(cin_table), _p(step), _stream())
    LAUNCHES[0] += 1
    check(rc, "dit_pre")
    return out


def dit_concat(x, cond, out, reps=1, cin_table=None, step=None):
    """out rows [(rep*B+b)*T + t][Cp] bf16 = (x[b,:,t]*c_in | cond[b,:,t] | 0): dit.py:160-165 (`torch.cat([x, input_concat_cond], dim=1)`)."""
    B, C, T = x.shape
    Dc, Cp = cond.shape[1], out.shape[1]
    assert x.dtype == torch.float32 and x.is_contiguous() and cond.dtype == torch.float32 and cond.is_contiguous()
    assert cond.shape[0] == B and cond.shape[2] == T and out.dtype == torch.bfloat16 and out.is_contiguous()
    rc = lib().b200sat_dit_concat(x.data_ptr(), cond.data_ptr(), out.data_ptr(), B, C, Dc, Cp, T, reps, _p(cin_table), _p(step), _stream())
    LAUNCHES[0] += 1
    check(rc, "dit_concat")
    return out


def dit_post(h, ld_batch, prepend, wconv, out, cfg=False, cfg_scale=1.0, scale_phi=0.0):
    B, C, T = out.shape
    rc = lib().b200sat_dit_post(h.data_ptr(), ld_batch, prepend, wconv.data_ptr(), out.data_ptr(), B, C, T, int(cfg),
                                float(cfg_scale), float(scale_phi), _stream())
    LAUNCHES[0] += 1
    check(rc, "dit_post")
    return out


def sampler_update(x, v, hist, noise, coef, step, advance=True):
    rc = lib().b200sat_sampler_update(x.data_ptr(), v.data_ptr(), hist.data_ptr(), _p(noise), coef.data_ptr(), step.data_ptr(),
                                      x.numel(), int(advance), _stream())
    LAUNCHES[0] += 2 if advance else 1
    check(rc, "sampler_update")
    return x


def step_set(step, value):
    LAUNCHES[0] += 1
    check(lib().b200sat_step_set(step.data_ptr(), int(value), _stream()), "step_set")


def attention_bwd(q, k, v, o, d_o, lse, dq, dk, dv, scale=None, rope=None):
    """All tensors bf16 [B, N, H, 64] views (arbitrary batch/seq/head strides); lse fp32 [B, Hq, Nq] from the forward.
    rope = (cos, sin) fp32 [N,16]: inverse rotation applied to dq, dk (self-attention)."""
    import ctypes
    B, Nq, Hq, D = q.shape
    _, Nk, Hkv, _ = k.shape
    if scale is None:
        scale = D ** -0.5
    st = []
    for t in (q, k, v, o, d_o, dq, dk, dv):
        _chk_bf16(t, "attention_bwd operand")
        st += [t.stride(0), t.stride(1), t.stride(2)]
    arr = (ctypes.c_long * 24)(*st)
    npad = (Nq + 63) // 64 * 64
    delta = torch.empty(B * Hq * (Nq + 2 * npad), device=q.device, dtype=torch.float32)   # padded -lse*log2e and -delta rows, then delta
    rc = lib().b200sat_attention_bwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), d_o.data_ptr(), lse.data_ptr(),
                                     delta.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), B, Hq, Hkv, Nq, Nk, arr, D,
                                     float(scale), _p(rope[0]) if rope else 0, _p(rope[1]) if rope else 0, _stream())
    LAUNCHES[0] += 3
    check(rc, "attention_bwd")


def layernorm_bwd(x, dy, gamma, dres=None, out=None, dgamma=None, eps=1e-5):
    rows, D = x.shape
    if out is None:
        out = torch.empty_like(x)
    rc = lib().b200sat_layernorm_bwd(x.data_ptr(), x.stride(0), dy.data_ptr(), dy.stride(0), gamma.data_ptr(), _p(dres),
                                     dres.stride(0) if dres is not None else 0, out.data_ptr(), out.stride(0), _p(dgamma), rows, D,
                                     float(eps), _stream())
    LAUNCHES[0] += 1
    check(rc, "layernorm_bwd")
    return out


def layernorm_mod_bwd(x, dy, gamma, mod_scale, rows_per_batch, dres=None, out=None, dp=None, eps=1e-5):
    """adaLN LayerNorm backward: gain gamma*(1+mod_scale[b]); dp fp32 [B, D] += per-batch sum of dy*xhat."""
    rows, D = x.shape
    if out is None:
        out = torch.empty_like(x)
    assert mod_scale.dtype == torch.float32 and dp.dtype == torch.float32 and dp.is_contiguous()
    rc = lib().b200sat_layernorm_mod_bwd(x.data_ptr(), x.stride(0), dy.data_ptr(), dy.stride(0), gamma.data_ptr(), mod_scale.data_ptr(),
                                         mod_scale.stride(0), rows_per_batch, _p(dres), dres.stride(0) if dres is not None else 0,
                                         out.data_ptr(), out.stride(0), dp.data_ptr(), rows, D, float(eps), _stream())
    LAUNCHES[0] += 1
    check(rc, "layernorm_mod_bwd")
    return out


def gate_bwd(dh, branch, gate, dbranch, dgate, rows_per_batch):
    """dbranch = dh * gate[b]; dgate[b] += sum_n dh * branch  (gate, dgate fp32 [B, D] contiguous)."""
    M, D = dh.shape
    assert gate.dtype == torch.float32 and gate.is_contiguous() and dgate.is_contiguous()
    rc = lib().b200sat_gate_bwd(dh.data_ptr(), dh.stride(0), branch.data_ptr(), branch.stride(0), gate.data_ptr(), dbranch.data_ptr(),
                                dbranch.stride(0), dgate.data_ptr(), rows_per_batch, M // rows_per_batch, D, _stream())
    LAUNCHES[0] += 1
    check(rc, "gate_bwd")
    return dbranch


def colsum(dy, out):
    M, N = dy.shape
    rc = lib().b200sat_colsum(dy.data_ptr(), dy.stride(0), out.data_ptr(), M, N, _stream())
    LAUNCHES[0] += 1
    check(rc, "colsum")
    return out

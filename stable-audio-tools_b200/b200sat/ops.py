"""Tensor-level wrappers over the C ABI.  Each wrapper validates shapes/dtypes, allocates outputs with torch (the
caching allocator owns all memory) and launches on torch's current stream."""
import torch
from ._lib import lib, check

GEMM_BIAS, GEMM_RESIDUAL, GEMM_SILU, GEMM_SWIGLU, GEMM_ROPE, GEMM_OUT_F32, GEMM_ROW_REMAP, GEMM_GATE = 1, 2, 4, 8, 16, 32, 64, 128


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
           row_remap=None, gate=None, force_bn=0):
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
    rc = lib().b200sat_gemm_bf16(x.data_ptr(), x.stride(0), w.data_ptr(), w.stride(0), out.data_ptr(), out.stride(0),
                                 M, N, K, flags, _p(bias), _p(residual), ldr, _p(rc_), _p(rs_), rseq, rdm, rdh, n_half,
                                 seg_in, seg_out, seg_off, _p(gate), force_bn, _stream())
    check(rc, "gemm_bf16")
    return out

"""GPU parity: tcgen05 GEMM + fused epilogues vs a plain torch fp32 reference of the same op (bf16-rounded inputs)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref_linear(x, w, bias=None):
    y = x.float() @ w.float().t()
    if bias is not None:
        y = y + bias
    return y


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 512, 256), (2050, 4608, 1536), (2050, 1536, 6144), (260, 768, 768),
                                   (2050, 64, 1536), (300, 1536, 64), (77, 200, 136)])
@pytest.mark.parametrize("bn", [0, 64, 128, 256, 2128, 2192, 2256])
def test_gemm_plain(M, N, K, bn):
    from b200sat import ops
    torch.manual_seed(0)
    x = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    out = ops.linear(x, w, force_bn=bn)
    torch.cuda.synchronize()
    ref = _ref_linear(x, w)
    err = (out.float() - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= 1e-2 * scale + 1e-3, (err, scale)


def test_gemm_bias_residual_f32():
    from b200sat import ops
    torch.manual_seed(1)
    M, N, K = 1000, 1536, 1536
    x = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    b = torch.randn(N, device="cuda")
    r = torch.randn(M, N, device="cuda").bfloat16()
    out = ops.linear(x, w, bias=b, residual=r)
    ref = (_ref_linear(x, w, b).bfloat16().float() + r.float())
    assert (out.float() - ref).abs().max().item() <= 2e-2 * ref.abs().max().item()
    out32 = ops.linear(x, w, bias=b, out_f32=True)
    ref32 = _ref_linear(x, w, b)
    assert (out32 - ref32).abs().max().item() <= 2e-3 * ref32.abs().max().item()


@pytest.mark.parametrize("bn", [256, 2256])
def test_gemm_swiglu(bn):
    from b200sat import ops
    torch.manual_seed(2)
    M, Nh, K = 2050, 6144, 1536
    x = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    w = (torch.randn(2 * Nh, K, device="cuda") * 0.03).bfloat16()
    b = torch.randn(2 * Nh, device="cuda") * 0.1
    out = ops.linear(x, w, bias=b, swiglu=True, force_bn=bn)
    u = _ref_linear(x, w, b)
    ref = u[:, :Nh] * torch.nn.functional.silu(u[:, Nh:])
    assert (out.float() - ref).abs().max().item() <= 2e-2 * ref.abs().max().item()


def test_gemm_rope():
    from b200sat import ops
    torch.manual_seed(3)
    B, S, d, dh = 2, 1025, 1536, 64
    x = (torch.randn(B * S, d, device="cuda") * 0.5).bfloat16()
    w = (torch.randn(3 * d, d, device="cuda") * 0.03).bfloat16()
    inv = 1.0 / (10000 ** (torch.arange(0, 32, 2, device="cuda").float() / 32))
    fr = torch.outer(torch.arange(S, device="cuda").float(), inv)
    cos, sin = fr.cos().contiguous(), fr.sin().contiguous()
    out = ops.linear(x, w, rope=(cos, sin, S, d, dh))
    qkv = _ref_linear(x, w).bfloat16().float().view(B, S, 3, d // dh, dh)
    c = torch.cat([cos, cos], -1)[None, :, None, None, :]
    s = torch.cat([sin, sin], -1)[None, :, None, None, :]
    t = qkv[..., :32]
    rot = torch.cat([-t[..., 16:], t[..., :16]], -1)
    ref = qkv.clone()
    ref[:, :, :2, :, :32] = (t * c + rot * s)[:, :, :2]
    ref = ref.view(B * S, 3 * d)
    assert (out.float() - ref).abs().max().item() <= 2e-2 * ref.abs().max().item()


def test_gemm_row_remap_silu():
    from b200sat import ops
    torch.manual_seed(4)
    B, T, K, N = 2, 1024, 64, 1536
    x = (torch.randn(B * T, K, device="cuda")).bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.1).bfloat16()
    out = torch.zeros(B * (T + 1), N, device="cuda", dtype=torch.bfloat16)
    ops.linear(x, w, out=out, row_remap=(T, T + 1, 1))
    ref = _ref_linear(x, w).view(B, T, N)
    got = out.view(B, T + 1, N)
    assert got[:, 0].abs().max().item() == 0
    assert (got[:, 1:].float() - ref).abs().max().item() <= 2e-2 * ref.abs().max().item()
    o2 = ops.linear(x, w, silu=True)
    r2 = torch.nn.functional.silu(_ref_linear(x, w))
    assert (o2.float() - r2).abs().max().item() <= 2e-2 * r2.abs().max().item()


@pytest.mark.parametrize("bn", [0, 128, 256, 2128, 2256])
@pytest.mark.parametrize("M,N,K", [(2050, 1536, 6144), (1000, 768, 1536), (300, 200, 136)])
def test_gemm_dgrad_b_mn_major(M, N, K, bn):
    """dX[M,N] = dY[M,K] @ W[K,N]  (W = nn.Linear weight [out=K, in=N], read in place as an MN-major B operand)."""
    from b200sat import ops
    torch.manual_seed(5)
    dy = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    w = (torch.randn(K, N, device="cuda") * 0.05).bfloat16()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(dy, w, out, M, N, K, b_mn=True, force_bn=bn)
    ref = dy.float() @ w.float()
    assert (out.float() - ref).abs().max().item() <= 1e-2 * ref.abs().max().item() + 1e-3


@pytest.mark.parametrize("bn", [0, 128, 256, 2128, 2256])
@pytest.mark.parametrize("M,N,K", [(1536, 1536, 2050), (12288, 1536, 8200), (200, 136, 300)])
def test_gemm_wgrad_both_mn_major_accumulate(M, N, K, bn):
    """dW[M=out, N=in] += dY[K=tokens, M]^T @ X[K, N]  (both operands token-major => MN-major), fp32 accumulation."""
    from b200sat import ops
    torch.manual_seed(6)
    dy = (torch.randn(K, M, device="cuda") * 0.5).bfloat16()
    x = (torch.randn(K, N, device="cuda") * 0.5).bfloat16()
    acc = torch.randn(M, N, device="cuda")
    ref = acc + dy.float().t() @ x.float()
    ops.gemm(dy, x, acc, M, N, K, a_mn=True, b_mn=True, accumulate=True, force_bn=bn)
    assert (acc - ref).abs().max().item() <= 1e-2 * ref.abs().max().item() + 1e-3


def test_gemm_swiglu_save_and_backward():
    from b200sat import ops
    torch.manual_seed(7)
    M, Nh, K = 1000, 6144, 1536
    x = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    w1 = (torch.randn(2 * Nh, K, device="cuda") * 0.03).bfloat16()
    b1 = torch.randn(2 * Nh, device="cuda") * 0.1
    u = torch.empty(M, 2 * Nh, device="cuda", dtype=torch.bfloat16)
    act = ops.linear(x, w1, bias=b1, swiglu=True, save_pre=u)
    uref = x.float() @ w1.float().t() + b1
    assert (u.float() - uref).abs().max().item() <= 2e-2 * uref.abs().max().item()
    # backward through act = a*silu(g) fused into the dgrad GEMM of the second linear: dact = dY @ W2
    w2 = (torch.randn(1536, Nh, device="cuda") * 0.03).bfloat16()
    dy = torch.randn(M, 1536, device="cuda").bfloat16()
    du = torch.empty(M, 2 * Nh, device="cuda", dtype=torch.bfloat16)
    ops.gemm(dy, w2, du, M, Nh, 1536, b_mn=True, swiglu_bwd_aux=u)
    uf = u.float().requires_grad_(True)
    a, g = uf[:, :Nh], uf[:, Nh:]
    (a * torch.nn.functional.silu(g)).backward(dy.float() @ w2.float())
    assert (du.float() - uf.grad).abs().max().item() <= 2e-2 * uf.grad.abs().max().item()

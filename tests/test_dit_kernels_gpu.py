"""GPU parity for the HBM-bound DiT kernels against plain torch fp32 references of the same op."""
import math
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def test_layernorm():
    from b200sat import ops
    torch.manual_seed(0)
    for rows, D in [(2050, 1536), (77, 768), (5, 4096), (1, 64)]:
        x = (torch.randn(rows, D, device="cuda") * 2 + 0.3).bfloat16()
        g = torch.randn(D, device="cuda")
        y = ops.layernorm(x, g)
        ref = F.layer_norm(x.float(), (D,), g, None, 1e-5)
        assert (y.float() - ref).abs().max().item() <= 2e-2 * ref.abs().max().item()


def test_layernorm_adaln():
    from b200sat import ops
    torch.manual_seed(1)
    B, N, D = 2, 513, 1536
    x = torch.randn(B * N, D, device="cuda").bfloat16()
    g = torch.randn(D, device="cuda")
    mod = torch.randn(B, 6 * D, device="cuda") * 0.5
    y = ops.layernorm(x, g, scale=mod[:, :D], shift=mod[:, D:2 * D], rows_per_batch=N)
    ref = F.layer_norm(x.float(), (D,), g, None, 1e-5).view(B, N, D)
    ref = ref * (1 + mod[:, None, :D]) + mod[:, None, D:2 * D]
    assert (y.float().view(B, N, D) - ref).abs().max().item() <= 3e-2 * ref.abs().max().item()


def test_small_linear_and_fourier():
    from b200sat import ops
    torch.manual_seed(2)
    M, K, N = 2, 1536, 1536
    x = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.03).bfloat16()
    b = torch.randn(N, device="cuda")
    add = torch.randn(M, N, device="cuda").bfloat16()
    y = ops.small_linear(x, w, bias=b, silu=True)
    ref = F.silu(x.float() @ w.float().t() + b)
    assert (y.float() - ref).abs().max().item() <= 2e-2 * ref.abs().max().item()
    y = ops.small_linear(x, w, bias=b, add=add)
    ref = x.float() @ w.float().t() + b + add.float()
    assert (y.float() - ref).abs().max().item() <= 2e-2 * ref.abs().max().item()
    t = torch.rand(4, device="cuda")
    fw = torch.randn(128, 1, device="cuda").bfloat16()
    ff = ops.fourier_features(t, fw)
    f = 2 * math.pi * t.bfloat16().float()[:, None] * fw.float().t()
    ref = torch.cat([f.cos(), f.sin()], -1)
    assert (ff.float() - ref).abs().max().item() <= 0.1  # bf16 phase rounding dominates


def test_dit_pre_post():
    from b200sat import ops
    torch.manual_seed(3)
    B, C, T = 2, 64, 1024
    x = torch.randn(B, C, T, device="cuda")
    w = (torch.randn(C, C, device="cuda") * 0.1).bfloat16()
    out = torch.empty(2 * B * T, C, device="cuda", dtype=torch.bfloat16)
    cin = torch.tensor([0.5], device="cuda")
    ops.dit_pre(x, w, out, reps=2, cin_table=cin)
    xb = (x * 0.5).bfloat16().float()
    ref = (torch.einsum("oc,bct->bot", w.float(), xb) + xb).transpose(1, 2)
    got = out.view(2, B, T, C).float()
    assert (got[0] - ref).abs().max().item() <= 2e-2 * ref.abs().max().item()
    assert torch.equal(got[0], got[1])
    # post: h [2B, T+1, C]
    h = torch.randn(2 * B, T + 1, C, device="cuda").bfloat16()
    o = torch.empty(B, C, T, device="cuda")
    ops.dit_post(h, (T + 1) * C, 1, w, o, cfg=True, cfg_scale=6.0, scale_phi=0.75)
    hh = h.float()[:, 1:].transpose(1, 2)
    y = torch.einsum("oc,bct->bot", w.float(), hh) + hh
    c_, u_ = y[:B], y[B:]
    cfg = u_ + (c_ - u_) * 6.0
    r = 0.75 * (cfg * (c_.std(dim=1, keepdim=True) / cfg.std(dim=1, keepdim=True))) + 0.25 * cfg
    assert (o - r).abs().max().item() <= 4e-2 * r.abs().max().item()
    o2 = torch.empty(2 * B, C, T, device="cuda")
    ops.dit_post(h, (T + 1) * C, 1, w, o2, cfg=False)
    assert (o2 - y).abs().max().item() <= 2e-2 * y.abs().max().item()


def test_sampler_update():
    from b200sat import ops
    torch.manual_seed(4)
    n = 2 * 64 * 1024
    x = torch.randn(n, device="cuda"); v = torch.randn(n, device="cuda")
    hist = torch.randn(3, n, device="cuda"); noise = torch.randn(5, n, device="cuda")
    coef = torch.randn(5, 8, device="cuda")
    step = torch.zeros(1, dtype=torch.int32, device="cuda")
    ops.step_set(step, 2)
    x0, h0 = x.clone(), hist.clone()
    ops.sampler_update(x, v, hist, noise, coef, step)
    c = coef[2]
    den = v * c[0] + x0 * c[1]
    ref = c[2] * x0 + c[3] * den + c[4] * h0[1] + c[5] * h0[0] + c[6] * noise[2]
    assert (x - ref).abs().max().item() <= 1e-4
    assert (hist[2] - den).abs().max().item() <= 1e-5
    assert int(step.item()) == 3

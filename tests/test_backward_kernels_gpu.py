"""GPU parity of the DiT backward kernels against torch autograd (fp32) on the same bf16 inputs."""
import math
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _attn_ref(q, k, v, do, rope=None):
    B, Nq, H, D = q.shape
    Hkv = k.shape[2]
    qf, kf, vf = (t.float().detach().requires_grad_(True) for t in (q, k, v))
    qq, kk = qf, kf
    if rope is not None:
        cos, sin = rope
        def rot(t, n):
            c = torch.cat([cos[:n], cos[:n]], -1)[None, :, None, :]; s = torch.cat([sin[:n], sin[:n]], -1)[None, :, None, :]
            tr = t[..., :32]
            r = torch.cat([-tr[..., 16:], tr[..., :16]], -1)
            return torch.cat([tr * c + r * s, t[..., 32:]], -1)
        qq, kk = rot(qf, Nq), rot(kf, k.shape[1])
    qh = qq.permute(0, 2, 1, 3)
    kh = kk.permute(0, 2, 1, 3).repeat_interleave(H // Hkv, 1)
    vh = vf.permute(0, 2, 1, 3).repeat_interleave(H // Hkv, 1)
    s = qh @ kh.transpose(-1, -2) / math.sqrt(D)
    o = (torch.softmax(s, -1) @ vh).permute(0, 2, 1, 3)
    o.backward(do.float())
    return o.detach(), qf.grad, kf.grad, vf.grad, qq.detach(), kk.detach()


@pytest.mark.parametrize("B,Nq,Nk,H,Hkv,use_rope", [(1, 128, 128, 1, 1, False), (2, 1025, 1025, 6, 6, True), (2, 1025, 130, 8, 4, False),
                                                     (1, 300, 77, 4, 2, False), (1, 513, 513, 3, 3, True)])
@pytest.mark.parametrize("v3_mask", ["3", "0"])      # 3: round-2 kernels (default); 0: the round-1 kernels kept for comparison
def test_attention_bwd(B, Nq, Nk, H, Hkv, use_rope, v3_mask, monkeypatch):
    from b200sat import ops
    monkeypatch.setenv("B200SAT_ATTN_BWD_V3", v3_mask)     # read by the library at every call
    torch.manual_seed(0)
    q = torch.randn(B, Nq, H, 64, device="cuda").bfloat16()
    k = torch.randn(B, Nk, Hkv, 64, device="cuda").bfloat16()
    v = torch.randn(B, Nk, Hkv, 64, device="cuda").bfloat16()
    do = torch.randn(B, Nq, H, 64, device="cuda").bfloat16()
    rope = None
    if use_rope:
        inv = 1.0 / (10000 ** (torch.arange(0, 32, 2, device="cuda").float() / 32))
        fr = torch.outer(torch.arange(max(Nq, Nk), device="cuda").float(), inv)
        rope = (fr.cos().contiguous(), fr.sin().contiguous())
    o_ref, dq_ref, dk_ref, dv_ref, q_rot, k_rot = _attn_ref(q, k, v, do, rope)
    # the kernels consume the ROTATED q,k (what the forward stored) and return gradients w.r.t. the unrotated ones
    qin, kin = (q_rot.bfloat16(), k_rot.bfloat16()) if use_rope else (q, k)
    lse = torch.empty(B, H, Nq, device="cuda")
    o = ops.attention(qin, kin, v, lse=lse)
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    ops.attention_bwd(qin, kin, v, o, do, lse, dq, dk, dv, rope=rope)
    torch.cuda.synchronize()
    for name, a, r in (("dq", dq, dq_ref), ("dk", dk, dk_ref), ("dv", dv, dv_ref)):
        assert torch.isfinite(a.float()).all(), name
        err = (a.float() - r).abs().max().item()
        assert err <= 3e-2 * r.abs().max().item() + 1e-3, (name, err, r.abs().max().item())


def test_layernorm_bwd_and_colsum():
    from b200sat import ops
    torch.manual_seed(1)
    for rows, D in [(2050, 1536), (77, 768), (8200, 1536)]:
        x = (torch.randn(rows, D, device="cuda") * 2 + 0.3).bfloat16()
        dy = torch.randn(rows, D, device="cuda").bfloat16()
        dres = torch.randn(rows, D, device="cuda").bfloat16()
        g = torch.randn(D, device="cuda")
        xf = x.float().requires_grad_(True); gf = g.clone().requires_grad_(True)
        F.layer_norm(xf, (D,), gf, None, 1e-5).backward(dy.float())
        dgamma = torch.zeros(D, device="cuda")
        out = ops.layernorm_bwd(x, dy, g, dres=dres, dgamma=dgamma)
        ref = xf.grad + dres.float()
        assert (out.float() - ref).abs().max().item() <= 2e-2 * ref.abs().max().item()
        assert (dgamma - gf.grad).abs().max().item() <= 2e-3 * gf.grad.abs().max().item() + 1e-2
        cs = torch.zeros(D, device="cuda")
        ops.colsum(dy, cs)
        r = dy.float().sum(0)
        assert (cs - r).abs().max().item() <= 1e-3 * r.abs().max().item() + 1e-2


@pytest.mark.parametrize("rows,N", [(300001, 64), (4097, 128), (33, 64), (1000, 32)])
def test_colsum_narrow_planes(rows, N):
    """Bias gradients of the 64 / 128-channel planes (discriminator feature maps, Oobleck blocks): the narrow-plane kernel (N/8 lanes per row,
    16-byte loads) against an fp64 column sum of the same bf16 values; accumulates (+=) into `out`."""
    from b200sat import ops
    g = torch.Generator(device="cuda").manual_seed(rows + N)
    dy = torch.randn(rows, N, device="cuda", generator=g).bfloat16()
    out = torch.full((N,), 0.5, device="cuda")
    ops.colsum(dy, out)
    torch.cuda.synchronize()
    want = dy.double().sum(0) + 0.5
    assert (out.double() - want).abs().max().item() <= 1e-5 * dy.double().abs().sum(0).max().item() + 1e-4

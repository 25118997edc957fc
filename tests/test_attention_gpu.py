"""GPU parity: tcgen05 flash attention vs torch fp32 softmax(QK^T/sqrt(d))V on the same bf16 inputs."""
import math
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(q, k, v):
    # q [B,Nq,H,D], k/v [B,Nk,Hkv,D]
    B, Nq, H, D = q.shape
    Hkv = k.shape[2]
    qf = q.float().permute(0, 2, 1, 3)
    kf = k.float().permute(0, 2, 1, 3).repeat_interleave(H // Hkv, dim=1)
    vf = v.float().permute(0, 2, 1, 3).repeat_interleave(H // Hkv, dim=1)
    s = qf @ kf.transpose(-1, -2) / math.sqrt(D)
    lse = torch.logsumexp(s, dim=-1)
    o = torch.softmax(s, dim=-1) @ vf
    return o.permute(0, 2, 1, 3), lse


@pytest.mark.parametrize("B,Nq,Nk,H,Hkv", [(1, 128, 128, 1, 1), (2, 1025, 1025, 24, 24), (2, 1025, 130, 24, 12),
                                            (1, 300, 77, 4, 2), (1, 513, 513, 3, 3), (2, 4097, 4097, 2, 2), (1, 1, 1, 2, 1)])
def test_attention_fwd(B, Nq, Nk, H, Hkv):
    from b200sat import ops
    torch.manual_seed(0)
    q = torch.randn(B, Nq, H, 64, device="cuda").bfloat16()
    k = torch.randn(B, Nk, Hkv, 64, device="cuda").bfloat16()
    v = torch.randn(B, Nk, Hkv, 64, device="cuda").bfloat16()
    lse = torch.zeros(B, H, Nq, device="cuda")
    o = ops.attention(q, k, v, lse=lse)
    torch.cuda.synchronize()
    ro, rl = _ref(q, k, v)
    assert torch.isfinite(o.float()).all()
    err = (o.float() - ro).abs().max().item()
    assert err <= 2e-2, err
    assert (lse - rl).abs().max().item() <= 2e-3


def test_attention_strided_qkv():
    """Heads read in place from a fused [B*N, 3*d] projection buffer (what the DiT engine does)."""
    from b200sat import ops
    torch.manual_seed(1)
    B, N, H = 2, 1025, 24
    qkv = torch.randn(B, N, 3, H, 64, device="cuda").bfloat16()
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    out = torch.empty(B, N, H, 64, device="cuda", dtype=torch.bfloat16)
    ops.attention(q, k, v, out=out)
    ro, _ = _ref(q, k, v)
    assert (out.float() - ro).abs().max().item() <= 2e-2


def test_attention_lazy_rescale_path():
    """Large-magnitude, growing scores: the running max moves by > 2^8 many times, exercising the TMEM rescale of O."""
    from b200sat import ops
    torch.manual_seed(2)
    B, N, H = 1, 777, 2
    q = (torch.randn(B, N, H, 64, device="cuda") * 4).bfloat16()
    k = torch.randn(B, N, H, 64, device="cuda")
    k = (k * torch.linspace(0.2, 6.0, N, device="cuda")[None, :, None, None]).bfloat16()   # later keys score higher
    v = torch.randn(B, N, H, 64, device="cuda").bfloat16()
    lse = torch.zeros(B, H, N, device="cuda")
    o = ops.attention(q, k, v, lse=lse)
    ro, rl = _ref(q, k, v)
    assert torch.isfinite(o.float()).all()
    assert (o.float() - ro).abs().max().item() <= 3e-2
    assert (lse - rl).abs().max().item() <= 2e-2 * rl.abs().max().item()

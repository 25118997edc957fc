"""GPU: the fused ResidualUnit kernel (csrc/residual_unit.cu) against (a) the two-launch path it replaces (same bf16 roundings: the
k7 output is rounded to bf16 in both, so results agree to one bf16 ulp of the accumulated value) and (b) an fp32 torch evaluation of
models/autoencoders.py:58-83 on the same bf16-rounded inputs and weights (tolerance 2e-2 of the output RMS: bf16 intermediate)."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _engine():
    from b200sat.autoencoder import OobleckEngine
    from b200sat.init import oobleck_state_dict
    g = torch.Generator(device="cuda").manual_seed(5)
    sd = oobleck_state_dict(torch.device("cuda"), g)
    return OobleckEngine(sd, precision="bf16", device="cuda"), sd


@pytest.mark.parametrize("B,T", [(2, 1000), (1, 4096), (3, 256), (1, 70000)])
def test_fused_residual_units_match_two_launch_path_and_torch(B, T, monkeypatch):
    from b200sat.autoencoder import _Planes
    eng, sd = _engine()
    blk = eng.enc["blocks"][0]
    g = torch.Generator(device="cuda").manual_seed(T)
    x = torch.randn(B, T, 128, device="cuda", generator=g) * 0.7
    raw = _Planes(B, T, 128, "cuda", False); raw.hi.copy_(x.bfloat16())
    s0 = blk["rus"][0]["s0"]
    xa = x.bfloat16().float()
    act = _Planes(B, T, 128, "cuda", False)
    act.hi.copy_((xa + s0.invb * torch.sin(xa * s0.a) ** 2).bfloat16())
    monkeypatch.setenv("B200SAT_FUSED_RU", "0")
    r0, a0 = eng._residual_units(raw, act, blk["rus"], blk["snake"])
    monkeypatch.setenv("B200SAT_FUSED_RU", "1")
    from b200sat import ops
    n0 = ops.LAUNCHES[0]
    r1, a1 = eng._residual_units(raw, act, blk["rus"], blk["snake"])
    assert ops.LAUNCHES[0] - n0 == 3, "three ResidualUnits = three launches"
    torch.cuda.synchronize()
    for name, p0, p1 in (("raw", r0.hi, r1.hi), ("act", a0.hi, a1.hi)):
        d = (p0.float() - p1.float()).abs().max().item()
        scale = p0.float().abs().max().item()
        print(f"\n[fused RU B={B} T={T}] {name}: max |two-launch - fused| = {d:.3e} (max |value| {scale:.2f})")
        assert d <= 2.0 ** -6 * scale, (name, d, scale)
    # fp32 torch evaluation of the three units on the same inputs (weights = the packed bf16 weights the kernels use)
    cur = raw.hi.float().transpose(1, 2)                                  # [B, C, T]
    def snake(t, s):
        return t + s.invb[None, :, None] * torch.sin(t * s.a[None, :, None]) ** 2
    for j, ru in enumerate(blk["rus"]):
        dil = (1, 3, 9)[j]
        w7 = ru["c7"].w_hi.float().view(128, 7, 128).permute(0, 2, 1).contiguous()    # packed [co][k*Cin + ci] -> [co, ci, k]
        w1 = ru["c1"].w_hi.float().view(128, 1, 128).permute(0, 2, 1).contiguous()
        h = snake(cur, ru["s0"]).bfloat16().float()
        h = F.conv1d(h, w7, ru["c7"].bias, dilation=dil, padding=3 * dil)
        h = snake(h, ru["s1"]).bfloat16().float()
        cur = (cur + F.conv1d(h, w1, ru["c1"].bias)).bfloat16().float()
    ref = cur.transpose(1, 2)
    err = (r1.hi.float() - ref).pow(2).mean().sqrt().item() / ref.pow(2).mean().sqrt().item()
    print(f"[fused RU B={B} T={T}] rel RMS vs fp32 torch on bf16-rounded operands: {err:.3e}")
    assert err <= 2e-2

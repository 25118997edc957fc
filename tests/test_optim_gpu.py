"""Fused AdamW + EMA + bf16 refresh kernel vs torch.optim.AdamW (CPU fp32) and the EMA recurrence."""
import pytest
import torch

pytestmark = pytest.mark.gpu


class _Flat:
    def __init__(self, n, n16, seed):
        g = torch.Generator().manual_seed(seed)
        self.flat = torch.randn(n, generator=g).cuda()
        self.flat_grad = torch.zeros(n, device="cuda")
        self._bf = torch.empty(n16, device="cuda", dtype=torch.bfloat16)
        self._p = {"w": torch.nn.Parameter(self.flat)}


def test_fused_adamw_ema_matches_torch_adamw():
    from b200sat.optim import FusedAdamWEMA, ema_decay_at
    n, n16 = 4096 + 8, 2048
    mdl = _Flat(n, n16, 0)
    ref_p = torch.nn.Parameter(mdl.flat.detach().cpu().clone())
    ref_opt = torch.optim.AdamW([ref_p], lr=3e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.05)
    ema_ref = ref_p.detach().clone()
    opt = FusedAdamWEMA(mdl, lr=3e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.05, ema=True)
    g = torch.Generator().manual_seed(1)
    for step in range(1, 6):
        grad = torch.randn(n, generator=g)
        mdl.flat_grad.copy_(grad)
        opt.step(grad_scale=0.5)
        ref_p.grad = grad * 0.5
        ref_opt.step()
        d = ema_decay_at(step)
        ema_ref = ema_ref * d + ref_p.detach() * (1 - d)
        torch.cuda.synchronize()
        assert torch.allclose(mdl.flat.cpu(), ref_p.detach(), rtol=2e-5, atol=2e-6), step
        assert torch.allclose(opt.ema.cpu(), ema_ref, rtol=2e-5, atol=2e-6), step
        assert torch.equal(mdl._bf.cpu(), ref_p.detach()[:n16].bfloat16()) or \
            (mdl._bf.float().cpu() - ref_p.detach()[:n16]).abs().max() <= 2e-2
    assert ema_decay_at(1) == 0.0 and ema_decay_at(2) == 0.0 and 0.4 < ema_decay_at(3) < 0.41 and ema_decay_at(10 ** 9) == 0.9999


def test_dit_train_model_steps_with_fused_optimizer():
    """Three optimizer steps on a small DiT: loss decreases and the bf16 working copy tracks the masters."""
    from oracle import dit as odit
    from b200sat.dit_train import DiTTrainModel, v_objective_loss
    from b200sat.optim import FusedAdamWEMA
    kw = dict(embed_dim=256, depth=2, num_heads=4, io_channels=64, cond_token_dim=128, global_cond_dim=256)
    model = DiTTrainModel(odit.make_state_dict(seed=3, **kw))
    opt = FusedAdamWEMA(model, lr=2e-3, weight_decay=0.0)
    g = torch.Generator(device="cuda").manual_seed(0)
    x0 = torch.randn(2, 64, 128, device="cuda", generator=g); nz = torch.randn(2, 64, 128, device="cuda", generator=g)
    t = torch.rand(2, device="cuda", generator=g); c = torch.randn(2, 9, 128, device="cuda", generator=g); ge = torch.randn(2, 256, device="cuda", generator=g)
    losses = []
    for _ in range(4):
        model.zero_grad()
        loss = v_objective_loss(model, x0, nz, t, c, ge)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert losses[-1] < losses[0], losses
    assert torch.equal(model._bf, model.flat[: model.stack_numel].bfloat16())
    assert set(opt.ema_state_dict()) == set(model._p)


def test_optimizer_in_backward_matches_the_monolithic_step():
    """GradAllReducer(optimizer=...): per-layer AdamW slices issued on the side stream during the backward give bit-identical masters,
    moments, EMA and bf16 working copy to one whole-buffer step after the backward (same kernel, same element-wise arithmetic)."""
    from oracle import dit as odit
    from b200sat.ddp import GradAllReducer
    from b200sat.dit_train import DiTTrainModel, v_objective_loss
    from b200sat.optim import FusedAdamWEMA
    kw = dict(embed_dim=256, depth=3, num_heads=4, io_channels=64, cond_token_dim=128, global_cond_dim=256)
    sd = odit.make_state_dict(seed=4, **kw)
    g = torch.Generator(device="cuda").manual_seed(1)
    x0 = torch.randn(2, 64, 128, device="cuda", generator=g); nz = torch.randn(2, 64, 128, device="cuda", generator=g)
    t = torch.rand(2, device="cuda", generator=g); c = torch.randn(2, 9, 128, device="cuda", generator=g); ge = torch.randn(2, 256, device="cuda", generator=g)
    runs = []
    for in_backward in (False, True):
        model = DiTTrainModel(sd)
        opt = FusedAdamWEMA(model, lr=2e-3, weight_decay=1e-2, ema=True)
        red = GradAllReducer(model, optimizer=opt if in_backward else None)
        assert (red.opt is not None) == in_backward
        for _ in range(3):
            model.zero_grad()
            loss = v_objective_loss(model, x0, nz, t, c, ge)
            red.begin_step()
            loss.backward()
            red.finish()
            opt.step()
        torch.cuda.synchronize()
        runs.append((model.flat.clone(), opt.m.clone(), opt.v.clone(), opt.ema.clone(), model._bf.clone(), float(loss)))
    for a, b in zip(runs[0], runs[1]):
        if torch.is_tensor(a):
            assert torch.equal(a, b)
        else:
            assert a == b

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


def _small_train_model(seed=4):
    from oracle import dit as odit
    from b200sat.dit_train import DiTTrainModel
    kw = dict(embed_dim=256, depth=3, num_heads=4, io_channels=64, cond_token_dim=128, global_cond_dim=256)
    return DiTTrainModel(odit.make_state_dict(seed=seed, **kw))


def test_optimizer_in_backward_slices_equal_the_monolithic_step():
    """GradAllReducer(optimizer=...): the per-layer AdamW/EMA/bf16 slices issued on the side stream cover the flat buffers exactly once
    and give bit-identical masters, moments, EMA and working copy to one whole-buffer launch (same kernel, same per-element arithmetic).
    Gradients are synthetic so that the comparison does not depend on the backward kernels' atomics."""
    from b200sat.ddp import GradAllReducer
    from b200sat.optim import FusedAdamWEMA
    runs = []
    for in_backward in (False, True):
        model = _small_train_model()
        opt = FusedAdamWEMA(model, lr=2e-3, weight_decay=1e-2, ema=True)
        red = GradAllReducer(model, optimizer=opt if in_backward else None)
        assert (red.opt is not None) == in_backward
        g = torch.Generator(device="cuda").manual_seed(7)
        for _ in range(3):
            model.flat_grad.copy_(torch.randn(model.flat_grad.shape, device="cuda", generator=g) * 1e-2)
            red.begin_step()
            if in_backward:
                for i in reversed(range(model.cfg.depth)):      # the order the backward finishes layers in
                    model.grad_ready_hook(i, model.layer_grad_slice(i))
            red.finish()
            opt.step()
        torch.cuda.synchronize()
        assert opt.t == 3
        runs.append((model.flat.clone(), opt.m.clone(), opt.v.clone(), opt.ema.clone(), model._bf.clone()))
    for a, b in zip(*runs):
        assert torch.equal(a, b)
    stack = runs[1][0][: runs[1][4].numel()]
    assert torch.equal(runs[1][4], stack.bfloat16())


def test_optimizer_in_backward_training_steps_track_the_monolithic_ones():
    """End to end (real backward, hooks fired by the layer loop): three training steps with the update folded into the backward follow
    the same loss trajectory as the plain order; a missing or doubled slice raises in `step()`."""
    from b200sat.ddp import GradAllReducer
    from b200sat.dit_train import v_objective_loss
    from b200sat.optim import FusedAdamWEMA
    g = torch.Generator(device="cuda").manual_seed(1)
    x0 = torch.randn(2, 64, 128, device="cuda", generator=g); nz = torch.randn(2, 64, 128, device="cuda", generator=g)
    t = torch.rand(2, device="cuda", generator=g); c = torch.randn(2, 9, 128, device="cuda", generator=g); ge = torch.randn(2, 256, device="cuda", generator=g)
    traj = []
    for in_backward in (False, True):
        model = _small_train_model()
        opt = FusedAdamWEMA(model, lr=2e-4, weight_decay=1e-2, ema=True)
        red = GradAllReducer(model, optimizer=opt if in_backward else None)
        losses = []
        for _ in range(3):
            model.zero_grad()
            loss = v_objective_loss(model, x0, nz, t, c, ge)
            red.begin_step()
            loss.backward()
            red.finish()
            opt.step()
            losses.append(float(loss))
        traj.append(losses)
        assert model._bf_fresh
    assert traj[0][0] == pytest.approx(traj[1][0], rel=1e-6)
    assert traj[0] == pytest.approx(traj[1], rel=5e-3)
    # a step that did not cover the whole buffer is an error, not a silent partial update
    opt.begin_step()
    opt.step_slice(0, model._layer_numel, torch.cuda.current_stream().cuda_stream)
    with pytest.raises(RuntimeError):
        opt.step()

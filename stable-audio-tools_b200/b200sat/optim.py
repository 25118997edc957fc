"""Fused AdamW + EMA step over the flat fp32 master buffer of a b200sat training model (one libb200sat launch per step).

Replaces, for `DiTTrainModel`, the three passes the reference makes per step: `torch.optim.AdamW.step` (training/utils.py:60-79),
`ema_pytorch.EMA.update` (training/diffusion.py:239-247, 489-491) and the fp32 -> bf16 weight cast of autocast.
EMA decay schedule = ema_pytorch 0.2.x `get_current_decay` (that package is not vendored in the reference: parity unpinned,
restated from its published formula): epoch = step - update_after_step - 1; 0 if epoch <= 0 else
clamp(1 - (1 + epoch/inv_gamma)^-power, min_value, beta).
"""
import torch

from ._lib import lib, check
from . import ops


def ema_decay_at(step, beta=0.9999, inv_gamma=1.0, power=0.75, update_after_step=1, min_value=0.0):
    epoch = step - update_after_step - 1
    if epoch <= 0:
        return 0.0
    value = 1.0 - (1.0 + epoch / inv_gamma) ** (-power)
    return float(min(max(value, min_value), beta))


class FlatParameters:
    """Re-homes a list of nn.Parameters into ONE flat fp32 buffer (and their .grad into a second one) so that the fused optimizer
    step is a single pass: `p.data` / `p.grad` become views, names / shapes / autograd behaviour are unchanged (AccumulateGrad adds
    in place into the existing .grad views).  Exposes the `.flat` / `.flat_grad` pair FusedAdamWEMA expects."""

    def __init__(self, params):
        self.params = [p for p in params]
        if not self.params:
            raise ValueError("no parameters")
        dev = self.params[0].device
        # every parameter starts on a 64-byte boundary: the kernels read biases / SnakeBeta vectors with 16-byte loads and the weight
        # packers may vectorise; the padding elements stay zero (zero gradient => AdamW leaves them at zero)
        al = lambda n: (n + 15) // 16 * 16
        total = sum(al(p.numel()) for p in self.params)
        self.flat = torch.zeros(total, device=dev, dtype=torch.float32)
        self.flat_grad = torch.zeros(total, device=dev, dtype=torch.float32)
        self._p = {}
        off = 0
        for i, p in enumerate(self.params):
            k = p.numel()
            view = self.flat[off:off + k].view(p.shape)
            view.copy_(p.detach())
            p.data = view
            p.grad = self.flat_grad[off:off + k].view(p.shape)
            self._p[str(i)] = p
            off += al(k)

    def zero_grad(self, set_to_none=False):
        self.flat_grad.zero_()

    def parameters(self):
        return iter(self.params)


class FusedAdamWEMA:
    """optimizer + EMA for a model exposing `.flat` / `.flat_grad` (fp32, same length) and optionally `._bf` (bf16 working copy of
    the first `stack_numel` elements)."""

    def __init__(self, model, lr=5e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-3, ema=True, ema_beta=0.9999, ema_power=0.75,
                 ema_inv_gamma=1.0, ema_update_after_step=1, ema_before_step=False):
        self.model = model
        self.ema_before_step = bool(ema_before_step)
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.n = model.flat.numel()
        if self.n % 4:
            raise ValueError("flat parameter buffer length must be a multiple of 4")
        self.m = torch.zeros_like(model.flat)
        self.v = torch.zeros_like(model.flat)
        self.ema = model.flat.detach().clone() if ema else None
        self.ema_cfg = dict(beta=ema_beta, inv_gamma=ema_inv_gamma, power=ema_power, update_after_step=ema_update_after_step)
        self.t = 0

    def zero_grad(self, set_to_none=False):
        self.model.flat_grad.zero_()

    @property
    def param_groups(self):
        return [{"lr": self.lr}]

    # ---- optimizer-in-backward (b200sat.ddp.GradAllReducer(optimizer=...)): the update of a finished layer's slice runs on the
    # reducer's side stream right after that layer's gradients are final (and all-reduced), overlapping the rest of the backward;
    # `step()` then only closes the step.  A layer's weights are not read again in the step once its backward has run.
    def begin_step(self):
        self.t += 1
        self._decay = ema_decay_at(self.t, **self.ema_cfg) if self.ema is not None else 0.0
        self._sliced = 0

    def step_slice(self, off, n, stream, grad_scale=1.0):
        """AdamW (+ EMA, + bf16 refresh where the slice lies inside the working copy) on elements [off, off + n) of the flat buffers."""
        mdl = self.model
        w16 = getattr(mdl, "_bf", None)
        n16 = 0
        w16_ptr = 0
        if w16 is not None and off < w16.numel():
            n16 = min(n, w16.numel() - off) // 4 * 4
            w16_ptr = w16.data_ptr() + 2 * off
        f = lambda t: t.data_ptr() + 4 * off
        rc = lib().b200sat_adamw_ema_step(f(mdl.flat), f(mdl.flat_grad), f(self.m), f(self.v), 0 if self.ema is None else f(self.ema), w16_ptr,
                                          n, n16, self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay, self.t | (1 << 24), self._decay,
                                          grad_scale, int(self.ema_before_step), stream)
        ops.LAUNCHES[0] += 1
        check(rc, "adamw_ema_step (slice)")
        self._sliced += n

    def end_step(self):
        """All slices of this step have been issued: the bf16 working copy is fresh if it was covered."""
        mdl = self.model
        if self._sliced != self.n:
            raise RuntimeError(f"optimizer-in-backward covered {self._sliced} of {self.n} elements")
        if getattr(mdl, "_bf", None) is not None:
            mdl._bf_fresh = True

    def step(self, grad_scale=1.0):
        mdl = self.model
        if getattr(self, "_sliced", 0):       # the reducer already applied this step slice by slice
            self.end_step()
            self._sliced = 0
            return
        self.t += 1
        decay = ema_decay_at(self.t, **self.ema_cfg) if self.ema is not None else 0.0
        w16 = getattr(mdl, "_bf", None)
        n16 = (w16.numel() // 4) * 4 if w16 is not None else 0
        rc = lib().b200sat_adamw_ema_step(mdl.flat.data_ptr(), mdl.flat_grad.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
                                          0 if self.ema is None else self.ema.data_ptr(), 0 if w16 is None else w16.data_ptr(), self.n, n16,
                                          self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay, self.t, decay, grad_scale,
                                          int(self.ema_before_step), torch.cuda.current_stream().cuda_stream)
        ops.LAUNCHES[0] += 1
        check(rc, "adamw_ema_step")
        if w16 is not None and n16 == w16.numel():
            mdl._bf_fresh = True      # the forward can skip its own fp32 -> bf16 refresh

    def ema_state_dict(self):
        """EMA weights under the reference parameter names."""
        out = {}
        base = self.model.flat.data_ptr()
        for n_, p in self.model._p.items():
            off, k = (p.data_ptr() - base) // 4, p.numel()       # parameters are views of `flat` (possibly padded apart)
            out[n_] = self.ema[off:off + k].view(p.shape).clone()
        return out

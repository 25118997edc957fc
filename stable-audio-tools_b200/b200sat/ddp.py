"""Data-parallel gradient exchange for the DiT training step (the reference's implicit Lightning-DDP all-reduce,
train.py:124-164): one process per GPU, NCCL over NVLink/NVSwitch.

B200-first shape of the problem: all gradients live in ONE flat fp32 buffer ordered by layer (b200sat.dit_train), the
backward finishes layers 23 -> 0, and each finished layer is a contiguous ~176 MB slice.  The reducer launches one
asynchronous all-reduce per finished layer on a side stream, so the exchange of layer i overlaps the backward kernels of
layers i-1 ... 0; buckets are layer-sized (>= 64 MB: NVSwitch cost is launch-latency, not link-count, bound).  The mean is
obtained by scaling the loss by 1/world_size before backward (no extra pass over 4.2 GB of gradients).

SM sharing (round-2 experiment, tools/ddp2_ab.sh, 2 x B200): the persistent GEMM kernels launch one CTA per SM, so an NCCL kernel
that holds SMs while a GEMM launches could push the GEMM's last CTAs into a second wave.  Two knobs exist for that: a dedicated
communicator limited to `nccl_ctas` CTAs (B200SAT_DDP_NCCL_CTAS) and a reduced persistent-grid size while reductions are in flight
(`b200sat_set_sm_limit`, B200SAT_DDP_SM_RESERVE).  Measured step times: defaults off 152.8 ms, 8 CTAs / 16 SMs 154.6 ms, 4 / 8 160.2 ms,
16 / 32 156.0 ms (single-GPU step on the same pool: ~150 ms) - no gain, so both default to OFF; the exposed cost of the exchange at
N = 2 is ~2-3 ms (the last layer's bucket and the optimizer step waiting for it), not SM contention.
"""
import os

import torch
import torch.distributed as dist


def _grad_group(nccl_ctas):
    """A dedicated NCCL communicator for the gradient exchange with a bounded CTA count; None = default group."""
    if not (dist.is_initialized() and dist.get_backend() == "nccl") or nccl_ctas <= 0:
        return None
    try:
        opts = dist.ProcessGroupNCCL.Options()
        opts.config.max_ctas = int(nccl_ctas)
        opts.config.min_ctas = min(int(nccl_ctas), 4)
        return dist.new_group(ranks=list(range(dist.get_world_size())), backend="nccl", pg_options=opts)
    except Exception:   # older torch / NCCL without per-communicator config: fall back to the default communicator
        return None


class GradAllReducer:
    def __init__(self, model, group=None, nccl_ctas=None, sm_reserve=None, optimizer=None):
        """optimizer: a FusedAdamWEMA over `model` -> optimizer-in-backward: each finished layer's slice is all-reduced (N > 1) and then
        updated on the side stream while the backward of the earlier layers is still running (works at N = 1 too: the 40 GB optimizer
        pass hides behind the backward instead of following it).  Call `begin_step()` before backward, `finish()` + `optimizer.step()` after."""
        self.model = model
        self.opt = optimizer if (optimizer is not None and model.flat_grad.is_cuda and os.environ.get("B200SAT_OPT_IN_BACKWARD", "1") != "0") else None
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.works = []
        self.cuda = model.flat_grad.is_cuda
        self.stream = torch.cuda.Stream() if self.cuda else None
        if nccl_ctas is None:
            nccl_ctas = int(os.environ.get("B200SAT_DDP_NCCL_CTAS", "0"))
        if sm_reserve is None:
            sm_reserve = int(os.environ.get("B200SAT_DDP_SM_RESERVE", "0"))
        self.sm_reserve = sm_reserve if (self.cuda and self.world > 1) else 0
        self.group = group
        self.nccl_ctas = 0
        if group is None and self.cuda and self.world > 1:
            g = _grad_group(nccl_ctas)
            if g is not None:
                self.group, self.nccl_ctas = g, nccl_ctas
        self._limited = False
        model.grad_ready_hook = self._on_layer if (self.world > 1 or self.opt is not None) else None

    @property
    def loss_scale(self):
        return 1.0 / self.world

    def _limit(self, on):
        if not self.sm_reserve or on == self._limited:
            return
        from ._lib import lib
        L = lib()
        if on:
            L.b200sat_set_sm_limit(0)
            L.b200sat_set_sm_limit(max(2, L.b200sat_num_sms() - self.sm_reserve))
        else:
            L.b200sat_set_sm_limit(0)
        self._limited = on

    def begin_step(self):
        if self.opt is not None:
            self.opt.begin_step()

    def _launch(self, t, off=None):
        """All-reduce `t` (N > 1) and, in optimizer-in-backward mode, update elements [off, off + t.numel()) right behind it."""
        if self.cuda:
            self.stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.stream):
                if self.world > 1:
                    w = dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
                    if self.opt is not None:
                        w.wait()                      # stream-side wait: the side stream continues after the collective, the host does not block
                    else:
                        self.works.append(w)
                if self.opt is not None:
                    self.opt.step_slice(off, t.numel(), self.stream.cuda_stream)
        else:
            self.works.append(dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def _on_layer(self, layer_index, grad_slice):
        self._launch(grad_slice, layer_index * grad_slice.numel())
        if self.world > 1:
            self._limit(True)      # every persistent kernel launched from here on leaves room for the collective

    def finish(self):
        """Call after loss.backward(): handles the non-stack parameters and joins the side stream."""
        if self.world == 1 and self.opt is None:
            return
        misc = self.model.misc_grad_slice()
        self._launch(misc, self.model.flat_grad.numel() - misc.numel())
        for w in self.works:
            w.wait()
        self.works.clear()
        if self.cuda:
            torch.cuda.current_stream().wait_stream(self.stream)
        self._limit(False)

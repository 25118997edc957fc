"""Data-parallel gradient exchange for the DiT training step (the reference's implicit Lightning-DDP all-reduce,
train.py:124-164): one process per GPU, NCCL over NVLink/NVSwitch.

B200-first shape of the problem: all gradients live in ONE flat fp32 buffer ordered by layer (b200sat.dit_train), the
backward finishes layers 23 -> 0, and each finished layer is a contiguous ~176 MB slice.  The reducer launches one
asynchronous all-reduce per finished layer on a side stream, so the exchange of layer i overlaps the backward kernels of
layers i-1 ... 0; buckets are layer-sized (>= 64 MB: NVSwitch cost is launch-latency, not link-count, bound).  The mean is
obtained by scaling the loss by 1/world_size before backward (no extra pass over 4.2 GB of gradients).
"""
import torch
import torch.distributed as dist


class GradAllReducer:
    def __init__(self, model, group=None):
        self.model = model
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.works = []
        self.cuda = model.flat_grad.is_cuda
        self.stream = torch.cuda.Stream() if self.cuda else None
        model.grad_ready_hook = self._on_layer if self.world > 1 else None

    @property
    def loss_scale(self):
        return 1.0 / self.world

    def _launch(self, t):
        if self.cuda:
            self.stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.stream):
                self.works.append(dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        else:
            self.works.append(dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def _on_layer(self, layer_index, grad_slice):
        self._launch(grad_slice)

    def finish(self):
        """Call after loss.backward(): reduces the non-stack parameters and joins the side stream."""
        if self.world == 1:
            return
        self._launch(self.model.misc_grad_slice())
        for w in self.works:
            w.wait()
        self.works.clear()
        if self.cuda:
            torch.cuda.current_stream().wait_stream(self.stream)

"""CPU, world_size 2, gloo: the layer-bucketed gradient all-reduce averages gradients across ranks exactly as
DistributedDataParallel would (sum of loss/world-scaled gradients), for the flat layer-major buffer layout."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class _FakeModel:
    def __init__(self, layers=3, per=7, misc=5):
        self.L, self.per = layers, per
        self.flat_grad = torch.zeros(layers * per + misc)
        self.grad_ready_hook = None

    def layer_grad_slice(self, i):
        return self.flat_grad[i * self.per:(i + 1) * self.per]

    def misc_grad_slice(self):
        return self.flat_grad[self.L * self.per:]

    def backward(self, rank, scale):
        # rank-dependent "local" gradients, written layer by layer in reverse order like the real backward
        self.flat_grad[self.L * self.per:] = (10.0 + rank) * scale
        for i in reversed(range(self.L)):
            self.layer_grad_slice(i).copy_(torch.arange(self.per, dtype=torch.float32) * (rank + 1) * (i + 1) * scale)
            if self.grad_ready_hook:
                self.grad_ready_hook(i, self.layer_grad_slice(i))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "stable-audio-tools_b200"))
    from b200sat.ddp import GradAllReducer
    m = _FakeModel()
    red = GradAllReducer(m)
    assert abs(red.loss_scale - 1.0 / world) < 1e-12
    m.backward(rank, red.loss_scale)
    red.finish()
    q.put((rank, m.flat_grad.clone()))
    dist.barrier()
    dist.destroy_process_group()


def test_layer_bucket_allreduce_world2():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # expected mean over ranks of the unscaled local gradients
    exp = torch.zeros_like(res[0])
    for r in range(2):
        m = _FakeModel(); m.backward(r, 1.0); exp += m.flat_grad / 2
    assert torch.allclose(res[0], exp, atol=1e-6) and torch.allclose(res[1], exp, atol=1e-6)

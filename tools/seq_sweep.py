"""BASELINE.json configs[4]: DiT training-step sequence sweep N_lat in {512, 1024, 2048, 4096}, per-GPU batch chosen to fill HBM
(analytic estimate of the saved-activation footprint, backed off on OOM), pre-encoded latents, fused AdamW + EMA, DDP when launched
under torchrun.  Prints one JSON object per sequence length: tokens/s, ms/step, algorithmic TFLOP/s vs the sustained bf16 peak, and
the attention share (self-attention forward + backward kernels timed alone at the same shape x 24 layers).

    python tools/seq_sweep.py [--mem-frac 0.85] [--steps 5]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/seq_sweep.py
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stable-audio-tools_b200"))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

D, DEPTH, HEADS, LC = 1536, 24, 24, 130
LINEAR_GFLOP_PER_TOKEN = 2.038      # linear layers + cross-attention per token, forward


def gflop_per_token(n_tok):
    return LINEAR_GFLOP_PER_TOKEN + 4.0 * n_tok * D * DEPTH / 1e9


def bytes_per_token():
    # saved bf16 activations per token per layer (dit_train._workspace): h, n1, qkv(3), a1, h1, n2, q2, a2, h2, n3, u(8), act(4) = 25 d
    return 25 * D * 2 * DEPTH + 2 * HEADS * 4 * DEPTH


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mem-frac", type=float, default=0.85)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--seqs", default="512,1024,2048,4096")
    args = ap.parse_args()
    import torch.distributed as dist
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from b200sat import init, ops
    from b200sat.ddp import GradAllReducer
    from b200sat.dit_train import DiTTrainModel, v_objective_loss
    from b200sat.optim import FusedAdamWEMA
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    peak = peaks.get("bf16_tflops_sustained", 1400.0)
    sd = init.dit_state_dict(seed=0, device=dev, dtype=torch.float32)
    model = DiTTrainModel(sd, device=dev)
    del sd
    opt = FusedAdamWEMA(model, lr=5e-5, weight_decay=1e-3, ema=True)
    red = GradAllReducer(model)
    total = torch.cuda.get_device_properties(dev).total_memory
    gd = torch.Generator(device=dev).manual_seed(3 + rank)
    for n_lat in [int(s) for s in args.seqs.split(",")]:
        n_tok = n_lat + 1
        fixed = torch.cuda.memory_allocated()
        budget = args.mem_frac * total - fixed
        B = max(1, int(budget / (bytes_per_token() * n_tok + 14 * D * 2 * n_tok)))
        res = None
        while B >= 1 and res is None:
            try:
                lat = torch.randn(B, 64, n_lat, device=dev, generator=gd)
                cross = torch.randn(B, LC, 768, device=dev, generator=gd)
                glob = torch.randn(B, D, device=dev, generator=gd)

                def step():
                    noise = torch.randn(lat.shape, device=dev, generator=gd)
                    t = torch.rand(B, device=dev, generator=gd)
                    model.zero_grad()
                    loss = v_objective_loss(model, lat, noise, t, cross, glob, cfg_dropout_prob=0.1)
                    (loss * red.loss_scale).backward()
                    red.finish()
                    opt.step()
                    return loss

                for _ in range(2):
                    step()
                if world > 1:
                    dist.barrier()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(args.steps):
                    loss = step()
                e1.record()
                if world > 1:
                    dist.barrier()
                torch.cuda.synchronize()
                ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
                if world > 1:
                    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
                ms = ms.item()
                # attention share: the self-attention kernels alone at this shape
                qkv = model._ws[(B, n_tok, LC)]["qkv"][0].view(B, n_tok, 3, HEADS, 64)
                o = torch.empty(B, n_tok, HEADS, 64, device=dev, dtype=torch.bfloat16)
                lse = torch.empty(B, HEADS, n_tok, device=dev)
                dqkv = model._ws[(B, n_tok, LC)]["dqkv"].view(B, n_tok, 3, HEADS, 64)
                do = torch.randn_like(o)

                def attn():
                    ops.attention(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], out=o, lse=lse)
                    ops.attention_bwd(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], o, do, lse, dqkv[:, :, 0], dqkv[:, :, 1], dqkv[:, :, 2])
                attn(); torch.cuda.synchronize()
                e0.record()
                for _ in range(3):
                    attn()
                e1.record(); torch.cuda.synchronize()
                ms_attn = e0.elapsed_time(e1) / 3
                fl = 3 * gflop_per_token(n_tok) * 1e9 * B * n_tok
                fl_attn = 3.5 * 4.0 * B * HEADS * n_tok * n_tok * 64
                res = {"n_lat": n_lat, "batch_per_gpu": B, "n_gpus": world, "ms_per_step": ms, "tokens_per_s": B * n_lat * world / (ms * 1e-3),
                       "tflops_per_gpu": fl / (ms * 1e-3) / 1e12, "frac_of_sustained_peak": fl / (ms * 1e-3) / 1e12 / peak,
                       "self_attention_ms_per_layer": ms_attn, "self_attention_share_of_step": DEPTH * ms_attn / ms,
                       "self_attention_tflops": fl_attn / (ms_attn * 1e-3) / 1e12, "mem_allocated_gb": torch.cuda.max_memory_allocated() / 2 ** 30,
                       "loss": float(loss)}
            except torch.OutOfMemoryError:
                B = int(B * 0.85)
            finally:
                lat = cross = glob = None
                model._ws.clear()
                torch.cuda.empty_cache()
                torch.cuda.reset_peak_memory_stats()
        if rank == 0:
            print(json.dumps(res), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

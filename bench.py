#!/usr/bin/env python
"""bench.py — BASELINE.json metric on B200: DiT 100-step dpmpp-3m-sde sampling (configs[1]) latent-steps/sec.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run, one rank per GPU)
  python bench.py --impl reference ...                     (the reference's CPU path = the oracle port, on the host cores)

A "step" is one pass of the hot path over one batch: a full 100-step sampling run of the Stable-Audio-Open-1.0 DiT
(d=1536, 24 layers, 24 heads, 1024 latents + 1 prepended token, cross-attention to 130x768 conditioning, CFG scale 7 =>
effective batch 2), random-init weights, synthetic conditioning.  Multi-GPU: the sampling loop of one sample is
sequential, so ranks are independent replicas (different seeds), no data-path collective ("replicas only", weak scaling).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "stable-audio-tools_b200"))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

SAMPLE_STEPS = 100
T_LAT, L_CTX, D_MODEL, DEPTH, HEADS = 1024, 130, 1536, 24, 24
CFG_SCALE, SIGMA_MIN, SIGMA_MAX, RHO = 7.0, 0.03, 1000.0, 1.0
BATCH = 1
# algorithmic work: 2.216 GFLOP per token forward at N=1025
GFLOP_PER_TOKEN = 2.216


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return dict(hbm=j["hbm_gbs"], bf16=j["bf16_tflops"], bf16_sustained=j.get("bf16_tflops_sustained", j["bf16_tflops"]), src="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        self.idx = gpu_index

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(",") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], 0.0, set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx = max(mx, float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if "Active" in v and "Not" not in v:
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def host_threads():
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            n = max(1, min(n, int(int(q) / int(per))))
    except Exception:
        pass
    return n


def oracle_cfg_step_seconds(n_steps, threads):
    """The reference's CPU path for this workload = the oracle port (torch fp32), one CFG denoising step per call."""
    from oracle import dit as odit
    torch.set_num_threads(threads)
    sd = odit.make_state_dict(embed_dim=D_MODEL, depth=DEPTH, num_heads=HEADS, seed=0)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(BATCH, 64, T_LAT, generator=g)
    c = torch.randn(BATCH, L_CTX, 768, generator=g)
    ge = torch.randn(BATCH, D_MODEL, generator=g)
    t = torch.full((BATCH,), 0.7)
    times = []
    with torch.no_grad():
        for _ in range(n_steps):
            t0 = time.perf_counter()
            odit.dit_forward(x, t, sd, DEPTH, c, ge, cfg_scale=CFG_SCALE)
            times.append(time.perf_counter() - t0)
    return times


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = host_threads()
    oracle_cfg_step_seconds(max(1, min(args.warmup, 1)), threads)  # warm-up (bounded: one CFG step)
    times = oracle_cfg_step_seconds(args.steps, threads)
    sec = sum(times) / len(times)
    val = BATCH / sec
    line = {
        "impl": "reference", "metric": "dit_sampling_latent_steps_per_sec", "value": val, "unit": "latent-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": val, "unit": "latent-steps/s", "cores": threads, "kind": "port",
                         "sample": "each step = 1 CFG denoising step (effective batch 2, N=1025) of the 100-step workload, torch fp32 oracle port"},
        "e2e": {"value": val, "unit": "latent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(n):
    return {"workload": "Stable-Audio-Open-1.0 DiT 100-step dpmpp-3m-sde sampling (BASELINE.json configs[1])",
            "model": "DiT d=1536 L=24 H=24 ff=6144 ctx=130x768 prepend", "seq_len": T_LAT, "batch_per_gpu": BATCH,
            "cfg_scale": CFG_SCALE, "sampler": "dpmpp-3m-sde", "sample_steps": SAMPLE_STEPS,
            "sigma_min": SIGMA_MIN, "sigma_max": SIGMA_MAX, "parallelism": f"replicas x{n} (independent seeds, no collective)",
            "l2": "no explicit flush: every denoising step streams 2.1 GB of bf16 weights (>> 126 MB L2)"}


def measure_train(args, dev, rank, world, dist):
    """BASELINE.json configs[2] (DiT part): v-objective training step, batch 8 x 1024 latents per GPU (pre-encoded latents,
    random T5-shaped conditioning), bf16 compute / fp32 master weights, AdamW, layer-bucketed NCCL all-reduce when N > 1."""
    from b200sat import init
    from b200sat.dit_train import DiTTrainModel, v_objective_loss
    from b200sat.ddp import GradAllReducer
    B = 8
    sd = init.dit_state_dict(embed_dim=D_MODEL, depth=DEPTH, num_heads=HEADS, seed=0, device=dev, dtype=torch.float32)
    model = DiTTrainModel(sd, device=dev)
    del sd
    from b200sat.optim import FusedAdamWEMA
    opt = FusedAdamWEMA(model, lr=5e-5, betas=(0.9, 0.999), weight_decay=1e-3, ema=True)   # stable_audio_2_0.json:95-102, EMA on
    red = GradAllReducer(model)
    g = torch.Generator().manual_seed(42 + rank)
    h_lat = torch.randn(B, 64, T_LAT, generator=g).pin_memory()
    h_cross = torch.randn(B, L_CTX, 768, generator=g).pin_memory()
    h_glob = torch.randn(B, D_MODEL, generator=g).pin_memory()
    h_loss = torch.zeros(1).pin_memory()
    gd = torch.Generator(device=dev).manual_seed(7 + rank)

    def step():
        lat = h_lat.to(dev, non_blocking=True); cross = h_cross.to(dev, non_blocking=True); glob = h_glob.to(dev, non_blocking=True)
        noise = torch.randn(lat.shape, device=dev, generator=gd)
        t = torch.rand(B, device=dev, generator=gd)
        model.zero_grad()
        loss = v_objective_loss(model, lat, noise, t, cross, glob, cfg_dropout_prob=0.1)
        (loss * red.loss_scale).backward()
        red.finish()
        opt.step()
        h_loss.copy_(loss.detach().reshape(1), non_blocking=True)

    for _ in range(3):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    k = max(3, min(args.steps * 2, 10))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k):
        step()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_step = ms.item() / k
    tokens = B * T_LAT * world
    flop = 3 * GFLOP_PER_TOKEN * 1e9 * B * (T_LAT + 1)
    pk = peaks()
    out = {"metric": "dit_training_latent_tokens_per_sec", "value": tokens / (ms_step * 1e-3), "unit": "latent-tokens/s", "ms_per_step": ms_step,
           "steps": k, "batch_per_gpu": B, "seq_len": T_LAT, "loss": float(h_loss.item()), "optimizer": "b200sat fused AdamW + EMA + bf16 weight refresh (one pass over the fp32 masters)",
           "pre_encoded": True, "includes": "H2D of latents+conditioning, fwd, bwd, layer-bucketed all-reduce, AdamW + EMA step, D2H loss",
           "tflops_per_gpu": flop / (ms_step * 1e-3) / 1e12, "frac_of_sustained_peak": flop / (ms_step * 1e-3) / 1e12 / pk["bf16_sustained"]}
    # same step with the frozen Oobleck encoder inside it (pre_encoded = False, training/diffusion.py:364-375): 8 x 47 s stereo
    # clips encoded one at a time (iterate_batch) in bf16, then the DiT step on the fresh latents
    try:
        from b200sat.autoencoder import OobleckEngine
        ae = OobleckEngine(_oobleck_state_dict(dev, torch.Generator(device=dev).manual_seed(3)), precision="bf16", device=dev)
        audio = (torch.randn(B, 2, T_LAT * 2048, device=dev, generator=gd).clamp(-1, 1) * 0.5)

        def step_enc():
            lat = ae.encode_audio(audio, noise=None, iterate_batch=True)
            cross = h_cross.to(dev, non_blocking=True); glob = h_glob.to(dev, non_blocking=True)
            noise = torch.randn(lat.shape, device=dev, generator=gd)
            t = torch.rand(B, device=dev, generator=gd)
            model.zero_grad()
            loss = v_objective_loss(model, lat, noise, t, cross, glob, cfg_dropout_prob=0.1)
            (loss * red.loss_scale).backward()
            red.finish()
            opt.step()

        step_enc(); torch.cuda.synchronize()
        e0.record()
        for _ in range(3):
            step_enc()
        e1.record(); torch.cuda.synchronize()
        ms_enc = e0.elapsed_time(e1) / 3
        out["with_frozen_encoder"] = {"ms_per_step": ms_enc, "value": tokens / world / (ms_enc * 1e-3) * world, "unit": "latent-tokens/s",
                                      "note": "per-rank time (not max-reduced); encoder = 8 x 5.16 TFLOP forward, bf16 single-pass convs"}
        del ae, audio
    except Exception as ex:  # keep the headline numbers if the secondary measurement fails
        out["with_frozen_encoder"] = {"error": repr(ex)[:200]}
    del model, opt
    torch.cuda.empty_cache()
    return out


def measure_ae_train(args, dev, rank, world, dist):
    """BASELINE.json configs[3], generator step of the warm-up phase (training/autoencoders.py:436-497 with `warmed_up` False: the
    discriminator is not evaluated): Oobleck encode -> VAE -> decode, MRSTFT sum/difference + left + right + KL, backward, AdamW.
    16 clips x 65536 samples per GPU (the config's 32 per GPU exceeds nothing but halves the steps timed; both fit)."""
    from b200sat.autoencoder_train import OobleckTrainModel
    from b200sat.stft_loss import SumAndDifferenceSTFTLoss, autoencoder_mrstft_terms
    B, T = 16, 65536
    g = torch.Generator(device=dev).manual_seed(11)
    model = OobleckTrainModel(_oobleck_state_dict(dev, g), device=dev)
    opt = torch.optim.AdamW(model.parameters(), lr=1.5e-4, betas=(0.8, 0.99), fused=True)
    fft, hop = [2048, 1024, 512, 256, 128, 64, 32], [512, 256, 128, 64, 32, 16, 8]
    loss_sd = SumAndDifferenceSTFTLoss(fft_sizes=fft, hop_sizes=hop, win_lengths=fft, perceptual_weighting=True, sample_rate=44100)
    gh = torch.Generator().manual_seed(42 + rank)
    h_audio = (torch.randn(B, 2, T, generator=gh).clamp(-1, 1) * 0.5).pin_memory()
    h_loss = torch.zeros(1).pin_memory()
    params = list(model.parameters())

    def step():
        reals = h_audio.to(dev, non_blocking=True)
        noise = torch.randn(B, 64, T // 2048, device=dev, generator=g)
        decoded, kl, _ = model(reals, noise)
        sd_, l_, r_ = autoencoder_mrstft_terms(loss_sd, decoded, reals)
        loss = sd_ + 0.5 * l_ + 0.5 * r_ + 1e-4 * kl
        opt.zero_grad(set_to_none=True)
        (loss / world).backward()
        if world > 1:   # one flat bucket: 156 M fp32 gradients
            flat = torch.cat([p.grad.view(-1) for p in params])
            dist.all_reduce(flat)
            off = 0
            for p in params:
                p.grad.copy_(flat[off:off + p.numel()].view_as(p)); off += p.numel()
        opt.step()
        h_loss.copy_(loss.detach().reshape(1), non_blocking=True)

    for _ in range(3):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    k = 4
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_step = ms.item() / k
    flop = B * 3 * 322.7e9
    pk = peaks()
    out = {"metric": "oobleck_generator_step_items_per_sec", "value": B * world / (ms_step * 1e-3), "unit": "clips/s", "ms_per_step": ms_step,
           "batch_per_gpu": B, "samples_per_clip": T, "loss": float(h_loss.item()),
           "includes": "H2D audio, encoder+VAE+decoder fwd, 4-term MRSTFT + KL, full backward, (all-reduce), AdamW(fused), D2H loss",
           "excludes": "adversarial / feature-matching terms (warm-up phase; see ae_adversarial for the post-warm-up steps)",
           "tflops_per_gpu": flop / (ms_step * 1e-3) / 1e12, "frac_of_sustained_peak": flop / (ms_step * 1e-3) / 1e12 / pk["bf16_sustained"]}
    del model, opt
    torch.cuda.empty_cache()
    return out


def measure_ae_adversarial(args, dev, rank, world, dist):
    """BASELINE.json configs[3] after warm-up (training/autoencoders.py:436-515): alternating discriminator / generator steps of the Oobleck
    autoencoder with the EncodecDiscriminator (hinge + feature matching, weights 0.1 / 5.0), MRSTFT sum/difference + L/R and KL.
    Minimal graphs: D step = AE forward (no grad) + D forward/backward on reals and fakes; G step = AE forward/backward + D forward on
    both + D data-gradient through the fake path.  8 clips x 65536 samples per GPU; two consecutive steps (one D, one G) are timed."""
    from b200sat.autoencoder_train import OobleckTrainModel
    from b200sat.discriminator import EncodecDiscriminatorTrain
    from b200sat.init import encodec_disc_state_dict
    from b200sat.stft_loss import SumAndDifferenceSTFTLoss, autoencoder_mrstft_terms
    B, T = 8, 65536
    g = torch.Generator(device=dev).manual_seed(21)
    ae = OobleckTrainModel(_oobleck_state_dict(dev, g), device=dev)
    disc = EncodecDiscriminatorTrain(encodec_disc_state_dict(dev, g), device=dev)
    opt_g = torch.optim.AdamW(ae.parameters(), lr=1.5e-4, betas=(0.8, 0.99), fused=True)
    opt_d = torch.optim.AdamW(disc.parameters(), lr=3e-4, betas=(0.8, 0.99), fused=True)
    fft, hop = [2048, 1024, 512, 256, 128, 64, 32], [512, 256, 128, 64, 32, 16, 8]
    loss_sd = SumAndDifferenceSTFTLoss(fft_sizes=fft, hop_sizes=hop, win_lengths=fft, perceptual_weighting=True, sample_rate=44100)
    reals = (torch.randn(B, 2, T, device=dev, generator=g).clamp(-1, 1) * 0.5)

    def allreduce(params):
        if world > 1:
            flat = torch.cat([p.grad.view(-1) for p in params])
            dist.all_reduce(flat)
            off = 0
            for p in params:
                p.grad.copy_(flat[off:off + p.numel()].view_as(p)); off += p.numel()

    def d_step():
        noise = torch.randn(B, 64, T // 2048, device=dev, generator=g)
        with torch.no_grad():
            decoded = ae(reals, noise)[0]
        dis = disc.discriminator_loss(reals, decoded)
        opt_d.zero_grad(set_to_none=True)
        (dis / world).backward()
        allreduce(list(disc.parameters()))
        opt_d.step()
        return dis

    def g_step():
        noise = torch.randn(B, 64, T // 2048, device=dev, generator=g)
        decoded, kl, _ = ae(reals, noise)
        sd_, l_, r_ = autoencoder_mrstft_terms(loss_sd, decoded, reals)
        adv, fm = disc.generator_terms(reals, decoded)
        loss = sd_ + 0.5 * l_ + 0.5 * r_ + 1e-4 * kl + 0.1 * adv + 5.0 * fm
        opt_g.zero_grad(set_to_none=True)
        (loss / world).backward()
        allreduce(list(ae.parameters()))
        opt_g.step()
        return loss

    for _ in range(2):      # two warm-up rounds: the caching allocator sees both steps' buffer sizes in both orders
        d_step(); g_step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    ev[0].record()
    d = d_step()
    ev[1].record()
    l = g_step()
    ev[2].record()
    d = d_step()
    ev[3].record()
    l = g_step()
    ev[4].record()
    torch.cuda.synchronize()
    t_d = 0.5 * (ev[0].elapsed_time(ev[1]) + ev[2].elapsed_time(ev[3]))
    t_g = 0.5 * (ev[1].elapsed_time(ev[2]) + ev[3].elapsed_time(ev[4]))
    ms = torch.tensor([(t_d + t_g) / 2], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_step = ms.item()
    flop = B * 2.86e12      # mean of the G step (2.44 TFLOP/item) and the D step (3.27 TFLOP/item), minimal graphs
    out = {"metric": "oobleck_adversarial_step_items_per_sec", "value": B * world / (ms_step * 1e-3), "unit": "clips/s", "ms_per_step": ms_step,
           "d_step_ms": t_d, "g_step_ms": t_g, "batch_per_gpu": B, "samples_per_clip": T, "dis_loss": float(d.detach()), "gen_loss": float(l.detach()),
           "includes": "one discriminator step and one generator step (mean), AdamW(fused) on each parameter group, (all-reduce)",
           "tflops_per_gpu": flop / (ms_step * 1e-3) / 1e12, "frac_of_sustained_peak": flop / (ms_step * 1e-3) / 1e12 / peaks()["bf16_sustained"]}
    del ae, disc, opt_g, opt_d
    torch.cuda.empty_cache()
    return out


def _graph_time_us(fn, reps=10, iters=3):
    """Kernel time with host launch overhead removed: capture `reps` calls in a CUDA graph, replay, CUDA events."""
    fn(); torch.cuda.synchronize()
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * iters)


def other_kernels(dev, pk):
    """Roofline rows for the other named kernels of the hot path (attention, Oobleck convs, MRSTFT, LayerNorm), timed live."""
    from b200sat import ops
    from b200sat.autoencoder import OobleckEngine
    from b200sat.stft_loss import SumAndDifferenceSTFTLoss, autoencoder_mrstft_terms
    out = []
    B, N, H = 2, T_LAT + 1, HEADS
    qkv = torch.randn(B, N, 3, H, 64, device=dev).bfloat16()
    o = torch.empty(B, N, H, 64, device=dev, dtype=torch.bfloat16)
    us = _graph_time_us(lambda: ops.attention(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], out=o), reps=20)
    fl = 4.0 * B * H * N * N * 64
    out.append({"kernel": "attention_fwd_tcgen05 (self-attention B=2 H=24 N=1025 dh=64)", "bound": "tensor", "achieved": fl / us / 1e6,
                "peak": pk["bf16"], "unit": "TFLOP/s", "frac": fl / us / 1e6 / pk["bf16"], "avg_launch_ms": us / 1e3})
    x = torch.randn(2050, D_MODEL, device=dev).bfloat16(); gm = torch.ones(D_MODEL, device=dev); y = torch.empty_like(x)
    us = _graph_time_us(lambda: ops.layernorm(x, gm, out=y), reps=20)
    by = 2.0 * x.numel() * 2
    out.append({"kernel": "layernorm_kernel (2050 x 1536 bf16)", "bound": "hbm", "achieved": by / us / 1e3, "peak": pk["hbm"], "unit": "GB/s",
                "frac": by / us / 1e3 / pk["hbm"], "avg_launch_ms": us / 1e3, "note": "12.6 MB working set is L2-resident: latency-, not HBM-bound"})
    # same kernels at shapes that leave L2 / the short-sequence regime: LayerNorm at the training batch over six rotating buffers
    # (302 MB in + out > 126 MB L2), self-attention at N = 4097 (BASELINE.json configs[4] seq sweep end point)
    xs = [torch.randn(8 * (T_LAT + 1), D_MODEL, device=dev).bfloat16() for _ in range(6)]
    ys = [torch.empty_like(a) for a in xs]

    def ln6():
        for a_, b_ in zip(xs, ys):
            ops.layernorm(a_, gm, out=b_)
    us = _graph_time_us(ln6, reps=4) / 6
    by = 2.0 * xs[0].numel() * 2
    out.append({"kernel": "layernorm_kernel (8200 x 1536 bf16, training batch, rotating buffers > L2)", "bound": "hbm", "achieved": by / us / 1e3,
                "peak": pk["hbm"], "unit": "GB/s", "frac": by / us / 1e3 / pk["hbm"], "avg_launch_ms": us / 1e3})
    del xs, ys
    N4 = 4097
    qkv4 = torch.randn(B, N4, 3, H, 64, device=dev).bfloat16()
    o4 = torch.empty(B, N4, H, 64, device=dev, dtype=torch.bfloat16)
    us = _graph_time_us(lambda: ops.attention(qkv4[:, :, 0], qkv4[:, :, 1], qkv4[:, :, 2], out=o4), reps=5)
    fl = 4.0 * B * H * N4 * N4 * 64
    out.append({"kernel": "attention_fwd_tcgen05 (self-attention B=2 H=24 N=4097 dh=64)", "bound": "tensor", "achieved": fl / us / 1e6,
                "peak": pk["bf16"], "unit": "TFLOP/s", "frac": fl / us / 1e6 / pk["bf16"], "avg_launch_ms": us / 1e3})
    del qkv4, o4
    # SnakeBeta backward stream (Oobleck training): 16 x 65536 x 128 elements, reads d_act / x / d_skip, writes d_raw (8 B per element)
    from b200sat._lib import lib as _lib, check as _check
    rows, C = 16 * 65536, 128
    pl = [torch.randn(rows, C, device=dev).bfloat16() for _ in range(4)]
    sa_, sb_ = torch.rand(C, device=dev) + 0.5, torch.rand(C, device=dev) + 0.5
    acc3 = [torch.zeros(C, device=dev) for _ in range(3)]
    st = torch.cuda.current_stream

    def snk():
        _check(_lib().b200sat_snake_bwd(pl[0].data_ptr(), pl[1].data_ptr(), pl[2].data_ptr(), sa_.data_ptr(), sb_.data_ptr(), pl[3].data_ptr(),
                                        acc3[0].data_ptr(), acc3[1].data_ptr(), acc3[2].data_ptr(), rows, C, st().cuda_stream), "snake_bwd")
    us = _graph_time_us(snk, reps=5)
    by = 8.0 * rows * C
    out.append({"kernel": "snake_bwd_kernel (16 x 65536 x 128, with skip add and alpha/beta/bias reductions)", "bound": "hbm", "achieved": by / us / 1e3,
                "peak": pk["hbm"], "unit": "GB/s", "frac": by / us / 1e3 / pk["hbm"], "avg_launch_ms": us / 1e3})
    del pl
    # Oobleck: random-init weights of the stable_audio_2_0_vae architecture, 47 s stereo clip (1024 latents)
    g = torch.Generator(device=dev).manual_seed(0)
    sd = _oobleck_state_dict(dev, g)
    for prec in ("bf16", "fp32x3"):
        eng = OobleckEngine(sd, precision=prec, device=dev)
        a = torch.randn(1, 2, T_LAT * 2048, device=dev) * 0.3
        z = eng.encode(a); torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record(); z = eng.encode(a); e1.record(); w = eng.decode(z); e2.record(); torch.cuda.synchronize()
        fl = 5.163e12
        for name, ms in (("OobleckEncoder fwd", e0.elapsed_time(e1)), ("OobleckDecoder fwd", e1.elapsed_time(e2))):
            out.append({"kernel": f"conv1d_tcgen05 stack: {name}, 47 s stereo clip, precision={prec}", "bound": "tensor", "achieved": fl / ms / 1e9,
                        "peak": pk["bf16_sustained"], "unit": "TFLOP/s (algorithmic; fp32x3 executes 3x the MMAs)", "frac": fl / ms / 1e9 / pk["bf16_sustained"], "ms": ms})
        del eng, a, z, w
        torch.cuda.empty_cache()
    FFT = [2048, 1024, 512, 256, 128, 64, 32]
    loss = SumAndDifferenceSTFTLoss(FFT, [n // 4 for n in FFT], FFT, perceptual_weighting=True, sample_rate=44100)
    reals = torch.randn(8, 2, 65536, device=dev) * 0.3; dec = reals + 0.05 * torch.randn_like(reals)
    autoencoder_mrstft_terms(loss, dec, reals); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        autoencoder_mrstft_terms(loss, dec, reals)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    gf = 0.55 * 8  # ~0.55 GFLOP fp32 per item per generator step
    out.append({"kernel": "MRSTFT (FIR + 7-resolution Stockham STFT + loss sums), 8 x 2 x 65536, all four generator-loss terms", "bound": "hbm",
                "achieved": 8 * 2 * 2 * 65536 * 4 / ms / 1e6, "peak": pk["hbm"], "unit": "GB/s", "frac": 8 * 2 * 2 * 65536 * 4 / ms / 1e6 / pk["hbm"], "ms": ms,
                "gflops_fp32": gf / ms, "reference_materialised_traffic_gbs": 8 * 120e6 / ms / 1e6,
                "note": "fused: 8.4 MB of waveforms in, 84 scalars out; bound by fp32 SIMT/shared memory, not HBM (frac is vs the HBM peak only "
                        "because the contract wants one; the reference moves ~120 MB per item through HBM for the same result, "
                        "reference_materialised_traffic_gbs is that traffic divided by our time)"})
    return out


def _oobleck_state_dict(dev, g):
    from b200sat.init import oobleck_state_dict
    return oobleck_state_dict(dev, g)


def run_ours(args):
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    # stdout carries exactly one line (the JSON): libraries that announce themselves on stdout (NCCL prints its version from C) are
    # sent to stderr until the result is printed
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from b200sat import init, ops, sampling
    from b200sat.generation import DiffusionCondModel, generate_diffusion_cond
    dev = torch.device("cuda", local)
    sd = init.dit_state_dict(embed_dim=D_MODEL, depth=DEPTH, num_heads=HEADS, seed=0, device=dev)
    model = DiffusionCondModel.from_state_dict(sd, device=dev)
    eng = model.engine
    del sd
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    noise = torch.randn(BATCH, 64, T_LAT, device=dev, generator=g)
    cross = torch.randn(BATCH, L_CTX, 768, device=dev, generator=g)
    glob = torch.randn(BATCH, D_MODEL, device=dev, generator=g)
    smp = model.sampler(BATCH, T_LAT, L_CTX, True, CFG_SCALE, 0.0)

    def one_sample():
        return sampling.sample_k_dpmpp_3m_sde(eng, noise, SAMPLE_STEPS, SIGMA_MIN, SIGMA_MAX, RHO, cross, glob, CFG_SCALE, 0.0, sampler=smp)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        out = one_sample()
    assert torch.isfinite(out).all(), "non-finite latents"
    # ---------------- device-resident timing (`value`)
    clocks = ClockSampler(local)
    barrier()
    if rank == 0:
        clocks.start()
    n0 = ops.LAUNCHES[0]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        one_sample()
    e1.record()
    barrier()
    launches = ops.LAUNCHES[0] - n0
    clk = clocks.stop() if rank == 0 else None
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = ms.item()
    value = SAMPLE_STEPS * BATCH * world * args.steps / (ms_total * 1e-3)

    # ---------------- end-to-end through the public API with HOST buffers (`e2e`)
    h_noise = noise.cpu().pin_memory(); h_cross = cross.cpu().pin_memory(); h_glob = glob.cpu().pin_memory()
    h_out = torch.empty(BATCH, 64, T_LAT).pin_memory()

    def one_e2e():
        lat = generate_diffusion_cond(model, steps=SAMPLE_STEPS, cfg_scale=CFG_SCALE, batch_size=BATCH,
                                      conditioning_tensors={"cross_attn_cond": h_cross, "global_cond": h_glob},
                                      sample_size=T_LAT * 2048, sampler_type="dpmpp-3m-sde", sigma_min=SIGMA_MIN,
                                      sigma_max=SIGMA_MAX, rho=RHO, noise=h_noise, return_latents=True, device=dev)
        h_out.copy_(lat, non_blocking=True)

    one_e2e()
    barrier()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        one_e2e()
    t1.record()
    barrier()
    ms2 = torch.tensor([t0.elapsed_time(t1)], device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_val = SAMPLE_STEPS * BATCH * world * args.steps / (ms2.item() * 1e-3)
    h2d = h_noise.numel() * 4 + h_cross.numel() * 4 + h_glob.numel() * 4
    d2h = h_out.numel() * 4

    train = None
    ae_train = None
    ae_adv = None
    if not args.no_train:
        del smp
        model._samplers.clear()
        torch.cuda.empty_cache()
        train = measure_train(args, dev, rank, world, dist)
        try:
            ae_train = measure_ae_train(args, dev, rank, world, dist)
        except Exception as ex:   # secondary measurement: never lose the headline line
            ae_train = {"error": repr(ex)[:300]}
            torch.cuda.empty_cache()
        try:
            ae_adv = measure_ae_adversarial(args, dev, rank, world, dist)
        except Exception as ex:
            ae_adv = {"error": repr(ex)[:300]}
            torch.cuda.empty_cache()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---------------- roofline of the dominant kernel: the FF1 SwiGLU GEMM (M=2050, N=2x6144, K=1536), timed live
    pk = peaks()
    M = 2 * BATCH * (T_LAT + 1)
    xin = torch.randn(M, D_MODEL, device=dev).bfloat16()
    outb = torch.empty(M, 4 * D_MODEL, device=dev, dtype=torch.bfloat16)
    ws = [(eng.w[f"transformer.layers.{i}.ff.ff.0.proj.weight"], eng.w[f"transformer.layers.{i}.ff.ff.0.proj.bias"]) for i in range(DEPTH)]
    for w_, b_ in ws[:3]:
        ops.linear(xin, w_, bias=b_, swiglu=True, out=outb)
    torch.cuda.synchronize()
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 4
    r0.record()
    for _ in range(reps):
        for w_, b_ in ws:  # 24 different 37.7 MB weight matrices: inputs larger than L2
            ops.linear(xin, w_, bias=b_, swiglu=True, out=outb)
    r1.record()
    torch.cuda.synchronize()
    k_ms = r0.elapsed_time(r1) / (reps * DEPTH)
    k_flop = 2.0 * M * (8 * D_MODEL) * D_MODEL
    achieved = k_flop / (k_ms * 1e-3) / 1e12
    roof = {"kernel": "gemm_bf16_tcgen05<256> (FF1 + SwiGLU epilogue, M=2050 N=12288 K=1536)", "bound": "tensor",
            "achieved": achieved, "peak": pk["bf16"], "unit": "TFLOP/s", "frac": achieved / pk["bf16"], "peak_source": pk["src"] + " burst (kernel timed alone)",
            "avg_launch_ms": k_ms, "traffic": 47.5e6,
            "traffic_source": "profiles/r1_ncu_full_summary.txt: dram read 44.2 MB + write 3.3 MB per launch (ncu --set full); algorithmic 69 MB"}
    other = other_kernels(dev, pk) if not args.no_train else None
    step_tflop = SAMPLE_STEPS * 2 * BATCH * (T_LAT + 1) * GFLOP_PER_TOKEN / 1e3
    whole = {"tflop_per_sample": step_tflop, "achieved_tflops": step_tflop * args.steps * world / (ms_total * 1e-3) / world,
             "frac_of_sustained_peak": step_tflop * args.steps / (ms_total * 1e-3) / pk["bf16_sustained"]}

    # ---------------- CPU baseline (oracle port) on a bounded sample
    threads = host_threads()
    cpu = None
    if not args.no_cpu_baseline:
        ts = oracle_cfg_step_seconds(2, threads)
        cpu_sec = min(ts)
        cpu = {"value": BATCH / cpu_sec, "unit": "latent-steps/s", "cores": threads, "kind": "port",
               "sample": "2 CFG denoising steps (of the 100-step workload; effective batch 2, N=1025), torch fp32 oracle port; best of 2"}

    line = {
        "metric": "dit_sampling_latent_steps_per_sec", "value": value, "unit": "latent-steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic (random-init weights, random conditioning)",
        "config": workload_config(world), "sample_seconds_100_steps": ms_total / args.steps / 1e3,
        "e2e": {"value": e2e_val, "unit": "latent-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "b200sat.generation.generate_diffusion_cond(host pinned noise/conditioning -> host latents)"},
        "gpu_launches": launches, "clocks": clk, "roofline": roof, "whole_step": whole, "cpu_baseline": cpu, "train": train, "ae_train": ae_train, "ae_adversarial": ae_adv, "other_kernels": other,
    }
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the secondary training-step measurement")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py — BASELINE.json metric on B200.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run, one rank per GPU)
  python bench.py --impl reference ...                     (the reference's own modules from baseline/_ref on the host cores)

Headline (`value`, every N): the Stable-Audio-Open-1.0 latent-diffusion TRAINING STEP of BASELINE.json configs[2] —
per GPU a batch of 8 x 47 s 44.1 kHz stereo clips (synthetic audio, 2 097 152 samples each), frozen Oobleck encoder ->
1024 latents per clip, random T5-shaped conditioning (128 x 768 + two number tokens), v-objective, bf16 compute with fp32
master weights, fused AdamW + EMA, and — for N > 1 — the data-parallel gradient all-reduce (NCCL over NVLink, layer-bucketed,
overlapped with the backward).  value = latent tokens/s over all ranks (8192 per GPU per step).  `e2e` is the same step fed from
pinned HOST buffers (audio + conditioning H2D inside the timed region, loss D2H).  This is the half of the two-part metric that
shards with a collective, so the driver's 1 -> 8 scaling curve measures the all-reduce.

Second half of the metric (`sample`, every N, compact): BASELINE.json configs[1], 100-step dpmpp-3m-sde sampling of the same DiT
(1024 latents, CFG 7 => effective batch 2), latent-steps/s and seconds per sample; ranks are independent replicas.

`gpu_reference`: the UNMODIFIED reference (baseline/_ref: its own create_model_from_config, DiffusionCondTrainingWrapper.training_step,
generate_diffusion_cond) on the same GPU(s) in the same run — the comparator BASELINE.md section 4 names (PyTorch + SDPA/flash, cuDNN,
torch.optim.AdamW, bf16 autocast).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "stable-audio-tools_b200"))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

SAMPLE_STEPS = 100
T_LAT, L_CTX, D_MODEL, DEPTH, HEADS = 1024, 130, 1536, 24, 24
T_AUDIO = T_LAT * 2048
CFG_SCALE, SIGMA_MIN, SIGMA_MAX, RHO = 7.0, 0.03, 1000.0, 1.0
TRAIN_BATCH = 8
GFLOP_PER_TOKEN = 2.216          # DiT forward per token at N = 1025
ENC_TFLOP_PER_CLIP = 5.163       # Oobleck encoder forward, 47 s stereo clip
CPU_SAMPLE_LAT = 256             # the CPU arm's bounded sample: one clip of 256 latents (524 288 samples) per step


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


def workload_config(n):
    return {"workload": "Stable-Audio-Open-1.0 latent-diffusion training step, v-objective (BASELINE.json configs[2])",
            "batch_per_gpu": TRAIN_BATCH, "global_batch": TRAIN_BATCH * n, "clip": "47.55 s stereo @ 44.1 kHz (2097152 samples)",
            "seq_len": T_LAT, "frozen_encoder_in_step": True, "conditioning": "random T5-shaped 128x768 + 2 number tokens, cfg_dropout 0.1",
            "optimizer": "AdamW lr 5e-5 wd 1e-3 + EMA (fused, fp32 masters)", "parallelism": f"dp{n}" + (" (NCCL all-reduce, layer buckets overlapped with backward)" if n > 1 else ""),
            "l2": "no explicit flush: every step streams > 20 GB (weights, saved activations, optimizer state) >> 126 MB L2"}


# ================================================================================================================
# the reference's own modules (baseline/_ref): CPU arm, cpu_baseline, GPU comparator
# ================================================================================================================
class _PromptOverride(torch.nn.Module):
    """MultiConditioner whose text branch is replaced by a fixed random embedding (the text encoder is outside the path and needs
    a checkpoint download); the number conditioners are the reference's own modules and run every step."""

    def __init__(self, inner, prompt):
        super().__init__()
        self.inner = inner
        self.register_buffer("prompt", prompt)

    def forward(self, meta, device):
        ct = self.inner(meta, device)
        b = len(meta)
        ct["prompt"] = (self.prompt[:b].to(ct["seconds_total"][0].dtype), torch.ones(b, self.prompt.shape[1], device=self.prompt.device, dtype=torch.bool))
        return ct


def _reference_trainer(device, batch, t_lat, with_encoder, seed=0):
    """Returns (step_fn(audio_or_latents) -> loss tensor, wrapper).  The step is the reference's DiffusionCondTrainingWrapper.training_step
    + backward + its configured AdamW + the EMA update, under bf16 autocast on CUDA (Lightning 'bf16-mixed'), fp32 on CPU."""
    import types
    from baseline import ref_loader, ref_models
    R = ref_loader.load(force_sdpa=True if torch.device(device).type == "cpu" else None)   # flash_attn has no CPU kernels
    T_ = ref_loader.load_training()
    cfg = ref_models.sao_config(pretransform=with_encoder)
    with torch.device(device):
        torch.manual_seed(seed)
        model = R.factory.create_model_from_config(cfg)
    ref_models.rerandomize_zero_init(model.model, seed=seed + 1)
    model = model.to(device)
    if model.pretransform is not None:
        with torch.no_grad():
            for n_, p in model.pretransform.named_parameters():
                if n_.endswith("weight_g"):
                    p.mul_(0.5)
    g = torch.Generator().manual_seed(seed + 2)
    model.conditioner = _PromptOverride(model.conditioner, torch.randn(batch, 128, 768, generator=g).to(device))
    tc = cfg["training"]
    wrap = T_.diffusion.DiffusionCondTrainingWrapper(model, use_ema=True, pre_encoded=not with_encoder, cfg_dropout_prob=tc["cfg_dropout_prob"],
                                                     optimizer_configs=tc["optimizer_configs"]).to(device)
    opt = wrap.configure_optimizers()[0]
    wrap.trainer = types.SimpleNamespace(optimizers=[opt])
    n_in = t_lat * 2048 if with_encoder else t_lat
    meta = [{"prompt": 0, "seconds_start": 0.0, "seconds_total": 47.0, "padding_mask": torch.ones(n_in, dtype=torch.bool)} for _ in range(batch)]
    is_cuda = torch.device(device).type == "cuda"

    hooks = {"post_backward": None, "meta": meta}

    def step(x):
        if is_cuda:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                loss = wrap.training_step((x, hooks["meta"]), 0)
        else:
            loss = wrap.training_step((x, hooks["meta"]), 0)
        loss.backward()
        if hooks["post_backward"] is not None:
            hooks["post_backward"]()
        opt.step()
        wrap.on_before_zero_grad()          # Lightning's order: EMA update before zero_grad (training/diffusion.py:489-491)
        opt.zero_grad(set_to_none=True)
        return loss.detach()

    step.hooks = hooks
    return step, wrap, R


def cpu_reference_train(n_steps, threads, budget_s=150.0, warmup=1):
    """The reference's CPU path on a bounded sample: one clip of CPU_SAMPLE_LAT latents per step (1/4 of a 47 s clip), full step
    (frozen encoder forward, DiT forward + checkpointed backward, AdamW, EMA), torch fp32.  Stops early when `budget_s` is spent."""
    torch.set_num_threads(threads)
    step, wrap, R = _reference_trainer("cpu", 1, CPU_SAMPLE_LAT, with_encoder=True)
    g = torch.Generator().manual_seed(5)
    audio = torch.randn(1, 2, CPU_SAMPLE_LAT * 2048, generator=g).clamp(-1, 1) * 0.5
    for _ in range(warmup):
        step(audio)
    times, t_start = [], time.perf_counter()
    for _ in range(n_steps):
        t0 = time.perf_counter()
        step(audio)
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start > budget_s:
            break
    return times, R.attention_backend


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = host_threads()
    times, attn = cpu_reference_train(args.steps, threads, budget_s=170.0, warmup=1)   # warm-up bounded to one step
    sec = sum(times) / len(times)
    val = CPU_SAMPLE_LAT / sec
    sample = (f"each step = the full training step on ONE clip of {CPU_SAMPLE_LAT} latents ({CPU_SAMPLE_LAT * 2048} samples; the workload has 8 x 1024 per GPU): "
              f"reference DiffusionCondTrainingWrapper.training_step (frozen Oobleck encode + DiT fwd/bwd with its checkpointing) + AdamW + EMA, torch fp32, "
              f"{attn} attention; {len(times)} of the requested {args.steps} steps fit the time budget")
    line = {
        "impl": "reference", "metric": "dit_training_latent_tokens_per_sec", "value": val, "unit": "latent-tokens/s",
        "n_gpus": args.gpus, "steps": len(times), "steps_requested": args.steps, "warmup": 1, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args.gpus),
        "cpu_baseline": {"value": val, "unit": "latent-tokens/s", "cores": threads, "torch_threads": torch.get_num_threads(), "kind": "reference", "sample": sample},
        "e2e": {"value": val, "unit": "latent-tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ================================================================================================================
# our arm
# ================================================================================================================
def _timed(fn, k, barrier, dev, dist, world):
    """K calls of fn bracketed by barrier + synchronize, CUDA events, max over ranks -> total ms."""
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k):
        fn()
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return ms.item()


def _oobleck_state_dict(dev, g):
    from b200sat.init import oobleck_state_dict
    return oobleck_state_dict(dev, g)


class OurTrainer:
    """The configs[2] step on the b200sat engines (public training API: OobleckEngine.encode_audio, DiTTrainModel, v_objective_loss,
    GradAllReducer, FusedAdamWEMA)."""

    def __init__(self, dev, rank, world):
        from b200sat import init
        from b200sat.autoencoder import OobleckEngine
        from b200sat.ddp import GradAllReducer
        from b200sat.dit_train import DiTTrainModel
        from b200sat.optim import FusedAdamWEMA
        self.dev, self.rank, self.world = dev, rank, world
        sd = init.dit_state_dict(embed_dim=D_MODEL, depth=DEPTH, num_heads=HEADS, seed=0, device=dev, dtype=torch.float32)
        self.model = DiTTrainModel(sd, device=dev)
        del sd
        self.opt = FusedAdamWEMA(self.model, lr=5e-5, betas=(0.9, 0.999), weight_decay=1e-3, ema=True)   # stable_audio_2_0.json:95-102
        self.red = GradAllReducer(self.model, optimizer=self.opt)     # optimizer-in-backward: slices updated behind the all-reduce
        self.ae = OobleckEngine(_oobleck_state_dict(dev, torch.Generator(device=dev).manual_seed(3)), precision="bf16", device=dev)
        B = TRAIN_BATCH
        g = torch.Generator().manual_seed(42 + rank)                      # train.py:30-33: seed + rank
        self.h_audio = [(torch.randn(B, 2, T_AUDIO, generator=g).clamp(-1, 1) * 0.5).pin_memory() for _ in range(2)]
        self.h_cross = torch.randn(B, L_CTX, 768, generator=g).pin_memory()
        self.h_glob = torch.randn(B, D_MODEL, generator=g).pin_memory()
        self.h_loss = torch.zeros(1).pin_memory()
        self.audio = self.h_audio[0].to(dev)
        self.cross, self.glob = self.h_cross.to(dev), self.h_glob.to(dev)
        self.lat = torch.randn(B, 64, T_LAT, device=dev)
        self.gd = torch.Generator(device=dev).manual_seed(7 + rank)
        self.sobol = torch.quasirandom.SobolEngine(1, scramble=True, seed=11 + rank)     # training/diffusion.py:256,383
        self.copy_stream = torch.cuda.Stream()
        self.d_audio = [torch.empty_like(self.audio) for _ in range(2)]
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]
        self.flip = 0

    def _dit_step(self, lat, cross, glob):
        from b200sat.dit_train import v_objective_loss
        B = lat.shape[0]
        noise = torch.randn(lat.shape, device=self.dev, generator=self.gd)
        t = self.sobol.draw(B)[:, 0].to(self.dev, non_blocking=True)
        self.model.zero_grad()
        loss = v_objective_loss(self.model, lat, noise, t, cross, glob, cfg_dropout_prob=0.1)
        self.red.begin_step()
        (loss * self.red.loss_scale).backward()
        self.red.finish()
        self.opt.step()
        return loss

    def step_resident(self):
        """Inputs already in HBM: frozen encoder (one clip at a time, autoencoders.py:470-474) -> VAE sample -> DiT step."""
        vae_noise = torch.randn(TRAIN_BATCH, 64, T_LAT, device=self.dev, generator=self.gd)
        lat = self.ae.encode_audio(self.audio, noise=vae_noise, iterate_batch=True)
        return self._dit_step(lat, self.cross, self.glob)

    def step_pre_encoded(self):
        return self._dit_step(self.lat, self.cross, self.glob)

    def prefetch(self):
        """H2D of the NEXT batch on the copy stream (what a pinned-memory DataLoader does), double-buffered."""
        i = self.flip
        with torch.cuda.stream(self.copy_stream):
            self.d_audio[i].copy_(self.h_audio[i], non_blocking=True)
            self.ready[i].record(self.copy_stream)

    def step_e2e(self):
        i = self.flip
        torch.cuda.current_stream().wait_event(self.ready[i])
        audio = self.d_audio[i]
        self.flip ^= 1
        self.prefetch()                                   # next batch's copy overlaps this step's compute
        cross = self.h_cross.to(self.dev, non_blocking=True); glob = self.h_glob.to(self.dev, non_blocking=True)
        vae_noise = torch.randn(TRAIN_BATCH, 64, T_LAT, device=self.dev, generator=self.gd)
        lat = self.ae.encode_audio(audio, noise=vae_noise, iterate_batch=True)
        loss = self._dit_step(lat, cross, glob)
        self.h_loss.copy_(loss.detach().reshape(1), non_blocking=True)

    def free(self):
        self.red._limit(False)
        del self.model, self.opt, self.ae, self.audio, self.d_audio
        torch.cuda.empty_cache()


def measure_sampling(args, dev, rank, world, dist, barrier):
    from b200sat import init, sampling
    from b200sat.generation import DiffusionCondModel, generate_diffusion_cond
    sd = init.dit_state_dict(embed_dim=D_MODEL, depth=DEPTH, num_heads=HEADS, seed=0, device=dev)
    model = DiffusionCondModel.from_state_dict(sd, device=dev)
    eng = model.engine
    del sd
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    noise = torch.randn(1, 64, T_LAT, device=dev, generator=g)
    cross = torch.randn(1, L_CTX, 768, device=dev, generator=g)
    glob = torch.randn(1, D_MODEL, device=dev, generator=g)
    smp = model.sampler(1, T_LAT, L_CTX, True, CFG_SCALE, 0.0)
    one = lambda: sampling.sample_k_dpmpp_3m_sde(eng, noise, SAMPLE_STEPS, SIGMA_MIN, SIGMA_MAX, RHO, cross, glob, CFG_SCALE, 0.0, sampler=smp)
    for _ in range(3):
        out = one()
    assert torch.isfinite(out).all(), "non-finite latents"
    k = max(3, min(args.steps, 8))
    ms = _timed(one, k, barrier, dev, dist, world)
    h_noise = noise.cpu().pin_memory(); h_cross = cross.cpu().pin_memory(); h_glob = glob.cpu().pin_memory()
    h_out = torch.empty(1, 64, T_LAT).pin_memory()

    def e2e():
        lat = generate_diffusion_cond(model, steps=SAMPLE_STEPS, cfg_scale=CFG_SCALE, batch_size=1,
                                      conditioning_tensors={"cross_attn_cond": h_cross, "global_cond": h_glob}, sample_size=T_AUDIO,
                                      sampler_type="dpmpp-3m-sde", sigma_min=SIGMA_MIN, sigma_max=SIGMA_MAX, rho=RHO, noise=h_noise,
                                      return_latents=True, device=dev)
        h_out.copy_(lat, non_blocking=True)

    e2e()
    ms2 = _timed(e2e, k, barrier, dev, dist, world)
    tflop = SAMPLE_STEPS * 2 * (T_LAT + 1) * GFLOP_PER_TOKEN / 1e3
    res = {"workload": "configs[1]: 100-step dpmpp-3m-sde, 1024 latents, CFG 7 (batch 2), bf16; replicas x%d" % world,
           "latent_steps_per_s": SAMPLE_STEPS * world * k / (ms * 1e-3), "seconds_per_sample": ms / k / 1e3,
           "e2e_latent_steps_per_s": SAMPLE_STEPS * world * k / (ms2 * 1e-3), "samples_timed": k,
           "frac_of_sustained_peak": tflop / (ms / k * 1e-3) / peaks()["bf16_sustained"]}
    del smp
    model._samplers.clear()
    return res, eng


def gpu_reference(args, dev, rank, world, dist, barrier, do_train=True, do_sample=True):
    """The unmodified reference on the same GPU(s), same run: bf16, its own attention dispatch, cuDNN convs, torch AdamW."""
    out = {"impl": "stable-audio-tools 0.0.19 (baseline/_ref, unmodified), torch %s" % torch.__version__, "compile": os.environ.get("ENABLE_TORCH_COMPILE", "0")}
    if do_train:
        try:
            step, wrap, R = _reference_trainer(dev, TRAIN_BATCH, T_LAT, with_encoder=True)
            out["attention"] = R.attention_backend
            g = torch.Generator().manual_seed(42 + rank)
            audio = (torch.randn(TRAIN_BATCH, 2, T_AUDIO, generator=g).clamp(-1, 1) * 0.5).to(dev)
            ddp = "none"
            if world > 1:
                from torch.nn.parallel import DistributedDataParallel as DDP
                plain = wrap.diffusion.model
                try:
                    wrap.diffusion.model = DDP(plain, device_ids=[dev.index], find_unused_parameters=True)   # train.py:147 default strategy
                    step(audio)
                    ddp = "torch DDP (ddp_find_unused_parameters_true, train.py's default multi-GPU strategy)"
                except Exception as ex:
                    # DDP's reducer and the reference's re-entrant per-layer checkpoint can disagree; fall back to one flat all-reduce
                    wrap.diffusion.model = plain
                    wrap.trainer.optimizers[0].zero_grad(set_to_none=True)
                    params = [p for p in plain.parameters() if p.requires_grad]

                    def allreduce():
                        gs = [p.grad for p in params if p.grad is not None]
                        flat = torch.cat([g_.reshape(-1) for g_ in gs])
                        dist.all_reduce(flat)
                        flat.div_(world)
                        off = 0
                        for g_ in gs:
                            g_.copy_(flat[off:off + g_.numel()].view_as(g_)); off += g_.numel()
                    step.hooks["post_backward"] = allreduce
                    ddp = "flat all-reduce after backward (torch DDP failed: %s)" % repr(ex)[:100]
            for _ in range(2):
                step(audio)
            k = 3
            ms = _timed(lambda: step(audio), k, barrier, dev, dist, world) / k
            out["train"] = {"tokens_per_s": TRAIN_BATCH * T_LAT * world / (ms * 1e-3), "ms_per_step": ms, "steps": k, "ddp": ddp,
                            "what": "DiffusionCondTrainingWrapper.training_step (pretransform.encode iterate_batch + DiT with its checkpointing) "
                                    "+ backward + torch AdamW + EMA, bf16 autocast, same batch/shape"}
            lat = torch.randn(TRAIN_BATCH, 64, T_LAT, device=dev)
            wrap.pre_encoded = True
            step.hooks["meta"] = [{"prompt": 0, "seconds_start": 0.0, "seconds_total": 47.0, "padding_mask": torch.ones(T_LAT, dtype=torch.bool)}
                                  for _ in range(TRAIN_BATCH)]
            step(lat)
            ms = _timed(lambda: step(lat), k, barrier, dev, dist, world) / k
            out["train_pre_encoded"] = {"tokens_per_s": TRAIN_BATCH * T_LAT * world / (ms * 1e-3), "ms_per_step": ms}
            del step, wrap, audio, lat
        except Exception as ex:
            out["train"] = {"error": repr(ex)[:300]}
        torch.cuda.empty_cache()
    if do_sample:
        try:
            from baseline import ref_loader, ref_models
            R = ref_loader.load()
            out["attention"] = R.attention_backend
            with torch.device(dev):
                torch.manual_seed(0)
                model = R.factory.create_model_from_config(ref_models.sao_config(pretransform=False))
            ref_models.rerandomize_zero_init(model.model, seed=1)
            model = model.to(device=dev, dtype=torch.bfloat16).eval().requires_grad_(False)
            ct = ref_models.conditioning_tensors(model, 1, device=dev, seed=3)
            ct = {k_: (v[0].to(torch.bfloat16), v[1]) for k_, v in ct.items()}

            def one():
                return R.generation.generate_diffusion_cond(model, steps=SAMPLE_STEPS, cfg_scale=CFG_SCALE, conditioning_tensors=ct, batch_size=1,
                                                            sample_size=T_LAT, seed=1 + rank, device=dev, sampler_type="dpmpp-3m-sde",
                                                            sigma_min=SIGMA_MIN, sigma_max=SIGMA_MAX, rho=RHO, return_latents=True)
            devnull = open(os.devnull, "w")
            saved = sys.stderr
            sys.stderr = devnull                      # tqdm bars of the reference loop
            try:
                one()
                k = 2
                ms = _timed(one, k, barrier, dev, dist, world) / k
            finally:
                sys.stderr = saved
            out["sample"] = {"latent_steps_per_s": SAMPLE_STEPS * world / (ms * 1e-3), "seconds_per_sample": ms / 1e3, "samples_timed": k,
                             "what": "reference generate_diffusion_cond -> sample_k(dpmpp-3m-sde) with the restated k-diffusion loop (k-diffusion is not "
                                     "installable here; i.i.d. noise instead of the Brownian tree), eager, bf16 weights"}
            del model
        except Exception as ex:
            out["sample"] = {"error": repr(ex)[:300]}
        torch.cuda.empty_cache()
    return out


def _ncu_dram_bytes(name):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed ncu --set full summary (profiles/)."""
    path = os.path.join(ROOT, "profiles", name)
    try:
        vals = {}
        for ln in open(path):
            f = ln.split()
            if len(f) >= 2 and f[0] in ("dram_read_MB", "dram_write_MB"):
                vals[f[0]] = float(f[1])
        return (vals["dram_read_MB"] + vals["dram_write_MB"]) * 1e6, "profiles/" + name
    except Exception:
        return None, None


def roofline_rows(dev, eng_w, pk):
    """Live microbenchmarks (CUDA events, inputs > L2) of the two kernels that dominate the headline step."""
    from b200sat import ops
    from b200sat.autoencoder import OobleckEngine, _Planes
    rows = []
    # (1) conv1d_tcgen05<128,0,2>, the k7 dilated conv of a ResidualUnit at C = 128, T = 2 097 152: the largest launch class of the encoder
    ae = OobleckEngine(_oobleck_state_dict(dev, torch.Generator(device=dev).manual_seed(3)), precision="bf16", device=dev)
    ru = ae.enc["blocks"][0]["rus"][0]
    x = _Planes(1, T_AUDIO, 128, dev, False); x.hi.normal_()
    h = _Planes(1, T_AUDIO, 128, dev, False)
    fn = lambda: ae._conv(x, ru["c7"], act=h, snake=ru["s1"], dil=1, pad=3)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    fl = 2.0 * T_AUDIO * 128 * 128 * 7
    traffic, traffic_src = _ncu_dram_bytes("r2_ncu_conv_k7_summary.txt")
    rows.append({"kernel": "conv1d_tcgen05<128,0,2> (ResidualUnit k7 conv + SnakeBeta epilogue, C=128, T=2097152, bf16)", "bound": "tensor", "achieved": fl / ms / 1e9,
                 "peak": pk["bf16"], "unit": "TFLOP/s", "frac": fl / ms / 1e9 / pk["bf16"], "avg_launch_ms": ms, "traffic": traffic,
                 "traffic_unit": "bytes per launch (dram read + write, ncu --set full capture of the same launch: %s); algorithmic bytes = %d" % (traffic_src, 2 * T_AUDIO * 128 * 2),
                 "peak_source": pk["src"] + " burst", "share_of_step": "conv1d_tcgen05<128,0,2> = 32 % of the step's kernel time (the frozen encoder; profiles/r2_launches_train_step_after_summary.txt)",
                 "hbm_GBps_algorithmic": 2.0 * T_AUDIO * 128 * 2 / ms / 1e6})
    del ae, x, h
    torch.cuda.empty_cache()
    # (2) gemm_bf16_tcgen05<256,2>: FF1 + SwiGLU at the training batch (M = 8200, N = 12288, K = 1536), 24 distinct weight matrices
    M = TRAIN_BATCH * (T_LAT + 1)
    xin = torch.randn(M, D_MODEL, device=dev).bfloat16()
    outb = torch.empty(M, 4 * D_MODEL, device=dev, dtype=torch.bfloat16)
    ws = eng_w
    for w_, b_ in ws[:3]:
        ops.linear(xin, w_, bias=b_, swiglu=True, out=outb)
    torch.cuda.synchronize()
    e0.record()
    for w_, b_ in ws:
        ops.linear(xin, w_, bias=b_, swiglu=True, out=outb)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / len(ws)
    fl = 2.0 * M * (8 * D_MODEL) * D_MODEL
    g_traffic, g_src = _ncu_dram_bytes("r2_ncu_gemm256_summary.txt")
    rows.append({"kernel": "gemm_bf16_tcgen05<256,2> (FF1 + SwiGLU epilogue, M=8200 N=12288 K=1536)", "bound": "tensor", "achieved": fl / ms / 1e9,
                 "peak": pk["bf16"], "unit": "TFLOP/s", "frac": fl / ms / 1e9 / pk["bf16"], "avg_launch_ms": ms, "traffic": g_traffic, "peak_source": pk["src"] + " burst",
                 "traffic_unit": "bytes per launch (dram read + write; ncu --set full capture of this GEMM inside the training step, where it also writes the "
                                 "pre-activation for the backward: %s); algorithmic bytes of that launch = %d" % (g_src, 2 * (M * D_MODEL + 8 * D_MODEL * D_MODEL + M * 4 * D_MODEL + M * 8 * D_MODEL)),
                 "share_of_step": "gemm_bf16_tcgen05<256,2> is the largest launch class of the step: 35 % of its kernel time, the encoder's conv1d_tcgen05<128,0,2> 32 % "
                                  "(profiles/r2_launches_train_step_after_summary.txt); this row is the class's largest launch, timed live; over its 480 launches per step "
                                  "(forward, data and weight gradients) the class averages 1.15 PF = 0.68 of burst in that launch list"})
    return rows


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
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from b200sat import ops

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    W = max(args.warmup, 3)
    tr = OurTrainer(dev, rank, world)
    for _ in range(W):
        loss = tr.step_resident()
    assert torch.isfinite(loss).all(), "non-finite training loss"
    # ---------------- headline: device-resident inputs (`value`)
    clocks = ClockSampler(local)
    barrier()
    if rank == 0:
        clocks.start()
    n0 = ops.LAUNCHES[0]
    ms_total = _timed(tr.step_resident, args.steps, barrier, dev, dist, world)
    launches = ops.LAUNCHES[0] - n0
    clk = clocks.stop() if rank == 0 else None
    ms_step = ms_total / args.steps
    tokens = TRAIN_BATCH * T_LAT * world
    value = tokens / (ms_step * 1e-3)
    # ---------------- end to end: host buffers in, loss out (`e2e`)
    tr.prefetch()
    tr.step_e2e()
    ms_e2e = _timed(tr.step_e2e, args.steps, barrier, dev, dist, world) / args.steps
    torch.cuda.current_stream().wait_event(tr.ready[tr.flip])
    h2d = tr.h_audio[0].numel() * 4 + tr.h_cross.numel() * 4 + tr.h_glob.numel() * 4 + TRAIN_BATCH * 4
    loss_val = float(tr.h_loss.item())
    # ---------------- the same step on pre-encoded latents (pre_encoded: true, training/diffusion.py:344)
    for _ in range(2):
        tr.step_pre_encoded()
    kpe = max(5, min(args.steps, 20))
    ms_pre = _timed(tr.step_pre_encoded, kpe, barrier, dev, dist, world) / kpe
    pk = peaks()
    flop_dit = 3 * GFLOP_PER_TOKEN * 1e9 * TRAIN_BATCH * (T_LAT + 1)
    flop_step = flop_dit + TRAIN_BATCH * ENC_TFLOP_PER_CLIP * 1e12
    nccl_cfg = {"optimizer_in_backward": tr.red.opt is not None, "nccl_ctas": tr.red.nccl_ctas, "sm_reserve": tr.red.sm_reserve}
    ff1 = [(tr.model._w(i, "ff.ff.0.proj.weight").clone(), tr.model._f32(i, "ff.ff.0.proj.bias").clone()) for i in range(DEPTH)] if rank == 0 and not args.quick else None
    tr.free()
    del tr
    torch.cuda.empty_cache()

    sample, eng = (None, None)
    if not args.no_sample:
        sample, eng = measure_sampling(args, dev, rank, world, dist, barrier)
        del eng
        torch.cuda.empty_cache()
    ref = None
    if not args.no_gpu_reference:
        ref = gpu_reference(args, dev, rank, world, dist, barrier, do_sample=not args.no_sample)
        if ref.get("train", {}).get("tokens_per_s"):
            ref["ours_over_reference_train"] = value / ref["train"]["tokens_per_s"]
        if ref.get("train_pre_encoded", {}).get("tokens_per_s"):
            ref["ours_over_reference_train_pre_encoded"] = tokens / (ms_pre * 1e-3) / ref["train_pre_encoded"]["tokens_per_s"]
        if sample and ref.get("sample", {}).get("latent_steps_per_s"):
            ref["ours_over_reference_sample"] = sample["latent_steps_per_s"] / ref["sample"]["latent_steps_per_s"]

    extras = {}
    if rank == 0 and world == 1 and not args.quick:
        from tools import bench_extras as bx
        for name, fn in (("ae_train", bx.measure_ae_train), ("ae_adversarial", bx.measure_ae_adversarial)):
            try:
                extras[name] = fn(args, dev, rank, world, dist)
            except Exception as ex:   # secondary measurement: never lose the headline line
                extras[name] = {"error": repr(ex)[:300]}
                torch.cuda.empty_cache()
        try:
            extras["other_kernels"] = bx.other_kernels(dev, pk)
        except Exception as ex:
            extras["other_kernels"] = {"error": repr(ex)[:300]}
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    roof_rows = []
    if ff1 is not None:
        try:
            roof_rows = roofline_rows(dev, ff1, pk)
        except Exception as ex:
            roof_rows = [{"error": repr(ex)[:300]}]
    # `roofline` = the largest launch class of the headline step (the DiT GEMMs, 35 % of its kernel time), `roofline_more` the runner-up (encoder convs, 32 %)
    if len(roof_rows) >= 2 and all("kernel" in r for r in roof_rows[:2]):
        roof_rows = [roof_rows[1], roof_rows[0]] + roof_rows[2:]
    roof = roof_rows[0] if roof_rows and "kernel" in roof_rows[0] else {"bound": "tensor", "achieved": None, "peak": pk["bf16"], "unit": "TFLOP/s", "frac": None, "traffic": None}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            threads = host_threads()
            ts, attn = cpu_reference_train(2, threads, budget_s=60.0, warmup=1)
            cpu = {"value": CPU_SAMPLE_LAT / min(ts), "unit": "latent-tokens/s", "cores": threads, "kind": "reference",
                   "sample": f"{len(ts)} full training steps on one clip of {CPU_SAMPLE_LAT} latents (best), reference modules from baseline/_ref, torch fp32, {attn}"}
        except Exception as ex:
            cpu = {"error": repr(ex)[:200]}
    line = {
        "metric": "dit_training_latent_tokens_per_sec", "value": value, "unit": "latent-tokens/s", "n_gpus": world,
        "steps": args.steps, "warmup": W, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic (random-init weights, random audio and conditioning)", "config": workload_config(world),
        "e2e": {"value": tokens / (ms_e2e * 1e-3), "unit": "latent-tokens/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "api": "b200sat OobleckEngine.encode_audio + DiTTrainModel/v_objective_loss + GradAllReducer + FusedAdamWEMA; pinned host audio/conditioning in (double-buffered), loss out"},
        "gpu_launches": launches, "clocks": clk, "loss": loss_val,
        "sample": sample,
        "gpu_reference": ref,
        "train_pre_encoded": {"tokens_per_s": tokens / (ms_pre * 1e-3), "ms_per_step": ms_pre, "frac_of_sustained_peak": flop_dit / (ms_pre * 1e-3) / 1e12 / pk["bf16_sustained"]},
        "whole_step": {"tflop_per_gpu": flop_step / 1e12, "achieved_tflops_per_gpu": flop_step / (ms_step * 1e-3) / 1e12,
                       "frac_of_sustained_peak": flop_step / (ms_step * 1e-3) / 1e12 / pk["bf16_sustained"], "ddp": nccl_cfg},
        "roofline": roof, "roofline_more": roof_rows[1:], "cpu_baseline": cpu,
    }
    line.update(extras)
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-reference", action="store_true")
    ap.add_argument("--no-sample", action="store_true", help="skip the sampling half of the metric")
    ap.add_argument("--quick", action="store_true", help="skip the secondary measurements (AE training steps, kernel rows)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

"""Secondary measurements printed by bench.py at N = 1 (kept out of the headline timed region): the Oobleck generator and
adversarial training steps (BASELINE.json configs[3]) and per-kernel roofline rows (attention, LayerNorm, SnakeBeta backward,
Oobleck conv stacks, MRSTFT).  Every function takes the device and returns JSON-serialisable dicts."""
import json
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
T_LAT, L_CTX, D_MODEL, DEPTH, HEADS = 1024, 130, 1536, 24, 24


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return dict(hbm=j["hbm_gbs"], bf16=j["bf16_tflops"], bf16_sustained=j.get("bf16_tflops_sustained", j["bf16_tflops"]), src="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, src="fallback")


def measure_ae_train(args, dev, rank, world, dist):
    """BASELINE.json configs[3], generator step of the warm-up phase (training/autoencoders.py:436-497 without a discriminator): Oobleck
    encode -> VAE -> decode, MRSTFT sum/difference + left + right + KL, backward, fused AdamW + EMA (AutoencoderTrainingStep).
    16 clips x 65536 samples per GPU."""
    from b200sat.ae_training import AutoencoderTrainingStep
    from b200sat.autoencoder_train import OobleckTrainModel
    B, T = 16, 65536
    g = torch.Generator(device=dev).manual_seed(11)
    model = OobleckTrainModel(_oobleck_state_dict(dev, g), device=dev)
    fft, hop = [2048, 1024, 512, 256, 128, 64, 32], [512, 256, 128, 64, 32, 16, 8]
    loss_config = {"spectral": {"type": "mrstft", "config": {"fft_sizes": fft, "hop_sizes": hop, "win_lengths": fft, "perceptual_weighting": True},
                                "weights": {"mrstft": 1.0}},
                   "time": {"type": "l1", "weights": {"l1": 0.0}}, "bottleneck": {"type": "kl", "weights": {"kl": 1e-4}}}
    opt = {"autoencoder": {"optimizer": {"type": "AdamW", "config": {"betas": [0.8, 0.99], "lr": 1.5e-4, "weight_decay": 1e-3}}}}
    allred = (lambda flat: dist.all_reduce(flat)) if world > 1 else None
    step = AutoencoderTrainingStep(model, None, loss_config=loss_config, optimizer_configs=opt, sample_rate=44100, use_ema=True,
                                   world_size=world, all_reduce=allred)
    gh = torch.Generator().manual_seed(42 + rank)
    h_audio = (torch.randn(B, 2, T, generator=gh).clamp(-1, 1) * 0.5).pin_memory()
    h_loss = torch.zeros(1).pin_memory()

    def one():
        reals = h_audio.to(dev, non_blocking=True)
        loss, _ = step.training_step(reals, vae_noise=torch.randn(B, 64, T // 2048, device=dev, generator=g))
        h_loss.copy_(loss.reshape(1), non_blocking=True)

    for _ in range(3):
        one()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    k = 4
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k):
        one()
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
           "includes": "H2D audio, encoder+VAE+decoder fwd, 4-term MRSTFT + KL, full backward, (all-reduce), fused AdamW + EMA, D2H loss",
           "excludes": "adversarial / feature-matching terms (warm-up phase; see ae_adversarial for the post-warm-up steps)",
           "tflops_per_gpu": flop / (ms_step * 1e-3) / 1e12, "frac_of_sustained_peak": flop / (ms_step * 1e-3) / 1e12 / pk["bf16_sustained"]}
    del model, step
    torch.cuda.empty_cache()
    return out


def measure_ae_adversarial(args, dev, rank, world, dist, B=32):
    """BASELINE.json configs[3] after warm-up (training/autoencoders.py:436-515): alternating discriminator / generator steps of the Oobleck
    autoencoder with the EncodecDiscriminator (hinge + feature matching, weights 0.1 / 5.0), MRSTFT sum/difference + L/R and KL, through
    b200sat.ae_training.AutoencoderTrainingStep (the reference's training_step semantics; fused AdamW + EMA per parameter group).
    32 clips x 65536 samples per GPU; two consecutive D/G rounds are timed after two warm-up rounds."""
    from b200sat.ae_training import AutoencoderTrainingStep
    from b200sat.autoencoder_train import OobleckTrainModel
    from b200sat.discriminator import EncodecDiscriminatorTrain
    from b200sat.init import encodec_disc_state_dict
    T = 65536
    g = torch.Generator(device=dev).manual_seed(21)
    ae = OobleckTrainModel(_oobleck_state_dict(dev, g), device=dev)
    disc = EncodecDiscriminatorTrain(encodec_disc_state_dict(dev, g), device=dev)
    fft, hop = [2048, 1024, 512, 256, 128, 64, 32], [512, 256, 128, 64, 32, 16, 8]
    loss_config = {"discriminator": {"type": "encodec", "config": {"filters": 64, "n_ffts": fft[:5], "hop_lengths": hop[:5], "win_lengths": fft[:5]},
                                     "weights": {"adversarial": 0.1, "feature_matching": 5.0}},
                   "spectral": {"type": "mrstft", "config": {"fft_sizes": fft, "hop_sizes": hop, "win_lengths": fft, "perceptual_weighting": True},
                                "weights": {"mrstft": 1.0}},
                   "time": {"type": "l1", "weights": {"l1": 0.0}}, "bottleneck": {"type": "kl", "weights": {"kl": 1e-4}}}
    opt = {"autoencoder": {"optimizer": {"type": "AdamW", "config": {"betas": [0.8, 0.99], "lr": 1.5e-4, "weight_decay": 1e-3}}},
           "discriminator": {"optimizer": {"type": "AdamW", "config": {"betas": [0.8, 0.99], "lr": 3e-4, "weight_decay": 1e-3}}}}
    allred = (lambda flat: dist.all_reduce(flat)) if world > 1 else None
    step = AutoencoderTrainingStep(ae, disc, loss_config=loss_config, optimizer_configs=opt, sample_rate=44100, warmup_steps=0, use_ema=True,
                                   world_size=world, all_reduce=allred)
    gh = torch.Generator().manual_seed(42 + rank)
    reals = (torch.randn(B, 2, T, generator=gh).clamp(-1, 1) * 0.5).to(dev)

    def one():
        return step.training_step(reals, vae_noise=torch.randn(B, 64, T // 2048, device=dev, generator=g))

    for _ in range(4):      # two warm-up rounds (G, D, G, D): the caching allocator sees both steps' buffer sizes in both orders
        one()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    logs = []
    ev[0].record()
    for i in range(4):
        logs.append(one())
        ev[i + 1].record()
    torch.cuda.synchronize()
    t_g = 0.5 * (ev[0].elapsed_time(ev[1]) + ev[2].elapsed_time(ev[3]))     # global_step even: generator
    t_d = 0.5 * (ev[1].elapsed_time(ev[2]) + ev[3].elapsed_time(ev[4]))
    ms = torch.tensor([(t_d + t_g) / 2], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_step = ms.item()
    flop = B * 2.86e12      # mean of the G step (2.44 TFLOP/item) and the D step (3.27 TFLOP/item), minimal graphs
    out = {"metric": "oobleck_adversarial_step_items_per_sec", "value": B * world / (ms_step * 1e-3), "unit": "clips/s", "ms_per_step": ms_step,
           "d_step_ms": t_d, "g_step_ms": t_g, "batch_per_gpu": B, "samples_per_clip": T, "gen_loss": float(logs[2][0]), "dis_loss": float(logs[3][0]),
           "includes": "one generator step and one discriminator step (mean) through AutoencoderTrainingStep: fused AdamW (+ EMA on the generator) per group, (all-reduce)",
           "tflops_per_gpu": flop / (ms_step * 1e-3) / 1e12, "frac_of_sustained_peak": flop / (ms_step * 1e-3) / 1e12 / peaks()["bf16_sustained"]}
    del ae, disc, step
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
    fp32_peak_tflops = 148 * 128 * 2 * 1.9e9 / 1e12      # 148 SMs x 128 FMA lanes x 2 flop x ~1.9 GHz
    out.append({"kernel": "MRSTFT (FIR + 7-resolution Stockham STFT + loss sums), 8 x 2 x 65536, all four generator-loss terms", "bound": "fp32 SIMT / shared memory",
                "achieved": gf / ms, "peak": fp32_peak_tflops, "unit": "TFLOP/s (fp32, algorithmic 0.55 GFLOP per item)", "frac": gf / ms / fp32_peak_tflops, "ms": ms,
                "hbm_GBps_algorithmic": 8 * 2 * 2 * 65536 * 4 / ms / 1e6, "reference_materialised_traffic_gbs": 8 * 120e6 / ms / 1e6,
                "note": "fused: 8.4 MB of waveforms in, 84 scalars out: the algorithmic HBM traffic is negligible, so the row is stated against the fp32 "
                        "FMA peak (nominal, 148 SMs x 128 lanes x 2 x 1.9 GHz); the reference moves ~120 MB per item through HBM for the same result "
                        "(reference_materialised_traffic_gbs = that traffic divided by our time)"})
    return out


def _oobleck_state_dict(dev, g):
    from b200sat.init import oobleck_state_dict
    return oobleck_state_dict(dev, g)



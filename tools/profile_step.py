"""One step of a workload between cudaProfilerStart/Stop, after untimed warm-up, for
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/x.csv python tools/profile_step.py train
Workloads: train (configs[2] step, frozen encoder + DiT fwd/bwd + optimizer), train_pre (pre-encoded), sample (3 denoising steps, eager),
ae (one 47 s encode + decode, bf16)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stable-audio-tools_b200"))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "train"
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    if what in ("train", "train_pre"):
        tr = bench.OurTrainer(dev, 0, 1)
        fn = tr.step_resident if what == "train" else tr.step_pre_encoded
        warm = 2
    elif what == "sample":
        from b200sat import init, sampling
        from b200sat.dit_engine import DiTEngine
        eng = DiTEngine(init.dit_state_dict(seed=0, device=dev), device=dev)
        g = torch.Generator(device=dev).manual_seed(1)
        noise = torch.randn(1, 64, 1024, device=dev, generator=g); cross = torch.randn(1, 130, 768, device=dev, generator=g)
        glob = torch.randn(1, 1536, device=dev, generator=g)
        fn = lambda: sampling.sample_k_dpmpp_3m_sde(eng, noise, 3, 0.03, 1000.0, 1.0, cross, glob, 7.0, 0.0, use_graph=False)
        warm = 1
    elif what == "ae_train":
        # one generator + one discriminator step of the assembled autoencoder training step, argv[2] (default 4) clips x 65536 samples
        from b200sat.ae_training import AutoencoderTrainingStep
        from b200sat.autoencoder_train import OobleckTrainModel
        from b200sat.discriminator import EncodecDiscriminatorTrain
        from b200sat.init import encodec_disc_state_dict
        g = torch.Generator(device=dev).manual_seed(21)
        ae = OobleckTrainModel(bench._oobleck_state_dict(dev, g), device=dev)
        disc = EncodecDiscriminatorTrain(encodec_disc_state_dict(dev, g), device=dev)
        fft, hop = [2048, 1024, 512, 256, 128, 64, 32], [512, 256, 128, 64, 32, 16, 8]
        lc = {"discriminator": {"type": "encodec", "config": {}, "weights": {"adversarial": 0.1, "feature_matching": 5.0}},
              "spectral": {"type": "mrstft", "config": {"fft_sizes": fft, "hop_sizes": hop, "win_lengths": fft, "perceptual_weighting": True}, "weights": {"mrstft": 1.0}},
              "time": {"type": "l1", "weights": {"l1": 0.0}}, "bottleneck": {"type": "kl", "weights": {"kl": 1e-4}}}
        oc = {"autoencoder": {"optimizer": {"type": "AdamW", "config": {"betas": [0.8, 0.99], "lr": 1.5e-4, "weight_decay": 1e-3}}},
              "discriminator": {"optimizer": {"type": "AdamW", "config": {"betas": [0.8, 0.99], "lr": 3e-4, "weight_decay": 1e-3}}}}
        st = AutoencoderTrainingStep(ae, disc, loss_config=lc, optimizer_configs=oc, use_ema=True)
        reals = torch.randn(int(sys.argv[2]) if len(sys.argv) > 2 else 4, 2, 65536, device=dev, generator=g).clamp(-1, 1) * 0.5

        def fn():
            st.training_step(reals); st.training_step(reals)
        warm = 1
    elif what == "ae":
        from b200sat.autoencoder import OobleckEngine
        ae = OobleckEngine(bench._oobleck_state_dict(dev, torch.Generator(device=dev).manual_seed(3)), precision="bf16", device=dev)
        a = torch.randn(1, 2, bench.T_AUDIO, device=dev) * 0.3
        fn = lambda: ae.decode(ae.encode(a))
        warm = 1
    else:
        raise SystemExit("unknown workload " + what)
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()

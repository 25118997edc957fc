"""Oobleck generator training step (warm-up phase: MRSTFT sum/diff + L/R + KL, no discriminator) at stable_audio_2_0_vae shapes.
Usage: python tools/ae_train_bench.py [B] [steps] [profile]   (T = 65536 samples per clip; BASELINE config 4 uses 32 per GPU)"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stable-audio-tools_b200"))
import torch
from b200sat import ops
from b200sat.init import oobleck_state_dict
from b200sat.autoencoder_train import OobleckTrainModel, generator_loss
from b200sat.stft_loss import SumAndDifferenceSTFTLoss

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
profile = len(sys.argv) > 3
dev = "cuda"
FFT, HOP = [2048, 1024, 512, 256, 128, 64, 32], [512, 256, 128, 64, 32, 16, 8]
g = torch.Generator(device=dev).manual_seed(1)
model = OobleckTrainModel(oobleck_state_dict(dev, g))
opt = torch.optim.AdamW(model.parameters(), lr=1.5e-4, betas=(0.8, 0.99), fused=True)
loss_sd = SumAndDifferenceSTFTLoss(fft_sizes=FFT, hop_sizes=HOP, win_lengths=FFT, perceptual_weighting=True, sample_rate=44100)
T = 65536
ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
res = []
for it in range(steps + (0 if profile else 2)):
    reals = (torch.randn(B, 2, T, device=dev, generator=g).clamp(-1, 1) * 0.5)
    noise = torch.randn(B, 64, T // 2048, device=dev, generator=g)
    n0 = ops.LAUNCHES[0]
    torch.cuda.synchronize()
    ev[0].record()
    decoded, kl, _ = model(reals, noise)
    ev[1].record()
    from b200sat.stft_loss import autoencoder_mrstft_terms
    sd_, l_, r_ = autoencoder_mrstft_terms(loss_sd, decoded, reals)
    loss = sd_ + 0.5 * l_ + 0.5 * r_ + 1e-4 * kl
    ev[2].record()
    opt.zero_grad(set_to_none=True)
    loss.backward()
    ev[3].record()
    opt.step()
    ev[4].record()
    torch.cuda.synchronize()
    t = [ev[i].elapsed_time(ev[i + 1]) for i in range(4)]
    res.append(dict(fwd=t[0], loss=t[1], bwd=t[2], opt=t[3], total=sum(t), loss_val=float(loss), launches=ops.LAUNCHES[0] - n0,
                    mem_gb=torch.cuda.max_memory_allocated() / 2**30))
    print(json.dumps(res[-1]), flush=True)
last = res[-1]
flop = B * 3 * 322.7e9
print(json.dumps(dict(B=B, T=T, ms_per_step=last["total"], items_per_s=B / last["total"] * 1e3, algorithmic_tflops=flop / last["total"] / 1e9,
                      note="AE fwd+bwd = 3 x 322.7 GFLOP per item (SURVEY 8d config 4, generator step without discriminator)")))

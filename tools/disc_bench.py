"""EncodecDiscriminator step timings at stable_audio_2_0_vae shapes: python tools/disc_bench.py [B]  (T = 65536)"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stable-audio-tools_b200"))
import torch
from b200sat import ops
from b200sat.init import encodec_disc_state_dict
from b200sat.discriminator import EncodecDiscriminatorTrain
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
model = EncodecDiscriminatorTrain(encodec_disc_state_dict(dev, g))
reals = torch.randn(B, 2, 65536, device=dev, generator=g).clamp(-1, 1) * 0.5
fakes = (reals + 0.1 * torch.randn(B, 2, 65536, device=dev, generator=g)).requires_grad_(True)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
for it in range(3):
    n0 = ops.LAUNCHES[0]
    torch.cuda.synchronize(); ev[0].record()
    adv, fm = model.generator_terms(reals, fakes)
    (0.1 * adv + 5.0 * fm).backward()
    ev[1].record()
    dis = model.discriminator_loss(reals, fakes.detach())
    dis.backward()
    ev[2].record(); torch.cuda.synchronize()
    fl_g = B * 3 * 491.5e9      # D(reals) fwd + D(fakes) fwd + dgrad
    fl_d = B * 3 * 983e9        # fwd + dgrad + wgrad on both
    tg, td = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
    print(json.dumps(dict(B=B, g_terms_ms=tg, g_tflops=fl_g / tg / 1e9, d_step_ms=td, d_tflops=fl_d / td / 1e9, launches=ops.LAUNCHES[0] - n0,
                          adv=float(adv), fm=float(fm), dis=float(dis), mem_gb=torch.cuda.max_memory_allocated() / 2**30)), flush=True)

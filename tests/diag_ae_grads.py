"""Per-parameter gradient agreement of the Oobleck training pass vs oracle autograd (diagnostic; prints in network order)."""
import math, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "stable-audio-tools_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from test_autoencoder_train_gpu import _oracle_step, _cos, _rel
from oracle import oobleck as oo
from b200sat.autoencoder_train import OobleckTrainModel

strides = (2, 4, 4, 8, 8)
sd = oo.make_state_dict(channels=64, strides=strides, seed=5)
g = torch.Generator().manual_seed(6)
B, T = 2, 4096
x = torch.randn(B, 2, T, generator=g) * 0.5
noise = torch.randn(B, 64, T // 2048, generator=g)
wy = torch.randn(B, 2, T, generator=g) / math.sqrt(T)
kl_w = float(sys.argv[1]) if len(sys.argv) > 1 else 0.05
y_ref, kl_ref, g_ref = _oracle_step(sd, x, noise, wy, kl_w, strides)
model = OobleckTrainModel(sd, strides=strides)
y, kl, _ = model(x.cuda(), noise.cuda())
((y * wy.cuda()).sum() + kl_w * kl).backward()
torch.cuda.synchronize()
print("fwd rel", _rel(y.detach().cpu(), y_ref), "kl", kl.item(), kl_ref.item())
for name in model.names:
    got = getattr(model, name.replace(".", "__")).grad.cpu().view(-1)
    ref = g_ref[name].view(-1)
    print(f"{name:55s} cos {_cos(got, ref):.5f} rel {_rel(got, ref):.4f} |ref| {ref.norm():.3e}")

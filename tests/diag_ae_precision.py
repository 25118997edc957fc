"""Where does the Oobleck fp32x3 error come from?  ours vs fp64 oracle, next to fp32 oracle vs fp64 oracle."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "stable-audio-tools_b200"))
from oracle import oobleck as oo
from b200sat.autoencoder import OobleckEngine
rms = lambda a: a.double().pow(2).mean().sqrt().item()
sd = oo.make_state_dict(seed=0)
sd64 = {k: v.double() for k, v in sd.items()}
g = torch.Generator().manual_seed(0)
x = torch.randn(2, 2, 65536, generator=g) * 0.5
nz = torch.randn(2, 64, 32, generator=g)
with torch.no_grad():
    e64 = oo.oobleck_encode(x.double(), sd64); e32 = oo.oobleck_encode(x, sd)
    lat64, _ = oo.vae_sample(e64, nz.double())
    d64 = oo.oobleck_decode(lat64, sd64); d32 = oo.oobleck_decode(lat64.float(), sd)
for prec in ("fp32x3", "bf16"):
    eng = OobleckEngine(sd, precision=prec)
    z, info = eng.encode(x.cuda(), noise=nz.cuda(), return_info=True)
    y = eng.decode(lat64.float().cuda()).cpu()
    print(prec, "enc: ours-vs-f64 %.3e  f32-vs-f64 %.3e | dec: ours-vs-f64 %.3e  f32-vs-f64 %.3e  (abs RMS err %.3e, signal RMS %.3f)" % (
        rms(info["mean_scale"].cpu().double() - e64) / rms(e64), rms(e32.double() - e64) / rms(e64),
        rms(y.double() - d64) / rms(d64), rms(d32.double() - d64) / rms(d64), rms(y.double() - d64), rms(d64)))

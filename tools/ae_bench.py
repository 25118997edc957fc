"""Oobleck encode/decode timings at BASELINE shapes (47 s clip = 1024 latents; 65536-sample clips)."""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stable-audio-tools_b200"))
from b200sat.init import oobleck_state_dict
from b200sat.autoencoder import OobleckEngine
sd = oobleck_state_dict("cuda", torch.Generator(device="cuda").manual_seed(0))
GF_PER_SAMPLE = 161.3e9 / 65536  # encoder (and decoder) forward flop per stereo sample
for prec in ("bf16", "fp32x3"):
    eng = OobleckEngine(sd, precision=prec)
    for (B, T) in ((8, 65536), (1, 2097152)):
        x = torch.randn(B, 2, T, device="cuda") * 0.3
        z = eng.encode(x); y = eng.decode(z); torch.cuda.synchronize()
        t0 = time.time(); z = eng.encode(x); torch.cuda.synchronize(); te = time.time() - t0
        t0 = time.time(); y = eng.decode(z); torch.cuda.synchronize(); td = time.time() - t0
        fl = GF_PER_SAMPLE * B * T
        print(f"{prec:7s} B={B} T={T}: encode {te*1e3:8.1f} ms ({fl/te/1e12:6.1f} TF/s algorithmic)  decode {td*1e3:8.1f} ms ({fl/td/1e12:6.1f} TF/s)  finite={torch.isfinite(y).all().item()}", flush=True)
        del x, z, y; torch.cuda.empty_cache()

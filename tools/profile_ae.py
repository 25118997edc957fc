import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stable-audio-tools_b200"))
from b200sat.init import oobleck_state_dict
from b200sat.autoencoder import OobleckEngine
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
eng = OobleckEngine(oobleck_state_dict("cuda", torch.Generator(device="cuda").manual_seed(0)), precision=prec)
x = torch.randn(1, 2, 2097152, device="cuda") * 0.3
z = eng.encode(x); torch.cuda.synchronize()
z = eng.encode(x); y = eng.decode(z); torch.cuda.synchronize(); print("ok")

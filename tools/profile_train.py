"""Small driver for ncu / timing: a few DiT training steps (bench configs[2] shapes)."""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stable-audio-tools_b200"))
from b200sat import init
from b200sat.dit_train import DiTTrainModel, v_objective_loss
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
depth = int(sys.argv[3]) if len(sys.argv) > 3 else 24
model = DiTTrainModel(init.dit_state_dict(depth=depth, dtype=torch.float32))
lat = torch.randn(B, 64, 1024, device="cuda"); c = torch.randn(B, 130, 768, device="cuda"); g = torch.randn(B, 1536, device="cuda")
for i in range(steps):
    torch.cuda.synchronize(); t0 = time.time()
    model.zero_grad()
    loss = v_objective_loss(model, lat, torch.randn_like(lat), torch.rand(B, device="cuda"), c, g)
    loss.backward()
    torch.cuda.synchronize(); print("step %d: %.1f ms  loss %.4f" % (i, (time.time() - t0) * 1e3, loss.item()), flush=True)

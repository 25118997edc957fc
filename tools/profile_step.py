"""Small driver for ncu: a few eager (non-graph) denoising steps of the bench workload (configs[1])."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stable-audio-tools_b200"))
from b200sat import init, sampling as bs
from b200sat.dit_engine import DiTEngine
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
eng = DiTEngine(init.dit_state_dict())
noise = torch.randn(1, 64, 1024, device="cuda"); c = torch.randn(1, 130, 768, device="cuda"); g = torch.randn(1, 1536, device="cuda")
out = bs.sample_k_dpmpp_3m_sde(eng, noise, steps=steps, cross_attn_cond=c, global_embed=g, cfg_scale=7.0, use_graph=False)
torch.cuda.synchronize()
print("ok", torch.isfinite(out).all().item())

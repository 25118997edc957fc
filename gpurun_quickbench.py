import sys, time, torch
sys.path.insert(0, 'stable-audio-tools_b200')
from b200sat.dit_engine import DiTEngine
from b200sat import sampling as bs, init
sd = init.dit_state_dict()
eng = DiTEngine(sd)
B, T, L = 1, 1024, 130
noise = torch.randn(B, 64, T, device='cuda'); c = torch.randn(B, L, 768, device='cuda'); g = torch.randn(B, 1536, device='cuda')
smp = bs.GraphSampler(eng, B, 64, T, L, True, 7.0, 0.0, True)
for i in range(3):
    torch.cuda.synchronize(); t0 = time.time()
    out = bs.sample_k_dpmpp_3m_sde(eng, noise, steps=100, cross_attn_cond=c, global_embed=g, cfg_scale=7.0, sampler=smp)
    torch.cuda.synchronize(); print('100 steps: %.3f s' % (time.time() - t0), torch.isfinite(out).all().item(), out.std().item())

import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stable-audio-tools_b200"))
from b200sat import ops
B, N, H = 4, 1025, 24
qkv = torch.randn(B, N, 3, H, 64, device="cuda").bfloat16()
out = torch.empty(B, N, H, 64, device="cuda", dtype=torch.bfloat16); lse = torch.empty(B, H, N, device="cuda")
do = torch.randn_like(out); dqkv = torch.empty_like(qkv)
for _ in range(2):
    ops.attention(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], out=out, lse=lse)
    ops.attention_bwd(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], out, do, lse, dqkv[:, :, 0], dqkv[:, :, 1], dqkv[:, :, 2])
torch.cuda.synchronize(); print("ok")

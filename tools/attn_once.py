"""One launch of each attention kernel at the training shape (B=8, H=24, N=1025) for `ncu --set full -k regex:att`."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stable-audio-tools_b200"))
import torch  # noqa: E402
from b200sat import ops  # noqa: E402
B, N, H = 8, 1025, 24
qkv = torch.randn(B, N, 3, H, 64, device="cuda").bfloat16()
o = torch.empty(B, N, H, 64, device="cuda", dtype=torch.bfloat16)
lse = torch.empty(B, H, N, device="cuda")
do = torch.randn(B, N, H, 64, device="cuda").bfloat16()
dqkv = torch.empty_like(qkv)
ops.attention(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], out=o, lse=lse)
ops.attention_bwd(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], o, do, lse, dqkv[:, :, 0], dqkv[:, :, 1], dqkv[:, :, 2])
torch.cuda.synchronize()

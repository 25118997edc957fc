"""One conv1d_fwd launch vs torch (debug helper): python tools/conv_unit.py C T K dil mode stride"""
import os, sys, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stable-audio-tools_b200"))
import torch, torch.nn.functional as F
from b200sat._lib import lib, check
C, T, K, dil, mode, s = [int(a) for a in sys.argv[1:7]] if len(sys.argv) > 6 else (128, 1024, 7, 3, 0, 1)
Cout = int(sys.argv[7]) if len(sys.argv) > 7 else C
B = 2
g = torch.Generator().manual_seed(0)
x = torch.randn(B, C, T, generator=g).bfloat16().float()
w = (torch.randn(Cout, C, K, generator=g) / math.sqrt(C * K)).bfloat16().float()
bias = torch.randn(Cout, generator=g)
pad = dil * (K - 1) // 2 if mode == 0 else math.ceil(s / 2)
ref = F.conv1d(x, w, bias, stride=s if mode == 1 else 1, padding=pad, dilation=dil)
xp = x.transpose(1, 2).contiguous().bfloat16().cuda()
wp = w.permute(0, 2, 1).reshape(Cout, K * C).contiguous().bfloat16().cuda()
out = torch.zeros(B, ref.shape[-1], Cout, device="cuda", dtype=torch.bfloat16)
bd = bias.cuda()
check(lib().b200sat_conv1d_fwd(xp.data_ptr(), 0, wp.data_ptr(), 0, bd.data_ptr(), 0, 0, out.data_ptr(), 0, 0, 0, 0, 0, B, T, C, Cout, K, dil, pad, s, mode, 1,
                               torch.cuda.current_stream().cuda_stream), "conv")
torch.cuda.synchronize()
got = out.float().cpu().transpose(1, 2)
print("rel err", ((got - ref).norm() / ref.norm()).item())

"""Tile-shape sweep of the tcgen05 GEMM on the DiT shapes (TFLOP/s per forced configuration)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stable-audio-tools_b200"))
from b200sat import ops
shapes = [("qkv", 2050, 4608, 1536), ("out", 2050, 1536, 1536), ("ff2", 2050, 1536, 6144), ("to_kv", 260, 1536, 768),
          ("train_qkv", 8200, 4608, 1536), ("train_out", 8200, 1536, 1536), ("train_ff2", 8200, 1536, 6144)]
cfgs = [0, 2256, 2192, 2128, 256, 128, 64]
for name, M, N, K in shapes:
    x = torch.randn(M, K, device="cuda").bfloat16()
    ws = [torch.randn(N, K, device="cuda").bfloat16() for _ in range(max(2, int(3e8 / (N * K * 2))))]  # cycle > L2
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    line = f"{name:10s} M={M} N={N} K={K}: "
    for c in cfgs:
        try:
            for w in ws[:2]:
                ops.linear(x, w, out=out, force_bn=c)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 3
            e0.record()
            for _ in range(reps):
                for w in ws:
                    ops.linear(x, w, out=out, force_bn=c)
            e1.record(); torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / (reps * len(ws))
            line += f" [{c}: {us:6.1f}us {2.0 * M * N * K / us / 1e6:6.0f}TF]"
        except Exception as ex:
            line += f" [{c}: {type(ex).__name__}]"
    print(line, flush=True)
# SwiGLU FF1
M, Nh, K = 2050, 6144, 1536
for M in (2050, 8200):
    x = torch.randn(M, K, device="cuda").bfloat16()
    ws = [(torch.randn(2 * Nh, K, device="cuda").bfloat16(), torch.randn(2 * Nh, device="cuda")) for _ in range(8)]
    out = torch.empty(M, Nh, device="cuda", dtype=torch.bfloat16)
    for c in (2256, 256):
        for w, b in ws[:2]:
            ops.linear(x, w, bias=b, swiglu=True, out=out, force_bn=c)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            for w, b in ws:
                ops.linear(x, w, bias=b, swiglu=True, out=out, force_bn=c)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 24
        print(f"ff1 swiglu M={M} cfg {c}: {us:6.1f} us {2.0 * M * 2 * Nh * K / us / 1e6:6.0f} TF", flush=True)

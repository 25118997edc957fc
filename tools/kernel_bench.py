"""Per-kernel timings with host overhead removed (each kernel captured 20x in a CUDA graph, replayed, CUDA events)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stable-audio-tools_b200"))
from b200sat import ops

def timeit(fn, reps=20, iters=5):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * iters)

H = 24
for B in (2, 8):
    N = 1025
    qkv = torch.randn(B, N, 3, H, 64, device="cuda").bfloat16()
    out = torch.empty(B, N, H, 64, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(B, H, N, device="cuda")
    us = timeit(lambda: ops.attention(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], out=out, lse=lse))
    fl = 4.0 * B * H * N * N * 64
    print(f"attn fwd self  B={B}: {us:7.1f} us  {fl / us / 1e6:6.0f} TF/s")
    kv = torch.randn(B, 130, 2, 12, 64, device="cuda").bfloat16()
    q = torch.randn(B, N, H, 64, device="cuda").bfloat16()
    us = timeit(lambda: ops.attention(q, kv[:, :, 0], kv[:, :, 1], out=out, lse=lse))
    print(f"attn fwd cross B={B}: {us:7.1f} us  {4.0 * B * H * N * 130 * 64 / us / 1e6:6.0f} TF/s")
    do = torch.randn_like(out); dqkv = torch.empty_like(qkv)
    ops.attention(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], out=out, lse=lse)
    us = timeit(lambda: ops.attention_bwd(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], out, do, lse, dqkv[:, :, 0], dqkv[:, :, 1], dqkv[:, :, 2]), reps=5)
    print(f"attn bwd self  B={B}: {us:7.1f} us  {2.5 * fl / us / 1e6:6.0f} TF/s (2.5x fwd flops)")
M = 2050
for name, N_, K_ in (("out 1536x1536", 1536, 1536), ("ff2 1536x6144", 1536, 6144), ("qkv 4608x1536", 4608, 1536)):
    x = torch.randn(M, K_, device="cuda").bfloat16()
    ws = [torch.randn(N_, K_, device="cuda").bfloat16() for _ in range(24)]
    o = torch.empty(M, N_, device="cuda", dtype=torch.bfloat16)
    for cfg in (0, 2256, 2192, 2128, 256, 128):
        st = {"i": 0}
        def f():
            w = ws[st["i"] % 24]; st["i"] += 1
            ops.linear(x, w, out=o, force_bn=cfg)
        us = timeit(f, reps=24)
        print(f"gemm {name} M={M} cfg {cfg}: {us:7.1f} us {2.0 * M * N_ * K_ / us / 1e6:6.0f} TF/s")
x = torch.randn(M, 1536, device="cuda").bfloat16(); g = torch.randn(1536, device="cuda"); o = torch.empty_like(x)
print(f"layernorm M={M}: {timeit(lambda: ops.layernorm(x, g, out=o)):6.1f} us")

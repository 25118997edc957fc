"""Attention backward at the training shapes (B = 8, N = 1025 self-attention; 130-key GQA cross-attention): CUDA-event timing of
the round-2 (v3) and round-1 (v2) kernels per half (B200SAT_ATTN_BWD_V3 bit 0 = dK/dV, bit 1 = dQ) and the forward kernel."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stable-audio-tools_b200"))
import torch  # noqa: E402
from b200sat import ops  # noqa: E402


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def main():
    dev = "cuda"
    B, N, H, L = 8, 1025, 24, 130
    qkv = torch.randn(B, N, 3, H, 64, device=dev).bfloat16()
    o = torch.empty(B, N, H, 64, device=dev, dtype=torch.bfloat16)
    lse = torch.empty(B, H, N, device=dev)
    do = torch.randn(B, N, H, 64, device=dev).bfloat16()
    dqkv = torch.empty_like(qkv)
    fwd = lambda: ops.attention(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], out=o, lse=lse)
    us_f = timeit(fwd)
    fl = 4.0 * B * H * N * N * 64
    print(f"self fwd  {us_f:8.1f} us  {fl / us_f / 1e6:7.1f} TFLOP/s")
    kv = torch.randn(B, L, 2, 12, 64, device=dev).bfloat16()
    q2 = torch.randn(B, N, H, 64, device=dev).bfloat16()
    o2 = torch.empty_like(q2); lse2 = torch.empty(B, H, N, device=dev)
    dq2 = torch.empty_like(q2); dkv = torch.empty_like(kv)
    ops.attention(q2, kv[:, :, 0], kv[:, :, 1], out=o2, lse=lse2)
    for w in ("0", "1", "2", "3"):
        os.environ["B200SAT_ATTN_BWD_V3"] = w
        us = timeit(lambda: ops.attention_bwd(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], o, do, lse, dqkv[:, :, 0], dqkv[:, :, 1], dqkv[:, :, 2]))
        print(f"self bwd  v3mask={w:2s} {us:8.1f} us  {2.5 * fl / us / 1e6:7.1f} TFLOP/s (algorithmic 2.5x fwd)")
        us = timeit(lambda: ops.attention_bwd(q2, kv[:, :, 0], kv[:, :, 1], o2, do, lse2, dq2, dkv[:, :, 0], dkv[:, :, 1]))
        print(f"cross bwd v3mask={w:2s} {us:8.1f} us")


if __name__ == "__main__":
    main()

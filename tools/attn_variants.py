"""Timing decomposition of the attention kernels: compile csrc/attention.cu (or attention_bwd.cu) with -DAT_VARIANT=<mask> (see the
switch in the source), link each against the other objects of the library, and time one launch shape per variant in its own process.
  python tools/attn_variants.py --build [attention|attention_bwd] 0 1 2 ...   (here: nvcc cross-compiles; the .so files travel under tools/micro/_variants/)
  python tools/attn_variants.py --time  [attention|attention_bwd] 0 1 2 ...   (on the GPU box)
Variants with the low five bits set do NOT compute attention; they exist to attribute time to the softmax phases."""
import os
import subprocess
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "stable-audio-tools_b200")
VDIR = os.path.join(ROOT, "tools", "micro", "_variants")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def build(masks, src="attention", macro="AT_VARIANT"):
    os.makedirs(VDIR, exist_ok=True)
    objs = [os.path.join(PKG, "build", f) for f in sorted(os.listdir(os.path.join(PKG, "build"))) if f.endswith(".o") and f != src + ".o"]
    for k in masks:
        o = os.path.join(VDIR, f"{src}_{k}.o")
        subprocess.check_call([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-I",
                               os.path.join(PKG, "csrc"), f"-D{macro}={k}", "-c", os.path.join(PKG, "csrc", src + ".cu"), "-o", o])
        subprocess.check_call([NVCC, "-shared", "-o", os.path.join(VDIR, f"lib_{src}_{k}.so"), o] + objs +
                              ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"])
        os.remove(o)
        print("built", k, flush=True)


def time_one(k, src="attention"):
    sys.path.insert(0, PKG)
    import torch
    from b200sat import _lib
    _lib.LIB_PATH = os.path.join(VDIR, f"lib_{src}_{k}.so")
    from b200sat import ops
    B, N, H = 8, 1025, 24
    torch.manual_seed(0)
    qkv = torch.randn(B, N, 3, H, 64, device="cuda").bfloat16()
    o = torch.empty(B, N, H, 64, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(B, H, N, device="cuda")
    do = torch.randn(B, N, H, 64, device="cuda").bfloat16()
    dqkv = torch.empty_like(qkv)
    if src == "attention":
        fn = lambda: ops.attention(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], out=o, lse=lse)
    else:
        ops.attention(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], out=o, lse=lse)
        fn = lambda: ops.attention_bwd(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], o, do, lse, dqkv[:, :, 0], dqkv[:, :, 1], dqkv[:, :, 2])
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            fn()
        b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / 20 * 1e3)
    flops = 4.0 * B * H * N * N * 64 * (1 if src == "attention" else 2.5)
    print(f"{src} variant {k:3d}: {min(ts):8.1f} us (median {sorted(ts)[2]:.1f})  {flops / min(ts) / 1e9:.3f} PFLOP/s-equivalent", flush=True)


if __name__ == "__main__" and sys.argv[1] != "--trace":
    mode = sys.argv[1]
    src = "attention"
    rest = sys.argv[2:]
    if rest and rest[0] in ("attention", "attention_bwd"):
        src, rest = rest[0], rest[1:]
    macro = "AT_VARIANT" if src == "attention" else "ATB_VARIANT"
    masks = [int(x) for x in rest]
    if mode == "--build":
        build(masks, src, macro)
    elif mode == "--time":
        for k in masks:
            subprocess.call([sys.executable, os.path.abspath(__file__), "--one", src, str(k)])
    elif mode == "--one":
        time_one(masks[0], src)


def trace(k=128):
    """AT_VARIANT bit 7: print the per-tile timeline (cycles since CTA entry) of the two traced CTAs."""
    sys.path.insert(0, PKG)
    import torch
    from b200sat import _lib
    _lib.LIB_PATH = os.path.join(VDIR, f"lib_attention_{k}.so")
    from b200sat import ops
    B, N, H = 8, 1025, 24
    torch.manual_seed(0)
    qkv = torch.randn(B, N, 3, H, 64, device="cuda").bfloat16()
    o = torch.empty(B, N, H, 64, device="cuda", dtype=torch.bfloat16)
    lse = torch.zeros(B, H, N, device="cuda")
    for _ in range(3):
        ops.attention(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], out=o, lse=lse)
    torch.cuda.synchronize()
    t = lse.view(torch.int32).flatten()[: 2 * 64 * 16].view(2, 64, 16).cpu().to(torch.int64) & 0xFFFFFFFF
    names = ["sm:pre-S", "S-ready", "ld+free", "max", "pv(j-1)", "exp+store", "fenced", "arrived", "mma:top", "S(j+1)-iss", "P-ready", "V-ready", "PV-issued"]
    for r in range(2):
        t0 = int(t[r, 0, 13])
        rel = lambda v: (int(v) - t0) & 0xFFFFFFFF
        print(f"--- CTA {r}: entry 0, softmax loop start {rel(t[r, 0, 14])}, end {rel(t[r, 0, 15])}")
        print("  j " + " ".join(f"{n:>10s}" for n in names))
        for j in range(17):
            print(f"{j:3d} " + " ".join(f"{rel(t[r, j, s]):10d}" for s in range(13)))


if __name__ == "__main__" and sys.argv[1] == "--trace":
    trace(int(sys.argv[2]) if len(sys.argv) > 2 else 128)

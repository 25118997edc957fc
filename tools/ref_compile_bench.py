"""The GPU comparator with ENABLE_TORCH_COMPILE=1 (models/utils.py:45-58 reads it at import, hence a separate process): the
unmodified reference's DiT training step on pre-encoded latents and its 100-step sampling, torch.compile'd blocks, bf16.
    ENABLE_TORCH_COMPILE=1 python tools/ref_compile_bench.py   (prints one JSON line; compile time is reported, not timed)"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "stable-audio-tools_b200"))
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    out = {"ENABLE_TORCH_COMPILE": os.environ.get("ENABLE_TORCH_COMPILE", "0"), "torch": torch.__version__}
    import argparse
    args = argparse.Namespace(steps=3, warmup=1)
    barrier = torch.cuda.synchronize
    t0 = time.time()
    try:
        step, wrap, R = bench._reference_trainer(dev, bench.TRAIN_BATCH, bench.T_LAT, with_encoder=False)
        lat = torch.randn(bench.TRAIN_BATCH, 64, bench.T_LAT, device=dev)
        for _ in range(3):
            step(lat)
        torch.cuda.synchronize()
        out["train_compile_s"] = time.time() - t0
        ms = bench._timed(lambda: step(lat), 5, barrier, dev, None, 1) / 5
        out["train_pre_encoded"] = {"ms_per_step": ms, "tokens_per_s": bench.TRAIN_BATCH * bench.T_LAT / (ms * 1e-3), "attention": R.attention_backend}
        del step, wrap, lat
    except Exception as ex:
        out["train_pre_encoded"] = {"error": repr(ex)[:400]}
    torch.cuda.empty_cache()
    t0 = time.time()
    try:
        ref = bench.gpu_reference(args, dev, 0, 1, None, barrier, do_train=False, do_sample=True)
        out["sample"] = ref.get("sample")
        out["sample_total_s_including_compile"] = time.time() - t0
    except Exception as ex:
        out["sample"] = {"error": repr(ex)[:400]}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()

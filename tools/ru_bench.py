"""ResidualUnit at the encoder's first block (C = 128, T = 2 097 152, B = 1, bf16): the k7 conv, the k1 conv and the fused kernel,
CUDA-event timing over rotating buffers (each plane is 0.5 GB, far beyond L2)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stable-audio-tools_b200"))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from b200sat.autoencoder import OobleckEngine, _Planes  # noqa: E402


def t(fn, reps=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def main():
    dev = torch.device("cuda", 0)
    T = bench.T_AUDIO
    ae = OobleckEngine(bench._oobleck_state_dict(dev, torch.Generator(device=dev).manual_seed(3)), precision="bf16", device=dev)
    blk = ae.enc["blocks"][0]
    raw = _Planes(1, T, 128, dev, False); raw.hi.normal_()
    act = _Planes(1, T, 128, dev, False); act.hi.normal_()
    h = _Planes(1, T, 128, dev, False); y = _Planes(1, T, 128, dev, False); ya = _Planes(1, T, 128, dev, False)
    fl7, fl1 = 2.0 * T * 128 * 128 * 7, 2.0 * T * 128 * 128
    for j, ru in enumerate(blk["rus"]):
        dil = (1, 3, 9)[j]
        us7 = t(lambda: ae._conv(act, ru["c7"], act=h, snake=ru["s1"], dil=dil, pad=3 * dil))
        us1 = t(lambda: ae._conv(h, ru["c1"], out=y, act=ya, snake=blk["snake"], res=raw))
        os.environ["B200SAT_FUSED_RU"] = "1"
        usf = t(lambda: ae._residual_units(raw, act, [ru], blk["snake"]))       # dil index 0 of a 1-element list => dil 1; timing only
        print(f"dil={dil}: k7 {us7:7.1f} us ({fl7 / us7 / 1e6:6.0f} TF/s)  k1 {us1:7.1f} us ({4 * T * 128 * 2 / us1 / 1e3:5.0f} GB/s)  "
              f"k7+k1 {us7 + us1:7.1f} us   fused(dil=1) {usf:7.1f} us ({(fl7 + fl1) / usf / 1e6:6.0f} TF/s, {4 * T * 128 * 2 / usf / 1e3:5.0f} GB/s)")


if __name__ == "__main__":
    main()

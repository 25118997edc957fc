"""Multi-resolution STFT losses on the fused libb200sat kernels — constructor / call surface of the reference's vendored
auraloss (`MultiResolutionSTFTLoss`, `SumAndDifferenceSTFTLoss`; training/losses/auraloss.py:451-615) for the option set
the shipped configs use: hann window, win_length == fft_size, w_sc = w_log_mag = 1, w_lin_mag = w_phs = 0, optional
A-weighting (`perceptual_weighting=True`), reduction 'mean', output 'loss'.

Forward and backward (gradients w.r.t. both arguments) run on the fused kernels; nothing but the waveforms, their filtered copies
and 3 sums per (row, resolution) ever exists in HBM.
"""
import math
import numpy as np
import torch

from ._lib import lib, check
from . import ops


def a_weighting_taps(fs, ntaps=101):
    """FIR A-weighting design exactly as the reference builds it (auraloss.py:118-149): analog prototype -> bilinear ->
    512-point frequency response -> least-squares FIR.  Host-side, once per loss object."""
    import scipy.signal
    f1, f2, f3, f4, a1000 = 20.598997, 107.65265, 737.86223, 12194.217, 1.9997
    nums = [(2 * np.pi * f4) ** 2 * (10 ** (a1000 / 20)), 0, 0, 0, 0]
    dens = np.polymul([1, 4 * np.pi * f4, (2 * np.pi * f4) ** 2], [1, 4 * np.pi * f1, (2 * np.pi * f1) ** 2])
    dens = np.polymul(np.polymul(dens, [1, 2 * np.pi * f3]), [1, 2 * np.pi * f2])
    b, a = scipy.signal.bilinear(nums, dens, fs=fs)
    w, h = scipy.signal.freqz(b, a, worN=512, fs=fs)
    return torch.tensor(scipy.signal.firls(ntaps, w, abs(h), fs=fs).astype("float32"))


class _STFTLossBase:
    def __init__(self, fft_sizes, hop_sizes, win_lengths, window="hann_window", w_sc=1.0, w_log_mag=1.0, w_lin_mag=0.0, w_phs=0.0,
                 sample_rate=None, perceptual_weighting=False, scale=None, scale_invariance=False, eps=1e-8, **kw):
        assert len(fft_sizes) == len(hop_sizes) == len(win_lengths)
        if window != "hann_window" or list(win_lengths) != list(fft_sizes):
            raise NotImplementedError("b200sat STFT loss: hann window with win_length == fft_size only")
        if w_lin_mag or w_phs or scale is not None or scale_invariance or w_sc != 1.0 or w_log_mag != 1.0:
            raise NotImplementedError("b200sat STFT loss: only the spectral-convergence + log-magnitude terms (weights 1, 1)")
        if perceptual_weighting and sample_rate is None:
            raise ValueError("`sample_rate` must be supplied when `perceptual_weighting = True`.")
        self.fft_sizes, self.hop_sizes, self.eps = list(fft_sizes), list(hop_sizes), float(eps)
        self.taps = a_weighting_taps(sample_rate) if perceptual_weighting else torch.ones(1)
        self._dev = {}

    def _tables(self, device):
        key = str(device)
        if key not in self._dev:
            tabs = []
            for n in self.fft_sizes:
                win = torch.hann_window(n, periodic=True, dtype=torch.float64)
                k = torch.arange(n // 2, dtype=torch.float64)
                tw = torch.stack([torch.cos(2 * math.pi * k / n), -torch.sin(2 * math.pi * k / n)], dim=1)
                tabs.append((win.float().to(device).contiguous(), tw.float().to(device).contiguous()))
            self._dev[key] = (self.taps.to(device).contiguous(), tabs)
        return self._dev[key]

    def _prefilter(self, x, mixd, taps):
        B, C, T = x.shape
        R = mixd.shape[0]
        out = torch.empty(B, R, T, device=x.device)
        check(lib().b200sat_stft_prefilter(x.data_ptr(), out.data_ptr(), mixd.data_ptr(), taps.data_ptr(), B, C, T, R, taps.numel(),
                                           torch.cuda.current_stream().cuda_stream), "stft_prefilter")
        ops.LAUNCHES[0] += 1
        return out

    def _accumulate_filtered(self, xf, yf):
        """xf, yf fp32 [B, R, T] -> acc [n_res, B*R, 3] float64 and per-resolution bin counts."""
        dev = xf.device
        B, R, T = xf.shape
        _, tabs = self._tables(dev)
        st = torch.cuda.current_stream().cuda_stream
        acc = torch.zeros(len(self.fft_sizes), B * R, 3, device=dev, dtype=torch.float64)
        counts = []
        for i, (n, hop) in enumerate(zip(self.fft_sizes, self.hop_sizes)):
            win, tw = tabs[i]
            check(lib().b200sat_stft_loss_accumulate(xf.data_ptr(), yf.data_ptr(), acc[i].data_ptr(), win.data_ptr(), tw.data_ptr(), B * R, T, n,
                                                     hop, self.eps, st), "stft_loss_accumulate")
            ops.LAUNCHES[0] += 1
            counts.append((n // 2 + 1) * (T // hop + 1))
        return acc, counts

    def group_losses(self, input, target, mix, groups):
        """Differentiable per-group losses.  mix [R, C]: rows of the mixing matrix = mono signals derived from the channels;
        groups: list of lists of row indices r; every group is one MultiResolutionSTFTLoss evaluation over (batch x its rows)."""
        if input.shape != target.shape or input.dim() != 3:
            raise ValueError("input and target must be [B, channels, T] with identical shapes")
        return _MRSTFTFn.apply(input, target, self, mix, groups)


class _MRSTFTFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y, obj, mix, groups):
        dev = x.device
        B, C, T = x.shape
        R = mix.shape[0]
        taps, tabs = obj._tables(dev)
        mixd = mix.to(dev, torch.float32).contiguous()
        xf = obj._prefilter(x.detach().float().contiguous(), mixd, taps)
        yf = obj._prefilter(y.detach().float().contiguous(), mixd, taps)
        acc, counts = obj._accumulate_filtered(xf, yf)
        acc4 = acc.view(len(obj.fft_sizes), B, R, 3)
        cnt = torch.tensor(counts, device=dev, dtype=torch.float64)
        out = []
        for rows in groups:
            a = acc4[:, :, rows, :].reshape(len(obj.fft_sizes), B * len(rows), 3)
            sc = torch.sqrt(a[..., 0]) / torch.sqrt(a[..., 1])          # per row ||Y-X||_F / ||Y||_F   (auraloss.py:181)
            lm = a[..., 2].sum(dim=1) / (cnt * a.shape[1])              # mean |log X - log Y|         (auraloss.py:219-223)
            out.append((sc.mean(dim=1) + lm).mean())                   # per resolution, then mean over resolutions (:437-443, :534)
        ctx.obj, ctx.groups, ctx.mixd, ctx.shape = obj, groups, mixd, (B, C, T, R)
        ctx.save_for_backward(xf, yf, acc4, cnt)
        return torch.stack(out).float()

    @staticmethod
    def backward(ctx, g):
        obj, groups, mixd = ctx.obj, ctx.groups, ctx.mixd
        B, C, T, R = ctx.shape
        xf, yf, acc4, cnt = ctx.saved_tensors
        dev = xf.device
        taps, tabs = obj._tables(dev)
        n_res = len(obj.fft_sizes)
        st = torch.cuda.current_stream().cuda_stream
        # per-row weights: d total / d L_group, divided by (n_res * rows in the group)
        wrow = torch.zeros(R, device=dev, dtype=torch.float64)
        for gi, rows in enumerate(groups):
            wrow[rows] = g[gi].double() / (n_res * B * len(rows))
        dxf = torch.zeros_like(xf); dyf = torch.zeros_like(yf)
        for i, (n, hop) in enumerate(zip(obj.fft_sizes, obj.hop_sizes)):
            S1, S2 = acc4[i, :, :, 0], acc4[i, :, :, 1]                  # [B, R]
            cs = wrow[None, :].expand(B, R)
            coef = torch.stack([cs / torch.sqrt(S1 * S2), cs * torch.sqrt(S1) / S2.pow(1.5), cs / cnt[i]], dim=-1).float().contiguous()
            win, tw = tabs[i]
            check(lib().b200sat_stft_loss_backward(xf.data_ptr(), yf.data_ptr(), dxf.data_ptr(), dyf.data_ptr(), coef.data_ptr(), win.data_ptr(),
                                                   tw.data_ptr(), B * R, T, n, hop, obj.eps, st), "stft_loss_backward")
            ops.LAUNCHES[0] += 1
        grads = []
        for need, d in ((ctx.needs_input_grad[0], dxf), (ctx.needs_input_grad[1], dyf)):
            if not need:
                grads.append(None)
                continue
            dx = torch.empty(B, C, T, device=dev)
            check(lib().b200sat_stft_prefilter_backward(d.data_ptr(), dx.data_ptr(), mixd.data_ptr(), taps.data_ptr(), B, C, T, R, taps.numel(), st),
                  "stft_prefilter_backward")
            ops.LAUNCHES[0] += 1
            grads.append(dx)
        return grads[0], grads[1], None, None, None


class MultiResolutionSTFTLoss(_STFTLossBase):
    def __call__(self, input, target):
        C = input.shape[1]
        return self.group_losses(input, target, torch.eye(C), [list(range(C))])[0]

    forward = __call__


class SumAndDifferenceSTFTLoss(_STFTLossBase):
    def __init__(self, fft_sizes, hop_sizes, win_lengths, window="hann_window", w_sum=1.0, w_diff=1.0, output="loss", **kw):
        super().__init__(fft_sizes, hop_sizes, win_lengths, window, **kw)
        self.w_sum, self.w_diff = w_sum, w_diff
        if output != "loss":
            raise NotImplementedError("output='full'")

    def __call__(self, input, target):
        if input.shape[1] != 2:
            raise ValueError(f"Input must be stereo: {input.shape[1]} channel(s).")
        l = self.group_losses(input, target, torch.tensor([[1.0, 1.0], [1.0, -1.0]]), [[0], [1]])
        return (self.w_sum * l[0] + self.w_diff * l[1]) / 2

    forward = __call__


def autoencoder_mrstft_terms(loss_sd, decoded, reals):
    """The four STFT terms of the autoencoder generator loss in ONE pass over the waveforms (training/autoencoders.py:185-194,
    training/losses/losses.py:107-113 argument swap: input = reals, target = decoded): returns (sum/difference, left, right),
    differentiable w.r.t. `decoded`."""
    l = loss_sd.group_losses(reals, decoded, torch.tensor([[1.0, 1.0], [1.0, -1.0], [1.0, 0.0], [0.0, 1.0]]), [[0], [1], [2], [3]])
    return (loss_sd.w_sum * l[0] + loss_sd.w_diff * l[1]) / 2, l[2], l[3]

"""GPU parity: fused multi-resolution STFT loss kernels vs the committed reference outputs (auraloss) and the oracle."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
FFT = [2048, 1024, 512, 256, 128, 64, 32]
HOP = [n // 4 for n in FFT]


def test_mrstft_golden():
    from b200sat.stft_loss import MultiResolutionSTFTLoss, SumAndDifferenceSTFTLoss
    z = np.load(os.path.join(G, "mrstft.npz"))
    x, y = torch.from_numpy(z["x"]).cuda(), torch.from_numpy(z["y"]).cuda()
    sd = SumAndDifferenceSTFTLoss(FFT, HOP, FFT, perceptual_weighting=True, sample_rate=44100)
    mr = MultiResolutionSTFTLoss(FFT, HOP, FFT, perceptual_weighting=True, sample_rate=44100)
    plain = MultiResolutionSTFTLoss(FFT, HOP, FFT)
    assert abs(sd(x, y).item() - float(z["loss_sd"])) <= 2e-5
    assert abs(mr(x[:, :1], y[:, :1]).item() - float(z["loss_l"])) <= 2e-5
    assert abs(plain(x, y).item() - float(z["loss_plain"])) <= 2e-5


def test_mrstft_config4_shape_vs_oracle():
    """4 x 2 x 65536 (the autoencoder training clip length): all four generator-loss STFT terms in one pass."""
    from oracle import stft_loss as ost
    from b200sat.stft_loss import SumAndDifferenceSTFTLoss, autoencoder_mrstft_terms
    g = torch.Generator().manual_seed(0)
    reals = torch.randn(4, 2, 65536, generator=g) * 0.3
    dec = reals + 0.05 * torch.randn(4, 2, 65536, generator=g)
    taps = ost.a_weighting_fir()
    ref_sd = ost.sum_and_difference_loss(reals, dec, FFT, HOP, taps).item()
    ref_l = ost.mrstft_loss(reals[:, :1], dec[:, :1], FFT, HOP, taps).item()
    ref_r = ost.mrstft_loss(reals[:, 1:], dec[:, 1:], FFT, HOP, taps).item()
    loss = SumAndDifferenceSTFTLoss(FFT, HOP, FFT, perceptual_weighting=True, sample_rate=44100)
    sd, l, r = autoencoder_mrstft_terms(loss, dec.cuda(), reals.cuda())
    for got, ref in ((sd, ref_sd), (l, ref_l), (r, ref_r)):
        assert abs(got.item() - ref) <= 2e-5 * max(1.0, abs(ref)), (got.item(), ref)


def test_mrstft_backward_golden_and_oracle():
    """Gradient w.r.t. the input against the reference-generated golden (auraloss autograd), and gradients w.r.t. the decoded
    audio of the 4-term autoencoder loss against oracle autograd."""
    from oracle import stft_loss as ost
    from b200sat.stft_loss import SumAndDifferenceSTFTLoss, autoencoder_mrstft_terms
    z = np.load(os.path.join(G, "mrstft.npz"))
    x = torch.from_numpy(z["x"]).cuda().requires_grad_(True)
    y = torch.from_numpy(z["y"]).cuda()
    sd = SumAndDifferenceSTFTLoss(FFT, HOP, FFT, perceptual_weighting=True, sample_rate=44100)
    loss = sd(x, y)
    loss.backward()
    ref = torch.from_numpy(z["grad_sd"]).cuda()
    rel = ((x.grad - ref).norm() / ref.norm()).item()
    # the fp32 reference gradient itself is 1.6e-3 away from the exact (fp64) one: the log-magnitude term divides by tiny bins.
    # Tolerance: within 3x of the reference's own fp32 error, measured against the fp64 oracle.
    x64 = torch.from_numpy(z["x"]).double().requires_grad_(True)
    ost.sum_and_difference_loss(x64, torch.from_numpy(z["y"]).double(), FFT, HOP, ost.a_weighting_fir().double()).backward()
    ref64 = x64.grad.float().cuda()
    e_ours = ((x.grad - ref64).norm() / ref64.norm()).item()
    e_ref32 = ((ref - ref64).norm() / ref64.norm()).item()
    print("grad rel err vs fp64 oracle: ours", e_ours, " fp32 reference", e_ref32, " ours vs fp32 golden", rel)
    assert abs(loss.item() - float(z["loss_sd"])) <= 2e-5 and e_ours <= 3 * e_ref32 + 1e-3
    # gradient through the target argument (what the autoencoder training uses: input = reals, target = decoded)
    g = torch.Generator().manual_seed(3)
    reals = torch.randn(2, 2, 16384, generator=g) * 0.3
    dec = (reals + 0.05 * torch.randn(2, 2, 16384, generator=g))
    taps = ost.a_weighting_fir()
    dref = dec.clone().requires_grad_(True)
    tot_ref = (1.0 * ost.sum_and_difference_loss(reals, dref, FFT, HOP, taps) + 0.5 * ost.mrstft_loss(reals[:, :1], dref[:, :1], FFT, HOP, taps)
               + 0.5 * ost.mrstft_loss(reals[:, 1:], dref[:, 1:], FFT, HOP, taps))
    tot_ref.backward()
    dgpu = dec.cuda().requires_grad_(True)
    l_sd, l_l, l_r = autoencoder_mrstft_terms(sd, dgpu, reals.cuda())
    tot = 1.0 * l_sd + 0.5 * l_l + 0.5 * l_r
    tot.backward()
    rel = ((dgpu.grad.cpu() - dref.grad).norm() / dref.grad.norm()).item()
    print("4-term generator STFT loss: value", tot.item(), tot_ref.item(), "grad rel", rel)
    assert abs(tot.item() - tot_ref.item()) <= 5e-5 and rel <= 8e-3   # both sides fp32: see the fp64 comparison above

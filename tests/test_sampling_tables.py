"""CPU tests (host logic): the per-step coefficient tables of b200sat.sampling reproduce the oracle samplers, and the
oracle's k-diffusion restatement passes closed-form known-answer checks (its parity is otherwise unpinned)."""
import math
import torch

from oracle import sampling as osamp
from b200sat import sampling as bs


def _toy_model(x, t, **kw):
    # any deterministic function of (x, t) works: the tables must reproduce the update rule, not the network
    return torch.tanh(x * 0.7) * (0.3 + t.view(-1, 1, 1)) - 0.1 * x


def _run_tables(noise, coef, cin, tt, model, noise_seq=None, init_scale=1.0):
    x = noise * init_scale
    hist = [torch.zeros_like(x) for _ in range(3)]
    for s in range(coef.shape[0]):
        c = coef[s]
        v = model(x * cin[s], tt[s].expand(x.shape[0]))
        den = v * c[0] + x * c[1]
        xn = c[2] * x + c[3] * den + c[4] * hist[(s + 2) % 3] + c[5] * hist[(s + 1) % 3]
        if noise_seq is not None:
            xn = xn + c[6] * noise_seq[s]
        hist[s % 3] = den
        x = xn
    return x


def test_sigmas_closed_form():
    s = osamp.get_sigmas_polyexponential(100, 0.03, 1000.0, 1.0)
    assert s.shape == (101,) and s[-1] == 0
    assert abs(s[0].item() - 1000.0) < 1e-2 and abs(s[99].item() - 0.03) < 1e-6
    # rho = 1: geometric sequence
    r = s[1:100] / s[:99]
    assert (r - r[0]).abs().max() < 1e-5
    assert torch.allclose(bs.get_sigmas_polyexponential(100, 0.03, 1000.0, 1.0), s)


def test_vdenoiser_known_answer():
    x = torch.full((1, 2, 3), 2.0)
    sigma = torch.tensor([1.0])
    out = osamp.v_denoiser(lambda xi, t: torch.ones_like(xi) * t.view(-1, 1, 1), x, sigma)
    # c_in = 1/sqrt2, t = atan(1)*2/pi = 0.5, c_out = -1/sqrt2, c_skip = 0.5
    assert torch.allclose(out, torch.full_like(x, 0.5 * (-1 / math.sqrt(2)) + 2.0 * 0.5), atol=1e-6)


def test_dpmpp3m_first_step_known_answer():
    # one step, eta = 0: x1 = exp(-h) x0 + (1 - exp(-h)) den  with h = ln(s0/s1)
    sig = torch.tensor([2.0, 1.0, 0.0])
    x0 = torch.ones(1, 1, 4)
    den_fn = lambda x, s: 3.0 * torch.ones_like(x)
    x = osamp.sample_dpmpp_3m_sde(den_fn, x0, sig[:2].clone(), eta=0.0)  # single interval 2 -> 1
    assert torch.allclose(x, 0.5 * x0 + 0.5 * 3.0, atol=1e-6)


def test_tables_match_oracle_dpmpp3m():
    torch.manual_seed(0)
    steps = 20
    noise = torch.randn(2, 4, 16)
    nseq = torch.randn(steps, 2, 4, 16)
    ref = osamp.sample_k_dpmpp_3m_sde(_toy_model, noise, steps=steps, sigma_min=0.03, sigma_max=1000.0, rho=1.0, noise_seq=nseq)
    sig = bs.get_sigmas_polyexponential(steps, 0.03, 1000.0, 1.0)
    coef, cin, tt = bs.dpmpp_3m_sde_tables(sig, 1.0, 1.0)
    got = _run_tables(noise, coef, cin, tt, _toy_model, nseq, init_scale=float(sig[0]))
    assert (got - ref).abs().max().item() <= 2e-4 * max(1.0, ref.abs().max().item())


def test_tables_match_oracle_vddim():
    torch.manual_seed(1)
    steps = 25
    noise = torch.randn(2, 4, 16)
    ref = osamp.sample_v_ddim(_toy_model, noise, steps)
    coef, cin, tt = bs.v_ddim_tables(steps, 1.0)
    got = _run_tables(noise, coef, cin, tt, _toy_model)
    assert (got - ref).abs().max().item() <= 1e-4

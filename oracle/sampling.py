"""Oracle: samplers (TEST INFRASTRUCTURE — see oracle/__init__.py).

PARITY UNPINNED for the k-diffusion pieces: `get_sigmas_polyexponential`, `VDenoiser` and `sample_dpmpp_3m_sde` live in
the pip dependency k-diffusion==0.1.1 (setup.py:16 of the reference), which is neither vendored in /root/reference nor
installed here.  They are restated from the published algorithm and anchored on the reference's
call site inference/sampling.py:351-387; closed-form known-answer tests live in tests/test_sampling_oracle.py.
The in-repo deterministic v-DDIM `sample` (inference/sampling.py:253-307) IS pinned against the reference.
"""
import math
import torch


def get_sigmas_polyexponential(n, sigma_min, sigma_max, rho=1.0):
    ramp = torch.linspace(1, 0, n) ** rho
    sigmas = torch.exp(ramp * (math.log(sigma_max) - math.log(sigma_min)) + math.log(sigma_min))
    return torch.cat([sigmas, sigmas.new_zeros([1])])


def v_denoiser(inner, x, sigma, **kw):
    """k_diffusion.external.VDenoiser.forward with sigma_data = 1."""
    s = sigma.view(-1, *([1] * (x.ndim - 1)))
    c_skip = 1.0 / (s ** 2 + 1.0)
    c_out = -s / (s ** 2 + 1.0) ** 0.5
    c_in = 1.0 / (s ** 2 + 1.0) ** 0.5
    t = sigma.atan() / math.pi * 2
    return inner(x * c_in, t, **kw) * c_out + x * c_skip


def sample_dpmpp_3m_sde(model, x, sigmas, noise_seq=None, eta=1.0, s_noise=1.0, extra_args=None):
    """k_diffusion.sampling.sample_dpmpp_3m_sde; `model(x, sigma)` is the denoiser.  noise_seq[i] replaces the
    BrownianTreeNoiseSampler draw of step i (unit-variance increments of disjoint intervals are i.i.d. N(0, I))."""
    extra_args = extra_args or {}
    s_in = x.new_ones([x.shape[0]])
    d1 = d2 = None
    h1 = h2 = None
    for i in range(len(sigmas) - 1):
        den = model(x, sigmas[i] * s_in, **extra_args)
        if sigmas[i + 1] == 0:
            x = den
            h = None
        else:
            t, s = -sigmas[i].log(), -sigmas[i + 1].log()
            h = s - t
            h_eta = h * (eta + 1)
            x = torch.exp(-h_eta) * x + (-h_eta).expm1().neg() * den
            if h2 is not None:
                r0, r1 = h1 / h, h2 / h
                d1_0 = (den - d1) / r0
                d1_1 = (d1 - d2) / r1
                dd1 = d1_0 + (d1_0 - d1_1) * r0 / (r0 + r1)
                dd2 = (d1_0 - d1_1) / (r0 + r1)
                phi_2 = h_eta.neg().expm1() / h_eta + 1
                phi_3 = phi_2 / h_eta - 0.5
                x = x + phi_2 * dd1 - phi_3 * dd2
            elif h1 is not None:
                r = h1 / h
                d = (den - d1) / r
                phi_2 = h_eta.neg().expm1() / h_eta + 1
                x = x + phi_2 * d
            if eta:
                nz = noise_seq[i] if noise_seq is not None else torch.randn_like(x)
                x = x + nz * sigmas[i + 1] * (-2 * h * eta).expm1().neg().sqrt() * s_noise
        d1, d2 = den, d1
        h1, h2 = h, h1
    return x


def sample_k_dpmpp_3m_sde(model_fn, noise, steps=100, sigma_min=0.03, sigma_max=1000.0, rho=1.0, noise_seq=None, **extra):
    """inference/sampling.py:351-387 for sampler_type == 'dpmpp-3m-sde' (init_data None)."""
    sigmas = get_sigmas_polyexponential(steps, sigma_min, sigma_max, rho)
    x = noise * sigmas[0]
    den = lambda x_, s_, **kw: v_denoiser(model_fn, x_, s_, **kw)
    return sample_dpmpp_3m_sde(den, x, sigmas, noise_seq=noise_seq, extra_args=extra)


def sample_v_ddim(model, x, steps, eta=0.0, sigma_max=1.0, **extra_args):
    """inference/sampling.py:253-307 (`sample`), cfg_pp False, dist_shift None."""
    ts = x.new_ones([x.shape[0]])
    t = torch.linspace(sigma_max, 0, steps + 1)[:-1]
    alphas, sigmas = torch.cos(t * math.pi / 2), torch.sin(t * math.pi / 2)
    pred = None
    for i in range(steps):
        v = model(x, ts * t[i], **extra_args)
        pred = x * alphas[i] - v * sigmas[i]
        eps = x * sigmas[i] + v * alphas[i]
        if i < steps - 1:
            ddim_sigma = eta * (sigmas[i + 1] ** 2 / sigmas[i] ** 2).sqrt() * (1 - alphas[i] ** 2 / alphas[i + 1] ** 2).sqrt()
            adjusted_sigma = (sigmas[i + 1] ** 2 - ddim_sigma ** 2).sqrt()
            x = pred * alphas[i + 1] + eps * adjusted_sigma
            if eta:
                x = x + torch.randn_like(x) * ddim_sigma
    return pred

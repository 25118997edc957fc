"""Load the UNMODIFIED reference (stable-audio-tools 0.0.19) from `baseline/_ref` for the reference arms of bench.py and for
the drop-in tests.  `baseline/_ref` is produced by

    python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target baseline/_ref /root/reference

(git-ignored, travels to the GPU box with the snapshot).  Nothing in here is product code and nothing in the product imports it.

The reference's pip dependencies that are absent from this image (no index access) are replaced by inert stand-ins so that the
reference's own modules import; none of them is on the measured path:
  alias_free_torch, einops_exts   only touched by options the Stable Audio configs do not enable
  k_diffusion                     the dpmpp samplers (the reference arm drives the real DiffusionTransformer with the restated
                                  k-diffusion update instead, see bench.py)
  pytorch_lightning, ema_pytorch, wandb, auraloss (pip), aeiou viz deps: only needed to import the training wrappers; the
                                  LightningModule stand-in is an nn.Module with `device`, `log_dict`, `trainer`.
"""
import importlib
import importlib.util
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def available():
    return os.path.isdir(os.path.join(REF_DIR, "stable_audio_tools"))


def _have(name):
    try:
        return importlib.util.find_spec(name) is not None
    except (ImportError, ValueError):
        return False


def _stub(name, **attrs):
    if name in sys.modules or _have(name):
        return sys.modules.get(name)
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__b200sat_stub__ = True
    sys.modules[name] = m
    return m


def _kdiffusion_standin():
    """k-diffusion==0.1.1 is a pip dependency that is neither vendored in the reference nor installable here.  So that the
    reference's OWN `sample_k(..., sampler_type="dpmpp-3m-sde")` (inference/sampling.py:331-387) can run for the GPU reference
    arm, the three entry points it calls are restated from the published algorithm (the same restatement
    as oracle/sampling.py, "parity unpinned").  The Brownian-tree noise sampler (torchsde) is replaced by i.i.d. N(0, I)
    increments, which have the same distribution over disjoint intervals and cost the reference arm less than the real tree."""
    if "k_diffusion" in sys.modules or _have("k_diffusion"):
        return
    import math
    import torch

    class VDenoiser(torch.nn.Module):
        def __init__(self, inner_model):
            super().__init__()
            self.inner_model = inner_model
            self.sigma_data = 1.0

        def get_scalings(self, sigma):
            c_skip = self.sigma_data ** 2 / (sigma ** 2 + self.sigma_data ** 2)
            c_out = -sigma * self.sigma_data / (sigma ** 2 + self.sigma_data ** 2) ** 0.5
            c_in = 1 / (sigma ** 2 + self.sigma_data ** 2) ** 0.5
            return c_skip, c_out, c_in

        def sigma_to_t(self, sigma):
            return sigma.atan() / math.pi * 2

        def forward(self, input, sigma, **kwargs):
            c_skip, c_out, c_in = [x[(...,) + (None,) * (input.ndim - x.ndim)] for x in self.get_scalings(sigma)]
            return self.inner_model(input * c_in, self.sigma_to_t(sigma), **kwargs) * c_out + input * c_skip

    def get_sigmas_polyexponential(n, sigma_min, sigma_max, rho=1.0, device="cpu"):
        ramp = torch.linspace(1, 0, n, device=device) ** rho
        sigmas = torch.exp(ramp * (math.log(sigma_max) - math.log(sigma_min)) + math.log(sigma_min))
        return torch.cat([sigmas, sigmas.new_zeros([1])])

    @torch.no_grad()
    def sample_dpmpp_3m_sde(model, x, sigmas, extra_args=None, callback=None, disable=None, eta=1.0, s_noise=1.0, noise_sampler=None):
        extra_args = {} if extra_args is None else extra_args
        if noise_sampler is None:
            noise_sampler = lambda sigma, sigma_next: torch.randn_like(x)
        s_in = x.new_ones([x.shape[0]])
        denoised_1, denoised_2 = None, None
        h_1, h_2 = None, None
        for i in range(len(sigmas) - 1):
            denoised = model(x, sigmas[i] * s_in, **extra_args)
            if callback is not None:
                callback({"x": x, "i": i, "sigma": sigmas[i], "sigma_hat": sigmas[i], "denoised": denoised})
            if sigmas[i + 1] == 0:
                x = denoised
            else:
                t, s = -sigmas[i].log(), -sigmas[i + 1].log()
                h = s - t
                h_eta = h * (eta + 1)
                x = torch.exp(-h_eta) * x + (-h_eta).expm1().neg() * denoised
                if h_2 is not None:
                    r0 = h_1 / h
                    r1 = h_2 / h
                    d1_0 = (denoised - denoised_1) / r0
                    d1_1 = (denoised_1 - denoised_2) / r1
                    d1 = d1_0 + (d1_0 - d1_1) * r0 / (r0 + r1)
                    d2 = (d1_0 - d1_1) / (r0 + r1)
                    phi_2 = h_eta.neg().expm1() / h_eta + 1
                    phi_3 = phi_2 / h_eta - 0.5
                    x = x + phi_2 * d1 - phi_3 * d2
                elif h_1 is not None:
                    r = h_1 / h
                    d = (denoised - denoised_1) / r
                    phi_2 = h_eta.neg().expm1() / h_eta + 1
                    x = x + phi_2 * d
                if eta:
                    x = x + noise_sampler(sigmas[i], sigmas[i + 1]) * sigmas[i + 1] * (-2 * h * eta).expm1().neg().sqrt() * s_noise
            denoised_1, denoised_2 = denoised, denoised_1
            h_1, h_2 = h, h_1
        return x

    K = types.ModuleType("k_diffusion")
    K.__b200sat_stub__ = True
    K.external = types.ModuleType("k_diffusion.external")
    K.external.VDenoiser = VDenoiser
    K.sampling = types.ModuleType("k_diffusion.sampling")
    K.sampling.get_sigmas_polyexponential = get_sigmas_polyexponential
    K.sampling.sample_dpmpp_3m_sde = sample_dpmpp_3m_sde
    sys.modules["k_diffusion"] = K
    sys.modules["k_diffusion.external"] = K.external
    sys.modules["k_diffusion.sampling"] = K.sampling


def install_stubs():
    import torch

    class _Unavailable:
        def __init__(self, *a, **k):
            raise NotImplementedError("dependency of the reference that is not installed in this image")

    _stub("alias_free_torch", Activation1d=_Unavailable)
    _kdiffusion_standin()
    _stub("einops_exts", rearrange_many=lambda *a, **k: None)
    _stub("wandb")
    _stub("webdataset")      # data/dataset.py imports it at module top; only the S3/WebDataset loaders use it
    _stub("audiotools")      # training/losses/semantic.py imports it at module top; only HubertLoss/PESQ-style metrics use it
    _stub("auraloss")
    # pytorch_lightning: the training wrappers subclass pl.LightningModule and use .device / .log_dict / .trainer / .all_gather
    if not _have("pytorch_lightning") and "pytorch_lightning" not in sys.modules:
        class LightningModule(torch.nn.Module):
            def __init__(self, *a, **k):
                super().__init__()
                self.automatic_optimization = True
                self.logged = {}
                self.trainer = None
                self.global_step = 0

            @property
            def device(self):
                return next(self.parameters()).device

            def log_dict(self, d, *a, **k):
                self.logged.update({k_: (v.detach() if torch.is_tensor(v) else v) for k_, v in d.items()})

            def log(self, k_, v, *a, **k):
                self.logged[k_] = v

            def all_gather(self, t):
                return t.unsqueeze(0)

            # manual-optimisation surface used by AutoencoderTrainingWrapper.training_step (training/autoencoders.py:449-515)
            def optimizers(self):
                opts = getattr(self, "_b200sat_optimizers", None)
                if opts is None:
                    cfg = self.configure_optimizers()
                    opts = cfg[0] if isinstance(cfg, tuple) else cfg
                    self._b200sat_optimizers = opts
                return opts[0] if len(opts) == 1 else opts

            def lr_schedulers(self):
                return None

            def manual_backward(self, loss):
                loss.backward()

        class Callback:
            pass

        pl = _stub("pytorch_lightning", LightningModule=LightningModule, Callback=Callback)
        util = _stub("pytorch_lightning.utilities")
        rz = _stub("pytorch_lightning.utilities.rank_zero", rank_zero_only=lambda f: f)
        lg = _stub("pytorch_lightning.loggers", WandbLogger=type("WandbLogger", (), {}), CometLogger=type("CometLogger", (), {}))
        pl.utilities = util; util.rank_zero = rz; pl.loggers = lg
    if not _have("ema_pytorch") and "ema_pytorch" not in sys.modules:
        class EMA(torch.nn.Module):
            """Stand-in with ema_pytorch's constructor signature; update() follows its published decay schedule
            (1 - (1 + step/inv_gamma)^-power clipped to [min_value, beta]) — unpinned, not on the measured path."""
            def __init__(self, model, ema_model=None, beta=0.9999, power=2 / 3, update_every=10, update_after_step=100, inv_gamma=1.0,
                         min_value=0.0, include_online_model=True, **kw):
                super().__init__()
                import copy
                self.beta, self.power, self.inv_gamma, self.min_value = beta, power, inv_gamma, min_value
                self.update_every, self.update_after_step = update_every, update_after_step
                self.online = [model]
                self.ema_model = (copy.deepcopy(model) if ema_model is None else ema_model).requires_grad_(False)
                self.step = 0

            @torch.no_grad()
            def update(self):
                self.step += 1
                if self.step % self.update_every:
                    return
                if self.step <= self.update_after_step:
                    for pe, po in zip(self.ema_model.parameters(), self.online[0].parameters()):
                        pe.copy_(po)
                    return
                ep = max(self.step - self.update_after_step - 1, 0)
                d = 0.0 if ep <= 0 else min(max(1 - (1 + ep / self.inv_gamma) ** -self.power, self.min_value), self.beta)
                for pe, po in zip(self.ema_model.parameters(), self.online[0].parameters()):
                    pe.lerp_(po.to(pe.dtype), 1.0 - d)

            def forward(self, *a, **k):
                return self.ema_model(*a, **k)

        _stub("ema_pytorch", EMA=EMA)


def load(force_sdpa=None):
    """Import the reference package from baseline/_ref.  force_sdpa: True -> disable the flash_attn dispatch (CPU runs);
    None -> disable it only when CUDA is unavailable or a probe call of flash_attn_func fails on this GPU."""
    if not available():
        raise RuntimeError("baseline/_ref is missing: run the pip install recorded in DESIGN.md section 1")
    install_stubs()
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import torch
    import stable_audio_tools.models.transformer as transformer
    if not hasattr(transformer, "_b200sat_flash_attn_func"):
        transformer._b200sat_flash_attn_func = getattr(transformer, "flash_attn_func", None)   # what the reference imported itself
    transformer.flash_attn_func = transformer._b200sat_flash_attn_func                         # every load() decides afresh
    attn = "flash_attn_func" if transformer.flash_attn_func is not None else "sdpa"
    if attn == "flash_attn_func":
        disable = force_sdpa
        if disable is None:
            disable = not torch.cuda.is_available()
            if not disable:
                try:
                    q = torch.randn(1, 128, 2, 64, device="cuda", dtype=torch.bfloat16)
                    transformer.flash_attn_func(q, q, q)
                    torch.cuda.synchronize()
                except Exception:
                    disable = True
        if disable:
            transformer.flash_attn_func = None
            attn = "sdpa"
    ns = types.SimpleNamespace(transformer=transformer, attention_backend=attn)
    for short, name in (("dit", "models.dit"), ("blocks", "models.blocks"), ("bottleneck", "models.bottleneck"),
                        ("autoencoders", "models.autoencoders"), ("factory", "models.factory"), ("diffusion", "models.diffusion"),
                        ("sampling", "inference.sampling"), ("generation", "inference.generation"), ("pretransforms", "models.pretransforms"),
                        ("discriminators", "models.discriminators")):
        setattr(ns, short, importlib.import_module("stable_audio_tools." + name))
    spec = importlib.util.spec_from_file_location("ref_auraloss", os.path.join(REF_DIR, "stable_audio_tools/training/losses/auraloss.py"))
    ns.auraloss = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ns.auraloss)
    return ns


def load_training(force_sdpa="keep"):
    """The Lightning training wrappers (training/diffusion.py, training/autoencoders.py) on the stand-ins above.
    force_sdpa="keep": leave the attention backend as the last load() decided."""
    if force_sdpa != "keep" or "stable_audio_tools.models.transformer" not in sys.modules:
        load(None if force_sdpa == "keep" else force_sdpa)
    # demo-time visualisation helpers (matplotlib, PIL, ...) are only called from the demo callbacks
    viz = lambda *a, **k: None
    if not _have("matplotlib"):
        m = types.ModuleType("stable_audio_tools.interface.aeiou")
        m.pca_point_cloud = m.audio_spectrogram_image = m.tokens_spectrogram_image = viz
        sys.modules["stable_audio_tools.interface.aeiou"] = m
    out = types.SimpleNamespace()
    out.diffusion = importlib.import_module("stable_audio_tools.training.diffusion")
    out.autoencoders = importlib.import_module("stable_audio_tools.training.autoencoders")
    return out

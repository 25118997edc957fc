"""`b200sat.install()` — route the reference's hot-path modules through libb200sat WITHOUT editing the reference.

After `install()`, these reference call sites run on the B200 kernels when their tensors are on a CUDA device and autograd is
off (inference / sampling / frozen pretransform); everything else — unsupported options, CPU tensors, autograd-tracked
calls — takes the ORIGINAL reference code (for unsupported options on CUDA under `strict=True` it raises instead):

  stable_audio_tools.models.dit.DiffusionTransformer.forward            -> DiTEngine.forward
  stable_audio_tools.models.autoencoders.OobleckEncoder.forward          -> OobleckEngine (encoder half, returns mean|scale)
  stable_audio_tools.models.autoencoders.OobleckDecoder.forward          -> OobleckEngine.decode
  stable_audio_tools.inference.sampling.sample_k (dpmpp-3m-sde, v-ddim)  -> CUDA-graph sampler (b200sat.sampling)
  stable_audio_tools.models.discriminators.EncodecDiscriminator.loss     -> b200sat.discriminator (WITH autograd: dis w.r.t. the
                                                                            discriminator parameters, adv / fm w.r.t. the fakes)

Engines are built lazily from `module.state_dict()` and rebuilt when any parameter's version counter changes (optimizer
step, load_state_dict).  `train.py` / `run_gradio.py` stay byte-identical: call `install()` from a `sitecustomize` /
`.pth` hook, or set `SAT_B200=1` and import `b200sat.autoinstall`.

Training through the reference wrappers is NOT rerouted by install() in round 1 (except the discriminator loss): the training path lives in
`b200sat.dit_train.DiTTrainModel` (same parameter names, flat fp32 master/grad buffers); see INTEGRATION.md.
"""
import functools
import importlib

import torch

_installed = {}
_TEST_TREAT_CPU_AS_DEVICE = False   # tests/test_install.py flips this to exercise the routing logic without a GPU


def _on_device(t):
    return t.is_cuda or _TEST_TREAT_CPU_AS_DEVICE


def _versions(module):
    return tuple(p._version for p in module.parameters()) + tuple(b._version for b in module.buffers())


class _EngineCache:
    """module -> engine, invalidated by parameter version counters."""

    def __init__(self, factory):
        self.factory = factory
        self.store = {}

    def get(self, module):
        key = id(module)
        ver = _versions(module)
        hit = self.store.get(key)
        if hit is None or hit[0] != ver:
            hit = (ver, self.factory(module))
            self.store[key] = hit
        return hit[1]


def _dit_supported(m, kwargs):
    """The option set DiTEngine implements (everything else must not be silently approximated)."""
    t = m.transformer
    bad = []
    if getattr(m, "patch_size", 1) != 1: bad.append("patch_size != 1")
    if getattr(m, "input_concat_dim", 0) != 0: bad.append("input_concat_cond")
    if getattr(m, "timestep_cond_type", "global") != "global": bad.append("timestep_cond_type")
    if getattr(t, "num_memory_tokens", 0): bad.append("memory tokens")
    if getattr(t, "use_sinusoidal_emb", False) or getattr(t, "use_abs_pos_emb", False): bad.append("absolute position embeddings")
    if getattr(t, "causal", False): bad.append("causal")
    if getattr(t, "sliding_window", None) is not None: bad.append("sliding window")
    blk = t.layers[0]
    if blk.self_attn.dim_heads != 64: bad.append("dim_heads != 64")
    if getattr(blk.self_attn, "qk_norm", "none") != "none" or getattr(blk.self_attn, "differential", False): bad.append("qk_norm / differential attention")
    if getattr(blk, "conformer", None) is not None: bad.append("conformer")
    for k in ("prepend_cond", "input_concat_cond", "mask", "exit_layer_ix", "negative_global_embed"):
        if kwargs.get(k) is not None: bad.append(k)
    if kwargs.get("return_info"): bad.append("return_info")
    if kwargs.get("cfg_interval", (0, 1)) not in ((0, 1), [0, 1], (0.0, 1.0)): bad.append("cfg_interval")
    return bad


def install(strict=False, engine_factories=None):
    """Patch the reference modules in place.  Returns the dict of original callables (also used by `uninstall`).
    engine_factories: test hook {"dit": f(module), "oobleck": f(state_dict, strides)} to substitute fake engines."""
    if _installed:
        return _installed
    dit_mod = importlib.import_module("stable_audio_tools.models.dit")
    ae_mod = importlib.import_module("stable_audio_tools.models.autoencoders")
    samp_mod = importlib.import_module("stable_audio_tools.inference.sampling")
    ef = engine_factories or {}

    def make_dit(m):
        from .dit_engine import DiTEngine
        return DiTEngine(m.state_dict(), device=next(m.parameters()).device)

    def make_ae(m):
        from .autoencoder import OobleckEngine
        pre = "encoder." if isinstance(m, ae_mod.OobleckEncoder) else "decoder."
        sd = {pre + k: v for k, v in m.state_dict().items()}
        strides = _oobleck_strides(m)
        return OobleckEngine(sd, strides=strides, device=next(m.parameters()).device, precision="fp32x3",
                             final_tanh=isinstance(m.layers[-1], torch.nn.Tanh))

    dit_cache = _EngineCache(ef.get("dit", make_dit))
    ae_cache = _EngineCache(ef.get("oobleck", make_ae))

    # ---------------------------------------------------------------- DiffusionTransformer.forward (dit.py:231-431)
    orig_dit_forward = dit_mod.DiffusionTransformer.forward

    @functools.wraps(orig_dit_forward)
    def dit_forward(self, x, t, cross_attn_cond=None, cross_attn_cond_mask=None, negative_cross_attn_cond=None,
                    negative_cross_attn_mask=None, global_embed=None, cfg_scale=1.0, cfg_dropout_prob=0.0, scale_phi=0.0, **kw):
        fast = _on_device(x) and not torch.is_grad_enabled() and negative_cross_attn_mask is None
        if fast:
            bad = _dit_supported(self, kw)
            if bad and strict:
                raise NotImplementedError("b200sat DiT engine does not implement: " + ", ".join(bad))
            fast = not bad
        if not fast:
            return orig_dit_forward(self, x, t, cross_attn_cond=cross_attn_cond, cross_attn_cond_mask=cross_attn_cond_mask,
                                    negative_cross_attn_cond=negative_cross_attn_cond, negative_cross_attn_mask=negative_cross_attn_mask,
                                    global_embed=global_embed, cfg_scale=cfg_scale, cfg_dropout_prob=cfg_dropout_prob, scale_phi=scale_phi, **kw)
        eng = dit_cache.get(self)
        out = eng.forward(x, t, cross_attn_cond=cross_attn_cond, global_embed=global_embed, cfg_scale=cfg_scale, scale_phi=scale_phi,
                          negative_cross_attn_cond=negative_cross_attn_cond)
        return out.to(next(self.parameters()).dtype)

    dit_mod.DiffusionTransformer.forward = dit_forward

    # ---------------------------------------------------------------- OobleckEncoder / OobleckDecoder .forward
    orig_enc_forward = ae_mod.OobleckEncoder.forward
    orig_dec_forward = ae_mod.OobleckDecoder.forward

    def enc_forward(self, x):
        if not (_on_device(x) and not torch.is_grad_enabled()):
            return orig_enc_forward(self, x)
        _z, info = ae_cache.get(self).encode(x, noise=None, return_info=True)
        return info["mean_scale"].to(x.dtype)          # [B, latent_dim, T/ratio]: what `self.layers(x)` returns (autoencoders.py:316)

    def dec_forward(self, x):
        if not (_on_device(x) and not torch.is_grad_enabled()):
            return orig_dec_forward(self, x)
        return ae_cache.get(self).decode(x).to(x.dtype)

    ae_mod.OobleckEncoder.forward = enc_forward
    ae_mod.OobleckDecoder.forward = dec_forward

    # ---------------------------------------------------------------- sample_k (sampling.py:331-412)
    orig_sample_k = samp_mod.sample_k

    @functools.wraps(orig_sample_k)
    def sample_k(model_fn, noise, init_data=None, steps=100, sampler_type="dpmpp-2m-sde", sigma_min=0.01, sigma_max=100, rho=1.0,
                 device="cuda", callback=None, cond_fn=None, **extra_args):
        dit = getattr(getattr(model_fn, "model", None), "__class__", None)
        ok = (noise.is_cuda and not torch.is_grad_enabled() and init_data is None and callback is None and cond_fn is None
              and sampler_type in ("dpmpp-3m-sde", "v-ddim") and dit is dit_mod.DiffusionTransformer
              and set(extra_args) <= {"cross_attn_cond", "cross_attn_mask", "global_cond", "cfg_scale", "batch_cfg", "rescale_cfg", "scale_phi",
                                      "negative_cross_attn_cond", "negative_cross_attn_mask", "cfg_interval"}
              and extra_args.get("negative_cross_attn_mask") is None and not _dit_supported(model_fn.model, {}))
        if not ok:
            return orig_sample_k(model_fn, noise, init_data, steps, sampler_type, sigma_min, sigma_max, rho, device, callback, cond_fn, **extra_args)
        from . import sampling
        eng = dit_cache.get(model_fn.model)
        kw = dict(cross_attn_cond=extra_args.get("cross_attn_cond"), global_embed=extra_args.get("global_cond"),
                  cfg_scale=extra_args.get("cfg_scale", 1.0), scale_phi=extra_args.get("scale_phi", 0.0))
        if sampler_type == "dpmpp-3m-sde":
            out = sampling.sample_k_dpmpp_3m_sde(eng, noise, steps, sigma_min, sigma_max, rho, **kw)
        else:
            out = sampling.sample_v_ddim(eng, noise, steps, sigma_max, **kw)
        return out.to(noise.dtype)

    samp_mod.sample_k = sample_k
    gen_mod = importlib.import_module("stable_audio_tools.inference.generation")
    orig_gen_sample_k = getattr(gen_mod, "sample_k", None)
    if orig_gen_sample_k is not None:
        gen_mod.sample_k = sample_k          # generation.py does `from .sampling import sample_k`

    # ---------------------------------------------------------------- EncodecDiscriminator.loss (discriminators.py:31-58)
    try:
        disc_mod = importlib.import_module("stable_audio_tools.models.discriminators")
    except Exception:   # optional dependency chain of the reference not importable: leave the discriminator alone
        disc_mod = None
    if disc_mod is not None:
        orig_disc_loss = disc_mod.EncodecDiscriminator.loss

        def _disc_supported(m):
            bad = []
            if getattr(m, "normalize_losses", False): bad.append("normalize_losses")
            if getattr(m, "loss_type", "hinge") != "hinge": bad.append("loss_type != hinge")
            for d in m.discriminators.discriminators:
                c0 = d.convs[0].conv
                if tuple(c0.weight_v.shape[:2]) != (64, 4): bad.append("filters != 64 or non-stereo input")
                if any(tuple(c.conv.stride) != (1, 1) for c in d.convs): bad.append("stride != (1,1)")
                if getattr(d, "spec_scale_pow", 0.0) != 0.0: bad.append("spec_scale_pow")
                if d.win_length != d.n_fft or not d.normalized: bad.append("win_length != n_fft / normalized=False")
            return sorted(set(bad))

        def disc_loss(self, reals, fakes):
            fast = _on_device(reals)
            if fast:
                bad = _disc_supported(self)
                if bad and strict:
                    raise NotImplementedError("b200sat discriminator does not implement: " + ", ".join(bad))
                fast = not bad
            if not fast:
                return orig_disc_loss(self, reals, fakes)
            fn = ef.get("disc_loss")
            if fn is None:
                from .discriminator import reference_discriminator_loss as fn
            dis, adv, fm = fn(self, reals, fakes)
            return dis, adv, fm

        disc_mod.EncodecDiscriminator.loss = disc_loss
        _installed["disc_loss"] = (disc_mod.EncodecDiscriminator, "loss", orig_disc_loss)

    _installed.update(dit_forward=(dit_mod.DiffusionTransformer, "forward", orig_dit_forward),
                      enc_forward=(ae_mod.OobleckEncoder, "forward", orig_enc_forward),
                      dec_forward=(ae_mod.OobleckDecoder, "forward", orig_dec_forward),
                      sample_k=(samp_mod, "sample_k", orig_sample_k))
    if orig_gen_sample_k is not None:
        _installed["gen_sample_k"] = (gen_mod, "sample_k", orig_gen_sample_k)
    return _installed


def uninstall():
    for owner, name, orig in _installed.values():
        setattr(owner, name, orig)
    _installed.clear()


def _oobleck_strides(m):
    """Recover the stride list from the module structure (EncoderBlock/DecoderBlock convs, autoencoders.py:233-283)."""
    strides = []
    for layer in m.layers:
        sub = getattr(layer, "layers", None)
        if sub is None:
            continue
        for l in sub:
            s = getattr(l, "stride", None)
            if s is not None and isinstance(s, tuple) and s[0] > 1:
                strides.append(s[0])
    if m.__class__.__name__ == "OobleckDecoder":
        strides = strides[::-1]
    return tuple(strides)

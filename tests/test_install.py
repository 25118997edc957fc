"""CPU: `b200sat.install()` re-routes the UNMODIFIED reference modules (imported from baseline/_ref, the pip-installed copy of
the reference) and keeps their semantics.  The engines are substituted by oracle-backed fakes so the routing, the option
screening, the fallbacks and the cache invalidation are tested without a GPU.  Skipped where the reference is absent."""
import pytest
import torch

from baseline import ref_loader as ref_harness

pytestmark = pytest.mark.skipif(not ref_harness.available(), reason="baseline/_ref not installed")


class _FakeDiT:
    builds = 0

    def __init__(self, module):
        from oracle import dit as odit
        _FakeDiT.builds += 1
        self.sd = {k: v.detach().clone() for k, v in module.state_dict().items()}
        self.depth = len(module.transformer.layers)
        self.gct = module.global_cond_type
        self.odit = odit

    def forward(self, x, t, cross_attn_cond=None, global_embed=None, cfg_scale=1.0, scale_phi=0.0, negative_cross_attn_cond=None,
                input_concat_cond=None):
        return self.odit.dit_forward(x, t, self.sd, self.depth, cross_attn_cond, global_embed, cfg_scale=cfg_scale, scale_phi=scale_phi,
                                     global_cond_type=self.gct, negative_cross_attn_cond=negative_cross_attn_cond,
                                     input_concat_cond=input_concat_cond)


class _FakeTrainer:
    """Stands in for b200sat.dit_train.ReferenceDiTTrainer on CPU: records the call and runs the reference's own forward."""
    calls = 0

    def __init__(self, module):
        self.module = module

    def forward(self, x, t, cross_attn_cond=None, global_embed=None, cfg_dropout_prob=0.0):
        import b200sat.install as inst
        _FakeTrainer.calls += 1
        return inst._installed["dit_forward"][2](self.module, x, t, cross_attn_cond=cross_attn_cond, global_embed=global_embed,
                                                  cfg_dropout_prob=cfg_dropout_prob)


class _FakeAE:
    def __init__(self, module):
        from oracle import oobleck as oo
        from b200sat.install import _oobleck_strides
        self.oo = oo
        self.enc = module.__class__.__name__ == "OobleckEncoder"
        pre = "encoder." if self.enc else "decoder."
        self.sd = {pre + k: v.detach().clone() for k, v in module.state_dict().items()}
        self.strides = _oobleck_strides(module)

    def encode(self, x, noise=None, return_info=False):
        ms = self.oo.oobleck_encode(x, self.sd, self.strides)
        return ms[:, : ms.shape[1] // 2], {"mean_scale": ms, "kl": torch.zeros(())}

    def decode(self, z):
        return self.oo.oobleck_decode(z, self.sd, self.strides)


@pytest.fixture
def installed():
    import b200sat.install as inst
    R = ref_harness.load()
    inst._TEST_TREAT_CPU_AS_DEVICE = True
    inst.install(strict=True, engine_factories={"dit": _FakeDiT, "oobleck": _FakeAE, "dit_train": _FakeTrainer})
    yield R, inst
    inst.uninstall()
    inst._TEST_TREAT_CPU_AS_DEVICE = False


def test_dit_forward_routing_cache_and_fallbacks(installed):
    R, inst = installed
    from oracle import dit as odit
    kw = dict(embed_dim=128, depth=2, num_heads=2, io_channels=64, cond_token_dim=64, global_cond_dim=128)
    m = R.dit.DiffusionTransformer(project_cond_tokens=False, transformer_type="continuous_transformer", **kw).eval()
    m.load_state_dict(odit.make_state_dict(seed=3, **kw))
    x = torch.randn(2, 64, 40); t = torch.rand(2); c = torch.randn(2, 5, 64); g = torch.randn(2, 128)
    orig = inst._installed["dit_forward"][2]
    _FakeDiT.builds = 0
    with torch.no_grad():
        ref = orig(m, x, t, cross_attn_cond=c, global_embed=g, cfg_scale=5.0, scale_phi=0.5)
        got = m(x, t, cross_attn_cond=c, global_embed=g, cfg_scale=5.0, scale_phi=0.5)       # routed through the (fake) engine
        got2 = m(x, t, cross_attn_cond=c, global_embed=g, cfg_scale=5.0, scale_phi=0.5)
    assert _FakeDiT.builds == 1, "engine must be cached across calls"
    assert torch.allclose(got, ref, atol=1e-5) and torch.equal(got, got2)
    with torch.no_grad():
        m.transformer.project_out.weight.mul_(1.5)                                            # bumps the parameter version counter
        got3 = m(x, t, cross_attn_cond=c, global_embed=g, cfg_scale=5.0, scale_phi=0.5)
        ref3 = orig(m, x, t, cross_attn_cond=c, global_embed=g, cfg_scale=5.0, scale_phi=0.5)
    assert _FakeDiT.builds == 2 and torch.allclose(got3, ref3, atol=1e-5), "engine must be rebuilt after a weight update"
    # autograd-tracked calls go to the training route (one trainer per module), never to the inference engine
    _FakeTrainer.calls = 0
    out = m(x, t, cross_attn_cond=c, global_embed=g)
    out2 = m(x, t, cross_attn_cond=c, global_embed=g, cfg_dropout_prob=0.0)
    assert out.requires_grad and _FakeDiT.builds == 2 and _FakeTrainer.calls == 2 and torch.allclose(out, out2)
    # a frozen model (requires_grad False everywhere) is inference even with grad mode on
    m.requires_grad_(False)
    out3 = m(x, t, cross_attn_cond=c, global_embed=g)
    assert not out3.requires_grad and _FakeTrainer.calls == 2
    m.requires_grad_(True)
    # unsupported options raise under strict=True instead of being approximated
    with torch.no_grad(), pytest.raises(NotImplementedError):
        m(x, t, cross_attn_cond=c, global_embed=g, return_info=True)


def test_oobleck_forward_routing(installed):
    R, inst = installed
    from oracle import oobleck as oo
    enc = R.autoencoders.OobleckEncoder(in_channels=2, channels=16, latent_dim=16, c_mults=[1, 2, 4], strides=[2, 4, 4], use_snake=True).eval()
    dec = R.autoencoders.OobleckDecoder(out_channels=2, channels=16, latent_dim=8, c_mults=[1, 2, 4], strides=[2, 4, 4], use_snake=True, final_tanh=False).eval()
    sd = oo.make_state_dict(channels=16, c_mults=(1, 2, 4), strides=(2, 4, 4), enc_latent=16, dec_latent=8, seed=4)
    enc.load_state_dict({k[len("encoder."):]: v for k, v in sd.items() if k.startswith("encoder.")})
    dec.load_state_dict({k[len("decoder."):]: v for k, v in sd.items() if k.startswith("decoder.")})
    assert inst._oobleck_strides(enc) == (2, 4, 4) and inst._oobleck_strides(dec) == (2, 4, 4)
    x = torch.randn(1, 2, 1024); z = torch.randn(1, 8, 32)
    with torch.no_grad():
        assert torch.allclose(enc(x), inst._installed["enc_forward"][2](enc, x), atol=1e-4)
        assert torch.allclose(dec(z), inst._installed["dec_forward"][2](dec, z), atol=1e-4)


def test_sample_k_falls_back_for_unsupported_samplers(installed):
    R, inst = installed
    import stable_audio_tools.inference.sampling as S
    assert S.sample_k is not inst._installed["sample_k"][2]
    toy = lambda x_, t_, **kw: torch.zeros_like(x_)
    out = S.sample_k(toy, torch.randn(1, 4, 8), steps=3, sampler_type="v-ddim", sigma_max=1.0, device="cpu")   # model_fn is not a DiT wrapper
    assert out.shape == (1, 4, 8)


def test_discriminator_loss_routing_and_fallbacks(installed):
    """EncodecDiscriminator.loss on the unmodified reference class: routed (with the reference's parameter names) for the supported
    option set, original code for normalize_losses / non-device tensors; the routed callable here is the oracle, fed from the module's
    own `named_parameters`, so the name mapping `discriminators.<module path>` is checked against the reference's result."""
    import stable_audio_tools.models.discriminators as RD
    from oracle import discriminator as od
    calls = []

    def fake(module, reals, fakes):
        sd = {"discriminators." + n: p for n, p in module.discriminators.named_parameters()}
        calls.append(len(sd))
        return od.discriminator_loss(reals, fakes, sd, n_ffts=(128, 256), hops=(32, 64))

    import b200sat.install as inst
    inst.uninstall()
    inst._TEST_TREAT_CPU_AS_DEVICE = True
    try:
        inst.install(engine_factories={"dit": _FakeDiT, "oobleck": _FakeAE, "disc_loss": fake})
        kw = dict(filters=64, in_channels=2, n_ffts=[128, 256], hop_lengths=[32, 64], win_lengths=[128, 256])
        m = RD.EncodecDiscriminator(**kw)
        g = torch.Generator().manual_seed(0)
        reals = torch.randn(1, 2, 2048, generator=g) * 0.3
        fakes = reals + 0.1 * torch.randn(1, 2, 2048, generator=g)
        got = m.loss(reals, fakes)
        assert calls == [36]
        ref = inst._installed["disc_loss"][2](m, reals, fakes)
        for a, b in zip(got, ref):
            assert abs(float(a) - float(b)) <= 1e-5 * max(1.0, abs(float(b)))
        m2 = RD.EncodecDiscriminator(normalize_losses=True, **kw)
        m2.loss(reals, fakes)
        assert calls == [36]                       # unsupported option -> original reference code
        inst._TEST_TREAT_CPU_AS_DEVICE = False
        m.loss(reals, fakes)
        assert calls == [36]                       # CPU tensors -> original reference code
    finally:
        inst._TEST_TREAT_CPU_AS_DEVICE = False
        inst.uninstall()


def _oracle_sampler(eng, noise, steps, sampler_type, sigma_min, sigma_max, rho, cross_attn_cond=None, global_embed=None, cfg_scale=1.0,
                    scale_phi=0.0, negative_cross_attn_cond=None, input_concat_cond=None):
    """engine_factories['sampler'] hook: the v-DDIM loop of b200sat.sampling's tables driven by the fake engine (CPU)."""
    from b200sat import sampling
    assert sampler_type == "v-ddim"
    coef, cin, tt = sampling.v_ddim_tables(steps, min(sigma_max, 1.0))
    x = noise.float()
    for i in range(steps):
        t = torch.full((x.shape[0],), float(tt[i]))
        extra = {} if input_concat_cond is None else {"input_concat_cond": input_concat_cond}
        v = eng.forward(x * cin[i], t, cross_attn_cond=cross_attn_cond, global_embed=global_embed, cfg_scale=cfg_scale, scale_phi=scale_phi,
                        negative_cross_attn_cond=negative_cross_attn_cond, **extra)
        den = v * coef[i, 0] + x * coef[i, 1]
        x = coef[i, 2] * x + coef[i, 3] * den
    return x


def test_generate_diffusion_cond_takes_the_fast_sampler_route():
    """The reference's OWN generate_diffusion_cond -> get_conditioning_inputs -> sample_k call chain (inference/generation.py:91-220):
    its kwargs always carry input_concat_cond / prepend_cond / prepend_cond_mask = None, which must not block the fast route; a
    negative prompt is forwarded; a cfg_interval other than (0, 1) falls back."""
    import b200sat.install as inst
    from baseline import ref_models
    R = ref_harness.load(force_sdpa=True)
    inst.uninstall()
    cfg = ref_models.sao_config(depth=2, embed_dim=128, num_heads=2, cond_token_dim=64, global_cond_dim=128)
    model = ref_models.build_diffusion_cond(R, cfg, seed=5)
    ct = ref_models.conditioning_tensors(model, 2, prompt_tokens=5, seed=1)
    neg = ref_models.conditioning_tensors(model, 2, prompt_tokens=5, seed=2)
    kw = dict(steps=5, cfg_scale=4.0, batch_size=2, sample_size=48, seed=11, device="cpu", sampler_type="v-ddim", return_latents=True)
    ref = R.generation.generate_diffusion_cond(model, conditioning_tensors=ct, **kw)
    ref_neg = R.generation.generate_diffusion_cond(model, conditioning_tensors=ct, negative_conditioning_tensors=neg, **kw)
    assert not torch.allclose(ref, ref_neg, atol=1e-3)
    inst._TEST_TREAT_CPU_AS_DEVICE = True
    try:
        inst.install(strict=True, engine_factories={"dit": _FakeDiT, "oobleck": _FakeAE, "dit_train": _FakeTrainer, "sampler": _oracle_sampler})
        n0 = dict(inst.STATS)
        got = R.generation.generate_diffusion_cond(model, conditioning_tensors=ct, **kw)
        assert inst.STATS["sample_k_fast"] == n0["sample_k_fast"] + 1 and inst.STATS["sample_k_ref"] == n0["sample_k_ref"]
        assert torch.allclose(got, ref, atol=2e-4), float((got - ref).abs().max())
        got_neg = R.generation.generate_diffusion_cond(model, conditioning_tensors=ct, negative_conditioning_tensors=neg, **kw)
        assert inst.STATS["sample_k_fast"] == n0["sample_k_fast"] + 2
        assert torch.allclose(got_neg, ref_neg, atol=2e-4), "negative prompt must reach the engine"
        # cfg_interval != (0, 1): strict raises; non-strict runs the reference loop and the reference model code
        with pytest.raises(NotImplementedError):
            R.generation.generate_diffusion_cond(model, conditioning_tensors=ct, cfg_interval=(0.2, 0.8), **kw)
        inst.uninstall()
        inst.install(strict=False, engine_factories={"dit": _FakeDiT, "oobleck": _FakeAE, "dit_train": _FakeTrainer, "sampler": _oracle_sampler})
        n1 = dict(inst.STATS)
        got_ci = R.generation.generate_diffusion_cond(model, conditioning_tensors=ct, cfg_interval=(0.2, 0.8), **kw)
        assert inst.STATS["sample_k_ref"] == n1["sample_k_ref"] + 1 and inst.STATS["dit_fast"] == n1["dit_fast"]
        inst.uninstall()
        ref_ci = R.generation.generate_diffusion_cond(model, conditioning_tensors=ct, cfg_interval=(0.2, 0.8), **kw)
        assert torch.allclose(got_ci, ref_ci, atol=2e-4)
    finally:
        inst._TEST_TREAT_CPU_AS_DEVICE = False
        inst.uninstall()


def test_generate_diffusion_cond_inpaint_reaches_the_engines():
    """Row f4: the reference's OWN generate_diffusion_cond_inpaint (inference/generation.py:222-405) on a model built with
    input_concat_ids = [inpaint_mask, inpaint_masked_input]: the mask / masked-input tensors it concatenates reach the sampler route as
    `input_concat_cond`; with init_audio (variation start) the reference's loop runs and only its model calls are routed.  Same latents
    as without install() in both cases."""
    import b200sat.install as inst
    from baseline import ref_models
    R = ref_harness.load(force_sdpa=True)
    inst.uninstall()
    cfg = ref_models.sao_config(depth=2, embed_dim=128, num_heads=2, cond_token_dim=64, global_cond_dim=128, sample_size=48)
    cfg["model"]["diffusion"]["config"]["input_concat_dim"] = 65
    cfg["model"]["diffusion"]["input_concat_ids"] = ["inpaint_mask", "inpaint_masked_input"]
    model = ref_models.build_diffusion_cond(R, cfg, seed=6)
    B, T = 2, 48
    ct = ref_models.conditioning_tensors(model, B, prompt_tokens=5, seed=1)
    g = torch.Generator().manual_seed(3)
    audio, init = torch.randn(64, T, generator=g), torch.randn(64, T, generator=g)
    mask = (torch.arange(T) >= 20).float().unsqueeze(0).repeat(B, 1)
    kw = dict(steps=4, cfg_scale=3.0, batch_size=B, sample_size=T, seed=9, device="cpu", sampler_type="v-ddim", return_latents=True,
              inpaint_audio=(44100, audio), inpaint_mask=mask)
    ref = R.generation.generate_diffusion_cond_inpaint(model, conditioning_tensors=dict(ct), **kw)
    ref_init = R.generation.generate_diffusion_cond_inpaint(model, conditioning_tensors=dict(ct), init_audio=(44100, init), init_noise_level=0.6, **kw)
    inst._TEST_TREAT_CPU_AS_DEVICE = True
    try:
        inst.install(strict=True, engine_factories={"dit": _FakeDiT, "oobleck": _FakeAE, "dit_train": _FakeTrainer, "sampler": _oracle_sampler})
        n0 = dict(inst.STATS)
        got = R.generation.generate_diffusion_cond_inpaint(model, conditioning_tensors=dict(ct), **kw)
        assert inst.STATS["sample_k_fast"] == n0["sample_k_fast"] + 1 and inst.STATS["sample_k_ref"] == n0["sample_k_ref"]
        assert torch.allclose(got, ref, atol=2e-4), float((got - ref).abs().max())
        n1 = dict(inst.STATS)
        got_init = R.generation.generate_diffusion_cond_inpaint(model, conditioning_tensors=dict(ct), init_audio=(44100, init), init_noise_level=0.6, **kw)
        assert inst.STATS["sample_k_ref"] == n1["sample_k_ref"] + 1 and inst.STATS["dit_fast"] == n1["dit_fast"] + 4   # reference loop, routed model
        assert torch.allclose(got_init, ref_init, atol=2e-4), float((got_init - ref_init).abs().max())
    finally:
        inst._TEST_TREAT_CPU_AS_DEVICE = False
        inst.uninstall()


def test_unsupported_architectures_fall_back_instead_of_approximating():
    import b200sat.install as inst
    R = ref_harness.load(force_sdpa=True)
    kw = dict(embed_dim=128, depth=1, num_heads=2, io_channels=64, cond_token_dim=64, global_cond_dim=128, project_cond_tokens=False,
              transformer_type="continuous_transformer")
    variants = {
        "layer_scale": R.dit.DiffusionTransformer(layer_scale=True, **kw),
        "io_channels": R.dit.DiffusionTransformer(**{**kw, "io_channels": 32}),
        "remove_norms": R.dit.DiffusionTransformer(remove_norms=True, **kw),
    }
    for name, m in variants.items():
        bad = inst._dit_supported(m, {})
        assert bad, f"{name}: must be reported as unsupported"
    ok = R.dit.DiffusionTransformer(**kw)
    assert inst._dit_supported(ok, {}) == []
    assert inst._dit_supported(ok, {"cfg_interval": (0.1, 1.0)}) == ["cfg_interval"]
    # inpainting models (input_concat_dim = mask + masked latents): routed only together with their concatenated conditioning
    inp = R.dit.DiffusionTransformer(input_concat_dim=65, **kw)
    assert inst._dit_supported(inp, {"input_concat_cond": torch.zeros(1, 65, 8)}) == []
    assert inst._dit_supported(inp, {}) != [] and inst._dit_supported(ok, {"input_concat_cond": torch.zeros(1, 65, 8)}) != []


def test_engine_cache_is_weak_and_tracks_dtype_and_storage():
    import gc
    import b200sat.install as inst
    built = []
    cache = inst._EngineCache(lambda m: built.append(1) or object())
    lin = torch.nn.Linear(4, 4)
    e0 = cache.get(lin)
    assert cache.get(lin) is e0 and len(built) == 1
    lin.half()
    assert cache.get(lin) is not e0 and len(built) == 2, "a dtype change must rebuild the engine"
    lin.float()
    assert len(cache.store) == 1
    del lin
    gc.collect()
    assert len(cache.store) == 0, "engines of collected modules must be dropped"


def test_fp32_model_outside_autocast_is_not_silently_run_in_bf16():
    import b200sat.install as inst
    m = torch.nn.Linear(4, 4)
    assert not inst._half_compute(m, fp32_models=False)
    assert inst._half_compute(m, fp32_models=True)
    assert inst._half_compute(m.bfloat16(), fp32_models=False)


def test_autoinstall_pth_hook(tmp_path):
    """b200sat.pth + SAT_B200=1: the first create_model_from_config call installs the routing; train.py is untouched."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import site, sys\n"
        f"site.addsitedir({os.path.join(root, 'stable-audio-tools_b200')!r})\n"
        f"sys.path.insert(0, {root!r})\n"
        "assert ('b200sat.autoinstall' in sys.modules) == (__import__('os').environ.get('SAT_B200') == '1')\n"
        "from baseline import ref_loader, ref_models\n"
        "R = ref_loader.load(force_sdpa=True)\n"
        "import b200sat.install as inst\n"
        "assert not inst._installed\n"
        "cfg = ref_models.sao_config(depth=1, embed_dim=128, num_heads=2, cond_token_dim=64, global_cond_dim=128)\n"
        "import stable_audio_tools.models.factory as F\n"
        "m = F.create_model_from_config(cfg)\n"
        "print('INSTALLED' if inst._installed else 'PLAIN')\n")
    for flag, want in (("1", "INSTALLED"), ("0", "PLAIN")):
        r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, SAT_B200=flag), capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-1500:]
        assert r.stdout.strip().splitlines()[-1] == want

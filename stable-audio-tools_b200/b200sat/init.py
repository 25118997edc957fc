"""Seeded random weights with the reference's parameter names/shapes, generated directly on the device (there is no
network for checkpoints; BASELINE.json's configs are measured on random-init weights of the named architecture).
Zero-initialised reference weights (to_out, ff out, pre/post conv: transformer.py:311-314,:366-367; dit.py:121-123) are drawn
non-zero so that no branch is an identity."""
import torch


def dit_state_dict(embed_dim=1536, depth=24, num_heads=24, io_channels=64, cond_token_dim=768, global_cond_dim=1536,
                   global_cond_type="prepend", seed=0, device="cuda", dtype=torch.bfloat16, std=0.02):
    g = torch.Generator(device=device).manual_seed(seed)
    r = lambda *s, sc=std: (torch.randn(*s, generator=g, device=device) * sc).to(dtype)
    d, ff = embed_dim, 4 * embed_dim
    dh = d // num_heads
    rot = max(dh // 2, 32)
    sd = {
        "timestep_features.weight": r(128, 1, sc=1.0),
        "to_timestep_embed.0.weight": r(d, 256), "to_timestep_embed.0.bias": r(d),
        "to_timestep_embed.2.weight": r(d, d), "to_timestep_embed.2.bias": r(d),
        "preprocess_conv.weight": r(io_channels, io_channels, 1, sc=0.05),
        "postprocess_conv.weight": r(io_channels, io_channels, 1, sc=0.05),
        "transformer.project_in.weight": r(d, io_channels, sc=0.1),
        "transformer.project_out.weight": r(io_channels, d),
        "transformer.rotary_pos_emb.inv_freq": (1.0 / (10000 ** (torch.arange(0, rot, 2, device=device).float() / rot))).to(dtype),
    }
    if cond_token_dim > 0:
        sd["to_cond_embed.0.weight"] = r(cond_token_dim, cond_token_dim, sc=0.04)
        sd["to_cond_embed.2.weight"] = r(cond_token_dim, cond_token_dim, sc=0.04)
    if global_cond_dim > 0:
        sd["to_global_embed.0.weight"] = r(d, global_cond_dim)
        sd["to_global_embed.2.weight"] = r(d, d)
    if global_cond_type == "adaLN":
        sd["transformer.global_cond_embedder.0.weight"] = r(d, d); sd["transformer.global_cond_embedder.0.bias"] = r(d)
        sd["transformer.global_cond_embedder.2.weight"] = r(6 * d, d); sd["transformer.global_cond_embedder.2.bias"] = r(6 * d)
    for i in range(depth):
        p = f"transformer.layers.{i}."
        for n in ("pre_norm", "cross_attend_norm", "ff_norm"):
            if n == "cross_attend_norm" and cond_token_dim == 0:
                continue
            sd[p + n + ".gamma"] = (1.0 + 0.1 * torch.randn(d, generator=g, device=device)).to(dtype)
            sd[p + n + ".beta"] = torch.zeros(d, device=device, dtype=dtype)
        sd[p + "self_attn.to_qkv.weight"] = r(3 * d, d)
        sd[p + "self_attn.to_out.weight"] = r(d, d)
        if cond_token_dim > 0:
            sd[p + "cross_attn.to_q.weight"] = r(d, d)
            sd[p + "cross_attn.to_kv.weight"] = r(2 * cond_token_dim, cond_token_dim, sc=0.04)
            sd[p + "cross_attn.to_out.weight"] = r(d, d)
        sd[p + "ff.ff.0.proj.weight"] = r(2 * ff, d); sd[p + "ff.ff.0.proj.bias"] = r(2 * ff)
        sd[p + "ff.ff.2.weight"] = r(d, ff); sd[p + "ff.ff.2.bias"] = r(d)
        if global_cond_type == "adaLN":
            sd[p + "to_scale_shift_gate"] = (torch.randn(6 * d, generator=g, device=device) / d ** 0.5).to(dtype)
    return sd


def oobleck_state_dict(dev, g, ch=128):
    """Random weights with the reference names/shapes of stable_audio_2_0_vae.json (autoencoders.py:285-362)."""
    import math
    sd = {}
    cm, strides = [1, 1, 2, 4, 8, 16], [2, 4, 4, 8, 8]

    def conv(p, cout, cin, k, bias=True, transpose=False):
        shape = (cin, cout, k) if transpose else (cout, cin, k)
        v = torch.randn(*shape, device=dev, generator=g) / math.sqrt(cin * k)
        sd[p + "weight_v"] = v
        sd[p + "weight_g"] = v.flatten(1).norm(dim=1).view(shape[0], 1, 1) * 0.7
        if bias:
            sd[p + "bias"] = 0.05 * torch.randn(cout, device=dev, generator=g)

    def snake(p, c):
        sd[p + "alpha"] = 0.3 * torch.randn(c, device=dev, generator=g); sd[p + "beta"] = 0.3 * torch.randn(c, device=dev, generator=g)

    def ru(p, c):
        snake(p + "layers.0.", c); conv(p + "layers.1.", c, c, 7); snake(p + "layers.2.", c); conv(p + "layers.3.", c, c, 1)

    n = len(strides)
    p = "encoder.layers."
    conv(p + "0.", cm[0] * ch, 2, 7)
    for i in range(n):
        ci, co = cm[i] * ch, cm[i + 1] * ch
        for j in range(3):
            ru(f"{p}{i + 1}.layers.{j}.", ci)
        snake(f"{p}{i + 1}.layers.3.", ci); conv(f"{p}{i + 1}.layers.4.", co, ci, 2 * strides[i])
    snake(f"{p}{n + 1}.", cm[-1] * ch); conv(f"{p}{n + 2}.", 128, cm[-1] * ch, 3)
    p = "decoder.layers."
    conv(p + "0.", cm[-1] * ch, 64, 7)
    for idx, i in enumerate(range(n, 0, -1)):
        ci, co = cm[i] * ch, cm[i - 1] * ch
        q = f"{p}{idx + 1}."
        snake(q + "layers.0.", ci); conv(q + "layers.1.", co, ci, 2 * strides[i - 1], transpose=True)
        for j in range(3):
            ru(f"{q}layers.{2 + j}.", co)
    snake(f"{p}{n + 1}.", cm[0] * ch); conv(f"{p}{n + 2}.", 2, cm[0] * ch, 7, bias=False)
    return sd


def encodec_disc_state_dict(dev, g, filters=64, in_channels=2, n_scales=5):
    """Random weights with the reference names/shapes of EncodecDiscriminator(filters=64) (models/encodec.py:76-92)."""
    import math
    sd = {}

    def conv(p, cout, cin, kh, kw):
        v = torch.randn(cout, cin, kh, kw, device=dev, generator=g) / math.sqrt(cin * kh * kw)
        sd[p + "weight_v"] = v
        sd[p + "weight_g"] = v.flatten(1).norm(dim=1).view(-1, 1, 1, 1).clone()
        sd[p + "bias"] = 0.05 * torch.randn(cout, device=dev, generator=g)

    for i in range(n_scales):
        pre = f"discriminators.discriminators.{i}."
        conv(pre + "convs.0.conv.", filters, 2 * in_channels, 3, 9)
        for j in range(1, 4):
            conv(f"{pre}convs.{j}.conv.", filters, filters, 3, 9)
        conv(pre + "convs.4.conv.", filters, filters, 3, 3)
        conv(pre + "conv_post.conv.", 1, filters, 3, 3)
    return sd

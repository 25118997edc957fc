"""Oracle: DiT denoiser (TEST INFRASTRUCTURE — see oracle/__init__.py).

Functional torch restatement operating on a flat state dict with the
reference's parameter names.  Citations are into
/root/reference/stable_audio_tools/models/.
"""
import math
import torch
import torch.nn.functional as F


def layer_norm(x, gamma, beta=None, eps=1e-5):
    # transformer.py:236-238 — F.layer_norm(x, gamma, beta(=0 buffer), eps=1e-5)
    return F.layer_norm(x, x.shape[-1:], weight=gamma, bias=beta, eps=eps)


def rope_freqs(seq_len, inv_freq):
    # transformer.py:122-138 — t=arange(N) fp32; freqs = outer(t, inv_freq); cat(freqs, freqs)
    t = torch.arange(seq_len, dtype=torch.float32)
    f = torch.einsum("i,j->ij", t, inv_freq.to(torch.float32))
    return torch.cat((f, f), dim=-1)


def rotate_half(x):
    # transformer.py:149-152
    d = x.shape[-1] // 2
    return torch.cat((-x[..., d:], x[..., :d]), dim=-1)


def apply_rope(t, freqs):
    # transformer.py:154-174 (fp32 inside; partial rotation over the first rot_dim dims)
    out_dtype = t.dtype
    rot = freqs.shape[-1]
    n = t.shape[-2]
    t = t.to(torch.float32)
    fr = freqs.to(torch.float32)[-n:, :]
    tr, tu = t[..., :rot], t[..., rot:]
    tr = tr * fr.cos() + rotate_half(tr) * fr.sin()
    return torch.cat((tr.to(out_dtype), tu.to(out_dtype)), dim=-1)


def sdpa(q, k, v):
    # transformer.py:440 — F.scaled_dot_product_attention, non-causal, no mask (math restated)
    scale = 1.0 / math.sqrt(q.shape[-1])
    s = torch.matmul(q.float(), k.float().transpose(-1, -2)) * scale
    p = torch.softmax(s, dim=-1)
    return torch.matmul(p, v.float()).to(q.dtype)


def split_heads(t, h):
    b, n, _ = t.shape
    return t.view(b, n, h, -1).permute(0, 2, 1, 3)


def merge_heads(t):
    b, h, n, d = t.shape
    return t.permute(0, 2, 1, 3).reshape(b, n, h * d)


def self_attention(x, sd, pre, freqs, dim_heads=64):
    # transformer.py:445-543 (fused to_qkv branch :481-482; RoPE :491-507; merge+to_out :526-534)
    h = x.shape[-1] // dim_heads
    qkv = F.linear(x, sd[pre + "to_qkv.weight"])
    q, k, v = qkv.chunk(3, dim=-1)
    q, k, v = (split_heads(t, h) for t in (q, k, v))
    if freqs is not None:
        qd = q.dtype
        q = apply_rope(q.float(), freqs).to(v.dtype)
        k = apply_rope(k.float(), freqs).to(v.dtype)
    o = sdpa(q, k, v)
    return F.linear(merge_heads(o), sd[pre + "to_out.weight"])


def cross_attention(x, ctx, sd, pre, dim_heads=64):
    # transformer.py:459-472 (to_q / to_kv), :406-411 (GQA repeat_interleave), no RoPE (:689)
    h = x.shape[-1] // dim_heads
    kvh = ctx.shape[-1] // dim_heads
    q = split_heads(F.linear(x, sd[pre + "to_q.weight"]), h)
    k, v = F.linear(ctx, sd[pre + "to_kv.weight"]).chunk(2, dim=-1)
    k, v = split_heads(k, kvh), split_heads(v, kvh)
    if h != kvh:
        k = k.repeat_interleave(h // kvh, dim=1)
        v = v.repeat_interleave(h // kvh, dim=1)
    o = sdpa(q, k, v)
    return F.linear(merge_heads(o), sd[pre + "to_out.weight"])


def feed_forward(x, sd, pre):
    # transformer.py:263-275 (GLU: proj -> chunk(value, gate) -> value*silu(gate)), :308 (linear_out)
    u = F.linear(x, sd[pre + "ff.0.proj.weight"], sd[pre + "ff.0.proj.bias"])
    a, g = u.chunk(2, dim=-1)
    return F.linear(a * F.silu(g), sd[pre + "ff.2.weight"], sd[pre + "ff.2.bias"])


def transformer_block(x, sd, pre, ctx, freqs, global_cond=None, dim_heads=64):
    # transformer.py:658-713
    g = lambda name: sd.get(pre + name)
    if global_cond is not None and (pre + "to_scale_shift_gate") in sd:
        # adaLN path :675-701
        ssg = (sd[pre + "to_scale_shift_gate"] + global_cond).unsqueeze(1)
        sc_s, sh_s, g_s, sc_f, sh_f, g_f = ssg.chunk(6, dim=-1)
        res = x
        h = layer_norm(x, g("pre_norm.gamma"), g("pre_norm.beta"))
        h = h * (1 + sc_s) + sh_s
        h = self_attention(h, sd, pre + "self_attn.", freqs, dim_heads)
        x = h * torch.sigmoid(1 - g_s) + res
        if ctx is not None and (pre + "cross_attn.to_q.weight") in sd:
            x = x + cross_attention(layer_norm(x, g("cross_attend_norm.gamma"), g("cross_attend_norm.beta")),
                                    ctx, sd, pre + "cross_attn.", dim_heads)
        res = x
        h = layer_norm(x, g("ff_norm.gamma"), g("ff_norm.beta"))
        h = h * (1 + sc_f) + sh_f
        h = feed_forward(h, sd, pre + "ff.")
        x = h * torch.sigmoid(1 - g_f) + res
        return x
    # plain path :703-712
    x = x + self_attention(layer_norm(x, g("pre_norm.gamma"), g("pre_norm.beta")), sd, pre + "self_attn.", freqs, dim_heads)
    if ctx is not None and (pre + "cross_attn.to_q.weight") in sd:
        x = x + cross_attention(layer_norm(x, g("cross_attend_norm.gamma"), g("cross_attend_norm.beta")),
                                ctx, sd, pre + "cross_attn.", dim_heads)
    x = x + feed_forward(layer_norm(x, g("ff_norm.gamma"), g("ff_norm.beta")), sd, pre + "ff.")
    return x


def continuous_transformer(x, sd, pre, depth, prepend_embeds=None, context=None, global_cond=None,
                           dim_heads=64, return_hidden=False):
    # transformer.py:796-865
    x = F.linear(x, sd[pre + "project_in.weight"])
    if prepend_embeds is not None:
        x = torch.cat((prepend_embeds, x), dim=-2)
    freqs = None
    if (pre + "rotary_pos_emb.inv_freq") in sd:
        freqs = rope_freqs(x.shape[1], sd[pre + "rotary_pos_emb.inv_freq"])
    if global_cond is not None and (pre + "global_cond_embedder.0.weight") in sd:
        # :767-773, :836-837
        gc = F.linear(global_cond, sd[pre + "global_cond_embedder.0.weight"], sd[pre + "global_cond_embedder.0.bias"])
        global_cond = F.linear(F.silu(gc), sd[pre + "global_cond_embedder.2.weight"], sd[pre + "global_cond_embedder.2.bias"])
    hidden = []
    for i in range(depth):
        x = transformer_block(x, sd, f"{pre}layers.{i}.", context, freqs, global_cond, dim_heads)
        if return_hidden:
            hidden.append(x)
    x = F.linear(x, sd[pre + "project_out.weight"])
    return (x, hidden) if return_hidden else x


def fourier_features(t, weight):
    # blocks.py:85-94
    f = 2 * math.pi * t @ weight.T
    return torch.cat([f.cos(), f.sin()], dim=-1)


def dit_inner(x, t, sd, depth, cross_attn_cond=None, global_embed=None, global_cond_type="prepend", dim_heads=64, input_concat_cond=None):
    """DiffusionTransformer._forward — dit.py:125-229.  x [B,C,T], t [B]; input_concat_cond [B, Dc, T'] is resized (nearest) and
    concatenated on the channel axis before the preprocess conv (dit.py:160-165)."""
    if input_concat_cond is not None:
        if input_concat_cond.shape[2] != x.shape[2]:
            input_concat_cond = F.interpolate(input_concat_cond, (x.shape[2],), mode="nearest")
        x = torch.cat([x, input_concat_cond], dim=1)
    if cross_attn_cond is not None:
        c = F.linear(cross_attn_cond, sd["to_cond_embed.0.weight"])
        cross_attn_cond = F.linear(F.silu(c), sd["to_cond_embed.2.weight"])
    if global_embed is not None:
        g = F.linear(global_embed, sd["to_global_embed.0.weight"])
        global_embed = F.linear(F.silu(g), sd["to_global_embed.2.weight"])
    te = fourier_features(t[:, None], sd["timestep_features.weight"])
    te = F.linear(te, sd["to_timestep_embed.0.weight"], sd["to_timestep_embed.0.bias"])
    te = F.linear(F.silu(te), sd["to_timestep_embed.2.weight"], sd["to_timestep_embed.2.bias"])
    global_embed = te if global_embed is None else global_embed + te
    prepend = None
    extra = {}
    if global_cond_type == "prepend":
        prepend = global_embed.unsqueeze(1)
    else:
        extra["global_cond"] = global_embed
    x = F.conv1d(x, sd["preprocess_conv.weight"]) + x
    x = x.transpose(1, 2)
    out = continuous_transformer(x, sd, "transformer.", depth, prepend_embeds=prepend,
                                 context=cross_attn_cond, dim_heads=dim_heads, **extra)
    plen = 0 if prepend is None else prepend.shape[1]
    out = out.transpose(1, 2)[:, :, plen:]
    return F.conv1d(out, sd["postprocess_conv.weight"]) + out


def dit_forward(x, t, sd, depth, cross_attn_cond=None, global_embed=None, cfg_scale=1.0, scale_phi=0.0,
                global_cond_type="prepend", dim_heads=64, negative_cross_attn_cond=None, input_concat_cond=None):
    """DiffusionTransformer.forward, inference branch — dit.py:231-431 (CFG :324-410; the concatenated conditioning is the same for
    both halves of the CFG batch, :336-337)."""
    dt = sd["preprocess_conv.weight"].dtype
    x, t = x.to(dt), t.to(dt)
    if input_concat_cond is not None:
        input_concat_cond = input_concat_cond.to(dt)
    if cross_attn_cond is not None:
        cross_attn_cond = cross_attn_cond.to(dt)
    if global_embed is not None:
        global_embed = global_embed.to(dt)
    if cfg_scale != 1.0 and cross_attn_cond is not None:
        bx = torch.cat([x, x], 0)
        bt = torch.cat([t, t], 0)
        bg = None if global_embed is None else torch.cat([global_embed, global_embed], 0)
        null = torch.zeros_like(cross_attn_cond) if negative_cross_attn_cond is None else negative_cross_attn_cond.to(dt)
        bc = torch.cat([cross_attn_cond, null], 0)
        bi = None if input_concat_cond is None else torch.cat([input_concat_cond, input_concat_cond], 0)
        out = dit_inner(bx, bt, sd, depth, bc, bg, global_cond_type, dim_heads, input_concat_cond=bi)
        cond, uncond = out.chunk(2, 0)
        cfg = uncond + (cond - uncond) * cfg_scale
        if scale_phi != 0.0:
            cs = cond.std(dim=1, keepdim=True)
            os_ = cfg.std(dim=1, keepdim=True)
            return scale_phi * (cfg * (cs / os_)) + (1 - scale_phi) * cfg
        return cfg
    return dit_inner(x, t, sd, depth, cross_attn_cond, global_embed, global_cond_type, dim_heads, input_concat_cond=input_concat_cond)


# ---------------------------------------------------------------------------
# v-objective training-step arithmetic — training/diffusion.py:381-449, losses.py:66-91
def v_objective_loss(model_fn, x0, noise, t):
    alpha = torch.cos(t * math.pi / 2)[:, None, None]
    sigma = torch.sin(t * math.pi / 2)[:, None, None]
    noised = x0 * alpha + noise * sigma
    target = noise * alpha - x0 * sigma
    out = model_fn(noised, t)
    return F.mse_loss(out, target), out, target


def make_state_dict(embed_dim=1536, depth=24, num_heads=24, io_channels=64, cond_token_dim=768,
                    global_cond_dim=1536, global_cond_type="prepend", seed=0, std=0.02, dtype=torch.float32, input_concat_dim=0):
    """Seeded random DiT weights with the reference's names/shapes (dit.py:12-123, transformer.py).
    Zero-initialised reference weights (to_out, ff out, pre/post conv) are re-randomised so parity is
    not vacuous."""
    g = torch.Generator().manual_seed(seed)
    r = lambda *s, sc=std: (torch.randn(*s, generator=g) * sc).to(dtype)
    d, ff = embed_dim, embed_dim * 4
    dh = embed_dim // num_heads
    sd = {
        "timestep_features.weight": r(128, 1, sc=1.0),
        "to_timestep_embed.0.weight": r(d, 256), "to_timestep_embed.0.bias": r(d),
        "to_timestep_embed.2.weight": r(d, d), "to_timestep_embed.2.bias": r(d),
        "preprocess_conv.weight": r(io_channels + input_concat_dim, io_channels + input_concat_dim, 1, sc=0.05),      # dit.py:88, :115
        "postprocess_conv.weight": r(io_channels, io_channels, 1, sc=0.05),
        "transformer.project_in.weight": r(d, io_channels + input_concat_dim, sc=0.1),
        "transformer.project_out.weight": r(io_channels, d),
    }
    rot = max(dh // 2, 32)
    sd["transformer.rotary_pos_emb.inv_freq"] = (1.0 / (10000 ** (torch.arange(0, rot, 2).float() / rot))).to(dtype)
    if cond_token_dim > 0:
        sd["to_cond_embed.0.weight"] = r(cond_token_dim, cond_token_dim, sc=0.04)
        sd["to_cond_embed.2.weight"] = r(cond_token_dim, cond_token_dim, sc=0.04)
    if global_cond_dim > 0:
        sd["to_global_embed.0.weight"] = r(d, global_cond_dim)
        sd["to_global_embed.2.weight"] = r(d, d)
    if global_cond_type == "adaLN":
        sd["transformer.global_cond_embedder.0.weight"] = r(d, d)
        sd["transformer.global_cond_embedder.0.bias"] = r(d)
        sd["transformer.global_cond_embedder.2.weight"] = r(6 * d, d)
        sd["transformer.global_cond_embedder.2.bias"] = r(6 * d)
    for i in range(depth):
        p = f"transformer.layers.{i}."
        for n in ("pre_norm", "cross_attend_norm", "ff_norm"):
            if n == "cross_attend_norm" and cond_token_dim == 0:
                continue
            sd[p + n + ".gamma"] = (1.0 + 0.1 * torch.randn(d, generator=g)).to(dtype)
            sd[p + n + ".beta"] = torch.zeros(d, dtype=dtype)
        sd[p + "self_attn.to_qkv.weight"] = r(3 * d, d)
        sd[p + "self_attn.to_out.weight"] = r(d, d)
        if cond_token_dim > 0:
            sd[p + "cross_attn.to_q.weight"] = r(d, d)
            sd[p + "cross_attn.to_kv.weight"] = r(2 * cond_token_dim, cond_token_dim, sc=0.04)
            sd[p + "cross_attn.to_out.weight"] = r(d, d)
        sd[p + "ff.ff.0.proj.weight"] = r(2 * ff, d)
        sd[p + "ff.ff.0.proj.bias"] = r(2 * ff)
        sd[p + "ff.ff.2.weight"] = r(d, ff)
        sd[p + "ff.ff.2.bias"] = r(d)
        if global_cond_type == "adaLN":
            sd[p + "to_scale_shift_gate"] = (torch.randn(6 * d, generator=g) / d ** 0.5).to(dtype)
    return sd

"""Inference engine for the DiT denoiser: the reference's DiffusionTransformer.forward
(models/dit.py:231-431 -> _forward :125-229 -> ContinuousTransformer.forward transformer.py:796-865 ->
TransformerBlock.forward :658-713) as a fixed sequence of libb200sat kernel launches on preallocated buffers.

No host synchronisation and no allocation happen inside `forward_into`, so one denoising step is CUDA-graph
capturable (used by b200sat.sampling).  Weights are taken from a state dict with the reference's parameter names.
"""
import math
import torch

from . import ops


class DiTConfig:
    def __init__(self, io_channels=64, embed_dim=1536, depth=24, num_heads=24, cond_token_dim=768, global_cond_dim=1536,
                 global_cond_type="prepend", project_cond_tokens=False, project_global_cond=True, input_concat_dim=0):
        self.io_channels, self.embed_dim, self.depth, self.num_heads = io_channels, embed_dim, depth, num_heads
        # channels concatenated to the input (inpainting: mask + masked latents, dit.py:86-88); rows are padded to a multiple of 8 channels
        self.input_concat_dim = int(input_concat_dim)
        self.dim_in_pad = (io_channels + self.input_concat_dim + 7) // 8 * 8
        if self.input_concat_dim < 0 or self.dim_in_pad > 256:
            raise NotImplementedError("b200sat DiT engine: input_concat_dim must keep io_channels + input_concat_dim <= 256")
        self.cond_token_dim, self.global_cond_dim, self.global_cond_type = cond_token_dim, global_cond_dim, global_cond_type
        self.dim_heads = embed_dim // num_heads
        self.cond_embed_dim = cond_token_dim if not project_cond_tokens else embed_dim
        if self.dim_heads != 64:
            raise NotImplementedError("b200sat DiT engine: only dim_heads == 64 is implemented")
        if io_channels != 64:
            raise NotImplementedError("b200sat DiT engine: only io_channels == 64 is implemented")

    @staticmethod
    def from_state_dict(sd, num_heads=None):
        d = sd["transformer.project_in.weight"].shape[0]
        io = sd["transformer.project_out.weight"].shape[0]
        icd = sd["transformer.project_in.weight"].shape[1] - io
        depth = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("transformer.layers."))
        ctd = sd["to_cond_embed.0.weight"].shape[1] if "to_cond_embed.0.weight" in sd else 0
        gcd = sd["to_global_embed.0.weight"].shape[1] if "to_global_embed.0.weight" in sd else 0
        gct = "adaLN" if "transformer.global_cond_embedder.0.weight" in sd else "prepend"
        pct = ("to_cond_embed.0.weight" in sd) and sd["to_cond_embed.0.weight"].shape[0] != ctd
        return DiTConfig(io, d, depth, num_heads or d // 64, ctd, gcd, gct, project_cond_tokens=pct, input_concat_dim=icd)


class DiTEngine:
    def __init__(self, state_dict, config=None, device="cuda", fuse_layernorm=True):
        self.cfg = config or DiTConfig.from_state_dict(state_dict)
        self.device = torch.device(device)
        # prepend path: LayerNorm is folded into the consumer GEMM (gamma into the weights, mean/rstd correction in the epilogue,
        # row statistics produced by the previous residual GEMM's epilogue) — no LayerNorm pass over HBM at all
        self.fuse_ln = bool(fuse_layernorm) and self.cfg.global_cond_type == "prepend"
        self.w = {}
        self.load_state_dict(state_dict)
        self._ws = {}
        self._rope = {}

    # ------------------------------------------------------------------ weights
    def load_state_dict(self, sd):
        dev = self.device
        w = {}
        for k, v in sd.items():
            if not torch.is_tensor(v):
                continue
            v = v.detach()
            if k.endswith(("gamma", "bias", "inv_freq")) or k.endswith(".beta"):
                w[k] = v.to(dev, torch.float32).contiguous()
            elif k.endswith("to_scale_shift_gate"):
                w[k] = v.to(dev, torch.bfloat16).contiguous()
            elif k.endswith("conv.weight"):
                w[k] = v.to(dev, torch.bfloat16).reshape(v.shape[0], v.shape[1]).contiguous()
            elif k.endswith("weight"):
                w[k] = v.to(dev, torch.bfloat16).contiguous()
        # biases follow the parameter dtype of a bf16 model (values representable in bf16); the kernels read fp32
        for k in list(w):
            if k.endswith("bias"):
                w[k] = w[k].bfloat16().float().contiguous()
        # all layers' cross-attention K/V projections stacked: one GEMM per forward instead of `depth` small ones
        kvs = [w[f"transformer.layers.{i}.cross_attn.to_kv.weight"] for i in range(self.cfg.depth) if f"transformer.layers.{i}.cross_attn.to_kv.weight" in w]
        if len(kvs) == self.cfg.depth and self.cfg.depth > 0:
            w["_all_to_kv.weight"] = torch.cat(kvs, dim=0).contiguous()
        if self.cfg.input_concat_dim > 0:
            # preprocess_conv [dim_in, dim_in] and project_in [d, dim_in] zero-padded to the row pitch of the concatenated input
            cp, di = self.cfg.dim_in_pad, self.cfg.io_channels + self.cfg.input_concat_dim
            wp = torch.zeros(cp, cp, device=dev, dtype=torch.bfloat16)
            wp[:di, :di] = w["preprocess_conv.weight"]
            wi = torch.zeros(self.cfg.embed_dim, cp, device=dev, dtype=torch.bfloat16)
            wi[:, :di] = w["transformer.project_in.weight"]
            w["_preprocess_pad.weight"], w["_project_in_pad.weight"] = wp, wi
        self.ln_fold = {}
        if getattr(self, "fuse_ln", False):
            for i in range(self.cfg.depth):
                p = f"transformer.layers.{i}."
                for norm, lin in (("pre_norm", "self_attn.to_qkv"), ("cross_attend_norm", "cross_attn.to_q"), ("ff_norm", "ff.ff.0.proj")):
                    if (p + lin + ".weight") not in w or (p + norm + ".gamma") not in w:
                        self.fuse_ln = False
                        break
                    wf = (w[p + lin + ".weight"].float() * w[p + norm + ".gamma"][None, :]).to(torch.bfloat16).contiguous()
                    self.ln_fold[p + lin] = (wf, wf.float().sum(dim=1).contiguous())
        self.w = w

    def rope_tables(self, seq):
        if seq not in self._rope:
            inv = self.w["transformer.rotary_pos_emb.inv_freq"].bfloat16().float()  # buffer of a bf16 model
            t = torch.arange(seq, device=self.device, dtype=torch.float32)
            fr = torch.outer(t, inv)
            self._rope[seq] = (fr.cos().contiguous(), fr.sin().contiguous())
        return self._rope[seq]

    # ------------------------------------------------------------------ workspace
    def workspace(self, Bx, T, L):
        key = (Bx, T, L)
        if key in self._ws:
            return self._ws[key]
        c = self.cfg
        d = c.embed_dim
        P = 1 if c.global_cond_type == "prepend" else 0
        N = T + P
        M = Bx * N
        dev = self.device
        bf = lambda *s: torch.empty(*s, device=dev, dtype=torch.bfloat16)
        ws = dict(
            N=N, M=M, P=P,
            h=bf(M, d), n=bf(M, d), qkv=bf(M, 3 * d), a=bf(M, d), q=bf(M, d), ff=bf(M, 4 * d),
            xin=bf(Bx * T, c.io_channels), o=bf(M, c.io_channels),
            ff_feat=bf(Bx, 256), te1=bf(Bx, d), ge1=bf(Bx, d), ge=bf(Bx, d), gl=bf(Bx, d),
        )
        if c.input_concat_dim > 0:
            ws.update(xcat=bf(Bx * T, c.dim_in_pad), xin_cat=bf(Bx * T, c.dim_in_pad))
        if L > 0:
            ws.update(ctx_in=bf(Bx * L, c.cond_token_dim), ctx1=bf(Bx * L, c.cond_embed_dim), ctx=bf(Bx * L, c.cond_embed_dim),
                      kv=bf(Bx * L, 2 * c.cond_embed_dim), kv_all=bf(Bx * L, c.depth * 2 * c.cond_embed_dim))
        if c.global_cond_type == "adaLN":
            ws.update(g1=bf(Bx, d), g6=bf(Bx, 6 * d))
        if self.fuse_ln:
            ws["stats"] = torch.zeros(3 * c.depth + 1, M, 2, device=dev, dtype=torch.float32)
        self._ws[key] = ws
        return ws

    # ------------------------------------------------------------------ forward
    def precompute_conditioning(self, ctx_in, global_in, Bx, T):
        """The step-invariant part of the forward (dit.py:140-168, transformer.py:469-472): conditioning-token embedding, the cross-attention
        K/V of every layer and the global-embedding MLP depend only on the prompt, not on x or t, so a sampler computes them ONCE per run and
        replays its step with `cond_cached=True`.  Returns False (nothing cached) when the per-layer K/V weights are not stacked."""
        w = self.w
        if ctx_in is not None and "_all_to_kv.weight" not in w:
            return False
        L = 0 if ctx_in is None else ctx_in.shape[0] // Bx
        ws = self.workspace(Bx, T, L)
        if ctx_in is not None:
            ops.linear(ctx_in, w["to_cond_embed.0.weight"], silu=True, out=ws["ctx1"])
            ctx = ops.linear(ws["ctx1"], w["to_cond_embed.2.weight"], out=ws["ctx"])
            ops.linear(ctx, w["_all_to_kv.weight"], out=ws["kv_all"])
        if global_in is not None:
            self._lin_small(global_in, w["to_global_embed.0.weight"], None, ws["ge1"], silu=True)
            self._lin_small(ws["ge1"], w["to_global_embed.2.weight"], None, ws["ge"])
        return True

    def forward_into(self, out, x, t, ctx_in, global_in, Bx, reps, cfg, cfg_scale, scale_phi,
                     cin_table=None, t_table_step=None, step=None, cond_cached=False, concat_in=None):
        """out fp32 [B,C,T];  x fp32 [B,C,T];  t fp32 [Bx] (or a [steps, Bx] table walked by *step);
        ctx_in bf16 [Bx*L, cond_token_dim] or None;  global_in bf16 [Bx, global_cond_dim] or None.
        cond_cached: `precompute_conditioning` already filled this signature's ctx / kv_all / ge buffers for the same ctx_in / global_in.
        concat_in: fp32 [B, input_concat_dim, T] (models with input_concat_dim > 0; the same tensor serves both CFG halves, dit.py:336-337)."""
        c, w = self.cfg, self.w
        d, H = c.embed_dim, c.num_heads
        B, C, T = x.shape
        L = 0 if ctx_in is None else ctx_in.shape[0] // Bx
        ws = self.workspace(Bx, T, L)
        N, M, P = ws["N"], ws["M"], ws["P"]
        h = ws["h"]
        # --- conditioning (dit.py:140-168)
        ctx = None
        if ctx_in is not None and cond_cached:
            ctx = ws["ctx"]
        elif ctx_in is not None:
            ops.linear(ctx_in, w["to_cond_embed.0.weight"], silu=True, out=ws["ctx1"])
            ctx = ops.linear(ws["ctx1"], w["to_cond_embed.2.weight"], out=ws["ctx"])
            if "_all_to_kv.weight" in w:   # K/V of the conditioning tokens for every layer (transformer.py:469-472), one launch
                ops.linear(ctx, w["_all_to_kv.weight"], out=ws["kv_all"])
        ops.fourier_features(t, w["timestep_features.weight"], out=ws["ff_feat"], step=step,
                             t_stride=Bx if step is not None else 0)
        self._lin_small(ws["ff_feat"], w["to_timestep_embed.0.weight"], w["to_timestep_embed.0.bias"], ws["te1"], silu=True)
        gdst = h.view(Bx, N, d)[:, 0, :] if P else ws["gl"]
        fuse = self.fuse_ln and P == 1 and Bx <= 8
        st = ws["stats"] if fuse else None
        if fuse:
            st.zero_()
        st0 = dict(stats=st[0], stats_stride=2 * N) if fuse else {}
        if global_in is not None:
            if not cond_cached:
                self._lin_small(global_in, w["to_global_embed.0.weight"], None, ws["ge1"], silu=True)
                self._lin_small(ws["ge1"], w["to_global_embed.2.weight"], None, ws["ge"])
            self._lin_small(ws["te1"], w["to_timestep_embed.2.weight"], w["to_timestep_embed.2.bias"], gdst, add=ws["ge"], **st0)
        else:
            self._lin_small(ws["te1"], w["to_timestep_embed.2.weight"], w["to_timestep_embed.2.bias"], gdst, **st0)
        # --- input stage (dit.py:193-195, transformer.py:811-819)
        if c.input_concat_dim > 0:
            if concat_in is None:
                raise ValueError("this DiT was built with input_concat_dim > 0: input_concat_cond is required (dit.py:160-165)")
            ops.dit_concat(x, concat_in, ws["xcat"], reps=reps, cin_table=cin_table, step=step)
            ops.linear(ws["xcat"], w["_preprocess_pad.weight"], residual=ws["xcat"], out=ws["xin_cat"])      # preprocess_conv(x) + x, dit.py:193
            ops.linear(ws["xin_cat"], w["_project_in_pad.weight"], out=h, row_remap=(T, N, P), out_stats=st[0] if fuse else None)
        else:
            if concat_in is not None:
                raise ValueError("input_concat_cond given to a DiT without input_concat_dim")
            ops.dit_pre(x, w["preprocess_conv.weight"], ws["xin"], reps=reps, cin_table=cin_table, step=step)
            ops.linear(ws["xin"], w["transformer.project_in.weight"], out=h, row_remap=(T, N, P), out_stats=st[0] if fuse else None)
        cos, sin = self.rope_tables(N)
        # --- adaLN modulation tables (transformer.py:836-837, :677)
        mods = None
        if c.global_cond_type == "adaLN":
            self._lin_small(ws["gl"], w["transformer.global_cond_embedder.0.weight"], w["transformer.global_cond_embedder.0.bias"], ws["g1"], silu=True)
            self._lin_small(ws["g1"], w["transformer.global_cond_embedder.2.weight"], w["transformer.global_cond_embedder.2.bias"], ws["g6"])
            mods = []
            for i in range(c.depth):
                m6 = (w[f"transformer.layers.{i}.to_scale_shift_gate"][None, :] + ws["g6"])  # bf16 [Bx, 6d]
                gates = torch.sigmoid((1 - m6).float())
                mods.append((m6.float().contiguous(), gates.contiguous()))
        rope = (cos, sin, N, d, c.dim_heads)
        for i in range(c.depth):
            p = f"transformer.layers.{i}."
            if mods is None:
                sc_s = sh_s = sc_f = sh_f = g_s = g_f = None
            else:
                m6, gt = mods[i]
                sc_s, sh_s, g_s = m6[:, 0:d], m6[:, d:2 * d], gt[:, 2 * d:3 * d].contiguous()
                sc_f, sh_f, g_f = m6[:, 3 * d:4 * d], m6[:, 4 * d:5 * d], gt[:, 5 * d:6 * d].contiguous()
            if fuse:
                # LayerNorm-free block: raw residual stream in, statistics travel through the GEMM epilogues
                wq, cq = self.ln_fold[p + "self_attn.to_qkv"]
                ops.linear(h, wq, out=ws["qkv"], rope=rope, ln=(st[3 * i], cq, 1e-5))
                qkv = ws["qkv"].view(Bx, N, 3, H, 64)
                ops.attention(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], out=ws["a"].view(Bx, N, H, 64))
                ops.linear(ws["a"], w[p + "self_attn.to_out.weight"], residual=h, out=h, out_stats=st[3 * i + 1])
                if ctx is not None and (p + "cross_attn.to_q.weight") in w:
                    wq2, cq2 = self.ln_fold[p + "cross_attn.to_q"]
                    ops.linear(h, wq2, out=ws["q"], ln=(st[3 * i + 1], cq2, 1e-5))
                    kvh = c.cond_embed_dim // 64
                    kv = ws["kv_all"].view(Bx, L, c.depth, 2, kvh, 64)[:, :, i]
                    ops.attention(ws["q"].view(Bx, N, H, 64), kv[:, :, 0], kv[:, :, 1], out=ws["a"].view(Bx, N, H, 64))
                    ops.linear(ws["a"], w[p + "cross_attn.to_out.weight"], residual=h, out=h, out_stats=st[3 * i + 2])
                    ff_stats = st[3 * i + 2]
                else:
                    ff_stats = st[3 * i + 1]
                wf, cf = self.ln_fold[p + "ff.ff.0.proj"]
                ops.linear(h, wf, bias=w[p + "ff.ff.0.proj.bias"], swiglu=True, out=ws["ff"], ln=(ff_stats, cf, 1e-5))
                ops.linear(ws["ff"], w[p + "ff.ff.2.weight"], bias=w[p + "ff.ff.2.bias"], residual=h, out=h, out_stats=st[3 * i + 3])
                continue
            # self-attention (transformer.py:704 / :678-686)
            ops.layernorm(h, w[p + "pre_norm.gamma"], scale=sc_s, shift=sh_s, rows_per_batch=N, out=ws["n"])
            ops.linear(ws["n"], w[p + "self_attn.to_qkv.weight"], out=ws["qkv"], rope=rope)
            qkv = ws["qkv"].view(Bx, N, 3, H, 64)
            ops.attention(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], out=ws["a"].view(Bx, N, H, 64))
            ops.linear(ws["a"], w[p + "self_attn.to_out.weight"], residual=h, out=h, gate=g_s,
                       row_remap=(N, 0, 0) if g_s is not None else None)
            # cross-attention (transformer.py:706-707)
            if ctx is not None and (p + "cross_attn.to_q.weight") in w:
                ops.layernorm(h, w[p + "cross_attend_norm.gamma"], out=ws["n"])
                ops.linear(ws["n"], w[p + "cross_attn.to_q.weight"], out=ws["q"])
                kvh = c.cond_embed_dim // 64
                if "_all_to_kv.weight" in w:
                    kv = ws["kv_all"].view(Bx, L, c.depth, 2, kvh, 64)[:, :, i]
                else:
                    ops.linear(ctx, w[p + "cross_attn.to_kv.weight"], out=ws["kv"])
                    kv = ws["kv"].view(Bx, L, 2, kvh, 64)
                ops.attention(ws["q"].view(Bx, N, H, 64), kv[:, :, 0], kv[:, :, 1], out=ws["a"].view(Bx, N, H, 64))
                ops.linear(ws["a"], w[p + "cross_attn.to_out.weight"], residual=h, out=h)
            # feed-forward (transformer.py:712 / :693-701)
            ops.layernorm(h, w[p + "ff_norm.gamma"], scale=sc_f, shift=sh_f, rows_per_batch=N, out=ws["n"])
            ops.linear(ws["n"], w[p + "ff.ff.0.proj.weight"], bias=w[p + "ff.ff.0.proj.bias"], swiglu=True, out=ws["ff"])
            ops.linear(ws["ff"], w[p + "ff.ff.2.weight"], bias=w[p + "ff.ff.2.bias"], residual=h, out=h, gate=g_f,
                       row_remap=(N, 0, 0) if g_f is not None else None)
        # --- output stage (transformer.py:859, dit.py:219-224, :398-408)
        ops.linear(h, w["transformer.project_out.weight"], out=ws["o"])
        ops.dit_post(ws["o"], N * c.io_channels, P, w["postprocess_conv.weight"], out, cfg=cfg, cfg_scale=cfg_scale,
                     scale_phi=scale_phi)
        return out

    def _lin_small(self, x, wt, bias, out, silu=False, add=None, stats=None, stats_stride=0):
        if x.shape[0] <= 8:
            ops.small_linear(x, wt, bias=bias, add=add, out=out, silu=silu, stats=stats, stats_stride=stats_stride)
        else:
            if out.stride(0) != out.shape[1] or add is not None and silu:
                raise NotImplementedError("conditioning MLP with batch > 8 into a strided destination")
            ops.linear(x, wt, bias=bias, residual=add, out=out, silu=silu)

    @torch.no_grad()
    def forward(self, x, t, cross_attn_cond=None, global_embed=None, cfg_scale=1.0, scale_phi=0.0,
                negative_cross_attn_cond=None, out=None, input_concat_cond=None):
        """Mirror of DiffusionTransformer.forward's inference branch (dit.py:231-431)."""
        B, C, T = x.shape
        dev = self.device
        x = x.to(dev, torch.float32).contiguous()
        cfg = cfg_scale != 1.0 and cross_attn_cond is not None
        reps = 2 if cfg else 1
        Bx = B * reps
        t = t.to(dev, torch.float32).reshape(B)
        ctx_in = None
        if cross_attn_cond is not None:
            cc = cross_attn_cond.to(dev, torch.bfloat16)
            if cfg:
                null = torch.zeros_like(cc) if negative_cross_attn_cond is None else negative_cross_attn_cond.to(dev, torch.bfloat16)
                cc = torch.cat([cc, null], 0)
            ctx_in = cc.reshape(Bx * cc.shape[1], cc.shape[2]).contiguous()
        g_in = None
        if global_embed is not None:
            g_in = global_embed.to(dev, torch.bfloat16)
            if cfg:
                g_in = torch.cat([g_in, g_in], 0)
            g_in = g_in.contiguous()
        if cfg:
            t = torch.cat([t, t], 0)
        t = t.contiguous()
        if out is None:
            out = torch.empty(B, C, T, device=dev, dtype=torch.float32)
        concat = None
        if input_concat_cond is not None:
            concat = self.prepare_concat(input_concat_cond, T)
        return self.forward_into(out, x, t, ctx_in, g_in, Bx, reps, cfg, float(cfg_scale), float(scale_phi), concat_in=concat)

    def prepare_concat(self, input_concat_cond, T):
        """dit.py:160-163 + :268-269: nearest-neighbour resize to the input length, cast to the model dtype (bf16), kept as fp32 values."""
        cc = input_concat_cond.to(self.device, torch.float32)
        if cc.shape[2] != T:
            cc = torch.nn.functional.interpolate(cc, (T,), mode="nearest")
        return cc.bfloat16().float().contiguous()

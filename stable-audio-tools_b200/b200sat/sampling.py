"""Sampling loops: the reference's `sample_k(..., sampler_type="dpmpp-3m-sde")` (inference/sampling.py:331-387, which calls
k-diffusion==0.1.1 `VDenoiser`, `get_sigmas_polyexponential`, `sample_dpmpp_3m_sde`) and the in-repo deterministic
v-DDIM `sample` (inference/sampling.py:253-307).

B200-first structure: every per-step scalar (c_in, c_out, c_skip, the multistep coefficients, the model timestep) is
known on the host before the loop starts, so they are precomputed into small device tables; one denoising step
(DiT forward + CFG + state update) is captured ONCE into a CUDA graph that reads a device-side step counter, and the
host loop only replays it.  No host<->device synchronisation happens inside the loop.
"""
import math
import torch

from . import ops


def get_sigmas_polyexponential(n, sigma_min, sigma_max, rho=1.0):
    """k_diffusion.sampling.get_sigmas_polyexponential: exp(linspace(1,0,n)**rho * (ln smax - ln smin) + ln smin) ++ [0]."""
    ramp = torch.linspace(1, 0, n, dtype=torch.float32) ** rho
    sig = torch.exp(ramp * (math.log(sigma_max) - math.log(sigma_min)) + math.log(sigma_min))
    return torch.cat([sig, sig.new_zeros(1)])


def dpmpp_3m_sde_tables(sigmas, eta=1.0, s_noise=1.0):
    """Per-step coefficient rows [c_out, c_skip, A, Bd, C1, C2, NZ, 0] such that
         den = v*c_out + x*c_skip ;  x' = A*x + Bd*den + C1*den_{-1} + C2*den_{-2} + NZ*noise
       reproduces VDenoiser + sample_dpmpp_3m_sde exactly (float64 arithmetic on the host).
       Also returns c_in[i] = 1/sqrt(sigma^2+1) and the model timestep t[i] = atan(sigma)*2/pi."""
    s = [float(v) for v in sigmas]
    n = len(s) - 1
    coef = torch.zeros(n, 8, dtype=torch.float64)
    cin = torch.zeros(n, dtype=torch.float64)
    tt = torch.zeros(n, dtype=torch.float64)
    h1 = h2 = None
    for i in range(n):
        sg, sn = s[i], s[i + 1]
        cin[i] = 1.0 / math.sqrt(sg * sg + 1.0)
        tt[i] = math.atan(sg) * 2.0 / math.pi
        coef[i, 0] = -sg / math.sqrt(sg * sg + 1.0)
        coef[i, 1] = 1.0 / (sg * sg + 1.0)
        if sn == 0:
            coef[i, 2], coef[i, 3] = 0.0, 1.0
            h = None
        else:
            t_, s_ = -math.log(sg), -math.log(sn)
            h = s_ - t_
            he = h * (eta + 1.0)
            A = math.exp(-he)
            Bd = -math.expm1(-he)
            C1 = C2 = 0.0
            if h2 is not None:
                r0, r1 = h1 / h, h2 / h
                phi2 = math.expm1(-he) / he + 1.0
                phi3 = phi2 / he - 0.5
                ca = phi2 * (1.0 + r0 / (r0 + r1)) - phi3 / (r0 + r1)
                cb = -phi2 * r0 / (r0 + r1) + phi3 / (r0 + r1)
                Bd += ca / r0
                C1 = -ca / r0 + cb / r1
                C2 = -cb / r1
            elif h1 is not None:
                r = h1 / h
                phi2 = math.expm1(-he) / he + 1.0
                Bd += phi2 / r
                C1 = -phi2 / r
            NZ = sn * math.sqrt(-math.expm1(-2.0 * h * eta)) * s_noise if eta else 0.0
            coef[i, 2:7] = torch.tensor([A, Bd, C1, C2, NZ], dtype=torch.float64)
        h1, h2 = h, h1
    return coef.float(), cin.float(), tt.float()


def v_ddim_tables(steps, sigma_max=1.0):
    """In-repo v-DDIM `sample` with eta = 0 (inference/sampling.py:253-307) in the same table form:
       pred = x*alpha - v*sigma (= den with c_out=-sigma, c_skip=alpha);  x' = (s'/s)*x + (a' - s'*a/s)*pred; last: pred."""
    t = torch.linspace(sigma_max, 0, steps + 1, dtype=torch.float64)[:-1]
    al, sg = torch.cos(t * math.pi / 2), torch.sin(t * math.pi / 2)
    coef = torch.zeros(steps, 8, dtype=torch.float64)
    for i in range(steps):
        coef[i, 0], coef[i, 1] = -sg[i], al[i]
        if i < steps - 1:
            coef[i, 2] = sg[i + 1] / sg[i]
            coef[i, 3] = al[i + 1] - sg[i + 1] * al[i] / sg[i]
        else:
            coef[i, 2], coef[i, 3] = 0.0, 1.0
    return coef.float(), torch.ones(steps), t.float()


class GraphSampler:
    """Runs `steps` denoising steps of a DiTEngine for a fixed (batch, length, conditioning-shape) signature."""

    def __init__(self, engine, B, C, T, L, has_global=True, cfg_scale=1.0, scale_phi=0.0, use_graph=True):
        self.e = engine
        dev = engine.device
        self.B, self.C, self.T, self.L = B, C, T, L
        self.cfg = cfg_scale != 1.0 and L > 0
        self.cfg_scale, self.scale_phi = float(cfg_scale), float(scale_phi)
        self.reps = 2 if self.cfg else 1
        self.Bx = B * self.reps
        self.x = torch.zeros(B, C, T, device=dev)
        self.v = torch.zeros(B, C, T, device=dev)
        self.hist = torch.zeros(3, B * C * T, device=dev)
        self.step = torch.zeros(1, dtype=torch.int32, device=dev)
        self.ctx = torch.zeros(self.Bx * L, engine.cfg.cond_token_dim, device=dev, dtype=torch.bfloat16) if L > 0 else None
        self.glob = torch.zeros(self.Bx, engine.cfg.global_cond_dim, device=dev, dtype=torch.bfloat16) if has_global else None
        # channel-concatenated conditioning (inpainting mask + masked latents): constant over the run, same for both CFG halves
        dc = engine.cfg.input_concat_dim
        self.concat = torch.zeros(B, dc, T, device=dev) if dc > 0 else None
        self.use_graph = use_graph
        # prompt-only work (conditioning embedding, every layer's cross-attention K/V, global-embedding MLP) runs once per run(), not per step
        self.cond_cached = L == 0 or "_all_to_kv.weight" in engine.w
        self.graph = None
        self.max_steps = 0
        self.coef = self.cin = self.tt = self.noise = None
        self.launches_per_step = 0

    def _alloc_tables(self, steps):
        if steps > self.max_steps:
            dev = self.e.device
            self.coef = torch.zeros(steps, 8, device=dev)
            self.cin = torch.zeros(steps, device=dev)
            self.tt = torch.zeros(steps, self.Bx, device=dev)
            self.noise = torch.zeros(steps, self.B * self.C * self.T, device=dev)
            self.max_steps = steps
            self.graph = None  # table pointers changed

    def _one_step(self):
        self.e.forward_into(self.v, self.x, self.tt, self.ctx, self.glob, self.Bx, self.reps, self.cfg, self.cfg_scale,
                            self.scale_phi, cin_table=self.cin, step=self.step, cond_cached=self.cond_cached, concat_in=self.concat)
        ops.sampler_update(self.x, self.v, self.hist, self.noise, self.coef, self.step, advance=True)

    def _ensure_graph(self):
        if not self.use_graph or self.graph is not None:
            return
        # warm-up on a side stream (sets kernel attributes, fills caches), then capture
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            ops.step_set(self.step, 0)
            n0 = ops.LAUNCHES[0]
            self._one_step()
            self.launches_per_step = ops.LAUNCHES[0] - n0
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._one_step()
        self.graph = g

    @torch.no_grad()
    def run(self, noise, coef, cin, tt, cross_attn_cond=None, global_embed=None, step_noise=None, init_scale=1.0,
            negative_cross_attn_cond=None, input_concat_cond=None, init_data=None, init_data_scale=1.0):
        """noise [B,C,T] (unit variance), tables from *_tables(); step_noise [steps,B,C,T] or None (drawn with torch.randn).
        Host tensors are accepted (copied with non_blocking=True from pinned memory).
        input_concat_cond [B, input_concat_dim, T'] for models built with input_concat_dim (resized like dit.py:160-163);
        init_data: start from init_data*init_data_scale + noise*init_scale (sampling.py:360-366 / :396-399: variations)."""
        steps = coef.shape[0]
        dev = self.e.device
        self._alloc_tables(steps)
        self._ensure_graph()  # the warm-up step scribbles on x/hist/step: do it before loading this run's state
        self.coef[:steps].copy_(coef, non_blocking=True)
        self.cin[:steps].copy_(cin, non_blocking=True)
        self.tt[:steps].copy_(tt.to(torch.float32)[:, None].expand(steps, self.Bx), non_blocking=True)
        if self.ctx is not None:
            cc = cross_attn_cond.to(dev, torch.bfloat16, non_blocking=True).reshape(self.B * self.L, -1)
            self.ctx[: self.B * self.L].copy_(cc)
            if self.cfg:
                if negative_cross_attn_cond is None:
                    self.ctx[self.B * self.L:].zero_()
                else:
                    self.ctx[self.B * self.L:].copy_(negative_cross_attn_cond.to(dev, torch.bfloat16).reshape(self.B * self.L, -1))
        if self.glob is not None:
            gg = global_embed.to(dev, torch.bfloat16, non_blocking=True)
            self.glob[: self.B].copy_(gg)
            if self.cfg:
                self.glob[self.B:].copy_(gg)
        if self.cond_cached:
            self.e.precompute_conditioning(self.ctx, self.glob, self.Bx, self.T)
        if self.concat is not None:
            if input_concat_cond is None:
                raise ValueError("this DiT was built with input_concat_dim > 0: input_concat_cond is required")
            self.concat.copy_(self.e.prepare_concat(input_concat_cond, self.T))
        elif input_concat_cond is not None:
            raise ValueError("input_concat_cond given to a DiT without input_concat_dim")
        self.x.copy_(noise.to(dev, torch.float32, non_blocking=True))
        if init_scale != 1.0:
            self.x.mul_(init_scale)
        if init_data is not None:
            self.x.add_(init_data.to(dev, torch.float32), alpha=float(init_data_scale))
        if step_noise is None:
            if bool((coef[:, 6] != 0).any()):
                torch.randn(self.noise[:steps].shape, out=self.noise[:steps])
        else:
            self.noise[:steps].copy_(step_noise.to(dev, torch.float32, non_blocking=True).reshape(steps, -1))
        self.hist.zero_()
        ops.step_set(self.step, 0)
        if self.graph is not None:
            for _ in range(steps):
                self.graph.replay()
            ops.LAUNCHES[0] += steps * self.launches_per_step  # kernels replayed from the captured step
        else:
            for _ in range(steps):
                self._one_step()
        return self.x


def sample_k_dpmpp_3m_sde(engine, noise, steps=100, sigma_min=0.03, sigma_max=1000.0, rho=1.0, cross_attn_cond=None,
                          global_embed=None, cfg_scale=1.0, scale_phi=0.0, eta=1.0, s_noise=1.0, step_noise=None,
                          sampler=None, use_graph=True, negative_cross_attn_cond=None, input_concat_cond=None, init_data=None):
    """`sample_k(model_fn, noise, steps=..., sampler_type='dpmpp-3m-sde', ...)` for a DiTEngine (inference/sampling.py:331-387)."""
    B, C, T = noise.shape
    L = 0 if cross_attn_cond is None else cross_attn_cond.shape[1]
    sig = get_sigmas_polyexponential(steps, sigma_min, sigma_max, rho)
    coef, cin, tt = dpmpp_3m_sde_tables(sig, eta, s_noise)
    if sampler is None:
        sampler = GraphSampler(engine, B, C, T, L, global_embed is not None, cfg_scale, scale_phi, use_graph)
    return sampler.run(noise, coef, cin, tt, cross_attn_cond, global_embed, step_noise, init_scale=float(sig[0]),
                       negative_cross_attn_cond=negative_cross_attn_cond, input_concat_cond=input_concat_cond, init_data=init_data).clone()


def sample_v_ddim(engine, noise, steps=100, sigma_max=1.0, cross_attn_cond=None, global_embed=None, cfg_scale=1.0,
                  scale_phi=0.0, sampler=None, use_graph=True, negative_cross_attn_cond=None, input_concat_cond=None, init_data=None):
    """`sample_k(..., sampler_type='v-ddim')` -> in-repo `sample(model, x, steps, eta=0)` (inference/sampling.py:253-307,:405-407)."""
    B, C, T = noise.shape
    L = 0 if cross_attn_cond is None else cross_attn_cond.shape[1]
    coef, cin, tt = v_ddim_tables(steps, min(sigma_max, 1.0))
    if sampler is None:
        sampler = GraphSampler(engine, B, C, T, L, global_embed is not None, cfg_scale, scale_phi, use_graph)
    sm = min(sigma_max, 1.0)
    a0, s0 = (math.cos(sm * math.pi / 2), math.sin(sm * math.pi / 2)) if init_data is not None else (1.0, 1.0)   # sampling.py:394-399
    return sampler.run(noise, coef, cin, tt, cross_attn_cond, global_embed, None, negative_cross_attn_cond=negative_cross_attn_cond,
                       input_concat_cond=input_concat_cond, init_scale=s0, init_data=init_data, init_data_scale=a0).clone()

"""Oobleck autoencoder training pass: encoder -> VAE bottleneck -> decoder forward with saved activations, and the full backward
(data, weight, bias, SnakeBeta alpha/beta and weight-norm g/v gradients) as libb200sat launches.

Reference: `AutoencoderTrainingWrapper.training_step` generator branch (training/autoencoders.py:367-527) on an `AudioAutoencoder`
built from OobleckEncoder / VAEBottleneck / OobleckDecoder (models/autoencoders.py:285-362, models/bottleneck.py:105-134), bf16
autocast.  Parameters keep the reference names and shapes (`encoder.layers...weight_g / weight_v / bias / alpha / beta`).

Layout: activations and their gradients are time-major bf16 planes [B, T, C] (csrc/conv1d.cu).  Backward of one conv node
(y_raw = conv(x_act) [+ skip];  y_act = snake(y_raw)):
    d_raw  = snake'(y_raw) * d_act + d_skip          b200sat_snake_bwd  (also d alpha, d beta, d bias)
    dW_tap = d_raw^T (x) x_act shifted by the tap    b200sat_conv_wgrad (tcgen05, both operands read in place)  -> b200sat_wn_bwd
    d_x    = conv with flipped / transposed weights  b200sat_wn_pack_dgrad + b200sat_conv1d_fwd
The reference re-computes every ResidualUnit under `checkpoint` (autoencoders.py:78-79); here the raw planes are kept instead
(about 26 GB at 32 x 65536 samples), which removes the second forward.
"""
import math
import torch
from torch import nn

from ._lib import lib, check
from . import ops
from .autoencoder import _Planes, _p, _s


class _TConv:
    """One weight-normed conv for the duration of a step: parameter tensors + packed weights."""

    def __init__(self, params, prefix, transposed, stride):
        self.prefix = prefix
        self.v = params[prefix + "weight_v"]
        self.g = params[prefix + "weight_g"]
        self.bias = params.get(prefix + "bias")
        self.transposed, self.stride = transposed, stride
        if transposed:
            self.Cin, self.Cout, self.K = self.v.shape
        else:
            self.Cout, self.Cin, self.K = self.v.shape
        self.small = (self.Cin % 64 != 0) or (self.Cout % 64 != 0)
        self.w_hi = self.inv_norm = self.w_f32 = None

    def pack(self):
        dev = self.v.device
        if self.small:   # 2-channel edge convs: fp32 dense weights for the SIMT kernels (1792 values)
            n = self.v.detach().flatten(1).norm(dim=1).view(-1, 1, 1)
            self.w_f32 = (self.g.detach().view(-1, 1, 1) * self.v.detach() / n).contiguous()
            return
        rows, cols = (self.stride * self.Cout, 2 * self.Cin) if self.transposed else (self.Cout, self.K * self.Cin)
        self.w_hi = torch.empty(rows, cols, device=dev, dtype=torch.bfloat16)
        self.inv_norm = torch.empty(self.v.shape[0], device=dev, dtype=torch.float32)
        check(lib().b200sat_wn_pack(self.v.data_ptr(), self.g.data_ptr(), self.inv_norm.data_ptr(), self.w_hi.data_ptr(), 0, self.Cout,
                                    self.Cin, self.K, int(self.transposed), self.stride, _s()), "wn_pack")
        ops.LAUNCHES[0] += 2

    def pack_dgrad(self, mode):
        out = torch.empty(self.Cout * self.Cin * self.K, device=self.v.device, dtype=torch.bfloat16)
        check(lib().b200sat_wn_pack_dgrad(self.v.data_ptr(), self.g.data_ptr(), self.inv_norm.data_ptr(), out.data_ptr(), self.Cout, self.Cin,
                                          self.K, mode, self.stride, _s()), "wn_pack_dgrad")
        ops.LAUNCHES[0] += 1
        return out


class _TSnake:
    def __init__(self, params, prefix):
        self.prefix = prefix
        self.alpha, self.beta = params[prefix + "alpha"], params[prefix + "beta"]
        self.a = torch.empty_like(self.alpha)
        self.invb = torch.empty_like(self.beta)
        check(lib().b200sat_snake_prep(self.alpha.data_ptr(), self.beta.data_ptr(), self.a.data_ptr(), self.invb.data_ptr(),
                                       self.alpha.numel(), _s()), "snake_prep")
        ops.LAUNCHES[0] += 1


def _plane(B, T, C, dev):
    return _Planes(B, T, C, dev, False)


class _Step:
    """Forward with a tape, then backward over it."""

    def __init__(self, params, strides, final_tanh=False):
        self.P = params
        self.strides = tuple(strides)
        self.dev = next(iter(params.values())).device
        self.tape = []
        if final_tanh:
            raise NotImplementedError("final_tanh decoder is not supported by the training pass")

    # ------------------------------------------------------------------ forward
    def _conv(self, x, conv, out=None, act=None, snake=None, res=None, dil=1, pad=0, mode=0):
        check(lib().b200sat_conv1d_fwd(x.hi.data_ptr(), 0, conv.w_hi.data_ptr(), 0, _p(conv.bias), _p(res.hi) if res else 0, 0,
                                       _p(out.hi) if out else 0, 0, _p(act.hi) if act else 0, 0, _p(snake.a) if snake else 0,
                                       _p(snake.invb) if snake else 0, x.B, x.T, x.C, conv.Cout, conv.K, dil, pad, conv.stride, mode, 1,
                                       _s()), "conv1d_fwd")
        ops.LAUNCHES[0] += 1

    def _tc(self, prefix, transposed=False, stride=1):
        c = _TConv(self.P, prefix, transposed, stride)
        c.pack()
        return c

    def _node(self, conv, x, x_src, y_raw, snake, res=None, dil=1, pad=0, mode=0, kind="conv"):
        self.tape.append(dict(conv=conv, x=x, src=x_src, y_raw=y_raw, snake=snake, res=res, dil=dil, pad=pad, mode=mode, kind=kind))
        return len(self.tape) - 1

    def _rus(self, raw, act, last, prefix, first_ru, next_snake):
        """three ResidualUnits (autoencoders.py:58-83); `last` = tape index of the node that produced (raw, act)."""
        B, T, C = raw.B, raw.T, raw.C
        P = self.P
        s0s = [None, _TSnake(P, f"{prefix}{first_ru + 1}.layers.0."), _TSnake(P, f"{prefix}{first_ru + 2}.layers.0.")]
        for j in range(3):
            q = f"{prefix}{first_ru + j}.layers."
            dil = (1, 3, 9)[j]
            c7, s1, c1 = self._tc(q + "1."), _TSnake(P, q + "2."), self._tc(q + "3.")
            h_raw, h_act = _plane(B, T, C, self.dev), _plane(B, T, C, self.dev)
            self._conv(act, c7, out=h_raw, act=h_act, snake=s1, dil=dil, pad=3 * dil)
            i7 = self._node(c7, act, last, h_raw, s1, dil=dil, pad=3 * dil)
            nxt = s0s[j + 1] if j < 2 else next_snake
            y_raw, y_act = _plane(B, T, C, self.dev), _plane(B, T, C, self.dev)
            self._conv(h_act, c1, out=y_raw, act=y_act, snake=nxt, res=raw)
            last = self._node(c1, h_act, i7, y_raw, nxt, res=last)
            raw, act = y_raw, y_act
        return raw, act, last

    def forward(self, audio, noise):
        P, dev = self.P, self.dev
        n = len(self.strides)
        x = audio.to(dev, torch.float32).contiguous()
        B, Cin, T = x.shape
        self.B = B
        self.audio = x
        # ---- encoder (autoencoders.py:285-317)
        pe = "encoder.layers."
        c0 = self._tc(pe + "0.")
        snk = _TSnake(P, f"{pe}1.layers.0.layers.0.")
        raw, act = _plane(B, T, c0.Cout, dev), _plane(B, T, c0.Cout, dev)
        check(lib().b200sat_conv_in(x.data_ptr(), c0.w_f32.data_ptr(), _p(c0.bias), snk.a.data_ptr(), snk.invb.data_ptr(), raw.hi.data_ptr(), 0,
                                    act.hi.data_ptr(), 0, B, Cin, T, c0.Cout, c0.K, c0.K // 2, _s()), "conv_in")
        ops.LAUNCHES[0] += 1
        last = self._node(c0, None, None, raw, snk, kind="conv_in")
        for i, s in enumerate(self.strides):
            q = f"{pe}{i + 1}.layers."
            blk_snake = _TSnake(P, q + "3.")
            raw, act, last = self._rus(raw, act, last, q, 0, blk_snake)
            down = self._tc(q + "4.", False, s)
            nxt = _TSnake(P, f"{pe}{i + 2}.layers.0.layers.0.") if i + 1 < n else _TSnake(P, f"{pe}{n + 1}.")
            pad = math.ceil(s / 2)
            T2 = (raw.T + 2 * pad - down.K) // s + 1
            o_raw, o_act = _plane(B, T2, down.Cout, dev), _plane(B, T2, down.Cout, dev)
            self._conv(act, down, out=o_raw, act=o_act, snake=nxt, pad=pad, mode=1)
            last = self._node(down, act, last, o_raw, nxt, pad=pad, mode=1)
            raw, act = o_raw, o_act
        co = self._tc(f"{pe}{n + 2}.")
        ms = _plane(B, act.T, co.Cout, dev)
        self._conv(act, co, out=ms, pad=co.K // 2)
        self.enc_out = self._node(co, act, last, ms, None, pad=co.K // 2)
        # ---- VAE bottleneck (bottleneck.py:105-134)
        L, Tl = co.Cout // 2, ms.T
        self.L, self.Tl, self.ms = L, Tl, ms
        self.noise = None if noise is None else noise.to(dev, torch.float32).contiguous()
        z = torch.empty(B, L, Tl, device=dev)
        kl = torch.zeros(1, device=dev)
        check(lib().b200sat_vae_sample(ms.hi.data_ptr(), 0, _p(self.noise), z.data_ptr(), 0, kl.data_ptr(), B, L, Tl, _s()), "vae_sample")
        ops.LAUNCHES[0] += 1
        # ---- decoder (autoencoders.py:320-362)
        pd = "decoder.layers."
        zp = _plane(B, Tl, L, dev)
        check(lib().b200sat_to_planes(z.data_ptr(), zp.hi.data_ptr(), 0, B, L, Tl, _s()), "to_planes")
        ops.LAUNCHES[0] += 1
        d0 = self._tc(pd + "0.")
        snk = _TSnake(P, f"{pd}1.layers.0.")
        raw, act = _plane(B, Tl, d0.Cout, dev), _plane(B, Tl, d0.Cout, dev)
        self._conv(zp, d0, out=raw, act=act, snake=snk, pad=d0.K // 2)
        last = self._node(d0, zp, None, raw, snk, pad=d0.K // 2, kind="dec_in")
        for i in range(n):
            s = self.strides[n - 1 - i]
            q = f"{pd}{i + 1}.layers."
            up = self._tc(q + "1.", True, s)
            pad = math.ceil(s / 2)
            T2 = (act.T - 1) * s - 2 * pad + up.K
            s0 = _TSnake(P, q + "2.layers.0.")
            r2, a2 = _plane(B, T2, up.Cout, dev), _plane(B, T2, up.Cout, dev)
            self._conv(act, up, out=r2, act=a2, snake=s0, pad=pad, mode=2)
            last = self._node(up, act, last, r2, s0, pad=pad, mode=2)
            nxt = _TSnake(P, f"{pd}{i + 2}.layers.0.") if i + 1 < n else _TSnake(P, f"{pd}{n + 1}.")
            raw, act, last = self._rus(r2, a2, last, q, 2, nxt)
        do = self._tc(f"{pd}{n + 2}.")
        y = torch.empty(B, do.Cout, act.T, device=dev)
        check(lib().b200sat_conv_out(act.hi.data_ptr(), 0, do.w_f32.data_ptr(), _p(do.bias), y.data_ptr(), B, do.Cin, act.T, do.Cout, do.K,
                                     do.K // 2, 0, _s()), "conv_out")
        ops.LAUNCHES[0] += 1
        self._node(do, act, last, None, None, pad=do.K // 2, kind="conv_out")
        return y, (kl / (B * Tl)).squeeze(0), z

    # ------------------------------------------------------------------ backward
    def _wn_small_bwd(self, conv, dw, grads):
        v, g = conv.v.detach(), conv.g.detach()
        nrm = v.flatten(1).norm(dim=1).view(-1, 1, 1)
        dot = (dw * v).sum(dim=(1, 2), keepdim=True)
        grads[conv.prefix + "weight_g"] = (dot / nrm).view_as(g)
        grads[conv.prefix + "weight_v"] = g.view(-1, 1, 1) / nrm * (dw - v * dot / (nrm * nrm))

    def _wgrad(self, e, d_raw, grads):
        conv, x, mode = e["conv"], e["x"], e["mode"]
        K, s = conv.K, conv.stride
        R, Cc = conv.v.shape[0], conv.v.shape[1]
        dwp = torch.zeros(K, R, Cc, device=self.dev, dtype=torch.float32)
        for k in range(K):
            if mode == 0:
                a = (d_raw, 1, 0, 0)
                b = (x, 1, 0, k * e["dil"] - e["pad"])
                t_iter = d_raw.T
            elif mode == 1:
                a = (d_raw, 1, 0, 0)
                b = (x, s, (k - e["pad"]) % s, (k - e["pad"]) // s)
                t_iter = d_raw.T
            else:
                a = (x, 1, 0, 0)
                b = (d_raw, s, (k - e["pad"]) % s, (k - e["pad"]) // s)
                t_iter = x.T
            check(lib().b200sat_conv_wgrad(a[0].hi.data_ptr(), a[0].C, a[0].T, a[1], a[2], a[3], b[0].hi.data_ptr(), b[0].C, b[0].T, b[1], b[2],
                                           b[3], dwp[k].data_ptr(), self.B, t_iter, _s()), "conv_wgrad")
        ops.LAUNCHES[0] += K
        dv = torch.empty_like(conv.v)
        dg = torch.empty_like(conv.g)
        check(lib().b200sat_wn_bwd(dwp.data_ptr(), conv.v.data_ptr(), conv.g.data_ptr(), conv.inv_norm.data_ptr(), dv.data_ptr(), dg.data_ptr(),
                                   R, Cc, K, _s()), "wn_bwd")
        ops.LAUNCHES[0] += 1
        grads[conv.prefix + "weight_v"], grads[conv.prefix + "weight_g"] = dv, dg

    def _dgrad(self, e, d_raw):
        conv, mode = e["conv"], e["mode"]
        w = conv.pack_dgrad(mode)
        x = e["x"]
        d_x = _plane(x.B, x.T, x.C, self.dev)
        if mode == 0:
            m2, pad2 = 0, e["dil"] * (conv.K - 1) - e["pad"]
        elif mode == 1:
            m2, pad2 = 2, e["pad"]
        else:
            m2, pad2 = 1, e["pad"]
        check(lib().b200sat_conv1d_fwd(d_raw.hi.data_ptr(), 0, w.data_ptr(), 0, 0, 0, 0, d_x.hi.data_ptr(), 0, 0, 0, 0, 0, d_raw.B, d_raw.T,
                                       conv.Cout, conv.Cin, conv.K, e["dil"], pad2, conv.stride, m2, 1, _s()), "conv1d dgrad")
        ops.LAUNCHES[0] += 1
        return d_x

    def backward(self, dY, dkl=None, dz_extra=None):
        """dY fp32 [B, Cout, T]: gradient of the decoded audio; dkl: 0-dim tensor, gradient of the KL scalar.
        Returns {reference parameter name: gradient}."""
        dev, B = self.dev, self.B
        grads = {}
        d_act, d_skip = {}, {}
        dY = dY.to(dev, torch.float32).contiguous()
        for idx in range(len(self.tape) - 1, -1, -1):
            e = self.tape[idx]
            conv, kind = e["conv"], e["kind"]
            if kind == "conv_out":
                x = e["x"]
                dw = torch.zeros_like(conv.w_f32)      # [A, C, K]
                check(lib().b200sat_edge_wgrad(x.hi.data_ptr(), dY.data_ptr(), dw.data_ptr(), B, x.T, x.C, conv.Cout, conv.K, e["pad"], -1,
                                               conv.K, x.C * conv.K, _s()), "edge_wgrad")
                self._wn_small_bwd(conv, dw, grads)
                if conv.bias is not None:
                    grads[conv.prefix + "bias"] = dY.sum(dim=(0, 2))
                wflip = conv.w_f32.flip(2).permute(1, 0, 2).contiguous()   # [C, A, K]: dgrad as a conv audio -> C
                d_x = _plane(B, x.T, x.C, dev)
                check(lib().b200sat_conv_in(dY.data_ptr(), wflip.data_ptr(), 0, 0, 0, d_x.hi.data_ptr(), 0, 0, 0, B, conv.Cout, x.T, x.C, conv.K,
                                            conv.K - 1 - e["pad"], _s()), "conv_out dgrad")
                ops.LAUNCHES[0] += 2
                d_act[e["src"]] = d_x
                continue
            # ---- gradient of the raw conv output
            y_raw, snake = e["y_raw"], e["snake"]
            dbias = torch.zeros(conv.Cout, device=dev) if conv.bias is not None else None
            if snake is not None:
                d_raw = _plane(y_raw.B, y_raw.T, y_raw.C, dev)
                da, db = torch.zeros_like(snake.alpha), torch.zeros_like(snake.beta)
                sk = d_skip.pop(idx, None)
                check(lib().b200sat_snake_bwd(d_act.pop(idx).hi.data_ptr(), y_raw.hi.data_ptr(), _p(sk.hi) if sk else 0, snake.a.data_ptr(),
                                              snake.invb.data_ptr(), d_raw.hi.data_ptr(), da.data_ptr(), db.data_ptr(), _p(dbias),
                                              y_raw.B * y_raw.T, y_raw.C, _s()), "snake_bwd")
                ops.LAUNCHES[0] += 1
                grads[snake.prefix + "alpha"], grads[snake.prefix + "beta"] = da, db
            else:   # encoder conv_out: gradient comes from the bottleneck
                d_raw = _plane(B, self.Tl, 2 * self.L, dev)
                dz = d_act.pop(idx, None)
                check(lib().b200sat_vae_sample_bwd(_p(dz.hi) if dz else 0, self.ms.hi.data_ptr(), _p(self.noise), _p(dkl),
                                                   1.0 / (B * self.Tl) if dkl is not None else 0.0, d_raw.hi.data_ptr(), B, self.L, self.Tl,
                                                   _s()), "vae_sample_bwd")
                ops.LAUNCHES[0] += 1
                if dbias is not None:
                    ops.colsum(d_raw.hi.view(-1, d_raw.C), dbias)
            if dbias is not None:
                grads[conv.prefix + "bias"] = dbias
            if e["res"] is not None:
                d_skip[e["res"]] = d_raw
            e["y_raw"] = None
            # ---- weight gradient, data gradient
            if kind == "conv_in":
                dw = torch.zeros_like(conv.w_f32)      # [C, A, K]
                check(lib().b200sat_edge_wgrad(d_raw.hi.data_ptr(), self.audio.data_ptr(), dw.data_ptr(), B, d_raw.T, d_raw.C, conv.Cin, conv.K,
                                               conv.K // 2, 1, conv.Cin * conv.K, conv.K, _s()), "edge_wgrad")
                ops.LAUNCHES[0] += 1
                self._wn_small_bwd(conv, dw, grads)
                continue
            self._wgrad(e, d_raw, grads)
            d_x = self._dgrad(e, d_raw)
            e["x"] = None
            if kind == "dec_in":
                d_act[self.enc_out] = d_x      # d z (plane layout) -> bottleneck backward at the encoder's last conv
            else:
                d_act[e["src"]] = d_x
        return grads


class _AEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, audio, noise, *params):
        P = {n: p.detach() for n, p in zip(model.names, params)}
        step = _Step(P, model.strides)
        y, kl, z = step.forward(audio, noise)
        ctx.step, ctx.model = step, model
        ctx.mark_non_differentiable(z)
        return y, kl, z

    @staticmethod
    def backward(ctx, dY, dkl, _dz):
        step, model = ctx.step, ctx.model
        ctx.step = None
        grads = step.backward(dY, dkl)
        return (None, None, None) + tuple(grads.get(n) for n in model.names)


class OobleckTrainModel(nn.Module):
    """Trainable Oobleck VAE (encoder + bottleneck + decoder) on the b200sat kernels.  `forward(audio, noise)` returns
    (decoded fp32 [B, C, T], kl scalar, latents) like AudioAutoencoder.encode(return_info=True) + decode (autoencoders.py:446-534);
    gradients reach `.grad` of parameters named exactly like the reference state dict (`reference_state_dict()`)."""

    def __init__(self, state_dict, strides=(2, 4, 4, 8, 8), device="cuda"):
        super().__init__()
        self.strides = tuple(strides)
        self.names = [k for k in state_dict if k.startswith(("encoder.", "decoder."))]
        for k in self.names:
            self.register_parameter(k.replace(".", "__"), nn.Parameter(state_dict[k].detach().to(device, torch.float32).clone().contiguous()))

    def reference_state_dict(self):
        return {k: getattr(self, k.replace(".", "__")).detach() for k in self.names}

    def forward(self, audio, noise=None):
        if not audio.is_cuda:
            raise RuntimeError("b200sat: OobleckTrainModel needs CUDA tensors (no CPU fallback)")
        params = [getattr(self, k.replace(".", "__")) for k in self.names]
        return _AEFn.apply(self, audio, noise, *params)


def generator_loss(model, loss_sd, reals, noise=None, kl_weight=1e-4, stft_weight=1.0):
    """Generator loss of the warm-up phase (no adversarial terms; training/autoencoders.py:185-231, 436-441, 497):
    stft_weight * (sum/difference MRSTFT + (left + right)/2) + kl_weight * kl."""
    from .stft_loss import autoencoder_mrstft_terms
    decoded, kl, _ = model(reals, noise)
    sd, left, right = autoencoder_mrstft_terms(loss_sd, decoded, reals)
    loss = stft_weight * (sd + 0.5 * left + 0.5 * right) + kl_weight * kl
    return loss, dict(mrstft_sd=sd.detach(), mrstft_l=left.detach(), mrstft_r=right.detach(), kl=kl.detach())

"""Encodec multi-scale STFT discriminator on libb200sat: forward (logits + feature maps), the hinge / feature-matching loss values,
and the GENERATOR-side backward (gradient of adv + feature-matching w.r.t. the decoded audio).

Reference: `EncodecDiscriminator` (models/discriminators.py:13-58) over `MultiScaleSTFTDiscriminator` / `DiscriminatorSTFT`
(models/encodec.py:38-138) as configured by stable_audio_2_0_vae.json:80-91 (filters 64, five scales, stereo, stride (1,1), hinge).
Layout and kernels: csrc/discriminator.cu (STFT front end, first / last conv, loss reductions, activation backward) and
`b200sat_conv2d_flat` (the four 64 -> 64 channel Conv2d layers per scale on the tcgen05 conv kernel).

The discriminator step (training/autoencoders.py:476-489: hinge loss on reals and fakes, gradients w.r.t. the discriminator's
weight-normed parameters) is `EncodecDiscriminatorTrain.discriminator_loss`: weight gradients of the 64 -> 64 convs are one
`b200sat_conv_wgrad` launch per tap on the flattened planes, the two SIMT layers have their own kernels, weight-norm backward is
`b200sat_wn_bwd`.
"""
import ctypes
import math
import os

import torch
from torch import nn

from ._lib import lib, check
from . import ops

LEAKY = 0.2
DILATIONS = (1, 2, 4)
# first conv on the tensor cores (default); B200SAT_DISC_CONV0=simt keeps the round-1 fp32 SIMT kernels
CONV0_TC = os.environ.get("B200SAT_DISC_CONV0", "tc") != "simt"
WGRAD_MODE = os.environ.get("B200SAT_DISC_WGRAD", "win")


def _s():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return 0 if t is None else t.data_ptr()


def _wn_dense(sd, p):
    v = sd[p + "weight_v"].float()
    g = sd[p + "weight_g"].float()
    return (g * v / v.flatten(1).norm(dim=1).view(-1, 1, 1, 1)).contiguous()


class _Scale:
    def __init__(self, sd, pre, n_fft, hop, dev):
        self.n, self.hop = n_fft, hop
        self.F = n_fft // 2 + 1
        self.Fp = self.F + 8
        win = torch.hann_window(n_fft, periodic=True, dtype=torch.float64)
        k = torch.arange(n_fft // 2, dtype=torch.float64)
        self.window = win.float().to(dev).contiguous()
        self.twiddle = torch.stack([torch.cos(2 * math.pi * k / n_fft), -torch.sin(2 * math.pi * k / n_fft)], dim=1).float().to(dev).contiguous()
        # first conv (4 -> 64) and conv_post (64 -> 1): dense weight-normalised fp32 weights for the SIMT kernels
        self.pre = pre
        self.raw = {k[len(pre):]: sd[k].float().to(dev) for k in sd if k.startswith(pre) and k.endswith(("weight_v", "weight_g", "bias"))}
        self.w0 = _wn_dense(sd, pre + "convs.0.conv.").to(dev).reshape(64, 4, 27).contiguous()
        self.b0 = sd[pre + "convs.0.conv.bias"].float().to(dev).contiguous()
        self.wp = _wn_dense(sd, pre + "conv_post.conv.").to(dev).reshape(64, 9).contiguous()      # [1,64,3,3] -> [c][tap]
        self.bp = sd[pre + "conv_post.conv.bias"].float().to(dev).contiguous()
        # the four 64 -> 64 convs: packed bf16 weights (forward and data-gradient layouts) + per-tap row shifts.  Entry 0 of the list
        # built here is the FIRST conv in its tensor-core form (CONV0_TC): frequency taps folded into channels (b200sat_disc_spec_pack),
        # v9[co][ci*9 + df][dt] = v[co][ci][dt][df], three taps = row shifts (-Fp, 0, +Fp); its zero rows leave the weight norm unchanged.
        self.convs = []
        self.cv0 = None
        for j in (range(0, 5) if CONV0_TC else range(1, 5)):
            q = f"{pre}convs.{j}.conv."
            v = sd[q + "weight_v"].float().to(dev)
            if j == 0:
                v9 = torch.zeros(64, 64, 3, device=dev)
                v9[:, :36] = v.reshape(64, 4, 3, 9).permute(0, 1, 3, 2).reshape(64, 36, 3)
                v = v9.view(64, 64, 3, 1)
            K = v.shape[2] * v.shape[3]
            v3 = v.reshape(64, 64, K).contiguous()
            g = sd[q + "weight_g"].float().to(dev).reshape(-1).contiguous()
            inv = torch.empty(64, device=dev)
            w_f = torch.empty(64, K * 64, device=dev, dtype=torch.bfloat16)
            check(lib().b200sat_wn_pack(v3.data_ptr(), g.data_ptr(), inv.data_ptr(), w_f.data_ptr(), 0, 64, 64, K, 0, 1, _s()), "wn_pack")
            w_d = torch.empty(64, K * 64, device=dev, dtype=torch.bfloat16)
            check(lib().b200sat_wn_pack_dgrad(v3.data_ptr(), g.data_ptr(), inv.data_ptr(), w_d.data_ptr(), 64, 64, K, 0, 1, _s()), "wn_pack_dgrad")
            if j == 0:
                offs = [-self.Fp, 0, self.Fp]
            elif j <= 3:
                d = DILATIONS[j - 1]
                offs = [(k_ // 9 - 1) * d * self.Fp + (k_ % 9 - 4) for k_ in range(27)]
            else:
                offs = [(k_ // 3 - 1) * self.Fp + (k_ % 3 - 1) for k_ in range(9)]
            cv = dict(w=w_f, wd=w_d, bias=sd[q + "bias"].float().to(dev).contiguous(), K=K, offs=(ctypes.c_int * K)(*offs),
                      offs_py=offs, v3=v3, g=g, inv=inv, name=f"convs.{j}.conv.")
            if j == 0:
                self.cv0 = cv
            else:
                self.convs.append(cv)
        ops.LAUNCHES[0] += 12 + (3 if CONV0_TC else 0)

    def frames(self, T):
        return (T - self.n) // self.hop + 1


class EncodecDiscriminatorEngine:
    def __init__(self, state_dict, n_ffts=(2048, 1024, 512, 256, 128), hop_lengths=(512, 256, 128, 64, 32), device="cuda",
                 prefix="discriminators.discriminators."):
        self.dev = torch.device(device)
        sd = {k: v.detach() for k, v in state_dict.items()}
        if sd[f"{prefix}0.convs.0.conv.weight_v"].shape[:2] != (64, 4):
            raise NotImplementedError("b200sat discriminator: filters=64, stereo input (4 spectrogram channels) only")
        self.scales = [_Scale(sd, f"{prefix}{i}.", n, h, self.dev) for i, (n, h) in enumerate(zip(n_ffts, hop_lengths))]

    # ------------------------------------------------------------------ forward of one scale
    def _flat_conv(self, sc, x, cv, out, bias, w):
        B, P, _ = x.shape
        check(lib().b200sat_conv2d_flat(x.data_ptr(), w.data_ptr(), _p(bias), out.data_ptr(), B, P, 64, 64, cv["K"], cv["offs"], sc.Fp, 4,
                                        4 + sc.F, LEAKY if bias is not None else 0.0, _s()), "conv2d_flat")
        ops.LAUNCHES[0] += 1
        return out

    def _scale_forward(self, sc, x):
        """x fp32 [B, 2, T] -> (logits fp32 [B, P], [5 feature-map planes bf16 [B, P, 64]], frames)."""
        B, C, T = x.shape
        fr = sc.frames(T)
        P = fr * sc.Fp
        spec = torch.zeros(B, P, 4, device=self.dev)
        check(lib().b200sat_disc_stft(x.data_ptr(), spec.data_ptr(), sc.window.data_ptr(), sc.twiddle.data_ptr(), B, T, sc.n, sc.hop, 0, _s()), "disc_stft")
        f0 = torch.empty(B, P, 64, device=self.dev, dtype=torch.bfloat16)
        if sc.cv0 is not None:
            s9 = torch.empty(B, P, 64, device=self.dev, dtype=torch.bfloat16)
            check(lib().b200sat_disc_spec_pack(spec.data_ptr(), s9.data_ptr(), B, fr, sc.F, 0, _s()), "disc_spec_pack")
            self._flat_conv(sc, s9, sc.cv0, f0, sc.cv0["bias"], sc.cv0["w"])
            spec = s9                 # what the weight gradient of the first conv reads
        else:
            check(lib().b200sat_disc_conv0(spec.data_ptr(), sc.w0.data_ptr(), sc.b0.data_ptr(), f0.data_ptr(), 0, 0, B, fr, sc.F, LEAKY, _s()), "disc_conv0")
        fmaps = [f0]
        for cv in sc.convs:
            fmaps.append(self._flat_conv(sc, fmaps[-1], cv, torch.empty_like(f0), cv["bias"], cv["w"]))
        logits = torch.empty(B, P, device=self.dev)
        check(lib().b200sat_disc_convpost(fmaps[-1].data_ptr(), sc.wp.data_ptr(), sc.bp.data_ptr(), logits.data_ptr(), B, fr, sc.F, _s()), "disc_convpost")
        ops.LAUNCHES[0] += 3
        self._last_spec = spec
        return logits, fmaps, fr

    @torch.no_grad()
    def forward(self, x):
        """Like EncodecDiscriminator.forward: (logits, features) per scale, in the reference's [B, 1, frames, F] / [B, 64, frames, F] shapes
        (materialised from the flattened planes - for inspection and tests; the losses below work on the planes directly)."""
        x = x.to(self.dev, torch.float32).contiguous()
        logits, feats = [], []
        for sc in self.scales:
            lg, fm, fr = self._scale_forward(sc, x)
            B = x.shape[0]
            logits.append(lg.view(B, fr, sc.Fp)[:, :, 4:4 + sc.F].unsqueeze(1).contiguous())
            feats.append([f.view(B, fr, sc.Fp, 64)[:, :, 4:4 + sc.F].permute(0, 3, 1, 2).float() for f in fm])
        return logits, feats

    # ------------------------------------------------------------------ losses
    def _loss_forward(self, reals, fakes, keep_all=False, need_fm=True):
        """keep_all: also keep the real-path logits and both spectrograms (needed by the discriminator-side backward).
        need_fm=False skips the 25 feature-matching L1 reductions (the discriminator step only uses the hinge loss; fm is returned as 0)."""
        reals = reals.to(self.dev, torch.float32).contiguous()
        fakes = fakes.to(self.dev, torch.float32).contiguous()
        B = reals.shape[0]
        ns = len(self.scales)
        hs = torch.zeros(ns, 3, device=self.dev, dtype=torch.float64)
        l1 = torch.zeros(ns, 5, device=self.dev, dtype=torch.float64)
        saved = []
        n_logit, n_feat = [], []
        for i, sc in enumerate(self.scales):
            lt, ft, fr = self._scale_forward(sc, reals)
            spec_r = self._last_spec
            lf, ff, _ = self._scale_forward(sc, fakes)
            spec_f = self._last_spec
            check(lib().b200sat_disc_hinge_sums(lt.data_ptr(), lf.data_ptr(), hs[i].data_ptr(), B, fr, sc.F, _s()), "disc_hinge_sums")
            for l in range(5 if need_fm else 0):
                check(lib().b200sat_disc_l1_sum(ft[l].data_ptr(), ff[l].data_ptr(), l1[i, l:].data_ptr(), ft[l].numel(), _s()), "disc_l1_sum")
            ops.LAUNCHES[0] += 6 if need_fm else 1
            n_logit.append(B * fr * sc.F)
            n_feat.append(B * 64 * fr * sc.F)
            saved.append((ft, ff, lf, fr, lt, spec_r, spec_f) if keep_all else (ft, ff, lf, fr))
        nl = torch.tensor(n_logit, device=self.dev, dtype=torch.float64)
        nf = torch.tensor(n_feat, device=self.dev, dtype=torch.float64)
        dis = ((hs[:, 0] + hs[:, 1]) / nl).sum() / ns
        adv = (-hs[:, 2] / nl).sum() / ns
        fm = (l1.sum(1) / nf / 5.0).sum() / ns
        return dis.float(), adv.float(), fm.float(), saved, n_logit, n_feat, fakes.shape

    @torch.no_grad()
    def loss_values(self, reals, fakes):
        """(dis_loss, adv_loss, feature_matching_distance) of EncodecDiscriminator.loss, values only."""
        d, a, f, *_ = self._loss_forward(reals, fakes)
        return d, a, f

    def generator_terms(self, reals, fakes):
        """(adv_loss, feature_matching_distance) differentiable w.r.t. `fakes` (the generator step's use of the discriminator,
        training/autoencoders.py:436-441, 497)."""
        return _GenFn.apply(fakes, reals, self)

    def _generator_backward(self, saved, n_logit, n_feat, shape, d_adv, d_fm):
        B, C, T = shape
        ns = len(self.scales)
        d_audio = torch.zeros(B, C, T, device=self.dev)
        st = _s()
        for i, sc in enumerate(self.scales):
            ft, ff, lf, fr = saved[i][:4]
            P = fr * sc.Fp
            g = torch.empty(B, P, device=self.dev)
            check(lib().b200sat_disc_logit_grad(lf.data_ptr(), g.data_ptr(), B, fr, sc.F, 0, d_adv / (n_logit[i] * ns), st), "disc_logit_grad")
            coef = d_fm / (n_feat[i] * 5.0 * ns)
            d_pre = torch.empty(B, P, 64, device=self.dev, dtype=torch.bfloat16)
            check(lib().b200sat_disc_act_bwd(0, g.data_ptr(), sc.wp.data_ptr(), ff[4].data_ptr(), ft[4].data_ptr(), coef, LEAKY, d_pre.data_ptr(), B, fr,
                                             sc.F, st), "disc_act_bwd")
            ops.LAUNCHES[0] += 2
            for l in range(4, 0, -1):
                cv = sc.convs[l - 1]
                d_in = self._flat_conv(sc, d_pre, cv, torch.empty_like(d_pre), None, cv["wd"])
                d_pre = torch.empty_like(d_in)
                check(lib().b200sat_disc_act_bwd(d_in.data_ptr(), 0, 0, ff[l - 1].data_ptr(), ft[l - 1].data_ptr(), coef, LEAKY, d_pre.data_ptr(), B, fr,
                                                 sc.F, st), "disc_act_bwd")
                ops.LAUNCHES[0] += 1
            dspec = torch.empty(B, P, 4, device=self.dev)
            if sc.cv0 is not None:
                ds9 = self._flat_conv(sc, d_pre, sc.cv0, torch.empty_like(d_pre), None, sc.cv0["wd"])
                check(lib().b200sat_disc_spec_pack(dspec.data_ptr(), ds9.data_ptr(), B, fr, sc.F, 1, st), "disc_spec_pack backward")
            else:
                check(lib().b200sat_disc_conv0(0, sc.w0.data_ptr(), 0, 0, d_pre.data_ptr(), dspec.data_ptr(), B, fr, sc.F, LEAKY, st), "disc_conv0 dgrad")
            check(lib().b200sat_disc_stft(d_audio.data_ptr(), dspec.data_ptr(), sc.window.data_ptr(), sc.twiddle.data_ptr(), B, T, sc.n, sc.hop, 1, st),
                  "disc_stft backward")
            ops.LAUNCHES[0] += 2
        return d_audio


class _GenFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, fakes, reals, eng):
        with torch.no_grad():
            _, adv, fm, saved, n_logit, n_feat, shape = eng._loss_forward(reals, fakes.detach())
        ctx.eng, ctx.saved, ctx.meta = eng, saved, (n_logit, n_feat, shape)
        return adv, fm

    @staticmethod
    def backward(ctx, d_adv, d_fm):
        n_logit, n_feat, shape = ctx.meta
        # the two upstream scalars are read back once (they are loss weights in practice): one host sync per generator step
        g = ctx.eng._generator_backward(ctx.saved, n_logit, n_feat, shape, float(d_adv), float(d_fm))
        ctx.saved = None
        return g, None, None


def _wn_small_bwd(v, g, dw):
    """weight_norm backward for the two small SIMT layers (a few thousand values): dw dense, shaped like v."""
    nrm = v.flatten(1).norm(dim=1).view(-1, 1, 1, 1)
    dot = (dw * v).sum(dim=(1, 2, 3), keepdim=True)
    return g / nrm * (dw - v * dot / (nrm * nrm)), dot / nrm


def _discriminator_backward(eng, saved, n_logit, B):
    """Gradients of the hinge discriminator loss w.r.t. every discriminator parameter from the activations `_loss_forward(keep_all=True)`
    kept: {reference name: grad}."""
    dev = eng.dev
    ns = len(eng.scales)
    st = _s()
    grads = {}
    for i, sc in enumerate(eng.scales):
        ft, ff, lf, fr, lt, spec_r, spec_f = saved[i]
        P = fr * sc.Fp
        dW0 = torch.zeros(64, 4, 27, device=dev)
        dW0c = torch.zeros(64, 3, 64, device=dev)     # tensor-core form of the first conv: [co][dt][ci*9 + df] (cat entry)
        dW0t = torch.zeros(3, 64, 64, device=dev)     # ... [dt][co][ci*9 + df] (win entry)
        db0 = torch.zeros(64, device=dev)
        dWp = torch.zeros(64, 9, device=dev)
        dbp = torch.zeros(1, device=dev)
        # weight gradients of the 64 -> 64 convs, B200SAT_DISC_WGRAD = win (default: one pass over the planes, csrc/disc_wgrad.cu, dW[tap][ca][cb])
        # | cat (four taps per tile, dWc[ca][tap][cb]) | tap (one launch per tap)
        wmode = WGRAD_MODE
        cat = wmode == "cat"
        dwps = [torch.zeros(64, cv["K"], 64, device=dev) if cat else torch.zeros(cv["K"], 64, 64, device=dev) for cv in sc.convs]
        dbs = [torch.zeros(64, device=dev) for _ in sc.convs]
        for lg, fm, spec, mode in ((lt, ft, spec_r, 1), (lf, ff, spec_f, 2)):
            g = torch.empty(B, P, device=dev)
            check(lib().b200sat_disc_logit_grad(lg.data_ptr(), g.data_ptr(), B, fr, sc.F, mode, 1.0 / (n_logit[i] * ns), st), "disc_logit_grad")
            check(lib().b200sat_disc_convpost_wgrad(g.data_ptr(), fm[4].data_ptr(), dWp.data_ptr(), dbp.data_ptr(), B, fr, sc.F, st), "disc_convpost_wgrad")
            d_pre = torch.empty(B, P, 64, device=dev, dtype=torch.bfloat16)
            check(lib().b200sat_disc_act_bwd(0, g.data_ptr(), sc.wp.data_ptr(), fm[4].data_ptr(), 0, 0.0, LEAKY, d_pre.data_ptr(), B, fr, sc.F, st), "disc_act_bwd")
            ops.LAUNCHES[0] += 3
            for l in range(4, 0, -1):
                cv = sc.convs[l - 1]
                if wmode == "win":
                    check(lib().b200sat_conv_wgrad_taps_win(d_pre.data_ptr(), fm[l - 1].data_ptr(), P, cv["offs"], cv["K"], dwps[l - 1].data_ptr(), B, st),
                          "conv_wgrad_taps_win")
                    ops.LAUNCHES[0] += 1
                elif cat:
                    check(lib().b200sat_conv_wgrad_taps_cat(d_pre.data_ptr(), fm[l - 1].data_ptr(), P, cv["offs"], cv["K"], dwps[l - 1].data_ptr(), B, st),
                          "conv_wgrad_taps_cat")
                    ops.LAUNCHES[0] += 1
                else:
                    check(lib().b200sat_conv_wgrad_taps(d_pre.data_ptr(), 64, fm[l - 1].data_ptr(), 64, P, cv["offs"], cv["K"], dwps[l - 1].data_ptr(), B, st),
                          "conv_wgrad_taps")
                    ops.LAUNCHES[0] += cv["K"]
                ops.colsum(d_pre.view(-1, 64), dbs[l - 1])
                d_in = eng._flat_conv(sc, d_pre, cv, torch.empty_like(d_pre), None, cv["wd"])
                d_pre = torch.empty_like(d_in)
                check(lib().b200sat_disc_act_bwd(d_in.data_ptr(), 0, 0, fm[l - 1].data_ptr(), 0, 0.0, LEAKY, d_pre.data_ptr(), B, fr, sc.F, st), "disc_act_bwd")
                ops.LAUNCHES[0] += 1
            if sc.cv0 is not None and wmode == "win":     # `spec` is the packed S9 plane: a 3-tap 64 -> 64 weight gradient
                check(lib().b200sat_conv_wgrad_taps_win(d_pre.data_ptr(), spec.data_ptr(), P, sc.cv0["offs"], 3, dW0t.data_ptr(), B, st), "conv_wgrad_taps_win")
            elif sc.cv0 is not None:
                check(lib().b200sat_conv_wgrad_taps_cat(d_pre.data_ptr(), spec.data_ptr(), P, sc.cv0["offs"], 3, dW0c.data_ptr(), B, st), "conv_wgrad_taps_cat")
            else:
                check(lib().b200sat_disc_conv0_wgrad(d_pre.data_ptr(), spec.data_ptr(), dW0.data_ptr(), B, fr, sc.F, st), "disc_conv0_wgrad")
            ops.colsum(d_pre.view(-1, 64), db0)
            ops.LAUNCHES[0] += 1
        # weight-norm backward
        pre = sc.pre
        for l, cv in enumerate(sc.convs):
            if cat:
                dwps[l] = dwps[l].permute(1, 0, 2).contiguous()      # [ca][tap][cb] -> [tap][ca][cb] (110 K floats)
            dv = torch.empty_like(cv["v3"])
            dg = torch.empty_like(cv["g"])
            check(lib().b200sat_wn_bwd(dwps[l].data_ptr(), cv["v3"].data_ptr(), cv["g"].data_ptr(), cv["inv"].data_ptr(), dv.data_ptr(), dg.data_ptr(), 64, 64,
                                       cv["K"], st), "wn_bwd")
            ops.LAUNCHES[0] += 1
            grads[pre + cv["name"] + "weight_v"] = dv.view_as(sc.raw[cv["name"] + "weight_v"])
            grads[pre + cv["name"] + "weight_g"] = dg.view_as(sc.raw[cv["name"] + "weight_g"])
            grads[pre + cv["name"] + "bias"] = dbs[l]
        v0, g0 = sc.raw["convs.0.conv.weight_v"], sc.raw["convs.0.conv.weight_g"]
        if sc.cv0 is not None:
            cv = sc.cv0
            dv9 = torch.empty_like(cv["v3"])
            dg0 = torch.empty_like(cv["g"])
            dw0 = dW0t if wmode == "win" else dW0c.permute(1, 0, 2).contiguous()
            check(lib().b200sat_wn_bwd(dw0.data_ptr(), cv["v3"].data_ptr(), cv["g"].data_ptr(), cv["inv"].data_ptr(),
                                       dv9.data_ptr(), dg0.data_ptr(), 64, 64, 3, st), "wn_bwd")
            ops.LAUNCHES[0] += 1
            dv0 = dv9[:, :36].reshape(64, 4, 9, 3).permute(0, 1, 3, 2).reshape(v0.shape).contiguous()     # [co][ci*9+df][dt] -> [co][ci][dt][df]
            dg0 = dg0.view_as(g0)
        else:
            dv0, dg0 = _wn_small_bwd(v0, g0, dW0.view_as(v0))
        grads[pre + "convs.0.conv.weight_v"], grads[pre + "convs.0.conv.weight_g"], grads[pre + "convs.0.conv.bias"] = dv0, dg0, db0
        vp, gp = sc.raw["conv_post.conv.weight_v"], sc.raw["conv_post.conv.weight_g"]
        dvp, dgp = _wn_small_bwd(vp, gp, dWp.view(1, 64, 3, 3))
        grads[pre + "conv_post.conv.weight_v"], grads[pre + "conv_post.conv.weight_g"], grads[pre + "conv_post.conv.bias"] = dvp, dgp, dbp
    return grads


def _discriminator_forward_backward(eng, reals, fakes):
    """Hinge discriminator loss and the gradients of every discriminator parameter: returns (dis, {reference name: grad})."""
    dis, _, _, saved, n_logit, _, shape = eng._loss_forward(reals, fakes, keep_all=True, need_fm=False)
    return dis, _discriminator_backward(eng, saved, n_logit, shape[0])


class _DiscDFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, reals, fakes, model, *params):
        sd = {n: p.detach() for n, p in zip(model.names, params)}
        eng = EncodecDiscriminatorEngine(sd, model.n_ffts, model.hop_lengths, device=params[0].device, prefix=model.prefix)
        with torch.no_grad():
            dis, grads = _discriminator_forward_backward(eng, reals.detach(), fakes.detach())
        ctx.grads, ctx.names = grads, model.names
        return dis

    @staticmethod
    def backward(ctx, d_dis):
        grads = ctx.grads
        ctx.grads = None
        return (None, None, None) + tuple(grads[n] * d_dis for n in ctx.names)


class _DiscLossFn(torch.autograd.Function):
    """EncodecDiscriminator.loss as ONE node: (dis, adv, fm) = f(reals, fakes, *params); dis is differentiable w.r.t. the parameters,
    adv / fm w.r.t. `fakes` - the contract the reference training step relies on (training/autoencoders.py:436-515)."""

    @staticmethod
    def forward(ctx, reals, fakes, names, n_ffts, hops, prefix, *params):
        sd = {n: p.detach() for n, p in zip(names, params)}
        eng = EncodecDiscriminatorEngine(sd, n_ffts, hops, device=params[0].device, prefix=prefix)
        with torch.no_grad():
            dis, adv, fm, saved, n_logit, n_feat, shape = eng._loss_forward(reals.detach(), fakes.detach(), keep_all=True)
        ctx.eng, ctx.saved, ctx.meta, ctx.names = eng, saved, (n_logit, n_feat, shape), names
        ctx.need_fakes = fakes.requires_grad
        ctx.need_params = any(p.requires_grad for p in params)
        ctx.set_materialize_grads(False)      # an output that took no part in the loss arrives as None, not as a zero tensor
        return dis, adv, fm

    @staticmethod
    def backward(ctx, d_dis, d_adv, d_fm):
        n_logit, n_feat, shape = ctx.meta
        # Which side runs is decided by which outputs actually received a gradient: the reference's generator step back-propagates only
        # adv / fm (-> fakes), its discriminator step only dis (-> parameters); running both every time doubled the backward.
        g_fakes = None
        if ctx.need_fakes and (d_adv is not None or d_fm is not None):
            g_fakes = ctx.eng._generator_backward(ctx.saved, n_logit, n_feat, shape, 0.0 if d_adv is None else float(d_adv),
                                                  0.0 if d_fm is None else float(d_fm))
        g_params = (None,) * len(ctx.names)
        if ctx.need_params and d_dis is not None:
            grads = _discriminator_backward(ctx.eng, ctx.saved, n_logit, shape[0])
            g_params = tuple(grads[n] * d_dis for n in ctx.names)
        ctx.saved = None
        return (None, g_fakes, None, None, None, None) + g_params


def reference_discriminator_loss(module, reals, fakes):
    """Drop-in body for `EncodecDiscriminator.loss(reals, fakes)` (models/discriminators.py:31-58) on an UNMODIFIED reference module:
    parameters are read from `module.discriminators` (old-style weight_norm: `...conv.weight_g|weight_v|bias`)."""
    msd = module.discriminators
    prefix = "discriminators."
    names, params = [], []
    for n, p in msd.named_parameters():
        names.append(prefix + n); params.append(p)
    n_ffts = tuple(d.n_fft for d in msd.discriminators)
    hops = tuple(d.hop_length for d in msd.discriminators)
    return _DiscLossFn.apply(reals, fakes, names, n_ffts, hops, "discriminators.discriminators.", *params)


class EncodecDiscriminatorTrain(nn.Module):
    """Trainable EncodecDiscriminator with the reference parameter names (`discriminators.discriminators.{i}.convs.{j}.conv.weight_g|v|bias`).
    `discriminator_loss(reals, fakes)` = hinge `dis_loss` differentiable w.r.t. the parameters (the D step);
    `generator_terms(reals, fakes)` = (adv_loss, feature_matching_distance) differentiable w.r.t. `fakes` (the G step)."""

    def __init__(self, state_dict, n_ffts=(2048, 1024, 512, 256, 128), hop_lengths=(512, 256, 128, 64, 32), device="cuda",
                 prefix="discriminators.discriminators."):
        super().__init__()
        self.n_ffts, self.hop_lengths, self.prefix = tuple(n_ffts), tuple(hop_lengths), prefix
        self.names = [k for k in state_dict if k.startswith(prefix) and k.endswith(("weight_v", "weight_g", "bias"))]
        for k in self.names:
            self.register_parameter(k.replace(".", "__"), nn.Parameter(state_dict[k].detach().to(device, torch.float32).clone().contiguous()))

    def reference_state_dict(self):
        return {k: getattr(self, k.replace(".", "__")).detach() for k in self.names}

    def engine(self):
        p0 = getattr(self, self.names[0].replace(".", "__"))
        return EncodecDiscriminatorEngine(self.reference_state_dict(), self.n_ffts, self.hop_lengths, device=p0.device, prefix=self.prefix)

    def discriminator_loss(self, reals, fakes):
        params = [getattr(self, k.replace(".", "__")) for k in self.names]
        return _DiscDFn.apply(reals, fakes, self, *params)

    def generator_terms(self, reals, fakes):
        return self.engine().generator_terms(reals, fakes)

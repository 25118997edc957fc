"""Oobleck autoencoder engine: OobleckEncoder / OobleckDecoder forward (models/autoencoders.py:285-362) + VAE bottleneck
(models/bottleneck.py:105-134) as a sequence of libb200sat launches.

Layout: activations live as time-major bf16 plane pairs (hi, lo) [B, T, C] (see csrc/conv1d.cu); the 2-channel audio ends
are fp32 [B, C, T] exactly like the reference tensors.  `precision="fp32x3"` runs every conv as three bf16 tensor-core
passes (hi*hi + hi*lo + lo*hi) for fp32-class accuracy (inference parity <= 1e-4 RMS); `precision="bf16"` is the single-pass
mode that bf16 autocast training uses.
"""
import math
import os

import torch

from ._lib import lib, check
from . import ops


def _s():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return 0 if t is None else t.data_ptr()


class _Planes:
    """A (hi, lo) pair of bf16 [B, T, C] buffers."""

    def __init__(self, B, T, C, device, with_lo):
        self.B, self.T, self.C = B, T, C
        self.hi = torch.empty(B, T, C, device=device, dtype=torch.bfloat16)
        self.lo = torch.empty(B, T, C, device=device, dtype=torch.bfloat16) if with_lo else None

    def float(self):
        v = self.hi.float()
        return v + self.lo.float() if self.lo is not None else v


class _Conv:
    """Packed weights of one weight-normed conv (autoencoders.py:23-27)."""

    def __init__(self, sd, prefix, transposed, stride, device, with_lo):
        if (prefix + "weight_g") in sd:
            v = sd[prefix + "weight_v"].detach().to(device, torch.float32).contiguous()
            g = sd[prefix + "weight_g"].detach().to(device, torch.float32).contiguous().view(-1)
        else:
            v = sd[prefix + "weight"].detach().to(device, torch.float32).contiguous()
            g = None
        b = sd.get(prefix + "bias")
        self.bias = None if b is None else b.detach().to(device, torch.float32).contiguous()
        self.transposed, self.stride = transposed, stride
        if transposed:
            self.Cin, self.Cout, self.K = v.shape
            rows, cols = stride * self.Cout, 2 * self.Cin
        else:
            self.Cout, self.Cin, self.K = v.shape
            rows, cols = self.Cout, self.K * self.Cin
        self.v, self.g = v, g
        self.small = (self.Cin % 64 != 0) or (self.Cout % 32 != 0)   # 2-channel audio ends -> SIMT kernels, fp32 weights
        if self.small:
            self.w_f32 = self._dense_weight()
            return
        self.w_hi = torch.empty(rows, cols, device=device, dtype=torch.bfloat16)
        self.w_lo = torch.empty(rows, cols, device=device, dtype=torch.bfloat16) if with_lo else None
        scratch = torch.empty(v.shape[0], device=device, dtype=torch.float32) if g is not None else None
        rc = lib().b200sat_wn_pack(v.data_ptr(), _p(g), _p(scratch), self.w_hi.data_ptr(), _p(self.w_lo), self.Cout, self.Cin,
                                   self.K, int(transposed), stride, _s())
        ops.LAUNCHES[0] += 2 if g is not None else 1
        check(rc, "wn_pack")

    def _dense_weight(self):
        if self.g is None:
            return self.v
        n = self.v.flatten(1).norm(dim=1).view(-1, 1, 1)
        return (self.g.view(-1, 1, 1) * self.v / n).contiguous()


class _Snake:
    def __init__(self, sd, prefix, device):
        al = sd[prefix + "alpha"].detach().to(device, torch.float32).contiguous()
        be = sd[prefix + "beta"].detach().to(device, torch.float32).contiguous()
        self.a = torch.empty_like(al)
        self.invb = torch.empty_like(be)
        check(lib().b200sat_snake_prep(al.data_ptr(), be.data_ptr(), self.a.data_ptr(), self.invb.data_ptr(), al.numel(), _s()), "snake_prep")
        ops.LAUNCHES[0] += 1


class OobleckEngine:
    def __init__(self, state_dict, strides=(2, 4, 4, 8, 8), device="cuda", precision="fp32x3", encoder_prefix="encoder.",
                 decoder_prefix="decoder.", final_tanh=False):
        if precision not in ("fp32x3", "bf16"):
            raise ValueError("precision must be 'fp32x3' or 'bf16'")
        self.device = torch.device(device)
        self.passes = 3 if precision == "fp32x3" else 1
        self.with_lo = self.passes == 3
        self.strides = tuple(strides)
        self.final_tanh = final_tanh
        self.enc = self._build_encoder(state_dict, encoder_prefix) if any(k.startswith(encoder_prefix) for k in state_dict) else None
        self.dec = self._build_decoder(state_dict, decoder_prefix) if any(k.startswith(decoder_prefix) for k in state_dict) else None

    # ------------------------------------------------------------------ construction
    def _ru(self, sd, p):
        dev, lo = self.device, self.with_lo
        return dict(s0=_Snake(sd, p + "layers.0.", dev), c7=_Conv(sd, p + "layers.1.", False, 1, dev, lo),
                    s1=_Snake(sd, p + "layers.2.", dev), c1=_Conv(sd, p + "layers.3.", False, 1, dev, lo))

    def _build_encoder(self, sd, pre):
        dev, lo = self.device, self.with_lo
        p = pre + "layers."
        n = len(self.strides)
        e = dict(conv_in=_Conv(sd, p + "0.", False, 1, dev, lo), blocks=[])
        for i, s in enumerate(self.strides):
            q = f"{p}{i + 1}.layers."
            e["blocks"].append(dict(rus=[self._ru(sd, f"{q}{j}.") for j in range(3)], snake=_Snake(sd, q + "3.", dev),
                                    down=_Conv(sd, q + "4.", False, s, dev, lo), stride=s))
        e["snake_out"] = _Snake(sd, f"{p}{n + 1}.", dev)
        e["conv_out"] = _Conv(sd, f"{p}{n + 2}.", False, 1, dev, lo)
        return e

    def _build_decoder(self, sd, pre):
        dev, lo = self.device, self.with_lo
        p = pre + "layers."
        n = len(self.strides)
        d = dict(conv_in=_Conv(sd, p + "0.", False, 1, dev, lo), blocks=[])
        for i in range(n):
            s = self.strides[n - 1 - i]
            q = f"{p}{i + 1}.layers."
            d["blocks"].append(dict(snake=_Snake(sd, q + "0.", dev), up=_Conv(sd, q + "1.", True, s, dev, lo),
                                    rus=[self._ru(sd, f"{q}{2 + j}.") for j in range(3)], stride=s))
        d["snake_out"] = _Snake(sd, f"{p}{n + 1}.", dev)
        d["conv_out"] = _Conv(sd, f"{p}{n + 2}.", False, 1, dev, lo)
        return d

    # ------------------------------------------------------------------ kernels
    def _conv(self, x, conv, out=None, act=None, snake=None, res=None, dil=1, pad=0, mode=0):
        """x: _Planes; returns nothing (fills out / act planes)."""
        B, T, Cin = x.B, x.T, x.C
        rc = lib().b200sat_conv1d_fwd(x.hi.data_ptr(), _p(x.lo), conv.w_hi.data_ptr(), _p(conv.w_lo), _p(conv.bias),
                                      _p(res.hi) if res else 0, _p(res.lo) if res else 0,
                                      _p(out.hi) if out else 0, _p(out.lo) if out else 0,
                                      _p(act.hi) if act else 0, _p(act.lo) if act else 0,
                                      _p(snake.a) if snake else 0, _p(snake.invb) if snake else 0,
                                      B, T, Cin, conv.Cout, conv.K, dil, pad, conv.stride, mode, self.passes, _s())
        ops.LAUNCHES[0] += 1
        check(rc, "conv1d_fwd")

    def _planes(self, B, T, C):
        return _Planes(B, T, C, self.device, self.with_lo)

    def _residual_units(self, x_raw, x_act, rus, next_snake):
        """Three ResidualUnits (autoencoders.py:58-83).  x_raw: input planes; x_act = snake_{ru0.s0}(x_raw).
        Returns (raw, act) of the last unit where act = next_snake(raw) (or None)."""
        B, T, C = x_raw.B, x_raw.T, x_raw.C
        # B200SAT_FUSED_RU=1: one launch per unit (csrc/residual_unit.cu).  Measured on B200 (tools/ru_bench.py, T = 2 097 152): fused 983 us
        # vs 386 + 447 us for the two launches - both are EPILOGUE-bound (SnakeBeta: ~10 instructions per element), so serialising the
        # two epilogues on the same warps costs more than the 1.1 GB of HBM traffic it saves.  Off by default.
        fused = self.passes == 1 and C == 128 and os.environ.get("B200SAT_FUSED_RU", "0") == "1"
        for j, ru in enumerate(rus):
            dil = (1, 3, 9)[j]
            if fused:
                # bf16 mode, C = 128: the whole unit in one launch (csrc/residual_unit.cu) - the k7 output never reaches HBM
                nxt = rus[j + 1]["s0"] if j + 1 < len(rus) else next_snake
                y_raw = self._planes(B, T, C)
                y_act = self._planes(B, T, C) if nxt is not None else None
                c7, c1, s1 = ru["c7"], ru["c1"], ru["s1"]
                rc = lib().b200sat_residual_unit_fwd(x_act.hi.data_ptr(), x_raw.hi.data_ptr(), c7.w_hi.data_ptr(), _p(c7.bias), s1.a.data_ptr(),
                                                     s1.invb.data_ptr(), c1.w_hi.data_ptr(), _p(c1.bias), _p(nxt.a) if nxt else 0,
                                                     _p(nxt.invb) if nxt else 0, y_raw.hi.data_ptr(), _p(y_act.hi) if y_act else 0,
                                                     B, T, C, dil, _s())
                ops.LAUNCHES[0] += 1
                check(rc, "residual_unit_fwd")
                x_raw, x_act = y_raw, y_act
                continue
            h = self._planes(B, T, C)                                   # snake_{s1}(conv7(x_act))
            self._conv(x_act, ru["c7"], act=h, snake=ru["s1"], dil=dil, pad=3 * dil)
            nxt = rus[j + 1]["s0"] if j + 1 < len(rus) else next_snake
            y_raw = self._planes(B, T, C)
            y_act = self._planes(B, T, C) if nxt is not None else None
            self._conv(h, ru["c1"], out=y_raw, act=y_act, snake=nxt, res=x_raw)   # conv1 + skip, then the next Snake
            x_raw, x_act = y_raw, y_act
        return x_raw, x_act

    # ------------------------------------------------------------------ forward passes
    @torch.no_grad()
    def encode(self, audio, noise=None, return_info=False):
        """audio fp32 [B, Cin, T] -> latents fp32 [B, L, T/ratio] (VAE sample with `noise`, or the mean when noise is None
        and... see below).  With return_info also returns {'kl', 'mean_scale'}."""
        e = self.enc
        dev = self.device
        x = audio.to(dev, torch.float32).contiguous()
        B, Cin, T = x.shape
        c0 = e["conv_in"]
        first = e["blocks"][0]["rus"][0]["s0"]
        raw, act = self._planes(B, T, c0.Cout), self._planes(B, T, c0.Cout)
        rc = lib().b200sat_conv_in(x.data_ptr(), c0.w_f32.data_ptr(), _p(c0.bias), first.a.data_ptr(), first.invb.data_ptr(),
                                   raw.hi.data_ptr(), _p(raw.lo), act.hi.data_ptr(), _p(act.lo), B, Cin, T, c0.Cout, c0.K, c0.K // 2, _s())
        ops.LAUNCHES[0] += 1
        check(rc, "conv_in")
        for bi, blk in enumerate(e["blocks"]):
            raw, act = self._residual_units(raw, act, blk["rus"], blk["snake"])
            s = blk["stride"]
            nxt = e["blocks"][bi + 1]["rus"][0]["s0"] if bi + 1 < len(e["blocks"]) else e["snake_out"]
            down = blk["down"]
            T2 = (raw.T + 2 * math.ceil(s / 2) - down.K) // s + 1
            need_raw = bi + 1 < len(e["blocks"])
            o_raw = self._planes(B, T2, down.Cout) if need_raw else None
            o_act = self._planes(B, T2, down.Cout)
            self._conv(act, down, out=o_raw, act=o_act, snake=nxt, pad=math.ceil(s / 2), mode=1)
            raw, act = o_raw, o_act
        co = e["conv_out"]
        ms = self._planes(B, act.T, co.Cout)
        self._conv(act, co, out=ms, pad=co.K // 2)
        L = co.Cout // 2
        Tl = ms.T
        z = torch.empty(B, L, Tl, device=dev)
        info_ms = torch.empty(B, 2 * L, Tl, device=dev) if return_info else None
        kl = torch.zeros(1, device=dev) if return_info else None
        nz = None if noise is None else noise.to(dev, torch.float32).contiguous()
        rc = lib().b200sat_vae_sample(ms.hi.data_ptr(), _p(ms.lo), _p(nz), z.data_ptr(), _p(info_ms), _p(kl), B, L, Tl, _s())
        ops.LAUNCHES[0] += 1
        check(rc, "vae_sample")
        if return_info:
            return z, {"kl": (kl / (B * Tl)).squeeze(0), "mean_scale": info_ms}
        return z

    @torch.no_grad()
    def decode(self, latents):
        """latents fp32 [B, L, Tl] -> audio fp32 [B, Cout, Tl * ratio]."""
        d = self.dec
        dev = self.device
        z = latents.to(dev, torch.float32).contiguous()
        B, L, Tl = z.shape
        zp = self._planes(B, Tl, L)
        rc = lib().b200sat_to_planes(z.data_ptr(), zp.hi.data_ptr(), _p(zp.lo), B, L, Tl, _s())
        ops.LAUNCHES[0] += 1
        check(rc, "to_planes")
        c0 = d["conv_in"]
        act = self._planes(B, Tl, c0.Cout)
        self._conv(zp, c0, act=act, snake=d["blocks"][0]["snake"], pad=c0.K // 2)
        for bi, blk in enumerate(d["blocks"]):
            s = blk["stride"]
            up = blk["up"]
            T2 = (act.T - 1) * s - 2 * math.ceil(s / 2) + up.K
            raw, a2 = self._planes(B, T2, up.Cout), self._planes(B, T2, up.Cout)
            self._conv(act, up, out=raw, act=a2, snake=blk["rus"][0]["s0"], pad=math.ceil(s / 2), mode=2)
            nxt = d["blocks"][bi + 1]["snake"] if bi + 1 < len(d["blocks"]) else d["snake_out"]
            raw, act = self._residual_units(raw, a2, blk["rus"], nxt)
        co = d["conv_out"]
        y = torch.empty(B, co.Cout, act.T, device=dev)
        rc = lib().b200sat_conv_out(act.hi.data_ptr(), _p(act.lo), co.w_f32.data_ptr(), _p(co.bias), y.data_ptr(), B, co.Cin, act.T,
                                    co.Cout, co.K, co.K // 2, int(self.final_tanh), _s())
        ops.LAUNCHES[0] += 1
        check(rc, "conv_out")
        return y


    # ------------------------------------------------------------------ orchestration of AudioAutoencoder (autoencoders.py:446-534, 601-732)
    @torch.no_grad()
    def encode_audio(self, audio, noise=None, iterate_batch=False):
        """`AudioAutoencoder.encode` + VAE bottleneck: optionally one item at a time (`iterate_batch`, autoencoders.py:470-474),
        which bounds the live activation set to one clip (1 GB planes per tensor for a 47 s clip)."""
        if not iterate_batch or audio.shape[0] == 1:
            return self.encode(audio, noise=noise)
        outs = []
        for i in range(audio.shape[0]):
            outs.append(self.encode(audio[i:i + 1], noise=None if noise is None else noise[i:i + 1]))
        return torch.cat(outs, dim=0)

    @torch.no_grad()
    def decode_audio(self, latents, chunked=False, overlap=32, chunk_size=128):
        """`AudioAutoencoder.decode_audio` (autoencoders.py:671-732): full decode, or overlap-discard chunked decode for long
        latents: chunks of `chunk_size` latents hopping by chunk_size - overlap; each chunk contributes its interior (half an
        overlap trimmed at every internal edge) so the zero-padded conv edges never reach the output."""
        if not chunked or latents.shape[-1] <= chunk_size:
            return self.decode(latents)
        ratio = 1
        for s_ in self.strides:
            ratio *= s_
        B, L, total = latents.shape
        hop = chunk_size - overlap
        starts = list(range(0, total - chunk_size + 1, hop))
        if starts[-1] != total - chunk_size:
            starts.append(total - chunk_size)
        out = None
        for ci, st in enumerate(starts):
            y = self.decode(latents[:, :, st:st + chunk_size])
            if out is None:
                out = torch.zeros(B, y.shape[1], total * ratio, device=y.device, dtype=y.dtype)
            lo = 0 if ci == 0 else (overlap // 2) * ratio
            hi = chunk_size * ratio if ci == len(starts) - 1 else (chunk_size - overlap // 2) * ratio
            out[:, :, st * ratio + lo: st * ratio + hi] = y[:, :, lo:hi]
        return out

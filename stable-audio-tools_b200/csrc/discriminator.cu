// b200sat — Encodec multi-scale STFT discriminator (models/encodec.py:38-138, models/discriminators.py:13-58): everything around the
// 64 -> 64 channel 2-D convs (which run on the tcgen05 conv kernel through b200sat_conv2d_flat).
//
// Layout.  One scale's activations are a flattened time-major plane [B, P, C], P = frames * Fp, Fp = F + 8: the F = n_fft/2 + 1 bins
// of a frame sit in columns [4, 4 + F) of its Fp-wide row group, the 4 + 4 pad columns hold zeros.  A 2-D tap (dt, df) of a conv with
// dilation (d, 1) is the row shift dt * d * Fp + df; frequency padding = the zero columns, time padding = out-of-range rows.
//   spectrogram   fp32 [B, P, 4]   channels (re ch0, re ch1, im ch0, im ch1) = torch.cat([z.real, z.imag], dim=1) for stereo
//   feature maps  bf16 [B, P, 64]  (post LeakyReLU; pad columns zero)
//   logits        fp32 [B, P]
#include "common.cuh"

namespace b200sat {

__device__ __forceinline__ bool col_valid(int p, int Fp, int F) {
  const int f = p % Fp;
  return f >= 4 && f < 4 + F;
}

// ---------------------------------------------------------------------------------------------------------------------------------
// STFT front end: torchaudio Spectrogram(n_fft, hop, win = n_fft, hann, normalized=True, center=False, power=None) (encodec.py:72-74).
// One warp = one frame; both audio channels ride one complex FFT (z = ch0 + i ch1) and are separated by conjugate symmetry.
__global__ void __launch_bounds__(256) disc_stft_fwd_kernel(const float* __restrict__ x, float* __restrict__ spec, const float* __restrict__ window,
                                                            const float2* __restrict__ twiddle, int T, int n, int log2n, int hop, int frames,
                                                            int frames_per_block, int Fp, float norm) {
  extern __shared__ float2 fft_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;
  float2* buf0 = fft_smem + static_cast<size_t>(warp) * 2 * n;
  float2* buf1 = buf0 + n;
  const int b = blockIdx.y;
  const float* x0 = x + static_cast<long>(b) * 2 * T;
  const float* x1 = x0 + T;
  const long P = static_cast<long>(frames) * Fp;
  float4* sp = reinterpret_cast<float4*>(spec) + static_cast<long>(b) * P;
  const int f_begin = blockIdx.x * frames_per_block;
  const int f_end = min(frames, f_begin + frames_per_block);
  const int half = n >> 1;
  for (int f = f_begin + warp; f < f_end; f += nwarps) {
    const int base = f * hop;
    for (int i = lane; i < n; i += 32) {
      const float w = __ldg(window + i);
      buf0[i] = make_float2(__ldg(x0 + base + i) * w, __ldg(x1 + base + i) * w);
    }
    __syncwarp();
    float2* in = buf0;
    float2* out = buf1;
    for (int s = 0; s < log2n; ++s) {
      const int Ns = 1 << s, tw_stride = half >> s;
      for (int j = lane; j < half; j += 32) {
        const int k = j & (Ns - 1);
        const float2 w = __ldg(twiddle + k * tw_stride);
        const float2 a = in[j], bb = in[j + half];
        const float2 bw = make_float2(bb.x * w.x - bb.y * w.y, bb.x * w.y + bb.y * w.x);
        const int j0 = ((j - k) << 1) + k;
        out[j0] = make_float2(a.x + bw.x, a.y + bw.y);
        out[j0 + Ns] = make_float2(a.x - bw.x, a.y - bw.y);
      }
      __syncwarp();
      float2* tmp = in; in = out; out = tmp;
    }
    for (int k = lane; k <= half; k += 32) {
      const float2 zk = in[k];
      const float2 zn = in[(n - k) & (n - 1)];
      const float xr = 0.5f * (zk.x + zn.x), xi = 0.5f * (zk.y - zn.y);
      const float yr = 0.5f * (zk.y + zn.y), yi = 0.5f * (zn.x - zk.x);
      sp[static_cast<long>(f) * Fp + 4 + k] = make_float4(xr * norm, yr * norm, xi * norm, yi * norm);
    }
    __syncwarp();
  }
}

// Backward: d spec -> d audio.  The per-bin gradients of both channels are Hermitian-extended and packed into one spectrum, one
// inverse-direction FFT returns d/d(ch0 * w) in the real part and d/d(ch1 * w) in the imaginary part; windowed, scattered with atomics.
__global__ void __launch_bounds__(256) disc_stft_bwd_kernel(const float* __restrict__ dspec, float* __restrict__ dx, const float* __restrict__ window,
                                                            const float2* __restrict__ twiddle, int T, int n, int log2n, int hop, int frames,
                                                            int frames_per_block, int Fp, float norm) {
  extern __shared__ float2 fft_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;
  float2* buf0 = fft_smem + static_cast<size_t>(warp) * 2 * n;
  float2* buf1 = buf0 + n;
  const int b = blockIdx.y;
  float* d0 = dx + static_cast<long>(b) * 2 * T;
  float* d1 = d0 + T;
  const long P = static_cast<long>(frames) * Fp;
  const float4* sp = reinterpret_cast<const float4*>(dspec) + static_cast<long>(b) * P;
  const int f_begin = blockIdx.x * frames_per_block;
  const int f_end = min(frames, f_begin + frames_per_block);
  const int half = n >> 1;
  for (int f = f_begin + warp; f < f_end; f += nwarps) {
    for (int k = lane; k <= half; k += 32) {
      const float4 g = __ldg(sp + static_cast<long>(f) * Fp + 4 + k);
      const float gxr = g.x * norm, gyr = g.y * norm, gxi = g.z * norm, gyi = g.w * norm;
      if (k == 0 || k == half) {
        buf0[k] = make_float2(gxr, gyr);
      } else {
        buf0[k] = make_float2(0.5f * (gxr - gyi), 0.5f * (gxi + gyr));
        buf0[n - k] = make_float2(0.5f * (gxr + gyi), 0.5f * (-gxi + gyr));
      }
    }
    __syncwarp();
    float2* in = buf0;
    float2* out = buf1;
    for (int s = 0; s < log2n; ++s) {
      const int Ns = 1 << s, tw_stride = half >> s;
      for (int j = lane; j < half; j += 32) {
        const int k = j & (Ns - 1);
        float2 w = __ldg(twiddle + k * tw_stride);
        w.y = -w.y;
        const float2 a = in[j], bb = in[j + half];
        const float2 bw = make_float2(bb.x * w.x - bb.y * w.y, bb.x * w.y + bb.y * w.x);
        const int j0 = ((j - k) << 1) + k;
        out[j0] = make_float2(a.x + bw.x, a.y + bw.y);
        out[j0 + Ns] = make_float2(a.x - bw.x, a.y - bw.y);
      }
      __syncwarp();
      float2* tmp = in; in = out; out = tmp;
    }
    const int base = f * hop;
    for (int i = lane; i < n; i += 32) {
      const float w = __ldg(window + i);
      const float2 z = in[i];
      atomicAdd(d0 + base + i, z.x * w);
      atomicAdd(d1 + base + i, z.y * w);
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------------------------------------------
// First conv (4 -> 64 channels, 3 x 9, encodec.py:77-79) + LeakyReLU.  thread = one output position with 64 accumulators; the
// 27 x 4 x 64 weights sit in shared memory as [tap][ci][co] and are read as broadcasts.
constexpr int D0_TAPS = 27;
__global__ void __launch_bounds__(128) disc_conv0_fwd_kernel(const float* __restrict__ spec, const float* __restrict__ w /*[64][4][27]*/,
                                                             const float* __restrict__ bias, __nv_bfloat16* __restrict__ out, int B, int frames,
                                                             int Fp, int F, float leaky) {
  __shared__ float sw[D0_TAPS * 4 * 64];
  __shared__ float sb[64];
  for (int i = threadIdx.x; i < D0_TAPS * 4 * 64; i += blockDim.x) {
    const int tap = i / 256, ci = (i / 64) % 4, co = i % 64;
    sw[i] = w[(co * 4 + ci) * D0_TAPS + tap];
  }
  if (threadIdx.x < 64) sb[threadIdx.x] = bias ? bias[threadIdx.x] : 0.f;
  __syncthreads();
  const long P = static_cast<long>(frames) * Fp;
  const long total = static_cast<long>(B) * P;
  for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total; idx += static_cast<long>(gridDim.x) * blockDim.x) {
    const int b = idx / P;
    const long p = idx % P;
    __nv_bfloat16* orow = out + idx * 64;
    if (!col_valid(static_cast<int>(p % Fp), Fp, F)) {
#pragma unroll
      for (int i = 0; i < 8; ++i) reinterpret_cast<uint4*>(orow)[i] = make_uint4(0, 0, 0, 0);
      continue;
    }
    float acc[64];
#pragma unroll
    for (int co = 0; co < 64; ++co) acc[co] = sb[co];
    const float4* sp = reinterpret_cast<const float4*>(spec) + static_cast<long>(b) * P;
#pragma unroll 1
    for (int tap = 0; tap < D0_TAPS; ++tap) {
      const long q = p + (tap / 9 - 1) * Fp + (tap % 9 - 4);
      if (q < 0 || q >= P) continue;
      const float4 v = __ldg(sp + q);
      const float* wt = sw + tap * 256;
#pragma unroll
      for (int c4 = 0; c4 < 16; ++c4) {
        const float4 w0 = *reinterpret_cast<const float4*>(wt + c4 * 4);
        const float4 w1 = *reinterpret_cast<const float4*>(wt + 64 + c4 * 4);
        const float4 w2 = *reinterpret_cast<const float4*>(wt + 128 + c4 * 4);
        const float4 w3 = *reinterpret_cast<const float4*>(wt + 192 + c4 * 4);
        acc[c4 * 4 + 0] += v.x * w0.x + v.y * w1.x + v.z * w2.x + v.w * w3.x;
        acc[c4 * 4 + 1] += v.x * w0.y + v.y * w1.y + v.z * w2.y + v.w * w3.y;
        acc[c4 * 4 + 2] += v.x * w0.z + v.y * w1.z + v.z * w2.z + v.w * w3.z;
        acc[c4 * 4 + 3] += v.x * w0.w + v.y * w1.w + v.z * w2.w + v.w * w3.w;
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      uint32_t pk[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float a0 = acc[i * 8 + 2 * e], a1 = acc[i * 8 + 2 * e + 1];
        a0 = a0 > 0.f ? a0 : a0 * leaky; a1 = a1 > 0.f ? a1 : a1 * leaky;
        pk[e] = pack_bf16(a0, a1);
      }
      reinterpret_cast<uint4*>(orow)[i] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
  }
}

// Data gradient of the first conv: d spec[p][ci] = sum_{tap, co} d_pre[p - off(tap)][co] * w[co][ci][tap].
__global__ void __launch_bounds__(128) disc_conv0_dgrad_kernel(const __nv_bfloat16* __restrict__ dpre, const float* __restrict__ w /*[64][4][27]*/,
                                                               float* __restrict__ dspec, int B, int frames, int Fp, int F) {
  __shared__ float4 sw[D0_TAPS * 64];   // [tap][co] -> (ci 0..3)
  for (int i = threadIdx.x; i < D0_TAPS * 64; i += blockDim.x) {
    const int tap = i / 64, co = i % 64;
    sw[i] = make_float4(w[(co * 4 + 0) * D0_TAPS + tap], w[(co * 4 + 1) * D0_TAPS + tap], w[(co * 4 + 2) * D0_TAPS + tap], w[(co * 4 + 3) * D0_TAPS + tap]);
  }
  __syncthreads();
  const long P = static_cast<long>(frames) * Fp;
  const long total = static_cast<long>(B) * P;
  for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total; idx += static_cast<long>(gridDim.x) * blockDim.x) {
    const int b = idx / P;
    const long p = idx % P;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (col_valid(static_cast<int>(p % Fp), Fp, F)) {
      const __nv_bfloat16* base = dpre + static_cast<long>(b) * P * 64;
#pragma unroll 1
      for (int tap = 0; tap < D0_TAPS; ++tap) {
        const long q = p - ((tap / 9 - 1) * Fp + (tap % 9 - 4));
        if (q < 0 || q >= P) continue;
        const uint4* row = reinterpret_cast<const uint4*>(base + q * 64);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint4 u = __ldg(row + i);
          const uint32_t uw[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 d = unpack_bf16(uw[e]);
            const float4 w0 = sw[tap * 64 + i * 8 + 2 * e], w1 = sw[tap * 64 + i * 8 + 2 * e + 1];
            acc.x += d.x * w0.x + d.y * w1.x; acc.y += d.x * w0.y + d.y * w1.y;
            acc.z += d.x * w0.z + d.y * w1.z; acc.w += d.x * w0.w + d.y * w1.w;
          }
        }
      }
    }
    reinterpret_cast<float4*>(dspec)[idx] = acc;
  }
}

// Last conv (64 -> 1, 3 x 3, no activation; encodec.py:88-90): logits[p] = b + sum_{tap, c} act[p + off][c] * w[c][tap].
__global__ void __launch_bounds__(128) disc_convpost_fwd_v1_kernel(const __nv_bfloat16* __restrict__ act, const float* __restrict__ w /*[1][64][9]*/,
                                                                const float* __restrict__ bias, float* __restrict__ logits, int B, int frames,
                                                                int Fp, int F) {
  __shared__ float sw[9 * 64];   // [tap][c]
  for (int i = threadIdx.x; i < 9 * 64; i += blockDim.x) sw[i] = w[(i % 64) * 9 + i / 64];
  __syncthreads();
  const long P = static_cast<long>(frames) * Fp;
  const long total = static_cast<long>(B) * P;
  const float bv = bias ? bias[0] : 0.f;
  for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total; idx += static_cast<long>(gridDim.x) * blockDim.x) {
    const int b = idx / P;
    const long p = idx % P;
    float acc = 0.f;
    const bool ok = col_valid(static_cast<int>(p % Fp), Fp, F);
    if (ok) {
      acc = bv;
      const __nv_bfloat16* base = act + static_cast<long>(b) * P * 64;
#pragma unroll 1
      for (int tap = 0; tap < 9; ++tap) {
        const long q = p + (tap / 3 - 1) * Fp + (tap % 3 - 1);
        if (q < 0 || q >= P) continue;
        const uint4* row = reinterpret_cast<const uint4*>(base + q * 64);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint4 u = __ldg(row + i);
          const uint32_t uw[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 d = unpack_bf16(uw[e]);
            acc += d.x * sw[tap * 64 + i * 8 + 2 * e] + d.y * sw[tap * 64 + i * 8 + 2 * e + 1];
          }
        }
      }
    }
    logits[idx] = acc;
  }
}

// ---------------------------------------------------------------------------------------------------------------------------------
// Loss reductions (discriminators.py:13-58).  sums[0] += sum relu(1 - lt), sums[1] += sum relu(1 + lf), sums[2] += sum lf over valid bins.
__global__ void __launch_bounds__(256) disc_hinge_sums_kernel(const float* __restrict__ lt, const float* __restrict__ lf, double* __restrict__ sums,
                                                              long total, int Fp, int F) {
  float s0 = 0.f, s1 = 0.f, s2 = 0.f;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
    if (!col_valid(static_cast<int>(i % Fp), Fp, F)) continue;   // P is a multiple of Fp, so i % Fp is the column for every batch entry
    if (lt) s0 += fmaxf(1.f - lt[i], 0.f);
    if (lf) { s1 += fmaxf(1.f + lf[i], 0.f); s2 += lf[i]; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(sums + 0, static_cast<double>(s0)); atomicAdd(sums + 1, static_cast<double>(s1)); atomicAdd(sums + 2, static_cast<double>(s2));
  }
}

// out[0] += sum |a - b| over two bf16 planes (feature matching, discriminators.py:24).  Pad columns are zero in both.
__global__ void __launch_bounds__(256) disc_l1_sum_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                                                          double* __restrict__ out, long n8) {
  float s = 0.f;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n8; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const uint4 ua = __ldg(reinterpret_cast<const uint4*>(a) + i), ub = __ldg(reinterpret_cast<const uint4*>(b) + i);
    const uint32_t aw[4] = {ua.x, ua.y, ua.z, ua.w}, bw[4] = {ub.x, ub.y, ub.z, ub.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 x = unpack_bf16(aw[e]), y = unpack_bf16(bw[e]);
      s += fabsf(x.x - y.x) + fabsf(x.y - y.y);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, static_cast<double>(s));
}

// d logits for the three uses of a logit map: mode 0 generator (adv = -mean lf): g = -scale; mode 1 discriminator on reals
// (relu(1 - l)): g = -scale [1 - l > 0]; mode 2 discriminator on fakes (relu(1 + l)): g = +scale [1 + l > 0].  Pad columns get 0.
__global__ void __launch_bounds__(256) disc_logit_grad_kernel(const float* __restrict__ logits, float* __restrict__ g, long total, int Fp, int F,
                                                              int mode, float scale) {
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
    float v = 0.f;
    if (col_valid(static_cast<int>(i % Fp), Fp, F)) {
      const float l = logits[i];
      v = (mode == 0) ? -scale : (mode == 1 ? ((1.f - l > 0.f) ? -scale : 0.f) : ((1.f + l > 0.f) ? scale : 0.f));
    }
    g[i] = v;
  }
}

// Backward through one feature map: d_post = d_in (from the next conv's dgrad, optional) + conv_post^T(d logits) (last layer only,
// optional) + fm_coef * sign(post - other) (feature matching: d/d post of mean |other - post|; optional);
// d_pre = d_post * (post > 0 ? 1 : leaky), zero in the pad columns.  thread = 8 channels of one position.
__global__ void __launch_bounds__(256) disc_act_bwd_kernel(const __nv_bfloat16* __restrict__ d_in, const float* __restrict__ d_logit,
                                                           const float* __restrict__ w_post /*[64][9]*/, const __nv_bfloat16* __restrict__ post,
                                                           const __nv_bfloat16* __restrict__ other, float fm_coef, float leaky,
                                                           __nv_bfloat16* __restrict__ d_pre, int B, int frames, int Fp, int F) {
  __shared__ float sw[9 * 64];   // [tap][c]
  if (d_logit) {
    for (int i = threadIdx.x; i < 9 * 64; i += blockDim.x) sw[i] = w_post[(i % 64) * 9 + i / 64];
  }
  __syncthreads();
  // 32-bit position arithmetic (entry point: B * P < 2^27); the 64-bit `/` and `%` of round 1 dominated this kernel's instruction count
  const int P = frames * Fp;
  const int total = B * P * 8;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int pos = idx >> 3;
    const int ch = idx & 7;
    const int b = pos / P;
    const int p = pos - b * P;
    uint4 o = make_uint4(0, 0, 0, 0);
    if (col_valid(p, Fp, F)) {
      float d[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) d[j] = 0.f;
      if (d_in) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(d_in) + idx);
        const uint32_t uw[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) { const float2 f = unpack_bf16(uw[e]); d[2 * e] = f.x; d[2 * e + 1] = f.y; }
      }
      if (d_logit) {
        const float* gl = d_logit + static_cast<size_t>(b) * P;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const int q = p - ((tap / 3 - 1) * Fp + (tap % 3 - 1));
          const int qc = min(max(q, 0), P - 1);
          const float gt = __ldg(gl + qc);
          const float g = (q == qc) ? gt : 0.f;
#pragma unroll
          for (int j = 0; j < 8; ++j) d[j] += g * sw[tap * 64 + ch * 8 + j];
        }
      }
      const uint4 up = __ldg(reinterpret_cast<const uint4*>(post) + idx);
      const uint32_t pw[4] = {up.x, up.y, up.z, up.w};
      float pv[8];
#pragma unroll
      for (int e = 0; e < 4; ++e) { const float2 f = unpack_bf16(pw[e]); pv[2 * e] = f.x; pv[2 * e + 1] = f.y; }
      if (other) {
        const uint4 uo = __ldg(reinterpret_cast<const uint4*>(other) + idx);
        const uint32_t ow[4] = {uo.x, uo.y, uo.z, uo.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = unpack_bf16(ow[e]);
          const float e0 = pv[2 * e] - f.x, e1 = pv[2 * e + 1] - f.y;
          d[2 * e] += fm_coef * ((e0 > 0.f) ? 1.f : ((e0 < 0.f) ? -1.f : 0.f));
          d[2 * e + 1] += fm_coef * ((e1 > 0.f) ? 1.f : ((e1 < 0.f) ? -1.f : 0.f));
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) d[j] *= (pv[j] > 0.f) ? 1.f : leaky;
      o = make_uint4(pack_bf16(d[0], d[1]), pack_bf16(d[2], d[3]), pack_bf16(d[4], d[5]), pack_bf16(d[6], d[7]));
    }
    reinterpret_cast<uint4*>(d_pre)[idx] = o;
  }
}


// ---------------------------------------------------------------------------------------------------------------------------------
// Weight gradients of the two SIMT layers (the D step).  Persistent blocks keep their partial sums in registers over all their
// positions and issue one atomic per weight at the end.
// first conv: dW[co][ci][tap] += sum_p d_pre[p][co] * spec[p + off(tap)][ci];  thread = (co, tap group of 7), float4 over ci.
__global__ void __launch_bounds__(256) disc_conv0_wgrad_kernel(const __nv_bfloat16* __restrict__ dpre, const float* __restrict__ spec,
                                                               float* __restrict__ dW /*[64][4][27]*/, int B, int frames, int Fp, int F) {
  const int co = threadIdx.x & 63, tg = threadIdx.x >> 6;
  float4 acc[7];
#pragma unroll
  for (int i = 0; i < 7; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  const long P = static_cast<long>(frames) * Fp;
  const long total = static_cast<long>(B) * P;
  // eight positions per iteration so that their loads are in flight together (one position per iteration is latency-bound)
  for (long i0 = static_cast<long>(blockIdx.x) * 8; i0 < total; i0 += static_cast<long>(gridDim.x) * 8) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const long idx = i0 + u;
      if (idx >= total) break;
      const long p = idx % P;
      if (!col_valid(static_cast<int>(p % Fp), Fp, F)) continue;
      const float d = __bfloat162float(dpre[idx * 64 + co]);
      const float4* sp = reinterpret_cast<const float4*>(spec) + (idx - p);
#pragma unroll
      for (int i = 0; i < 7; ++i) {
        const int tap = tg + 4 * i;
        if (tap < D0_TAPS) {
          const long q = p + (tap / 9 - 1) * Fp + (tap % 9 - 4);
          if (q >= 0 && q < P) {
            const float4 v = __ldg(sp + q);
            acc[i].x += d * v.x; acc[i].y += d * v.y; acc[i].z += d * v.z; acc[i].w += d * v.w;
          }
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    const int tap = tg + 4 * i;
    if (tap < D0_TAPS) {
      atomicAdd(dW + (co * 4 + 0) * D0_TAPS + tap, acc[i].x); atomicAdd(dW + (co * 4 + 1) * D0_TAPS + tap, acc[i].y);
      atomicAdd(dW + (co * 4 + 2) * D0_TAPS + tap, acc[i].z); atomicAdd(dW + (co * 4 + 3) * D0_TAPS + tap, acc[i].w);
    }
  }
}

// conv_post: dW[c][tap] += sum_p g[p] * act[p + off(tap)][c];  dbias += sum_p g[p].   thread = (c, tap group of 3)
__global__ void __launch_bounds__(256) disc_convpost_wgrad_v1_kernel(const float* __restrict__ g, const __nv_bfloat16* __restrict__ act,
                                                                  float* __restrict__ dW /*[64][9]*/, float* __restrict__ dbias, int B, int frames,
                                                                  int Fp, int F) {
  const int c = threadIdx.x & 63, tg = threadIdx.x >> 6;
  float acc[3] = {0.f, 0.f, 0.f};
  float gsum = 0.f;
  const long P = static_cast<long>(frames) * Fp;
  const long total = static_cast<long>(B) * P;
  for (long i0 = static_cast<long>(blockIdx.x) * 8; i0 < total; i0 += static_cast<long>(gridDim.x) * 8) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const long idx = i0 + u;
      if (idx >= total) break;
      const long p = idx % P;
      const float gv = __ldg(g + idx);
      if (gv == 0.f) continue;          // pad columns and inactive hinge positions
      gsum += gv;
      const __nv_bfloat16* base = act + (idx - p) * 64;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int tap = tg + 4 * i;
        if (tap < 9) {
          const long q = p + (tap / 3 - 1) * Fp + (tap % 3 - 1);
          if (q >= 0 && q < P) acc[i] += gv * __bfloat162float(base[q * 64 + c]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int tap = tg + 4 * i;
    if (tap < 9) atomicAdd(dW + c * 9 + tap, acc[i]);
  }
  if (threadIdx.x == 0 && dbias) atomicAdd(dbias, gsum);
}

// ---------------------------------------------------------------------------------------------------------------------------------
// Round 2: row-coalesced versions of the two conv_post kernels.  The v1 kernels above give one thread a whole 128-byte activation row
// (32 different cache lines per warp load: 72 such loads per output position made the forward L1-wavefront bound at 1.2 ms per scale
// and batch 32) or one channel (2-byte loads).  Here EIGHT lanes share a row (one 16-byte chunk each, so a warp load touches four
// full lines) and the per-row dot products are finished with three shuffles.
__global__ void __launch_bounds__(256, 2) disc_convpost_fwd_kernel(const __nv_bfloat16* __restrict__ act, const float* __restrict__ w /*[1][64][9]*/,
                                                                   const float* __restrict__ bias, float* __restrict__ logits, int B, int frames,
                                                                   int Fp, int F) {
  const int lane = threadIdx.x & 31, l8 = lane & 7, sub = lane >> 3;
  // this lane's 9 x 8 weights stay in registers: reading them from shared memory per row made the kernel shared-memory bound
  // (ncu: 152 M shared wavefronts for 4.1 M rows, 0.75 ms); taps are loaded three at a time so the register budget allows two blocks per SM
  float wr[9][8];
#pragma unroll
  for (int tap = 0; tap < 9; ++tap)
#pragma unroll
    for (int j = 0; j < 8; ++j) wr[tap][j] = __ldg(w + (l8 * 8 + j) * 9 + tap);
  // 32-bit position arithmetic (the entry point checks B * P < 2^27)
  const int P = frames * Fp;
  const int total = B * P;
  const float bv = bias ? bias[0] : 0.f;
  const int stride = gridDim.x * (blockDim.x >> 3);
  for (int base = (blockIdx.x * blockDim.x + (threadIdx.x & ~31)) >> 3; base < total; base += stride) {   // warp-uniform
    const int idx = base + sub;
    const bool in = idx < total;
    const int p = in ? idx % P : 0;
    const bool ok = in && col_valid(p, Fp, F);
    float acc = 0.f;
    if (ok) {
      const __nv_bfloat16* rowbase = act + static_cast<size_t>(idx - p) * 64 + l8 * 8;
#pragma unroll
      for (int g3 = 0; g3 < 3; ++g3) {
        // three independent loads (row index clamped, contribution masked): a bounds branch per tap would serialise them behind each other
        uint4 u[3];
        float m[3];
#pragma unroll
        for (int t = 0; t < 3; ++t) {
          const int q = p + (g3 - 1) * Fp + (t - 1);
          const int qc = min(max(q, 0), P - 1);
          m[t] = (q == qc) ? 1.f : 0.f;
          u[t] = __ldg(reinterpret_cast<const uint4*>(rowbase + static_cast<size_t>(qc) * 64));
        }
#pragma unroll
        for (int t = 0; t < 3; ++t) {
          const int tap = g3 * 3 + t;
          const float2 d0 = unpack_bf16(u[t].x), d1 = unpack_bf16(u[t].y), d2 = unpack_bf16(u[t].z), d3 = unpack_bf16(u[t].w);
          acc += m[t] * (d0.x * wr[tap][0] + d0.y * wr[tap][1] + d1.x * wr[tap][2] + d1.y * wr[tap][3] + d2.x * wr[tap][4] + d2.y * wr[tap][5] +
                         d3.x * wr[tap][6] + d3.y * wr[tap][7]);
        }
      }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if (in && l8 == 0) logits[idx] = ok ? acc + bv : 0.f;
  }
}

// conv_post weight gradient, gather form: the activation row r is read ONCE and meets the nine logit gradients g[r - off(tap)]:
//   dW[c][tap] += sum_r act[r][c] * g[r - off(tap)]          dbias += sum_p g[p]
// lane = 8 channels of one row (72 fp32 partial sums per thread), four rows per warp step; partial sums meet in shared memory once.
__global__ void __launch_bounds__(256) disc_convpost_wgrad_kernel(const float* __restrict__ g, const __nv_bfloat16* __restrict__ act,
                                                                  float* __restrict__ dW /*[64][9]*/, float* __restrict__ dbias, int B, int frames,
                                                                  int Fp, int F) {
  __shared__ float red[9 * 64];
  __shared__ float red_g;
  const int lane = threadIdx.x & 31, l8 = lane & 7, sub = lane >> 3;
  for (int i = threadIdx.x; i < 9 * 64; i += blockDim.x) red[i] = 0.f;
  if (threadIdx.x == 0) red_g = 0.f;
  __syncthreads();
  float acc[9][8];
#pragma unroll
  for (int tap = 0; tap < 9; ++tap)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[tap][j] = 0.f;
  float gsum = 0.f;
  const int P = frames * Fp;          // 32-bit position arithmetic (entry point: B * P < 2^27)
  const int total = B * P;
  const int stride = gridDim.x * (blockDim.x >> 3);
  for (int idx = ((blockIdx.x * blockDim.x + threadIdx.x) >> 3); idx < total; idx += stride) {
    const int p = idx % P;
    if (!col_valid(p, Fp, F)) continue;     // activation rows and logit gradients are zero in the pad columns
    const float* gb = g + (idx - p);
    if (l8 == 0) gsum += __ldg(gb + p);
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(act + static_cast<size_t>(idx) * 64 + l8 * 8));
    const float2 d0 = unpack_bf16(u.x), d1 = unpack_bf16(u.y), d2 = unpack_bf16(u.z), d3 = unpack_bf16(u.w);
    const float a[8] = {d0.x, d0.y, d1.x, d1.y, d2.x, d2.y, d3.x, d3.y};
    float gv[9];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {     // independent loads: clamp the row, mask the value
      const int q = p - ((tap / 3 - 1) * Fp + (tap % 3 - 1));
      const int qc = min(max(q, 0), P - 1);
      const float t = __ldg(gb + qc);
      gv[tap] = (q == qc) ? t : 0.f;
    }
#pragma unroll
    for (int tap = 0; tap < 9; ++tap)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[tap][j] += gv[tap] * a[j];
  }
  // the four row groups of a warp hold the same channels: fold them, then one shared-memory add per (warp, tap, channel)
#pragma unroll
  for (int tap = 0; tap < 9; ++tap)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v = acc[tap][j];
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if (sub == 0) atomicAdd(&red[tap * 64 + l8 * 8 + j], v);
    }
  gsum += __shfl_xor_sync(0xffffffffu, gsum, 8);
  gsum += __shfl_xor_sync(0xffffffffu, gsum, 16);
  if (lane == 0) atomicAdd(&red_g, gsum);
  __syncthreads();
  for (int i = threadIdx.x; i < 9 * 64; i += blockDim.x) atomicAdd(dW + (i % 64) * 9 + i / 64, red[i]);
  if (threadIdx.x == 0 && dbias) atomicAdd(dbias, red_g);
}

// ---------------------------------------------------------------------------------------------------------------------------------
// Round 2: the first conv (4 -> 64 channels, 3 x 9; encodec.py:77-79) on the tensor cores.  Its nine frequency taps are folded into
// channels once per spectrogram:   S9[row][ci*9 + df] = spec[row + df - 4][ci]   (36 of 64 bf16 channels, the rest zero)
// which turns the layer into a 3-tap (dt = -1, 0, 1 -> row shifts -Fp, 0, Fp) 64 -> 64 channel flattened conv: forward and data gradient
// run on b200sat_conv2d_flat, the weight gradient on b200sat_conv_wgrad_taps_cat, weight-norm on b200sat_wn_pack / b200sat_wn_bwd - the
// same entries as the other four layers.  (The fp32 SIMT kernels above took 2.6 / 5 / 10 ms per scale at batch 32: forward / data /
// weight gradient.)  The spectrogram is rounded to bf16 here - what the reference's Conv2d does under bf16 autocast.
__global__ void __launch_bounds__(256) disc_spec_pack_kernel(const float* __restrict__ spec, __nv_bfloat16* __restrict__ s9, int total, int P,
                                                             int Fp, int F) {
  // thread = 8 channels of one position; 32-bit position arithmetic (entry point: B * P < 2^27)
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total * 8; idx += gridDim.x * blockDim.x) {
    const int pos = idx >> 3;
    const int ch8 = idx & 7;
    const int p = pos % P;
    uint4 o = make_uint4(0, 0, 0, 0);
    if (ch8 < 5 && col_valid(p, Fp, F)) {
      const float* sb = spec + static_cast<size_t>(pos - p) * 4;
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = ch8 * 8 + j;
        const int ci = c / 9, df = c - 9 * ci;
        const int q = p + df - 4;
        const int qc = min(max(q, 0), P - 1);
        const float t = __ldg(sb + qc * 4 + (c < 36 ? ci : 0));          // unconditional load, masked: the eight loads stay independent
        v[j] = (c < 36 && q == qc) ? t : 0.f;
      }
      o = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
    }
    reinterpret_cast<uint4*>(s9)[idx] = o;
  }
}

// d spec[q][ci] = sum_df dS9[q - df + 4][ci*9 + df]   (the transpose of the packing above); block = 128 consecutive rows of one item
__global__ void __launch_bounds__(128) disc_spec_unpack_kernel(const __nv_bfloat16* __restrict__ ds9, float* __restrict__ dspec, long P, int Fp, int F) {
  __shared__ __align__(16) __nv_bfloat16 sm[136][40];
  const long r0 = static_cast<long>(blockIdx.x) * 128;
  const __nv_bfloat16* base = ds9 + static_cast<long>(blockIdx.y) * P * 64;
  for (int i = threadIdx.x; i < 136 * 5; i += 128) {
    const int rr = i / 5, ck = i % 5;
    const long q = r0 - 4 + rr;
    uint4 u = make_uint4(0, 0, 0, 0);
    if (q >= 0 && q < P) u = __ldg(reinterpret_cast<const uint4*>(base + q * 64 + ck * 8));
    *reinterpret_cast<uint4*>(&sm[rr][ck * 8]) = u;
  }
  __syncthreads();
  const long q = r0 + threadIdx.x;
  if (q >= P) return;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  if (col_valid(static_cast<int>(q % Fp), Fp, F)) {
#pragma unroll
    for (int ci = 0; ci < 4; ++ci)
#pragma unroll
      for (int df = 0; df < 9; ++df) acc[ci] += __bfloat162float(sm[threadIdx.x + 8 - df][ci * 9 + df]);
  }
  reinterpret_cast<float4*>(dspec)[static_cast<long>(blockIdx.y) * P + q] = make_float4(acc[0], acc[1], acc[2], acc[3]);
}

static int fft_launch_cfg(int n_fft, int* log2n, int* warps, int* smem) {
  int l = 0;
  while ((1 << l) < n_fft) ++l;
  if ((1 << l) != n_fft || n_fft < 32 || n_fft > 4096) return -1;
  int w = 8;
  while (w > 1 && w * 2 * n_fft * 8 > 160 * 1024) w >>= 1;
  *log2n = l; *warps = w; *smem = w * 2 * n_fft * 8;
  return 0;
}

}  // namespace b200sat

using namespace b200sat;

static inline int grid_for(long total, int threads, int per_sm) {
  long g = (total + threads - 1) / threads;
  const long cap = static_cast<long>(num_sms()) * per_sm;
  if (g > cap) g = cap;
  return static_cast<int>(g < 1 ? 1 : g);
}

extern "C" int b200sat_disc_stft(const float* x, float* spec, const float* window, const float* twiddle, int B, int T, int n_fft, int hop,
                                 int backward, void* stream) {
  if (!x || !spec || !window || !twiddle || B <= 0 || T < n_fft || hop <= 0) { set_last_error("disc_stft: bad arguments"); return B200SAT_EINVAL; }
  int log2n, warps, smem;
  if (fft_launch_cfg(n_fft, &log2n, &warps, &smem)) { set_last_error("disc_stft: n_fft must be a power of two in [32, 4096]"); return B200SAT_EUNSUPPORTED; }
  const int frames = (T - n_fft) / hop + 1;
  const int Fp = n_fft / 2 + 1 + 8;
  // normalized=True: divide by sqrt(sum w^2); for the periodic hann window sum w^2 = 3 n / 8
  const float norm = 1.0f / sqrtf(0.375f * static_cast<float>(n_fft));
  static bool attr = false;
  if (!attr) {
    B200SAT_CHECK_CUDA(cudaFuncSetAttribute(disc_stft_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    B200SAT_CHECK_CUDA(cudaFuncSetAttribute(disc_stft_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    attr = true;
  }
  const int fpb = warps * 4;
  dim3 grid((frames + fpb - 1) / fpb, B);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!backward)
    disc_stft_fwd_kernel<<<grid, warps * 32, smem, s>>>(x, spec, window, reinterpret_cast<const float2*>(twiddle), T, n_fft, log2n, hop, frames, fpb, Fp, norm);
  else   // x = d audio (accumulated, fp32 [B,2,T]), spec = d spec
    disc_stft_bwd_kernel<<<grid, warps * 32, smem, s>>>(spec, const_cast<float*>(x), window, reinterpret_cast<const float2*>(twiddle), T, n_fft, log2n, hop,
                                                       frames, fpb, Fp, norm);
  B200SAT_CHECK_CUDA(cudaGetLastError());
  return B200SAT_OK;
}

extern "C" int b200sat_disc_conv0(const float* spec, const float* w, const float* bias, void* out, const void* dpre, float* dspec, int B, int frames,
                                  int F, float leaky, void* stream) {
  if (!w || B <= 0 || frames <= 0 || F <= 0 || (!(spec && out) && !(dpre && dspec))) { set_last_error("disc_conv0: bad arguments"); return B200SAT_EINVAL; }
  const int Fp = F + 8;
  const long total = static_cast<long>(B) * frames * Fp;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (spec && out) disc_conv0_fwd_kernel<<<grid_for(total, 128, 8), 128, 0, s>>>(spec, w, bias, static_cast<__nv_bfloat16*>(out), B, frames, Fp, F, leaky);
  if (dpre && dspec) disc_conv0_dgrad_kernel<<<grid_for(total, 128, 8), 128, 0, s>>>(static_cast<const __nv_bfloat16*>(dpre), w, dspec, B, frames, Fp, F);
  B200SAT_CHECK_CUDA(cudaGetLastError());
  return B200SAT_OK;
}

static bool disc_post_v1() {
  static const bool v = [] { const char* e = getenv("B200SAT_DISC_POST_V1"); return e && atoi(e) != 0; }();
  return v;
}

extern "C" int b200sat_disc_convpost(const void* act, const float* w, const float* bias, float* logits, int B, int frames, int F, void* stream) {
  if (!act || !w || !logits || B <= 0 || frames <= 0 || F <= 0) { set_last_error("disc_convpost: bad arguments"); return B200SAT_EINVAL; }
  const int Fp = F + 8;
  const long total = static_cast<long>(B) * frames * Fp;
  if (disc_post_v1() || total >= (1L << 27))
    disc_convpost_fwd_v1_kernel<<<grid_for(total, 128, 8), 128, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(act), w, bias, logits,
                                                                                                       B, frames, Fp, F);
  else
    disc_convpost_fwd_kernel<<<grid_for(total * 8, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(act), w, bias, logits,
                                                                                                        B, frames, Fp, F);
  B200SAT_CHECK_CUDA(cudaGetLastError());
  return B200SAT_OK;
}

/* spec fp32 [B, P, 4] <-> S9 bf16 [B, P, 64] (frequency taps of the first conv folded into channels; see disc_spec_pack_kernel):
 * backward == 0: s9 = pack(spec);  backward != 0: spec (= d spec) = pack^T(s9 (= d S9)). */
extern "C" int b200sat_disc_spec_pack(float* spec, void* s9, int B, int frames, int F, int backward, void* stream) {
  if (!spec || !s9 || B <= 0 || frames <= 0 || F <= 0) { set_last_error("disc_spec_pack: bad arguments"); return B200SAT_EINVAL; }
  const int Fp = F + 8;
  const long P = static_cast<long>(frames) * Fp;
  const long total = static_cast<long>(B) * P;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!backward) {
    if (total >= (1L << 27)) { set_last_error("disc_spec_pack: more than 2^27 positions"); return B200SAT_EUNSUPPORTED; }
    disc_spec_pack_kernel<<<grid_for(total * 8, 256, 8), 256, 0, s>>>(spec, static_cast<__nv_bfloat16*>(s9), static_cast<int>(total), static_cast<int>(P), Fp, F);
  } else {
    dim3 grid(static_cast<unsigned>((P + 127) / 128), B);
    disc_spec_unpack_kernel<<<grid, 128, 0, s>>>(static_cast<const __nv_bfloat16*>(s9), spec, P, Fp, F);
  }
  B200SAT_CHECK_CUDA(cudaGetLastError());
  return B200SAT_OK;
}

extern "C" int b200sat_disc_hinge_sums(const float* lt, const float* lf, double* sums, int B, int frames, int F, void* stream) {
  if ((!lt && !lf) || !sums || B <= 0 || frames <= 0 || F <= 0) { set_last_error("disc_hinge_sums: bad arguments"); return B200SAT_EINVAL; }
  const long total = static_cast<long>(B) * frames * (F + 8);
  disc_hinge_sums_kernel<<<grid_for(total, 256, 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(lt, lf, sums, total, F + 8, F);
  B200SAT_CHECK_CUDA(cudaGetLastError());
  return B200SAT_OK;
}

extern "C" int b200sat_disc_l1_sum(const void* a, const void* b, double* out, long n, void* stream) {
  if (!a || !b || !out || n <= 0 || (n & 7)) { set_last_error("disc_l1_sum: bad arguments (n % 8 == 0)"); return B200SAT_EINVAL; }
  disc_l1_sum_kernel<<<grid_for(n / 8, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(a),
                                                                                             static_cast<const __nv_bfloat16*>(b), out, n / 8);
  B200SAT_CHECK_CUDA(cudaGetLastError());
  return B200SAT_OK;
}

extern "C" int b200sat_disc_logit_grad(const float* logits, float* g, int B, int frames, int F, int mode, float scale, void* stream) {
  if (!logits || !g || B <= 0 || frames <= 0 || F <= 0 || mode < 0 || mode > 2) { set_last_error("disc_logit_grad: bad arguments"); return B200SAT_EINVAL; }
  const long total = static_cast<long>(B) * frames * (F + 8);
  disc_logit_grad_kernel<<<grid_for(total, 256, 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(logits, g, total, F + 8, F, mode, scale);
  B200SAT_CHECK_CUDA(cudaGetLastError());
  return B200SAT_OK;
}

extern "C" int b200sat_disc_act_bwd(const void* d_in, const float* d_logit, const float* w_post, const void* post, const void* other, float fm_coef,
                                    float leaky, void* d_pre, int B, int frames, int F, void* stream) {
  if (!post || !d_pre || (!d_in && !d_logit && !other) || (d_logit && !w_post) || B <= 0 || frames <= 0 || F <= 0) {
    set_last_error("disc_act_bwd: bad arguments"); return B200SAT_EINVAL;
  }
  const long total = static_cast<long>(B) * frames * (F + 8) * 8;
  if (total >= (1L << 30)) { set_last_error("disc_act_bwd: more than 2^27 positions"); return B200SAT_EUNSUPPORTED; }
  disc_act_bwd_kernel<<<grid_for(total, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(d_in), d_logit, w_post, static_cast<const __nv_bfloat16*>(post), static_cast<const __nv_bfloat16*>(other), fm_coef,
      leaky, static_cast<__nv_bfloat16*>(d_pre), B, frames, F + 8, F);
  B200SAT_CHECK_CUDA(cudaGetLastError());
  return B200SAT_OK;
}

extern "C" int b200sat_disc_conv0_wgrad(const void* dpre, const float* spec, float* dW, int B, int frames, int F, void* stream) {
  if (!dpre || !spec || !dW || B <= 0 || frames <= 0 || F <= 0) { set_last_error("disc_conv0_wgrad: bad arguments"); return B200SAT_EINVAL; }
  const long total = static_cast<long>(B) * frames * (F + 8);
  const long cap = static_cast<long>(num_sms()) * 8;
  disc_conv0_wgrad_kernel<<<static_cast<int>((total + 7) / 8 < cap ? (total + 7) / 8 : cap), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(dpre), spec, dW, B, frames, F + 8, F);
  B200SAT_CHECK_CUDA(cudaGetLastError());
  return B200SAT_OK;
}

extern "C" int b200sat_disc_convpost_wgrad(const float* g, const void* act, float* dW, float* dbias, int B, int frames, int F, void* stream) {
  if (!g || !act || !dW || B <= 0 || frames <= 0 || F <= 0) { set_last_error("disc_convpost_wgrad: bad arguments"); return B200SAT_EINVAL; }
  const long total = static_cast<long>(B) * frames * (F + 8);
  if (disc_post_v1() || total >= (1L << 27)) {
    const long cap = static_cast<long>(num_sms()) * 8;
    disc_convpost_wgrad_v1_kernel<<<static_cast<int>((total + 7) / 8 < cap ? (total + 7) / 8 : cap), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        g, static_cast<const __nv_bfloat16*>(act), dW, dbias, B, frames, F + 8, F);
  } else {
    disc_convpost_wgrad_kernel<<<grid_for(total * 8, 256, 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(g, static_cast<const __nv_bfloat16*>(act), dW, dbias,
                                                                                                          B, frames, F + 8, F);
  }
  B200SAT_CHECK_CUDA(cudaGetLastError());
  return B200SAT_OK;
}

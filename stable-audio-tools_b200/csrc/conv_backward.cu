// b200sat — backward pass of the Oobleck conv stack: the streaming / reduction / packing kernels that sit between the tensor-core
// launches.  The heavy parts reuse the forward machinery:
//   data gradient   = b200sat_conv1d_fwd on the output gradient with re-packed weights (a conv's dgrad is a conv with flipped taps,
//                     a strided conv's dgrad is a transposed conv and vice versa)                      -> wn_pack_dgrad below
//   weight gradient = b200sat_conv_wgrad (gemm.cu), one tcgen05 launch per tap reading both activation planes in place
// This file: SnakeBeta backward fused with the skip-connection add and the bias / alpha / beta reductions, weight-norm backward,
// the 2-channel edge layers' weight gradients, and the VAE bottleneck backward.
// Reference: autograd of models/autoencoders.py:23-27, :58-83, :233-283, :285-362, models/blocks.py:291-329, models/bottleneck.py:105-134
// as run by AutoencoderTrainingWrapper.training_step (training/autoencoders.py:367-527).
#include "common.cuh"

namespace b200sat {

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    w[i] = *reinterpret_cast<const uint32_t*>(&h);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// y = x + invb * sin^2(a x)   (a = exp(alpha), invb = 1/(exp(beta)+1e-9)):
//   d_raw   = d_skip + d_act * (1 + invb * a * sin(2 a x))
//   d_alpha = sum d_act * invb * sin(2 a x) * a * x                  (alpha is log-scale: da/dalpha = a)
//   d_beta  = -sum d_act * sin^2(a x) * invb^2 * exp(beta)
//   d_bias  = sum d_raw                                             (bias of the conv that produced x)
// One pass over HBM: reads d_act, x (and d_skip), writes d_raw; VEC channels per thread (4 keeps the register count low enough
// for 4 blocks per SM), two rows in flight per thread, rows strided over the grid.
template <int VEC>
__device__ __forceinline__ void load_vec(const __nv_bfloat16* p, float* f) {
  if (VEC == 8) {
    unpack8(*reinterpret_cast<const uint4*>(p), f);
  } else {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    f[0] = __uint_as_float(u.x << 16); f[1] = __uint_as_float(u.x & 0xffff0000u);
    f[2] = __uint_as_float(u.y << 16); f[3] = __uint_as_float(u.y & 0xffff0000u);
  }
}
template <int VEC>
__device__ __forceinline__ void store_vec(__nv_bfloat16* p, const float* f) {
  if (VEC == 8) {
    *reinterpret_cast<uint4*>(p) = pack8(f);
  } else {
    const __nv_bfloat162 h0 = __floats2bfloat162_rn(f[0], f[1]), h1 = __floats2bfloat162_rn(f[2], f[3]);
    *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
  }
}

template <int VEC>
__global__ void __launch_bounds__(256, (VEC == 4) ? 4 : 2) snake_bwd_kernel(const __nv_bfloat16* __restrict__ d_act, const __nv_bfloat16* __restrict__ x_raw,
                                                        const __nv_bfloat16* __restrict__ d_skip, const float* __restrict__ sa,
                                                        const float* __restrict__ sib, __nv_bfloat16* __restrict__ d_raw,
                                                        float* __restrict__ dalpha, float* __restrict__ dbeta, float* __restrict__ dbias,
                                                        long rows, int C) {
  __shared__ float red[3 * VEC][257];
  const int gpr = C / VEC;        // threads per row
  const int rpi = 256 / gpr;      // rows per block iteration
  const int cg = threadIdx.x % gpr, rs = threadIdx.x / gpr;
  float a2[VEC], iba[VEC], ga[VEC], gc[VEC], gbias[VEC];
  float gg[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    const float a = sa[cg * VEC + j];
    a2[j] = 2.f * a;
    iba[j] = sib[cg * VEC + j] * a;
    ga[j] = gc[j] = gbias[j] = gg[j] = 0.f;
  }
  const long stride = static_cast<long>(gridDim.x) * rpi;
  for (long r = static_cast<long>(blockIdx.x) * rpi + rs; r < rows; r += 2 * stride) {
    const long r1 = r + stride;
    const bool two = r1 < rows;
    const size_t off0 = static_cast<size_t>(r) * C + cg * VEC, off1 = static_cast<size_t>(two ? r1 : r) * C + cg * VEC;
    float g0[VEC], x0[VEC], k0[VEC], g1[VEC], x1[VEC], k1[VEC];
    load_vec<VEC>(d_act + off0, g0); load_vec<VEC>(x_raw + off0, x0);
    load_vec<VEC>(d_act + off1, g1); load_vec<VEC>(x_raw + off1, x1);
    if (d_skip) { load_vec<VEC>(d_skip + off0, k0); load_vec<VEC>(d_skip + off1, k1); }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (h == 1 && !two) break;
      float* g = h ? g1 : g0; float* x = h ? x1 : x0; float* sk = h ? k1 : k0;
      float o[VEC];
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        float s2, c2;
        fast_sincos(a2[j] * x[j], &s2, &c2);            // sin(2ax), cos(2ax); sin^2(ax) = (1 - cos 2ax) / 2
        const float t = g[j] * iba[j] * s2;
        const float dr = g[j] + t + (d_skip ? sk[j] : 0.f);
        ga[j] = fmaf(t, x[j], ga[j]);
        gg[j] += g[j];
        gc[j] = fmaf(g[j], c2, gc[j]);
        gbias[j] += dr;
        o[j] = dr;
      }
      store_vec<VEC>(d_raw + (h ? off1 : off0), o);
    }
  }
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    red[j][threadIdx.x] = ga[j];
    red[VEC + j][threadIdx.x] = 0.5f * (gg[j] - gc[j]);   // sum d_act * sin^2(a x)
    red[2 * VEC + j][threadIdx.x] = gbias[j];
  }
  __syncthreads();
  // thread (v, cg) sums over the rpi row slots
  for (int i = threadIdx.x; i < 3 * VEC * gpr; i += 256) {
    const int v = i / gpr, g2 = i % gpr;
    float s = 0.f;
    for (int r2 = 0; r2 < rpi; ++r2) s += red[v][r2 * gpr + g2];
    const int kind = v / VEC, ch = g2 * VEC + (v % VEC);
    if (kind == 0) atomicAdd(dalpha + ch, s);
    else if (kind == 1) { const float ibv = sib[ch]; atomicAdd(dbeta + ch, -s * ibv * ibv * (1.0f / ibv - 1e-9f)); }
    else if (dbias) atomicAdd(dbias + ch, s);
  }
}

// Weights for the data-gradient convolution, weight-normalised and packed for conv1d.cu (bf16):
//   mode 0 (conv, stride 1):   P[ci][k'*Cout + co]         = w[co,ci,K-1-k']       -> run as mode 0 with pad' = dil*(K-1) - pad
//   mode 1 (strided conv):     P[ph*Cin + ci][j*Cout + co] = w[co,ci,ph + s*j]     -> run as mode 2 (transposed) Cout -> Cin
//   mode 2 (transposed conv):  P[ci][k*Cout + co]          = wt[ci,co,k]           -> run as mode 1 (strided)    Cout -> Cin
// v fp32 [Cout,Cin,K] (modes 0/1) or [Cin,Cout,K] (mode 2); g / inv_norm per dim-0 row (null g = plain weights).
__global__ void wn_pack_dgrad_kernel(const float* __restrict__ v, const float* __restrict__ g, const float* __restrict__ inv_norm,
                                     __nv_bfloat16* __restrict__ out, int Cout, int Cin, int K, int mode, int stride) {
  const long n = static_cast<long>(Cout) * Cin * K;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * blockDim.x) {
    int co, ci, k;
    if (mode == 1) {
      const long row = i / (2 * Cout);
      const int rem = i % (2 * Cout);
      const int ph = row / Cin;
      ci = row % Cin;
      co = rem % Cout;
      k = ph + stride * (rem / Cout);
    } else {
      ci = i / (static_cast<long>(K) * Cout);
      const int rem = i % (static_cast<long>(K) * Cout);
      co = rem % Cout;
      k = rem / Cout;
      if (mode == 0) k = K - 1 - k;
    }
    float w;
    if (mode == 2) { w = v[(static_cast<long>(ci) * Cout + co) * K + k]; if (g) w *= g[ci] * inv_norm[ci]; }
    else { w = v[(static_cast<long>(co) * Cin + ci) * K + k]; if (g) w *= g[co] * inv_norm[co]; }
    out[i] = __float2bfloat16_rn(w);
  }
}

// weight_norm backward (w = g v / ||v|| over all dims but 0): with dot = <dw_r, v_r>, n = ||v_r||:
//   dg_r = dot / n;   dv_r = (g/n) (dw_r - v_r dot / n^2).
// dwp is the tap-major scratch the wgrad launches fill, [K][R][Cc]; v, dv are [R][Cc][K].  Null g: plain weights (dv = dw).
__global__ void __launch_bounds__(256) wn_bwd_kernel(const float* __restrict__ dwp, const float* __restrict__ v, const float* __restrict__ g,
                                                     const float* __restrict__ inv_norm, float* __restrict__ dv, float* __restrict__ dg,
                                                     int R, int Cc, int K) {
  __shared__ float red[8];
  __shared__ float dot_s;
  const int r = blockIdx.x;
  const long inner = static_cast<long>(Cc) * K;
  const float* vr = v + r * inner;
  float dot = 0.f;
  if (g) {
    for (long e = threadIdx.x; e < inner; e += 256) {
      const int c = e / K, k = e % K;
      dot += dwp[(static_cast<long>(k) * R + r) * Cc + c] * vr[e];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dot;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int i = 0; i < 8; ++i) s += red[i];
      dot_s = s;
    }
    __syncthreads();
    dot = dot_s;
  }
  const float inv = g ? inv_norm[r] : 1.f;
  const float scale = g ? g[r] * inv : 1.f;
  const float proj = g ? dot * inv * inv : 0.f;
  for (long e = threadIdx.x; e < inner; e += 256) {
    const int c = e / K, k = e % K;
    const float dw = dwp[(static_cast<long>(k) * R + r) * Cc + c];
    dv[r * inner + e] = scale * (dw - vr[e] * proj);
  }
  if (g && dg && threadIdx.x == 0) dg[r] = dot * inv;
}

// Weight gradient of the two audio-channel edge layers (SIMT; the tensor cores would idle on A <= 8 channels):
//   dW[c*stride_c + a*stride_a + k] += sum_{b,t} plane[b,t,c] * sig[b,a,t + sign*(k - pad)]
// encoder conv_in  (autoencoders.py:303):     plane = d(conv output) [B,T,C], sig = audio,      sign +1, W [C,A,K]
// decoder conv_out (autoencoders.py:355-357): plane = layer input    [B,T,C], sig = d(audio),   sign -1, W [A,C,K]
// Persistent blocks keep their partial sums in registers over all their time tiles and issue one atomic per weight at the end.
constexpr int EW_TT = 64;
constexpr int EW_MAXJ = 16;
__global__ void __launch_bounds__(256) edge_wgrad_kernel(const __nv_bfloat16* __restrict__ plane, const float* __restrict__ sig,
                                                         float* __restrict__ dW, int B, int T, int C, int A, int K, int pad, int sign,
                                                         long stride_c, long stride_a) {
  extern __shared__ float sm[];
  const int span = EW_TT + 2 * (K - 1);
  float* ssig = sm;                                                       // [A][span]
  __nv_bfloat16* spl = reinterpret_cast<__nv_bfloat16*>(sm + A * span);   // [EW_TT][C]
  const int G = 256 / C;
  const int c = threadIdx.x % C, grp = threadIdx.x / C;
  const int AK = A * K;
  float acc[EW_MAXJ];
  int base[EW_MAXJ];
#pragma unroll
  for (int i = 0; i < EW_MAXJ; ++i) {
    acc[i] = 0.f;
    const int j = grp + G * i;
    const int a = (j < AK) ? j / K : 0, k = (j < AK) ? j % K : pad;
    base[i] = a * span + (K - 1) + sign * (k - pad);
  }
  const int tiles_per_item = (T + EW_TT - 1) / EW_TT;
  const int tiles = B * tiles_per_item;
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int b = tile / tiles_per_item, t0 = (tile % tiles_per_item) * EW_TT;
    __syncthreads();
    for (int i = threadIdx.x; i < A * span; i += 256) {
      const int a = i / span, t = t0 + (i % span) - (K - 1);
      ssig[i] = (t >= 0 && t < T) ? sig[(static_cast<long>(b) * A + a) * T + t] : 0.f;
    }
    for (int i = threadIdx.x; i < EW_TT * C / 8; i += 256) {
      const int tt = (i * 8) / C, cc = (i * 8) % C;
      uint4 u = make_uint4(0, 0, 0, 0);
      if (t0 + tt < T) u = *reinterpret_cast<const uint4*>(plane + (static_cast<size_t>(b) * T + t0 + tt) * C + cc);
      *reinterpret_cast<uint4*>(spl + tt * C + cc) = u;
    }
    __syncthreads();
#pragma unroll 4
    for (int tt = 0; tt < EW_TT; ++tt) {
      const float pv = __bfloat162float(spl[tt * C + c]);
#pragma unroll
      for (int i = 0; i < EW_MAXJ; ++i) acc[i] = fmaf(pv, ssig[base[i] + tt], acc[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < EW_MAXJ; ++i) {
    const int j = grp + G * i;
    if (j < AK) atomicAdd(dW + c * stride_c + (j / K) * stride_a + (j % K), acc[i]);
  }
}

// VAE bottleneck backward (bottleneck.py:105-113): z = noise * sigma + mean, sigma = softplus(s) + 1e-4,
// kl = mean_{b,t} sum_c (mean^2 + sigma^2 - log sigma^2 - 1).  With g = dL/dz and kls = dL/dkl / (B*T):
//   d_mean = g + kls * 2 mean;   d_s = (g * noise + kls * (2 sigma - 2 / sigma)) * sigmoid(s)
// dz: bf16 plane [B,T,L] (the decoder's first conv dgrad), ms: forward planes [B,T,2L]; d_ms: bf16 plane [B,T,2L].
__global__ void vae_sample_bwd_kernel(const __nv_bfloat16* __restrict__ dz, const __nv_bfloat16* __restrict__ ms,
                                      const float* __restrict__ noise, const float* __restrict__ kl_grad, float kl_scale,
                                      __nv_bfloat16* __restrict__ d_ms, int B, int L, int T) {
  const long n = static_cast<long>(B) * T * L;
  const float kls = kl_scale * (kl_grad ? *kl_grad : 1.f);
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int c = i % L;
    const long bt = i / L;
    const int t = bt % T;
    const int b = bt / T;
    const size_t off = static_cast<size_t>(bt) * 2 * L;
    const float mean = __bfloat162float(ms[off + c]);
    const float sc = __bfloat162float(ms[off + L + c]);
    const float sp = (sc > 20.f) ? sc : log1pf(expf(sc));
    const float sigma = sp + 1e-4f;
    const float dsp = (sc > 20.f) ? 1.f : 1.f / (1.f + expf(-sc));
    const float g = dz ? __bfloat162float(dz[i]) : 0.f;
    const float nz = noise ? noise[(static_cast<long>(b) * L + c) * T + t] : 0.f;
    d_ms[off + c] = __float2bfloat16_rn(g + kls * 2.f * mean);
    d_ms[off + L + c] = __float2bfloat16_rn((g * nz + kls * (2.f * sigma - 2.f / sigma)) * dsp);
  }
}

}  // namespace b200sat

using namespace b200sat;

extern "C" int b200sat_snake_bwd(const void* d_act, const void* x_raw, const void* d_skip, const float* snake_a, const float* snake_invb,
                                 void* d_raw, float* dalpha, float* dbeta, float* dbias, long rows, int C, void* stream) {
  if (!d_act || !x_raw || !snake_a || !snake_invb || !d_raw || !dalpha || !dbeta || rows <= 0) { set_last_error("snake_bwd: bad arguments"); return B200SAT_EINVAL; }
  if (C < 64 || C > 2048 || (C & (C - 1))) { set_last_error("snake_bwd: C must be a power of two in [64, 2048]"); return B200SAT_EUNSUPPORTED; }
  const int vec = (C <= 1024) ? 4 : 8;
  const int rpi = 256 / (C / vec);
  long blocks = (rows + 2 * rpi - 1) / (2 * rpi);
  const long cap = static_cast<long>(num_sms()) * (vec == 4 ? 4 : 2);
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  auto* da = static_cast<const __nv_bfloat16*>(d_act); auto* xr = static_cast<const __nv_bfloat16*>(x_raw);
  auto* ds = static_cast<const __nv_bfloat16*>(d_skip); auto* dr = static_cast<__nv_bfloat16*>(d_raw);
  if (vec == 4) snake_bwd_kernel<4><<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(da, xr, ds, snake_a, snake_invb, dr, dalpha, dbeta, dbias, rows, C);
  else snake_bwd_kernel<8><<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(da, xr, ds, snake_a, snake_invb, dr, dalpha, dbeta, dbias, rows, C);
  B200SAT_CHECK_CUDA(cudaGetLastError());
  return B200SAT_OK;
}

extern "C" int b200sat_wn_pack_dgrad(const float* v, const float* g, const float* inv_norm, void* out, int Cout, int Cin, int K, int mode,
                                     int stride, void* stream) {
  if (!v || !out || Cout <= 0 || Cin <= 0 || K <= 0 || mode < 0 || mode > 2) { set_last_error("wn_pack_dgrad: bad arguments"); return B200SAT_EINVAL; }
  if (g && !inv_norm) { set_last_error("wn_pack_dgrad: weight-norm needs the 1/||v|| rows of the forward pack"); return B200SAT_EINVAL; }
  if (mode != 0 && K != 2 * stride) { set_last_error("wn_pack_dgrad: strided / transposed convs need K == 2*stride"); return B200SAT_EUNSUPPORTED; }
  const long n = static_cast<long>(Cout) * Cin * K;
  const int grid = static_cast<int>((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096);
  wn_pack_dgrad_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(v, g, inv_norm, static_cast<__nv_bfloat16*>(out), Cout, Cin, K,
                                                                             mode, stride);
  B200SAT_CHECK_CUDA(cudaGetLastError());
  return B200SAT_OK;
}

extern "C" int b200sat_wn_bwd(const float* dw_taps, const float* v, const float* g, const float* inv_norm, float* dv, float* dg, int R, int Cc,
                              int K, void* stream) {
  if (!dw_taps || !v || !dv || R <= 0 || Cc <= 0 || K <= 0 || (g && (!inv_norm || !dg))) { set_last_error("wn_bwd: bad arguments"); return B200SAT_EINVAL; }
  wn_bwd_kernel<<<R, 256, 0, static_cast<cudaStream_t>(stream)>>>(dw_taps, v, g, inv_norm, dv, dg, R, Cc, K);
  B200SAT_CHECK_CUDA(cudaGetLastError());
  return B200SAT_OK;
}

extern "C" int b200sat_edge_wgrad(const void* plane, const float* sig, float* dW, int B, int T, int C, int A, int K, int pad, int sign,
                                  long stride_c, long stride_a, void* stream) {
  if (!plane || !sig || !dW || B <= 0 || T <= 0 || A <= 0 || K <= 0 || (sign != 1 && sign != -1)) { set_last_error("edge_wgrad: bad arguments"); return B200SAT_EINVAL; }
  if (C < 8 || C > 256 || 256 % C) { set_last_error("edge_wgrad: C must divide 256"); return B200SAT_EUNSUPPORTED; }
  const int G = 256 / C;
  if ((A * K + G - 1) / G > EW_MAXJ) { set_last_error("edge_wgrad: too many (channel, tap) pairs per thread"); return B200SAT_EUNSUPPORTED; }
  const int span = EW_TT + 2 * (K - 1);
  const int smem = A * span * 4 + EW_TT * C * 2;
  if (smem > 48 * 1024) { set_last_error("edge_wgrad: tile does not fit in 48 KB of shared memory"); return B200SAT_EUNSUPPORTED; }
  if ((A * span * 4) % 16) { set_last_error("edge_wgrad: unsupported tap count (tile alignment)"); return B200SAT_EUNSUPPORTED; }
  const int tiles = B * ((T + EW_TT - 1) / EW_TT);
  const int cap = num_sms() * 4;
  edge_wgrad_kernel<<<tiles < cap ? tiles : cap, 256, smem, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(plane), sig, dW, B, T, C, A, K, pad, sign, stride_c, stride_a);
  B200SAT_CHECK_CUDA(cudaGetLastError());
  return B200SAT_OK;
}

extern "C" int b200sat_vae_sample_bwd(const void* dz, const void* ms, const float* noise, const float* kl_grad, float kl_scale, void* d_ms,
                                      int B, int L, int T, void* stream) {
  if (!ms || !d_ms || B <= 0 || L <= 0 || T <= 0) { set_last_error("vae_sample_bwd: bad arguments"); return B200SAT_EINVAL; }
  const long n = static_cast<long>(B) * T * L;
  const int grid = static_cast<int>((n + 255) / 256 < 2048 ? (n + 255) / 256 : 2048);
  vae_sample_bwd_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(dz),
                                                                             static_cast<const __nv_bfloat16*>(ms), noise, kl_grad, kl_scale,
                                                                             static_cast<__nv_bfloat16*>(d_ms), B, L, T);
  B200SAT_CHECK_CUDA(cudaGetLastError());
  return B200SAT_OK;
}

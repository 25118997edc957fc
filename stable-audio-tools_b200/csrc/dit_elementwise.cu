// b200sat — HBM-bound DiT kernels: LayerNorm (+adaLN modulate), small-M linears for the conditioning MLPs,
// Fourier timestep features, the pre/post 1x1 convs with their transposes, classifier-free-guidance combine.
// bf16 rounding points mirror the reference's bf16 eager path so that bf16-vs-bf16 parity is tight.
#include "common.cuh"
#include <cstring>

namespace b200sat {

// ---------------------------------------------------------------------------------------------------------
// LayerNorm over the last dim (transformer.py:236-238: F.layer_norm(x, gamma, beta=0, eps)), optional adaLN
// modulate h*(1+scale)+shift (transformer.py:680-682, :695-697).  One warp per row, values kept in registers.
template <int MAXC>
__global__ void __launch_bounds__(256, (MAXC <= 8) ? 4 : 1) layernorm_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, const float* __restrict__ scale,
                                                        const float* __restrict__ shift, __nv_bfloat16* __restrict__ y,
                                                        int rows, int D, long ldx, long ldy, int rows_per_batch, long ld_mod,
                                                        float eps) {
  griddep_launch();
  griddep_wait();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const __nv_bfloat16* xr = x + static_cast<long>(row) * ldx;
  const int nchunk = D / 8;  // 16-byte chunks
  // the row stays packed (bf16x2) in registers and is unpacked on the fly in each pass: 4 words per chunk instead of 8 floats keeps
  // the kernel at <= 64 registers (4 blocks / SM); the fp32-resident version ran 2 blocks / SM at 31 % of the HBM roofline
  uint4 xp[MAXC];
  float sum = 0.f;
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    const int ch = lane + c * 32;
    if (ch < nchunk) {
      xp[c] = __ldg(reinterpret_cast<const uint4*>(xr) + ch);
      const uint32_t w[4] = {xp[c].x, xp[c].y, xp[c].z, xp[c].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) { const float2 f = unpack_bf16(w[j]); sum += f.x + f.y; }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / D;
  float sq = 0.f;
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    if (lane + c * 32 < nchunk) {
      const uint32_t w[4] = {xp[c].x, xp[c].y, xp[c].z, xp[c].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) { const float2 f = unpack_bf16(w[j]); const float d0 = f.x - mean, d1 = f.y - mean; sq += d0 * d0 + d1 * d1; }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq / D + eps);
  __nv_bfloat16* yr = y + static_cast<long>(row) * ldy;
  const float* sc = scale ? scale + static_cast<long>(row / rows_per_batch) * ld_mod : nullptr;
  const float* sh = shift ? shift + static_cast<long>(row / rows_per_batch) * ld_mod : nullptr;
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    const int ch = lane + c * 32;
    if (ch < nchunk) {
      float o[8];
      // parameters as two 16-byte loads per 8 columns (the scalar version issued 8 LDGs per parameter per chunk)
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma) + ch * 2), g1 = __ldg(reinterpret_cast<const float4*>(gamma) + ch * 2 + 1);
      const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const uint32_t xw[4] = {xp[c].x, xp[c].y, xp[c].z, xp[c].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16(xw[j]);
        o[2 * j] = (f.x - mean) * rstd * gm[2 * j];
        o[2 * j + 1] = (f.y - mean) * rstd * gm[2 * j + 1];
      }
      if (beta) {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta) + ch * 2), b1 = __ldg(reinterpret_cast<const float4*>(beta) + ch * 2 + 1);
        const float bt[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += bt[j];
      }
      if (sc) {
        const float4 s0 = __ldg(reinterpret_cast<const float4*>(sc) + ch * 2), s1 = __ldg(reinterpret_cast<const float4*>(sc) + ch * 2 + 1);
        const float4 h0 = __ldg(reinterpret_cast<const float4*>(sh) + ch * 2), h1 = __ldg(reinterpret_cast<const float4*>(sh) + ch * 2 + 1);
        const float sv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w}, hv[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float t = bf16_round(o[j]);
          t = bf16_round(t * bf16_round(1.0f + sv[j]));
          o[j] = t + hv[j];
        }
      }
      uint4 u;
      u.x = pack_bf16(o[0], o[1]); u.y = pack_bf16(o[2], o[3]); u.z = pack_bf16(o[4], o[5]); u.w = pack_bf16(o[6], o[7]);
      reinterpret_cast<uint4*>(yr)[ch] = u;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Small-M linear (M <= 8): y[m,n] = act(x[m,:] . w[n,:] + b[n]) (+ add[m,n]).  One warp per output column.
// Used for the timestep / global conditioning MLPs (dit.py:41-76, :140-168) and the adaLN embedder
// (transformer.py:767-773), all of which have M = batch.
__global__ void __launch_bounds__(256) small_linear_kernel(const __nv_bfloat16* __restrict__ x, long ldx,
                                                           const __nv_bfloat16* __restrict__ w, long ldw,
                                                           const float* __restrict__ bias, const __nv_bfloat16* __restrict__ add,
                                                           long ldadd, void* __restrict__ y, long ldy, int M, int N, int K,
                                                           int act_silu, int out_f32, int act_sigmoid_1m, float* __restrict__ stats,
                                                           long stats_stride) {
  griddep_launch();
  griddep_wait();
  const int n = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (n >= N) return;
  const int lane = threadIdx.x & 31;
  float acc[8];
#pragma unroll
  for (int m = 0; m < 8; ++m) acc[m] = 0.f;
  const __nv_bfloat16* wr = w + static_cast<long>(n) * ldw;
  for (int k0 = lane * 8; k0 < K; k0 += 256) {
    const uint4 wu = __ldg(reinterpret_cast<const uint4*>(wr + k0));
    const uint32_t ww[4] = {wu.x, wu.y, wu.z, wu.w};
    float wf[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float2 f = unpack_bf16(ww[j]); wf[2 * j] = f.x; wf[2 * j + 1] = f.y; }
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      if (m < M) {
        const uint4 xu = __ldg(reinterpret_cast<const uint4*>(x + m * ldx + k0));
        const uint32_t xw[4] = {xu.x, xu.y, xu.z, xu.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float2 f = unpack_bf16(xw[j]); acc[m] += f.x * wf[2 * j] + f.y * wf[2 * j + 1]; }
      }
    }
  }
#pragma unroll
  for (int m = 0; m < 8; ++m) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[m] += __shfl_xor_sync(0xffffffffu, acc[m], o);
  }
  if (lane == 0) {
    for (int m = 0; m < M; ++m) {
      float t = acc[m] + (bias ? bias[n] : 0.f);
      t = bf16_round(t);
      if (act_silu) t = bf16_round(silu_f(t));
      if (add) t = bf16_round(t + __bfloat162float(add[m * ldadd + n]));
      if (act_sigmoid_1m) t = 1.0f / (1.0f + __expf(-(bf16_round(1.0f - t))));  // sigmoid(1 - g), transformer.py:684
      if (out_f32) reinterpret_cast<float*>(y)[m * ldy + n] = t;
      else reinterpret_cast<__nv_bfloat16*>(y)[m * ldy + n] = __float2bfloat16_rn(t);
      if (stats) {  // (sum, sum of squares) of the bf16 output row, for a LayerNorm folded into the consumer GEMM
        const float r_ = bf16_round(t);
        atomicAdd(stats + m * stats_stride, r_);
        atomicAdd(stats + m * stats_stride + 1, r_ * r_);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Fourier timestep features (blocks.py:85-94): f = 2*pi*t*w ; out = [cos f | sin f], bf16 rounding as the eager path.
__global__ void fourier_features_kernel(const float* __restrict__ t, const __nv_bfloat16* __restrict__ w,
                                        __nv_bfloat16* __restrict__ out, int B, int half, const int* __restrict__ step,
                                        int t_stride) {
  griddep_launch();
  griddep_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * half) return;
  const int b = i / half, j = i % half;
  const float tv = t[(step ? *step * t_stride : 0) + b];
  const float a = bf16_round(6.283185307179586f * bf16_round(tv));
  const float f = bf16_round(a * __bfloat162float(w[j]));
  out[b * 2 * half + j] = __float2bfloat16_rn(cosf(f));
  out[b * 2 * half + half + j] = __float2bfloat16_rn(sinf(f));
}

// ---------------------------------------------------------------------------------------------------------
// DiT input: xb = bf16(x * c_in); y = bf16(conv1x1(xb)) + xb; transpose [B,C,T] -> rows [(rep*B+b)*T + t, C] bf16
// (dit.py:193-195 preprocess_conv + residual + 'b c t -> b t c'; the CFG batch duplication of dit.py:330-331 is `reps`).
template <int C>
__global__ void __launch_bounds__(256) dit_pre_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ wconv,
                                                      __nv_bfloat16* __restrict__ out, int B, int T, int reps,
                                                      const float* __restrict__ cin_table, const int* __restrict__ step) {
  __shared__ float sx[C][65];
  __shared__ float sw[C][C + 1];
  griddep_launch();
  griddep_wait();
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * 64;
  const float c_in = cin_table ? cin_table[step ? *step : 0] : 1.0f;
  for (int i = threadIdx.x; i < C * C; i += 256) sw[i / C][i % C] = __bfloat162float(wconv[i]);
  for (int i = threadIdx.x; i < C * 64; i += 256) {
    const int c = i / 64, tt = i % 64;
    const int t = t0 + tt;
    sx[c][tt] = (t < T) ? bf16_round(x[(static_cast<long>(b) * C + c) * T + t] * c_in) : 0.f;
  }
  __syncthreads();
  // thread -> (tt, 16 output channels)
  const int tt = threadIdx.x & 63;
  const int cg = threadIdx.x >> 6;  // 0..3
  const int t = t0 + tt;
  if (t >= T) return;
  float o[C / 4];
#pragma unroll
  for (int i = 0; i < C / 4; ++i) {
    const int co = cg * (C / 4) + i;
    float acc = 0.f;
#pragma unroll 8
    for (int ci = 0; ci < C; ++ci) acc += sw[co][ci] * sx[ci][tt];
    o[i] = bf16_round(acc) + sx[co][tt];
  }
  for (int rep = 0; rep < reps; ++rep) {
    __nv_bfloat16* dst = out + (static_cast<long>(rep * B + b) * T + t) * C + cg * (C / 4);
#pragma unroll
    for (int i = 0; i < C / 4; i += 2) *reinterpret_cast<uint32_t*>(dst + i) = pack_bf16(o[i], o[i + 1]);
  }
}

// DiT input with channel-concatenated conditioning (inpainting: dit.py:160-165, `x = torch.cat([x, input_concat_cond], dim=1)`):
//   out[(rep*B + b)*T + t][c] = bf16(x[b][c][t] * c_in)        c <  C            (the sampler's state, rescaled per step)
//                               bf16(cond[b][c - C][t])         C <= c < C + Dc   (the same conditioning for both CFG halves, dit.py:336-337)
//                               0                               C + Dc <= c < Cp  (Cp = row pitch, a multiple of 8 for the GEMM's K)
// The 1x1 preprocess_conv + residual then run as ONE tcgen05 GEMM with the residual epilogue on these rows (its weight zero-padded to
// [Cp, Cp]); this kernel is only the transposition.  block = 32 time steps x all channels.
__global__ void __launch_bounds__(256) dit_concat_kernel(const float* __restrict__ x, const float* __restrict__ cond, __nv_bfloat16* __restrict__ out,
                                                         int B, int C, int Dc, int Cp, int T, int reps, const float* __restrict__ cin_table,
                                                         const int* __restrict__ step) {
  extern __shared__ float sx[];   // [Cp][33]
  griddep_launch();
  griddep_wait();
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * 32;
  const float c_in = cin_table ? cin_table[step ? *step : 0] : 1.0f;
  for (int i = threadIdx.x; i < Cp * 32; i += 256) {
    const int c = i >> 5, tt = i & 31;
    const int t = t0 + tt;
    float v = 0.f;
    if (t < T) {
      if (c < C) v = x[(static_cast<long>(b) * C + c) * T + t] * c_in;
      else if (c < C + Dc) v = cond[(static_cast<long>(b) * Dc + (c - C)) * T + t];
    }
    sx[c * 33 + tt] = v;
  }
  __syncthreads();
  const int pairs = Cp >> 1;
  for (int i = threadIdx.x; i < 32 * pairs; i += 256) {
    const int tt = i / pairs, cp = i % pairs;
    const int t = t0 + tt;
    if (t >= T) continue;
    const uint32_t v = pack_bf16(sx[(2 * cp) * 33 + tt], sx[(2 * cp + 1) * 33 + tt]);
    for (int rep = 0; rep < reps; ++rep)
      *reinterpret_cast<uint32_t*>(out + (static_cast<long>(rep * B + b) * T + t) * Cp + 2 * cp) = v;
  }
}

// DiT output: rows [bb, P + t, C] bf16 -> o[bb, c, t]; y = bf16(conv1x1(o)) + o  (dit.py:219-224), then classifier-free
// guidance over the (cond | uncond) batch halves: cfg = u + (c - u)*s, optional std rescale (dit.py:398-408).
// Writes fp32 v[B, C, T].
template <int C>
__global__ void __launch_bounds__(256) dit_post_kernel(const __nv_bfloat16* __restrict__ h, long ld_batch, int prepend,
                                                       const __nv_bfloat16* __restrict__ wconv, float* __restrict__ out, int B,
                                                       int T, int cfg, float cfg_scale, float scale_phi) {
  constexpr int TT = 32;        // time steps per block
  constexpr int CPT = C / 8;    // channels per thread (8 warps x CPT = C)
  __shared__ float so[2][C][TT + 1];
  __shared__ float sy[2][C][TT + 1];
  __shared__ __nv_bfloat16 sw[C][C];
  griddep_launch();
  griddep_wait();
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * TT;
  const int nb = cfg ? 2 : 1;
  for (int i = threadIdx.x; i < C * C; i += 256) sw[i / C][i % C] = wconv[i];
  for (int half = 0; half < nb; ++half) {
    const __nv_bfloat16* src = h + static_cast<long>(half * B + b) * ld_batch;
    for (int i = threadIdx.x; i < C * TT; i += 256) {
      const int tt = i / C, c = i % C;
      const int t = t0 + tt;
      so[half][c][tt] = (t < T) ? __bfloat162float(src[static_cast<long>(prepend + t) * C + c]) : 0.f;
    }
  }
  __syncthreads();
  const int tt = threadIdx.x & 31;
  const int cg = threadIdx.x >> 5;  // warp index: weights are a warp-wide broadcast
  for (int half = 0; half < nb; ++half) {
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
      const int co = cg * CPT + i;
      float acc = 0.f;
#pragma unroll 8
      for (int ci = 0; ci < C; ++ci) acc += __bfloat162float(sw[co][ci]) * so[half][ci][tt];
      sy[half][co][tt] = bf16_round(bf16_round(acc) + so[half][co][tt]);
    }
  }
  __syncthreads();
  const int t = t0 + tt;
  if (!cfg) {
    if (t < T) {
#pragma unroll
      for (int i = 0; i < CPT; ++i) {
        const int co = cg * CPT + i;
        out[(static_cast<long>(b) * C + co) * T + t] = sy[0][co][tt];
      }
    }
    return;
  }
  // CFG combine (bf16 eager rounding points), then per-(b,t) unbiased std over channels for the rescale
  float ratio = 1.f;
#pragma unroll
  for (int i = 0; i < CPT; ++i) {
    const int co = cg * CPT + i;
    const float c_ = sy[0][co][tt], u_ = sy[1][co][tt];
    so[0][co][tt] = bf16_round(u_ + bf16_round(bf16_round(c_ - u_) * cfg_scale));
  }
  __syncthreads();
  if (scale_phi != 0.f) {
    float mc = 0.f, mo = 0.f;
    for (int c = 0; c < C; ++c) { mc += sy[0][c][tt]; mo += so[0][c][tt]; }
    mc /= C; mo /= C;
    float vc = 0.f, vo = 0.f;
    for (int c = 0; c < C; ++c) {
      const float dc = sy[0][c][tt] - mc, d_o = so[0][c][tt] - mo;
      vc += dc * dc; vo += d_o * d_o;
    }
    const float sc = bf16_round(sqrtf(vc / (C - 1))), sof = bf16_round(sqrtf(vo / (C - 1)));
    ratio = bf16_round(sc / sof);
  }
  if (t < T) {
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
      const int co = cg * CPT + i;
      float v = so[0][co][tt];
      if (scale_phi != 0.f) v = bf16_round(bf16_round(scale_phi * bf16_round(v * ratio)) + bf16_round((1.f - scale_phi) * v));
      out[(static_cast<long>(b) * C + co) * T + t] = v;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Sampler state update (k-diffusion VDenoiser + DPM-Solver++(3M) SDE step, or any linear multistep rule):
//   den   = v * c_out + x * c_skip                                   (k_diffusion/external.py VDenoiser)
//   x_new = A*x + Bd*den + C1*d1 + C2*d2 + NZ*noise                  (coefficients precomputed per step on the host)
// history ring of 3 denoised tensors indexed by the device-side step counter; the counter is advanced here so that a
// CUDA graph of one sampling step can be replayed for every step.
__global__ void sampler_update_kernel(float* __restrict__ x, const float* __restrict__ v, float* __restrict__ hist,
                                      const float* __restrict__ noise, const float* __restrict__ coef, const int* __restrict__ step,
                                      long n) {
  griddep_launch();
  griddep_wait();
  const int s = *step;
  const float* c = coef + s * 8;
  const float c_out = c[0], c_skip = c[1], A = c[2], Bd = c[3], C1 = c[4], C2 = c[5], NZ = c[6];
  float* hcur = hist + static_cast<long>(s % 3) * n;
  const float* h1 = hist + static_cast<long>((s + 2) % 3) * n;
  const float* h2 = hist + static_cast<long>((s + 1) % 3) * n;
  const float* nz = noise ? noise + static_cast<long>(s) * n : nullptr;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const float xv = x[i];
    const float den = v[i] * c_out + xv * c_skip;
    float xn = A * xv + Bd * den;
    if (C1 != 0.f) xn += C1 * h1[i];
    if (C2 != 0.f) xn += C2 * h2[i];
    if (nz && NZ != 0.f) xn += NZ * nz[i];
    hcur[i] = den;
    x[i] = xn;
  }
}
__global__ void step_advance_kernel(int* step) { griddep_launch(); griddep_wait(); *step += 1; }
__global__ void step_set_kernel(int* step, int v) { griddep_launch(); griddep_wait(); *step = v; }

}  // namespace b200sat

using namespace b200sat;

extern "C" int b200sat_layernorm_fwd(const void* x, long ldx, const float* gamma, const float* beta, const float* scale,
                                     const float* shift, long ld_mod, int rows_per_batch, void* y, long ldy, int rows, int D,
                                     float eps, void* stream) {
  if (!x || !gamma || !y || rows <= 0 || D <= 0) { set_last_error("layernorm: bad arguments"); return B200SAT_EINVAL; }
  if (D % 8 || ldx % 8 || ldy % 8) { set_last_error("layernorm: D and leading dims must be multiples of 8"); return B200SAT_EINVAL; }
  if (D > 4096) { set_last_error("layernorm: D > 4096 not implemented"); return B200SAT_EUNSUPPORTED; }
  if ((scale != nullptr) != (shift != nullptr)) { set_last_error("layernorm: scale and shift go together"); return B200SAT_EINVAL; }
  const int grid = (rows + 7) / 8;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int rpb = rows_per_batch > 0 ? rows_per_batch : rows;
  if ((reinterpret_cast<uintptr_t>(gamma) & 15) || (beta && (reinterpret_cast<uintptr_t>(beta) & 15)) ||
      (scale && ((reinterpret_cast<uintptr_t>(scale) & 15) || (reinterpret_cast<uintptr_t>(shift) & 15) || (ld_mod & 3)))) {
    set_last_error("layernorm: gamma / beta / scale / shift must be 16-byte aligned (row stride of scale/shift a multiple of 4)"); return B200SAT_EINVAL;
  }
  if (D <= 1536)
    B200SAT_CHECK_CUDA(launch_k(layernorm_kernel<6>, dim3(grid), dim3(256), 0, s, 1, static_cast<const __nv_bfloat16*>(x), gamma, beta, scale, shift,
                                             static_cast<__nv_bfloat16*>(y), rows, D, ldx, ldy, rpb, ld_mod, eps));
  else if (D <= 2048)
    B200SAT_CHECK_CUDA(launch_k(layernorm_kernel<8>, dim3(grid), dim3(256), 0, s, 1, static_cast<const __nv_bfloat16*>(x), gamma, beta, scale, shift,
                                             static_cast<__nv_bfloat16*>(y), rows, D, ldx, ldy, rpb, ld_mod, eps));
  else
    B200SAT_CHECK_CUDA(launch_k(layernorm_kernel<16>, dim3(grid), dim3(256), 0, s, 1, static_cast<const __nv_bfloat16*>(x), gamma, beta, scale, shift,
                                              static_cast<__nv_bfloat16*>(y), rows, D, ldx, ldy, rpb, ld_mod, eps));
  B200SAT_CHECK_CUDA(cudaGetLastError());
  return B200SAT_OK;
}

extern "C" int b200sat_small_linear(const void* x, long ldx, const void* w, long ldw, const float* bias, const void* add,
                                    long ldadd, void* y, long ldy, int M, int N, int K, int act_silu, int out_f32,
                                    int act_sigmoid_1m, float* stats, long stats_stride, void* stream) {
  if (!x || !w || !y || M <= 0 || M > 8 || N <= 0 || K <= 0) { set_last_error("small_linear: bad arguments (1 <= M <= 8)"); return B200SAT_EINVAL; }
  if (K % 8 || ldx % 8 || ldw % 8) { set_last_error("small_linear: K, ldx, ldw must be multiples of 8"); return B200SAT_EINVAL; }
  B200SAT_CHECK_CUDA(launch_k(small_linear_kernel, dim3((N + 7) / 8), dim3(256), 0, static_cast<cudaStream_t>(stream), 1, 
      static_cast<const __nv_bfloat16*>(x), ldx, static_cast<const __nv_bfloat16*>(w), ldw, bias,
      static_cast<const __nv_bfloat16*>(add), ldadd, y, ldy, M, N, K, act_silu, out_f32, act_sigmoid_1m, stats, stats_stride));
  B200SAT_CHECK_CUDA(cudaGetLastError());
  return B200SAT_OK;
}

extern "C" int b200sat_fourier_features(const float* t, const void* w, void* out, int B, int half, const int* step,
                                        int t_stride, void* stream) {
  if (!t || !w || !out || B <= 0 || half <= 0) { set_last_error("fourier_features: bad arguments"); return B200SAT_EINVAL; }
  const int n = B * half;
  B200SAT_CHECK_CUDA(launch_k(fourier_features_kernel, dim3((n + 127) / 128), dim3(128), 0, static_cast<cudaStream_t>(stream), 1, 
      t, static_cast<const __nv_bfloat16*>(w), static_cast<__nv_bfloat16*>(out), B, half, step, t_stride));
  B200SAT_CHECK_CUDA(cudaGetLastError());
  return B200SAT_OK;
}

extern "C" int b200sat_dit_pre(const float* x, const void* wconv, void* out, int B, int C, int T, int reps,
                               const float* cin_table, const int* step, void* stream) {
  if (!x || !wconv || !out || B <= 0 || T <= 0 || reps <= 0) { set_last_error("dit_pre: bad arguments"); return B200SAT_EINVAL; }
  if (C != 64) { set_last_error("dit_pre: only io_channels == 64 is implemented"); return B200SAT_EUNSUPPORTED; }
  dim3 grid((T + 63) / 64, B);
  B200SAT_CHECK_CUDA(launch_k(dit_pre_kernel<64>, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(stream), 1, x, static_cast<const __nv_bfloat16*>(wconv),
                                                                         static_cast<__nv_bfloat16*>(out), B, T, reps,
                                                                         cin_table, step));
  B200SAT_CHECK_CUDA(cudaGetLastError());
  return B200SAT_OK;
}

extern "C" int b200sat_dit_concat(const float* x, const float* cond, void* out, int B, int C, int Dc, int Cp, int T, int reps, const float* cin_table,
                                  const int* step, void* stream) {
  if (!x || !cond || !out || B <= 0 || T <= 0 || reps <= 0 || C <= 0 || Dc <= 0) { set_last_error("dit_concat: bad arguments"); return B200SAT_EINVAL; }
  if (Cp < C + Dc || (Cp & 7) || Cp > 256) { set_last_error("dit_concat: Cp must be a multiple of 8 in [C + Dc, 256]"); return B200SAT_EINVAL; }
  dim3 grid((T + 31) / 32, B);
  B200SAT_CHECK_CUDA(launch_k(dit_concat_kernel, dim3(grid), dim3(256), static_cast<size_t>(Cp) * 33 * sizeof(float), static_cast<cudaStream_t>(stream), 1, x,
                              cond, static_cast<__nv_bfloat16*>(out), B, C, Dc, Cp, T, reps, cin_table, step));
  B200SAT_CHECK_CUDA(cudaGetLastError());
  return B200SAT_OK;
}

extern "C" int b200sat_dit_post(const void* h, long ld_batch, int prepend, const void* wconv, float* out, int B, int C, int T,
                                int cfg, float cfg_scale, float scale_phi, void* stream) {
  if (!h || !wconv || !out || B <= 0 || T <= 0) { set_last_error("dit_post: bad arguments"); return B200SAT_EINVAL; }
  if (C != 64) { set_last_error("dit_post: only io_channels == 64 is implemented"); return B200SAT_EUNSUPPORTED; }
  dim3 grid((T + 31) / 32, B);
  B200SAT_CHECK_CUDA(launch_k(dit_post_kernel<64>, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(stream), 1, static_cast<const __nv_bfloat16*>(h), ld_batch,
                                                                          prepend, static_cast<const __nv_bfloat16*>(wconv),
                                                                          out, B, T, cfg, cfg_scale, scale_phi));
  B200SAT_CHECK_CUDA(cudaGetLastError());
  return B200SAT_OK;
}

extern "C" int b200sat_sampler_update(float* x, const float* v, float* hist, const float* noise, const float* coef, int* step,
                                      long n, int advance, void* stream) {
  if (!x || !v || !hist || !coef || !step || n <= 0) { set_last_error("sampler_update: bad arguments"); return B200SAT_EINVAL; }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int grid = static_cast<int>((n + 1023) / 1024 < 1184 ? (n + 1023) / 1024 : 1184);
  B200SAT_CHECK_CUDA(launch_k(sampler_update_kernel, dim3(grid), dim3(256), 0, s, 1, x, v, hist, noise, coef, step, n));
  if (advance) B200SAT_CHECK_CUDA(launch_k(step_advance_kernel, dim3(1), dim3(1), 0, s, 1, step));
  B200SAT_CHECK_CUDA(cudaGetLastError());
  return B200SAT_OK;
}

extern "C" int b200sat_step_set(int* step, int value, void* stream) {
  if (!step) { set_last_error("step_set: null"); return B200SAT_EINVAL; }
  B200SAT_CHECK_CUDA(launch_k(step_set_kernel, dim3(1), dim3(1), 0, static_cast<cudaStream_t>(stream), 1, step, value));
  B200SAT_CHECK_CUDA(cudaGetLastError());
  return B200SAT_OK;
}

// b200sat — multi-resolution STFT loss without ever writing a spectrogram to HBM.
//
// Replaces, for MultiResolutionSTFTLoss / SumAndDifferenceSTFTLoss (stable_audio_tools/training/losses/auraloss.py:451-615):
//   FIRFilter.forward (A-weighting, 101 taps, :155-169) + SumAndDifference (:44-73)        -> stft_prefilter_kernel
//   STFTLoss.stft (torch.stft: reflect pad, periodic hann, rFFT -> sqrt(clamp(|.|^2, eps)), :368-395),
//   SpectralConvergenceLoss (:171-181), STFTMagnitudeLoss (log, L1, :183-223)             -> stft_loss_kernel
//
// One warp = one frame: a warp-cooperative Stockham radix-2 FFT in shared memory of the COMPLEX signal z = x_w + i*y_w
// (input frame and target frame at once); the two real spectra are separated with X[k] = (Z[k] + conj Z[n-k])/2,
// Y[k] = (Z[k] - conj Z[n-k])/(2i).  Each warp folds its bins into three running sums per signal row
//   S1 += (|Y|-|X|)^2   S2 += |Y|^2   S3 += |log|X| - log|Y||
// (fp32 per warp, fp64 atomics per block).  HBM traffic = the filtered waveforms (read ~4x from L2 because hop = n/4).
#include "common.cuh"

namespace b200sat {

// out[b, r, t] = sum_k taps[k] * (sum_c mix[r, c] * x[b, c, t + k - ntaps/2])     zero padding (F.conv1d padding=ntaps//2)
__global__ void __launch_bounds__(256) stft_prefilter_kernel(const float* __restrict__ x, float* __restrict__ out,
                                                             const float* __restrict__ mix, const float* __restrict__ taps,
                                                             int B, int C, int T, int R, int ntaps) {
  extern __shared__ float sm[];
  float* s_taps = sm;               // ntaps
  float* s_x = sm + ntaps;          // C * (1024 + ntaps - 1)
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * 1024;
  const int half = ntaps / 2;
  const int span = 1024 + ntaps - 1;
  for (int i = threadIdx.x; i < ntaps; i += 256) s_taps[i] = taps[i];
  for (int i = threadIdx.x; i < C * span; i += 256) {
    const int c = i / span, tt = i % span;
    const int t = t0 + tt - half;
    s_x[i] = (t >= 0 && t < T) ? x[(static_cast<long>(b) * C + c) * T + t] : 0.f;
  }
  __syncthreads();
  for (int tt = threadIdx.x; tt < 1024; tt += 256) {
    const int t = t0 + tt;
    if (t >= T) break;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c = 0; c < C; ++c) {
      float f = 0.f;
      for (int k = 0; k < ntaps; ++k) f += s_taps[k] * s_x[c * span + tt + k];
      for (int r = 0; r < R; ++r) acc[r] += mix[r * C + c] * f;
    }
    for (int r = 0; r < R; ++r) out[(static_cast<long>(b) * R + r) * T + t] = acc[r];
  }
}

__device__ __forceinline__ int reflect_index(int t, int T) {
  if (t < 0) t = -t;
  if (t >= T) t = 2 * (T - 1) - t;
  return t;
}

// grid: (frame groups, rows); block: W warps; dynamic smem: W * 2 * n * sizeof(float2)
__global__ void __launch_bounds__(256) stft_loss_kernel(const float* __restrict__ xf, const float* __restrict__ yf,
                                                        double* __restrict__ acc, const float* __restrict__ window,
                                                        const float2* __restrict__ twiddle, int T, int n, int log2n, int hop,
                                                        int frames, int frames_per_block, float eps) {
  extern __shared__ float2 fft_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;
  float2* buf0 = fft_smem + static_cast<size_t>(warp) * 2 * n;
  float2* buf1 = buf0 + n;
  const int row = blockIdx.y;
  const float* xs = xf + static_cast<long>(row) * T;
  const float* ys = yf + static_cast<long>(row) * T;
  const int f_begin = blockIdx.x * frames_per_block;
  const int f_end = min(frames, f_begin + frames_per_block);
  float s1 = 0.f, s2 = 0.f, s3 = 0.f;
  const int half = n >> 1;
  for (int f = f_begin + warp; f < f_end; f += nwarps) {
    const int base = f * hop - half;
    for (int i = lane; i < n; i += 32) {
      const int t = reflect_index(base + i, T);
      const float w = __ldg(window + i);
      buf0[i] = make_float2(__ldg(xs + t) * w, __ldg(ys + t) * w);
    }
    __syncwarp();
    float2* in = buf0;
    float2* out = buf1;
    for (int s = 0; s < log2n; ++s) {
      const int Ns = 1 << s;
      const int tw_stride = half >> s;  // n / (2 Ns)
      for (int j = lane; j < half; j += 32) {
        const int k = j & (Ns - 1);
        const float2 w = __ldg(twiddle + k * tw_stride);
        const float2 a = in[j];
        const float2 b = in[j + half];
        const float2 bw = make_float2(b.x * w.x - b.y * w.y, b.x * w.y + b.y * w.x);
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
      const float xm = sqrtf(fmaxf(xr * xr + xi * xi, eps));
      const float ym = sqrtf(fmaxf(yr * yr + yi * yi, eps));
      const float d = ym - xm;
      s1 += d * d;
      s2 += ym * ym;
      s3 += fabsf(logf(xm) - logf(ym));
    }
    __syncwarp();
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    s3 += __shfl_xor_sync(0xffffffffu, s3, o);
  }
  if (lane == 0) {
    atomicAdd(acc + row * 3 + 0, static_cast<double>(s1));
    atomicAdd(acc + row * 3 + 1, static_cast<double>(s2));
    atomicAdd(acc + row * 3 + 2, static_cast<double>(s3));
  }
}


// ---------------------------------------------------------------------------------------------------------
// Backward of one resolution.  Per frame the forward transform is recomputed (nothing was stored), the per-bin gradients of the
// two magnitudes are formed from the per-row coefficients
//     dL/dX = a1 (X - Y) + a3 sgn(log X - log Y) / X          a1 = c_sc / sqrt(S1 S2),  a2 = c_sc sqrt(S1) / S2^1.5,  a3 = c_log
//     dL/dY = a1 (Y - X) - a2 Y - a3 sgn(log X - log Y) / Y
// turned into spectrum gradients (d|Z|/dRe = Re/|Z| where not clamped), Hermitian-extended so that ONE inverse-direction complex
// FFT returns d/dx_w in the real part and d/dy_w in the imaginary part, windowed, and scattered (reflect-folded) into the
// gradients of the filtered waveforms with fp32 atomics (frames overlap 4x).
__global__ void __launch_bounds__(256) stft_loss_bwd_kernel(const float* __restrict__ xf, const float* __restrict__ yf,
                                                            float* __restrict__ dxf, float* __restrict__ dyf,
                                                            const float* __restrict__ coef, const float* __restrict__ window,
                                                            const float2* __restrict__ twiddle, int T, int n, int log2n, int hop,
                                                            int frames, int frames_per_block, float eps) {
  extern __shared__ float2 fft_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;
  float2* buf0 = fft_smem + static_cast<size_t>(warp) * 2 * n;
  float2* buf1 = buf0 + n;
  const int row = blockIdx.y;
  const float* xs = xf + static_cast<long>(row) * T;
  const float* ys = yf + static_cast<long>(row) * T;
  float* dxs = dxf + static_cast<long>(row) * T;
  float* dys = dyf + static_cast<long>(row) * T;
  const float a1 = coef[row * 3 + 0], a2 = coef[row * 3 + 1], a3 = coef[row * 3 + 2];
  const int f_begin = blockIdx.x * frames_per_block;
  const int f_end = min(frames, f_begin + frames_per_block);
  const int half = n >> 1;
  for (int f = f_begin + warp; f < f_end; f += nwarps) {
    const int base = f * hop - half;
    for (int i = lane; i < n; i += 32) {
      const int t = reflect_index(base + i, T);
      const float w = __ldg(window + i);
      buf0[i] = make_float2(__ldg(xs + t) * w, __ldg(ys + t) * w);
    }
    __syncwarp();
    float2* in = buf0;
    float2* out = buf1;
    for (int s = 0; s < log2n; ++s) {
      const int Ns = 1 << s, tw_stride = half >> s;
      for (int j = lane; j < half; j += 32) {
        const int k = j & (Ns - 1);
        const float2 w = __ldg(twiddle + k * tw_stride);
        const float2 a = in[j], b = in[j + half];
        const float2 bw = make_float2(b.x * w.x - b.y * w.y, b.x * w.y + b.y * w.x);
        const int j0 = ((j - k) << 1) + k;
        out[j0] = make_float2(a.x + bw.x, a.y + bw.y);
        out[j0 + Ns] = make_float2(a.x - bw.x, a.y - bw.y);
      }
      __syncwarp();
      float2* tmp = in; in = out; out = tmp;
    }
    // spectrum gradients -> Hermitian-extended packed spectrum in `out`
    for (int k = lane; k <= half; k += 32) {
      const float2 zk = in[k];
      const float2 zn = in[(n - k) & (n - 1)];
      const float xr = 0.5f * (zk.x + zn.x), xi = 0.5f * (zk.y - zn.y);
      const float yr = 0.5f * (zk.y + zn.y), yi = 0.5f * (zn.x - zk.x);
      const float px = xr * xr + xi * xi, py = yr * yr + yi * yi;
      const float xm = sqrtf(fmaxf(px, eps)), ym = sqrtf(fmaxf(py, eps));
      const float lg = logf(xm) - logf(ym);
      const float sg = (lg > 0.f) ? 1.f : ((lg < 0.f) ? -1.f : 0.f);
      const float gX = a1 * (xm - ym) + a3 * sg / xm;
      const float gY = a1 * (ym - xm) - a2 * ym - a3 * sg / ym;
      const float sx = (px > eps) ? gX / xm : 0.f, sy = (py > eps) ? gY / ym : 0.f;
      float gxr = sx * xr, gxi = sx * xi, gyr = sy * yr, gyi = sy * yi;   // dL/dRe, dL/dIm of the one-sided bins
      if (k == 0 || k == half) {
        // purely real basis vector: only the real part of the bin gradient reaches the signal
        out[k] = make_float2(gxr, gyr);
      } else {
        // H[k] = G/2, H[n-k] = conj(G)/2 for each signal; packed as Hx + i*Hy
        out[k] = make_float2(0.5f * (gxr - gyi), 0.5f * (gxi + gyr));
        out[n - k] = make_float2(0.5f * (gxr + gyi), 0.5f * (-gxi + gyr));
      }
    }
    __syncwarp();
    // inverse-direction transform (conjugated twiddles), unnormalised: z[t] = sum_k H[k] e^{+2 pi i k t / n}
    { float2* tmp = in; in = out; out = tmp; }
    for (int s = 0; s < log2n; ++s) {
      const int Ns = 1 << s, tw_stride = half >> s;
      for (int j = lane; j < half; j += 32) {
        const int k = j & (Ns - 1);
        float2 w = __ldg(twiddle + k * tw_stride);
        w.y = -w.y;
        const float2 a = in[j], b = in[j + half];
        const float2 bw = make_float2(b.x * w.x - b.y * w.y, b.x * w.y + b.y * w.x);
        const int j0 = ((j - k) << 1) + k;
        out[j0] = make_float2(a.x + bw.x, a.y + bw.y);
        out[j0 + Ns] = make_float2(a.x - bw.x, a.y - bw.y);
      }
      __syncwarp();
      float2* tmp = in; in = out; out = tmp;
    }
    for (int i = lane; i < n; i += 32) {
      const int t = reflect_index(base + i, T);
      const float w = __ldg(window + i);
      const float2 z = in[i];
      atomicAdd(dxs + t, z.x * w);
      atomicAdd(dys + t, z.y * w);
    }
    __syncwarp();
  }
}

// Backward of stft_prefilter_kernel: dx[b,c,t] = sum_r mix[r,c] sum_k taps[k] * dout[b,r,t - k + ntaps/2]   (zero outside [0,T))
__global__ void __launch_bounds__(256) stft_prefilter_bwd_kernel(const float* __restrict__ dout, float* __restrict__ dx,
                                                                 const float* __restrict__ mix, const float* __restrict__ taps,
                                                                 int B, int C, int T, int R, int ntaps) {
  extern __shared__ float sm[];
  float* s_taps = sm;
  float* s_d = sm + ntaps;   // R * (1024 + ntaps - 1)
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * 1024;
  const int half = ntaps / 2;
  const int span = 1024 + ntaps - 1;
  for (int i = threadIdx.x; i < ntaps; i += 256) s_taps[i] = taps[i];
  for (int i = threadIdx.x; i < R * span; i += 256) {
    const int r = i / span, tt = i % span;
    const int t = t0 + tt - half;
    s_d[i] = (t >= 0 && t < T) ? dout[(static_cast<long>(b) * R + r) * T + t] : 0.f;
  }
  __syncthreads();
  for (int tt = threadIdx.x; tt < 1024; tt += 256) {
    const int t = t0 + tt;
    if (t >= T) break;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int r = 0; r < R; ++r) {
      float f = 0.f;
      for (int k = 0; k < ntaps; ++k) f += s_taps[k] * s_d[r * span + tt + (ntaps - 1 - k)];   // dout[t - k + half]
      for (int c = 0; c < C && c < 4; ++c) acc[c] += mix[r * C + c] * f;
    }
    for (int c = 0; c < C && c < 4; ++c) dx[(static_cast<long>(b) * C + c) * T + t] = acc[c];
  }
}

}  // namespace b200sat

using namespace b200sat;

extern "C" int b200sat_stft_prefilter(const float* x, float* out, const float* mix, const float* taps, int B, int C, int T, int R,
                                      int ntaps, void* stream) {
  if (!x || !out || !mix || !taps || B <= 0 || C <= 0 || T <= 0 || R <= 0 || R > 4 || ntaps <= 0 || !(ntaps & 1)) {
    set_last_error("stft_prefilter: bad arguments (1 <= R <= 4, ntaps odd)");
    return B200SAT_EINVAL;
  }
  const int smem = (ntaps + C * (1024 + ntaps - 1)) * 4;
  if (smem > 48 * 1024) { set_last_error("stft_prefilter: too many channels/taps for 48 KB of shared memory"); return B200SAT_EUNSUPPORTED; }
  dim3 grid((T + 1023) / 1024, B);
  stft_prefilter_kernel<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(x, out, mix, taps, B, C, T, R, ntaps);
  B200SAT_CHECK_CUDA(cudaGetLastError());
  return B200SAT_OK;
}

extern "C" int b200sat_stft_loss_accumulate(const float* xf, const float* yf, double* acc, const float* window, const float* twiddle,
                                            int rows, int T, int n_fft, int hop, float eps, void* stream) {
  if (!xf || !yf || !acc || !window || !twiddle || rows <= 0 || T <= 0 || hop <= 0) { set_last_error("stft_loss: bad arguments"); return B200SAT_EINVAL; }
  int log2n = 0;
  while ((1 << log2n) < n_fft) ++log2n;
  if ((1 << log2n) != n_fft || n_fft < 32 || n_fft > 4096) { set_last_error("stft_loss: n_fft must be a power of two in [32, 4096]"); return B200SAT_EUNSUPPORTED; }
  if (T <= n_fft / 2) { set_last_error("stft_loss: signal shorter than the reflect padding"); return B200SAT_EINVAL; }
  const int frames = T / hop + 1;
  int warps = 8;
  while (warps > 1 && warps * 2 * n_fft * 8 > 160 * 1024) warps >>= 1;
  const int smem = warps * 2 * n_fft * 8;
  static int smem_set = 0;
  if (smem > smem_set) {
    B200SAT_CHECK_CUDA(cudaFuncSetAttribute(stft_loss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    smem_set = 160 * 1024;
  }
  int fpb = warps * 4;   // frames per block
  dim3 grid((frames + fpb - 1) / fpb, rows);
  stft_loss_kernel<<<grid, warps * 32, smem, static_cast<cudaStream_t>(stream)>>>(xf, yf, acc, window, reinterpret_cast<const float2*>(twiddle),
                                                                                 T, n_fft, log2n, hop, frames, fpb, eps);
  B200SAT_CHECK_CUDA(cudaGetLastError());
  return B200SAT_OK;
}

extern "C" int b200sat_stft_loss_backward(const float* xf, const float* yf, float* dxf, float* dyf, const float* coef, const float* window,
                                          const float* twiddle, int rows, int T, int n_fft, int hop, float eps, void* stream) {
  if (!xf || !yf || !dxf || !dyf || !coef || !window || !twiddle || rows <= 0 || T <= 0 || hop <= 0) { set_last_error("stft_loss_backward: bad arguments"); return B200SAT_EINVAL; }
  int log2n = 0;
  while ((1 << log2n) < n_fft) ++log2n;
  if ((1 << log2n) != n_fft || n_fft < 32 || n_fft > 4096) { set_last_error("stft_loss_backward: n_fft must be a power of two in [32, 4096]"); return B200SAT_EUNSUPPORTED; }
  const int frames = T / hop + 1;
  int warps = 8;
  while (warps > 1 && warps * 2 * n_fft * 8 > 160 * 1024) warps >>= 1;
  const int smem = warps * 2 * n_fft * 8;
  static bool attr = false;
  if (!attr) { B200SAT_CHECK_CUDA(cudaFuncSetAttribute(stft_loss_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024)); attr = true; }
  const int fpb = warps * 4;
  dim3 grid((frames + fpb - 1) / fpb, rows);
  stft_loss_bwd_kernel<<<grid, warps * 32, smem, static_cast<cudaStream_t>(stream)>>>(xf, yf, dxf, dyf, coef, window, reinterpret_cast<const float2*>(twiddle),
                                                                                     T, n_fft, log2n, hop, frames, fpb, eps);
  B200SAT_CHECK_CUDA(cudaGetLastError());
  return B200SAT_OK;
}

extern "C" int b200sat_stft_prefilter_backward(const float* dout, float* dx, const float* mix, const float* taps, int B, int C, int T, int R,
                                               int ntaps, void* stream) {
  if (!dout || !dx || !mix || !taps || B <= 0 || C <= 0 || C > 4 || T <= 0 || R <= 0 || R > 4 || ntaps <= 0 || !(ntaps & 1)) {
    set_last_error("stft_prefilter_backward: bad arguments (C, R <= 4, ntaps odd)");
    return B200SAT_EINVAL;
  }
  const int smem = (ntaps + R * (1024 + ntaps - 1)) * 4;
  if (smem > 48 * 1024) { set_last_error("stft_prefilter_backward: shared memory"); return B200SAT_EUNSUPPORTED; }
  dim3 grid((T + 1023) / 1024, B);
  stft_prefilter_bwd_kernel<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(dout, dx, mix, taps, B, C, T, R, ntaps);
  B200SAT_CHECK_CUDA(cudaGetLastError());
  return B200SAT_OK;
}

// b200sat — HBM-bound kernels of the DiT backward pass: LayerNorm backward (+ residual-stream gradient add, + dgamma),
// column sums for bias gradients.  (The GEMM, SwiGLU and attention backward live in gemm.cu / attention_bwd.cu.)
#include "common.cuh"

namespace b200sat {

// LayerNorm backward for y = (x - mean) * rstd * gamma  (transformer.py:236-238, beta is a zero buffer).
//   g = dy * gamma;  dx = rstd * (g - mean(g) - xhat * mean(g * xhat));  dgamma += sum_rows dy * xhat
// dx_out = dres + dx fuses the residual-stream gradient add of the pre-norm block (transformer.py:703-712).
// Persistent grid: each warp walks rows, keeps its dgamma partials in registers; one atomicAdd per column per block.
template <int MAXC>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                                                            const float* __restrict__ gamma, const __nv_bfloat16* __restrict__ dres,
                                                            __nv_bfloat16* __restrict__ dx_out, float* __restrict__ dgamma, int rows,
                                                            int D, long ldx, long lddy, long ldr, long ldo, float eps) {
  extern __shared__ float s_dg[];  // [D]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nchunk = D / 8;
  for (int i = threadIdx.x; i < D; i += blockDim.x) s_dg[i] = 0.f;
  __syncthreads();
  float dg[MAXC][8];
#pragma unroll
  for (int c = 0; c < MAXC; ++c)
#pragma unroll
    for (int j = 0; j < 8; ++j) dg[c][j] = 0.f;

  for (int row = blockIdx.x * 8 + warp; row < rows; row += gridDim.x * 8) {
    const __nv_bfloat16* xr = x + static_cast<long>(row) * ldx;
    const __nv_bfloat16* dr = dy + static_cast<long>(row) * lddy;
    float xv[MAXC][8], gv[MAXC][8];
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      const int ch = lane + c * 32;
      if (ch < nchunk) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(xr) + ch);
        const uint4 w = __ldg(reinterpret_cast<const uint4*>(dr) + ch);
        const uint32_t uw[4] = {u.x, u.y, u.z, u.w}, ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = unpack_bf16(uw[j]), d = unpack_bf16(ww[j]);
          xv[c][2 * j] = f.x; xv[c][2 * j + 1] = f.y;
          gv[c][2 * j] = d.x; gv[c][2 * j + 1] = d.y;   // dy for now
          sum += f.x + f.y;
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / D;
    float sq = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c)
      if (lane + c * 32 < nchunk)
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float d = xv[c][j] - mean; sq += d * d; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq / D + eps);
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      const int ch = lane + c * 32;
      if (ch < nchunk) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = (xv[c][j] - mean) * rstd;
          const float dyv = gv[c][j];
          dg[c][j] += dyv * xh;
          const float g = dyv * __ldg(gamma + ch * 8 + j);
          xv[c][j] = xh; gv[c][j] = g;
          sg += g; sgx += g * xh;
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { sg += __shfl_xor_sync(0xffffffffu, sg, o); sgx += __shfl_xor_sync(0xffffffffu, sgx, o); }
    const float mg = sg / D, mgx = sgx / D;
    const __nv_bfloat16* rr = dres ? dres + static_cast<long>(row) * ldr : nullptr;
    __nv_bfloat16* orow = dx_out + static_cast<long>(row) * ldo;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      const int ch = lane + c * 32;
      if (ch < nchunk) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = rstd * (gv[c][j] - mg - xv[c][j] * mgx);
        if (rr) {
          const uint4 u = __ldg(reinterpret_cast<const uint4*>(rr) + ch);
          const uint32_t uw[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) { const float2 f = unpack_bf16(uw[j]); o[2 * j] += f.x; o[2 * j + 1] += f.y; }
        }
        uint4 u;
        u.x = pack_bf16(o[0], o[1]); u.y = pack_bf16(o[2], o[3]); u.z = pack_bf16(o[4], o[5]); u.w = pack_bf16(o[6], o[7]);
        reinterpret_cast<uint4*>(orow)[ch] = u;
      }
    }
  }
  if (dgamma) {
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      const int ch = lane + c * 32;
      if (ch < nchunk)
#pragma unroll
        for (int j = 0; j < 8; ++j) atomicAdd(&s_dg[ch * 8 + j], dg[c][j]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < D; i += blockDim.x) atomicAdd(dgamma + i, s_dg[i]);
  }
}

// out[n] += sum_m dY[m, n]  (bias gradients).  grid (ceil(N/512), row chunks); thread = 2 adjacent columns.
__global__ void __launch_bounds__(256) colsum_kernel(const __nv_bfloat16* __restrict__ dy, long ld, float* __restrict__ out, int M, int N,
                                                     int rows_per_block) {
  const int col = (blockIdx.x * 256 + threadIdx.x) * 2;
  if (col >= N) return;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(M, r0 + rows_per_block);
  float a0 = 0.f, a1 = 0.f;
  for (int r = r0; r < r1; ++r) {
    const uint32_t u = __ldg(reinterpret_cast<const uint32_t*>(dy + static_cast<long>(r) * ld + col));
    const float2 f = unpack_bf16(u);
    a0 += f.x; a1 += f.y;
  }
  atomicAdd(out + col, a0);
  if (col + 1 < N) atomicAdd(out + col + 1, a1);
}

}  // namespace b200sat

using namespace b200sat;

extern "C" int b200sat_layernorm_bwd(const void* x, long ldx, const void* dy, long lddy, const float* gamma, const void* dres, long ldr,
                                     void* dx_out, long ldo, float* dgamma, int rows, int D, float eps, void* stream) {
  if (!x || !dy || !gamma || !dx_out || rows <= 0 || D <= 0) { set_last_error("layernorm_bwd: bad arguments"); return B200SAT_EINVAL; }
  if (D % 8 || ldx % 8 || lddy % 8 || ldo % 8 || (dres && ldr % 8)) { set_last_error("layernorm_bwd: D and leading dims must be multiples of 8"); return B200SAT_EINVAL; }
  if (D > 2048) { set_last_error("layernorm_bwd: D > 2048 not implemented"); return B200SAT_EUNSUPPORTED; }
  int grid = (rows + 7) / 8;
  const int cap = 2 * num_sms();
  if (grid > cap) grid = cap;
  layernorm_bwd_kernel<8><<<grid, 256, D * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(dy), gamma, static_cast<const __nv_bfloat16*>(dres),
      static_cast<__nv_bfloat16*>(dx_out), dgamma, rows, D, ldx, lddy, ldr, ldo, eps);
  B200SAT_CHECK_CUDA(cudaGetLastError());
  return B200SAT_OK;
}

extern "C" int b200sat_colsum(const void* dy, long ld, float* out, int M, int N, void* stream) {
  if (!dy || !out || M <= 0 || N <= 0 || (N % 2) || (ld % 2)) { set_last_error("colsum: bad arguments (N, ld even)"); return B200SAT_EINVAL; }
  const int rpb = 256;
  dim3 grid((N + 511) / 512, (M + rpb - 1) / rpb);
  colsum_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(dy), ld, out, M, N, rpb);
  B200SAT_CHECK_CUDA(cudaGetLastError());
  return B200SAT_OK;
}

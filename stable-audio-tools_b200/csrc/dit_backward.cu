// b200sat — HBM-bound kernels of the DiT backward pass: LayerNorm backward (+ residual-stream gradient add, + dgamma),
// column sums for bias gradients.  (The GEMM, SwiGLU and attention backward live in gemm.cu / attention_bwd.cu.)
#include "common.cuh"

namespace b200sat {

// LayerNorm backward for y = (x - mean) * rstd * gamma  (transformer.py:236-238, beta is a zero buffer).
//   g = dy * gamma;  dx = rstd * (g - mean(g) - xhat * mean(g * xhat));  dgamma += sum_rows dy * xhat
// dx_out = dres + dx fuses the residual-stream gradient add of the pre-norm block (transformer.py:703-712).
// Persistent grid: each warp walks rows, keeps its dgamma partials in registers; one atomicAdd per column per block.
// MOD (adaLN, transformer.py:680-697: y = LN(x; gamma) * (1 + scale_b) + shift_b): the effective gain is gamma * (1 + scale[b, :]); the grid
// is (blocks per batch entry, B) so a block only sees rows of one batch entry, and `dgamma` receives P[b, :] = sum_n dy * xhat PER BATCH
// ENTRY, from which the host forms dgamma = sum_b (1 + scale_b) P_b and dscale_b = gamma * P_b (dshift_b = per-batch column sums of dy).
template <int MAXC, bool MOD>
__global__ void __launch_bounds__(256, 2) layernorm_bwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                                                               const float* __restrict__ gamma, const __nv_bfloat16* __restrict__ dres,
                                                               __nv_bfloat16* __restrict__ dx_out, float* __restrict__ dgamma, int rows,
                                                               int D, long ldx, long lddy, long ldr, long ldo, float eps,
                                                               const float* __restrict__ mod_scale, long ld_mod, int rows_per_batch) {
  extern __shared__ float s_dg[];  // [D]
  griddep_launch();
  griddep_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nchunk = D / 8;
  for (int i = threadIdx.x; i < D; i += blockDim.x) s_dg[i] = 0.f;
  __syncthreads();
  float dg[MAXC][8];
#pragma unroll
  for (int c = 0; c < MAXC; ++c)
#pragma unroll
    for (int j = 0; j < 8; ++j) dg[c][j] = 0.f;

  const int bidx = MOD ? blockIdx.y : 0;
  const int row_begin = MOD ? bidx * rows_per_batch : 0;
  const int row_end = MOD ? row_begin + rows_per_batch : rows;
  const float* ms = MOD ? mod_scale + static_cast<long>(bidx) * ld_mod : nullptr;
  if (MOD) dgamma += static_cast<long>(bidx) * D;
  for (int row = row_begin + blockIdx.x * 8 + warp; row < row_end; row += gridDim.x * 8) {
    const __nv_bfloat16* xr = x + static_cast<long>(row) * ldx;
    const __nv_bfloat16* dr = dy + static_cast<long>(row) * lddy;
    // x and dy stay packed (bf16x2) in registers: 2 x MAXC x 4 words instead of 2 x MAXC x 8 floats (ncu r1: the fp32 version
    // needed 238 registers => 8 warps/SM => 1.3 TB/s)
    uint4 xp[MAXC], dp[MAXC];
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      const int ch = lane + c * 32;
      if (ch < nchunk) {
        xp[c] = __ldg(reinterpret_cast<const uint4*>(xr) + ch);
        dp[c] = __ldg(reinterpret_cast<const uint4*>(dr) + ch);
        const uint32_t uw[4] = {xp[c].x, xp[c].y, xp[c].z, xp[c].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float2 f = unpack_bf16(uw[j]); sum += f.x + f.y; }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / D;
    float sq = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      if (lane + c * 32 < nchunk) {
        const uint32_t uw[4] = {xp[c].x, xp[c].y, xp[c].z, xp[c].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float2 f = unpack_bf16(uw[j]); const float d0 = f.x - mean, d1 = f.y - mean; sq += d0 * d0 + d1 * d1; }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq / D + eps);
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      const int ch = lane + c * 32;
      if (ch < nchunk) {
        const uint32_t uw[4] = {xp[c].x, xp[c].y, xp[c].z, xp[c].w}, ww[4] = {dp[c].x, dp[c].y, dp[c].z, dp[c].w};
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma) + ch * 2), g1 = __ldg(reinterpret_cast<const float4*>(gamma) + ch * 2 + 1);
        float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        if (MOD) {
          const float4 s0 = __ldg(reinterpret_cast<const float4*>(ms) + ch * 2), s1 = __ldg(reinterpret_cast<const float4*>(ms) + ch * 2 + 1);
          const float sm[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
          for (int j = 0; j < 8; ++j) gm[j] *= 1.f + sm[j];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = unpack_bf16(uw[j]), d = unpack_bf16(ww[j]);
          const float xh0 = (f.x - mean) * rstd, xh1 = (f.y - mean) * rstd;
          dg[c][2 * j] += d.x * xh0; dg[c][2 * j + 1] += d.y * xh1;
          const float q0 = d.x * gm[2 * j], q1 = d.y * gm[2 * j + 1];
          sg += q0 + q1; sgx += q0 * xh0 + q1 * xh1;
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
        const uint32_t uw[4] = {xp[c].x, xp[c].y, xp[c].z, xp[c].w}, ww[4] = {dp[c].x, dp[c].y, dp[c].z, dp[c].w};
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma) + ch * 2), g1 = __ldg(reinterpret_cast<const float4*>(gamma) + ch * 2 + 1);
        float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        if (MOD) {
          const float4 s0 = __ldg(reinterpret_cast<const float4*>(ms) + ch * 2), s1 = __ldg(reinterpret_cast<const float4*>(ms) + ch * 2 + 1);
          const float sm[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
          for (int j = 0; j < 8; ++j) gm[j] *= 1.f + sm[j];
        }
        uint4 ru = make_uint4(0, 0, 0, 0);
        if (rr) ru = __ldg(reinterpret_cast<const uint4*>(rr) + ch);
        const uint32_t rw[4] = {ru.x, ru.y, ru.z, ru.w};
        uint32_t ow[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = unpack_bf16(uw[j]), d = unpack_bf16(ww[j]), rs = unpack_bf16(rw[j]);
          const float xh0 = (f.x - mean) * rstd, xh1 = (f.y - mean) * rstd;
          const float o0 = rstd * (d.x * gm[2 * j] - mg - xh0 * mgx) + rs.x;
          const float o1 = rstd * (d.y * gm[2 * j + 1] - mg - xh1 * mgx) + rs.y;
          ow[j] = pack_bf16(o0, o1);
        }
        reinterpret_cast<uint4*>(orow)[ch] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
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
  griddep_launch();
  griddep_wait();
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

// The same for narrow planes (N <= 128 columns, e.g. the 64-channel discriminator feature maps: 4 M rows x 64): the kernel above keeps one warp
// per block busy with 4-byte loads (245 us for a 0.56 GB plane).  Here N/8 lanes share a row (16-byte loads), a block walks 256/(N/8) rows per
// iteration, partial sums meet in shared memory.
__global__ void __launch_bounds__(256) colsum_narrow_kernel(const __nv_bfloat16* __restrict__ dy, float* __restrict__ out, int M, int N) {
  __shared__ float red[128];
  griddep_launch();
  griddep_wait();
  const int lpr = N >> 3;                       // lanes per row
  const int rows_per_iter = 256 / lpr;
  const int sub = threadIdx.x / lpr, c8 = threadIdx.x % lpr;
  if (threadIdx.x < 128) red[threadIdx.x] = 0.f;
  __syncthreads();
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (sub < rows_per_iter) {
    for (long r = static_cast<long>(blockIdx.x) * rows_per_iter + sub; r < M; r += static_cast<long>(gridDim.x) * rows_per_iter) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(dy + r * N) + c8);
      const float2 f0 = unpack_bf16(u.x), f1 = unpack_bf16(u.y), f2 = unpack_bf16(u.z), f3 = unpack_bf16(u.w);
      acc[0] += f0.x; acc[1] += f0.y; acc[2] += f1.x; acc[3] += f1.y; acc[4] += f2.x; acc[5] += f2.y; acc[6] += f3.x; acc[7] += f3.y;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&red[c8 * 8 + j], acc[j]);
  }
  __syncthreads();
  if (threadIdx.x < N) atomicAdd(out + threadIdx.x, red[threadIdx.x]);
}

// adaLN gate backward (transformer.py:690-701: x = x + branch * sigmoid(1 - gate_b)):  dbranch = dh * g_b,  dg_b += sum_n dh * branch.
// grid (row blocks, B); thread = 8 channels; the per-batch column sums stay in registers until one atomic per channel per block.
__global__ void __launch_bounds__(256) gate_bwd_kernel(const __nv_bfloat16* __restrict__ dh, long ldh, const __nv_bfloat16* __restrict__ branch,
                                                       long ldb, const float* __restrict__ gate, __nv_bfloat16* __restrict__ dbranch, long ldo,
                                                       float* __restrict__ dgate, int rows_per_batch, int D) {
  griddep_launch();
  griddep_wait();
  const int b = blockIdx.y;
  const int cg = threadIdx.x;
  if (cg * 8 >= D) return;
  const float4 q0 = __ldg(reinterpret_cast<const float4*>(gate + static_cast<long>(b) * D) + cg * 2);
  const float4 q1 = __ldg(reinterpret_cast<const float4*>(gate + static_cast<long>(b) * D) + cg * 2 + 1);
  const float gv[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  const long base = static_cast<long>(b) * rows_per_batch;
  for (int r = blockIdx.x; r < rows_per_batch; r += gridDim.x) {
    const uint4 ud = __ldg(reinterpret_cast<const uint4*>(dh + (base + r) * ldh) + cg);
    const uint4 ub = __ldg(reinterpret_cast<const uint4*>(branch + (base + r) * ldb) + cg);
    const uint32_t dw[4] = {ud.x, ud.y, ud.z, ud.w}, bw[4] = {ub.x, ub.y, ub.z, ub.w};
    uint32_t ow[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 d = unpack_bf16(dw[j]), br = unpack_bf16(bw[j]);
      acc[2 * j] = fmaf(d.x, br.x, acc[2 * j]);
      acc[2 * j + 1] = fmaf(d.y, br.y, acc[2 * j + 1]);
      ow[j] = pack_bf16(d.x * gv[2 * j], d.y * gv[2 * j + 1]);
    }
    reinterpret_cast<uint4*>(dbranch + (base + r) * ldo)[cg] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) atomicAdd(dgate + static_cast<long>(b) * D + cg * 8 + j, acc[j]);
}

}  // namespace b200sat

using namespace b200sat;

extern "C" int b200sat_layernorm_mod_bwd(const void* x, long ldx, const void* dy, long lddy, const float* gamma, const float* mod_scale,
                                         long ld_mod, int rows_per_batch, const void* dres, long ldr, void* dx_out, long ldo, float* dp,
                                         int rows, int D, float eps, void* stream) {
  if (!x || !dy || !gamma || !mod_scale || !dx_out || !dp || rows <= 0 || D <= 0 || rows_per_batch <= 0 || rows % rows_per_batch) {
    set_last_error("layernorm_mod_bwd: bad arguments"); return B200SAT_EINVAL;
  }
  if (D % 8 || ldx % 8 || lddy % 8 || ldo % 8 || (dres && ldr % 8) || ld_mod % 4) { set_last_error("layernorm_mod_bwd: D and leading dims must be multiples of 8"); return B200SAT_EINVAL; }
  if (D > 2048) { set_last_error("layernorm_mod_bwd: D > 2048 not implemented"); return B200SAT_EUNSUPPORTED; }
  const int batches = rows / rows_per_batch;
  int gx = (rows_per_batch + 7) / 8;
  const int cap = (2 * num_sms() + batches - 1) / batches;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  dim3 grid(gx, batches);
  if (D <= 1536)
    B200SAT_CHECK_CUDA(launch_k(layernorm_bwd_kernel<6, true>, grid, dim3(256), D * sizeof(float), static_cast<cudaStream_t>(stream), 1,
        static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(dy), gamma, static_cast<const __nv_bfloat16*>(dres),
        static_cast<__nv_bfloat16*>(dx_out), dp, rows, D, ldx, lddy, ldr, ldo, eps, mod_scale, ld_mod, rows_per_batch));
  else
    B200SAT_CHECK_CUDA(launch_k(layernorm_bwd_kernel<8, true>, grid, dim3(256), D * sizeof(float), static_cast<cudaStream_t>(stream), 1,
        static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(dy), gamma, static_cast<const __nv_bfloat16*>(dres),
        static_cast<__nv_bfloat16*>(dx_out), dp, rows, D, ldx, lddy, ldr, ldo, eps, mod_scale, ld_mod, rows_per_batch));
  B200SAT_CHECK_CUDA(cudaGetLastError());
  return B200SAT_OK;
}

extern "C" int b200sat_gate_bwd(const void* dh, long ldh, const void* branch, long ldb, const float* gate, void* dbranch, long ldo, float* dgate,
                                int rows_per_batch, int batches, int D, void* stream) {
  if (!dh || !branch || !gate || !dbranch || !dgate || rows_per_batch <= 0 || batches <= 0 || D <= 0) { set_last_error("gate_bwd: bad arguments"); return B200SAT_EINVAL; }
  if (D % 8 || D > 2048 || ldh % 8 || ldb % 8 || ldo % 8) { set_last_error("gate_bwd: D <= 2048, D and leading dims multiples of 8"); return B200SAT_EUNSUPPORTED; }
  const int threads = ((D / 8 + 31) / 32) * 32;
  int gx = rows_per_batch;
  const int cap = (8 * num_sms() + batches - 1) / batches;
  if (gx > cap) gx = cap;
  B200SAT_CHECK_CUDA(launch_k(gate_bwd_kernel, dim3(gx, batches), dim3(threads), 0, static_cast<cudaStream_t>(stream), 1,
      static_cast<const __nv_bfloat16*>(dh), ldh, static_cast<const __nv_bfloat16*>(branch), ldb, gate, static_cast<__nv_bfloat16*>(dbranch), ldo,
      dgate, rows_per_batch, D));
  B200SAT_CHECK_CUDA(cudaGetLastError());
  return B200SAT_OK;
}

extern "C" int b200sat_layernorm_bwd(const void* x, long ldx, const void* dy, long lddy, const float* gamma, const void* dres, long ldr,
                                     void* dx_out, long ldo, float* dgamma, int rows, int D, float eps, void* stream) {
  if (!x || !dy || !gamma || !dx_out || rows <= 0 || D <= 0) { set_last_error("layernorm_bwd: bad arguments"); return B200SAT_EINVAL; }
  if (D % 8 || ldx % 8 || lddy % 8 || ldo % 8 || (dres && ldr % 8)) { set_last_error("layernorm_bwd: D and leading dims must be multiples of 8"); return B200SAT_EINVAL; }
  if (D > 2048) { set_last_error("layernorm_bwd: D > 2048 not implemented"); return B200SAT_EUNSUPPORTED; }
  int grid = (rows + 7) / 8;
  const int cap = 2 * num_sms();
  if (grid > cap) grid = cap;
  if (D <= 1536)
    B200SAT_CHECK_CUDA(launch_k(layernorm_bwd_kernel<6, false>, dim3(grid), dim3(256), D * sizeof(float), static_cast<cudaStream_t>(stream), 1, 
        static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(dy), gamma, static_cast<const __nv_bfloat16*>(dres),
        static_cast<__nv_bfloat16*>(dx_out), dgamma, rows, D, ldx, lddy, ldr, ldo, eps, static_cast<const float*>(nullptr), 0L, 0));
  else
    B200SAT_CHECK_CUDA(launch_k(layernorm_bwd_kernel<8, false>, dim3(grid), dim3(256), D * sizeof(float), static_cast<cudaStream_t>(stream), 1, 
        static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(dy), gamma, static_cast<const __nv_bfloat16*>(dres),
        static_cast<__nv_bfloat16*>(dx_out), dgamma, rows, D, ldx, lddy, ldr, ldo, eps, static_cast<const float*>(nullptr), 0L, 0));
  B200SAT_CHECK_CUDA(cudaGetLastError());
  return B200SAT_OK;
}

extern "C" int b200sat_colsum(const void* dy, long ld, float* out, int M, int N, void* stream) {
  if (!dy || !out || M <= 0 || N <= 0 || (N % 2) || (ld % 2)) { set_last_error("colsum: bad arguments (N, ld even)"); return B200SAT_EINVAL; }
  // enough row chunks to fill the machine even for narrow matrices
  if (N <= 128 && (N & 7) == 0 && ld == N && (256 % (N >> 3)) == 0) {
    const int rows_per_iter = 256 / (N >> 3);
    long g = (static_cast<long>(M) + rows_per_iter * 8 - 1) / (rows_per_iter * 8);      // >= 8 iterations per block before adding blocks
    const long cap = static_cast<long>(num_sms()) * 8;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    B200SAT_CHECK_CUDA(launch_k(colsum_narrow_kernel, dim3(static_cast<unsigned>(g)), dim3(256), 0, static_cast<cudaStream_t>(stream), 1,
                                static_cast<const __nv_bfloat16*>(dy), out, M, N));
    B200SAT_CHECK_CUDA(cudaGetLastError());
    return B200SAT_OK;
  }
  int rpb = 256;
  while (rpb > 32 && static_cast<long>((N + 511) / 512) * ((M + rpb - 1) / rpb) < 2L * num_sms()) rpb >>= 1;
  dim3 grid((N + 511) / 512, (M + rpb - 1) / rpb);
  B200SAT_CHECK_CUDA(launch_k(colsum_kernel, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(stream), 1, static_cast<const __nv_bfloat16*>(dy), ld, out, M, N, rpb));
  B200SAT_CHECK_CUDA(cudaGetLastError());
  return B200SAT_OK;
}

// Microbenchmark: tcgen05.ld (TMEM -> registers) throughput per SM on sm_100a for the 32x32b shape at .x16 / .x32 / .x64 / .x128,
// with 4 / 8 / 16 reading warps per CTA (one CTA per SM) and 1 or 2 loads in flight per warp before tcgen05.wait::ld.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_ld_bench tools/micro/tmem_ld_bench.cu && ./tmem_ld_bench
// Used to decide what bounds the attention kernels: every 128 x 64 score tile is read out of TMEM once per kernel.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

template <int X>
__device__ __forceinline__ uint32_t ld(uint32_t taddr);
#define LD_BODY(X, ...)                                                                                                     \
  template <>                                                                                                               \
  __device__ __forceinline__ uint32_t ld<X>(uint32_t taddr) {                                                               \
    uint32_t v[X];                                                                                                          \
    __VA_ARGS__                                                                                                             \
    uint32_t a = 0;                                                                                                         \
    for (int i = 0; i < X; ++i) a ^= v[i];                                                                                  \
    return a;                                                                                                               \
  }
#define R4(b) "%" #b ", %" #b "+1"
LD_BODY(16, asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                           "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                         : "r"(taddr)
                         : "memory");)
LD_BODY(32, asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                           "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                           "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                           "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                         : "r"(taddr)
                         : "memory");)
// x64 = two x32 loads issued back to back without a wait in between (same registers count as one x64)
template <>
__device__ __forceinline__ uint32_t ld<64>(uint32_t taddr) { return ld<32>(taddr) ^ ld<32>(taddr + 32); }

__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

template <int X, int DEPTH>
__global__ void __launch_bounds__(576, 1) bench(uint32_t* out, long long* clk, int reps, int nwarps) {
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tptr)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tptr;
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = 0, t1 = 0;
  if (warp >= 2 && warp < 2 + nwarps) {
    t0 = clock64();
    const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const int span = X * DEPTH;
    for (int r = 0; r < reps; ++r) {
      const uint32_t col = (static_cast<uint32_t>(r) * span + (warp >> 2) * 64) & (511 & ~(span - 1));
#pragma unroll
      for (int d = 0; d < DEPTH; ++d) acc ^= ld<X>(base + lane_off + ((col + d * X) & 511 & ~(X - 1)));
      wait_ld();
    }
    t1 = clock64();
    if (threadIdx.x == 64) clk[blockIdx.x] = t1 - t0 + (acc == 0x12345u);   // warp 2 lane 0: cycles for its own `reps` loads (acc keeps the loads live)
  }
  __syncthreads();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(512) : "memory");
}

template <int X, int DEPTH>
void run(int nwarps) {
  const int reps = 4096, grid = 148;
  uint32_t* out; long long* clk;
  cudaMalloc(&out, grid * 576 * 4); cudaMalloc(&clk, grid * 8);
  bench<X, DEPTH><<<grid, 576>>>(out, clk, 16, nwarps);
  cudaError_t e0 = cudaDeviceSynchronize();
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  bench<X, DEPTH><<<grid, 576>>>(out, clk, reps, nwarps);
  cudaEventRecord(b);
  cudaError_t e1 = cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, a, b);
  long long h[148];
  cudaError_t e2 = cudaMemcpy(h, clk, grid * 8, cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < grid; ++i) avg += h[i];
  avg /= grid;
  const double bytes = double(nwarps) * reps * X * DEPTH * 32 * 4;
  // avg = cycles warp 2 spent in its loop (clock64 inside the reading warp); ms = the whole launch by CUDA events (cross-check)
  printf("32x32b.x%-3d loads-in-flight %d  warps %2d : %7.1f B/clk/SM  (%.0f clk per warp iteration of %d B; kernel %.3f ms => %.1f B/ns/SM)  err=%s/%s/%s\n", X,
         DEPTH, nwarps, bytes / avg, avg / reps, X * DEPTH * 128, ms, bytes / (ms * 1e6), cudaGetErrorString(e0), cudaGetErrorString(e1),
         cudaGetErrorString(e2));
  cudaFree(out); cudaFree(clk);
}

int main() {
  for (int nw : {4, 8, 16}) {
    run<16, 1>(nw); run<32, 1>(nw); run<32, 2>(nw); run<64, 1>(nw);
  }
  return 0;
}

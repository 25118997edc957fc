// b200sat — flash attention backward on tcgen05 (non-causal, dh = 64, GQA-aware, ragged tails, optional inverse RoPE).
//
// Backward of the attention the reference reaches through flash_attn_func / F.scaled_dot_product_attention
// (stable_audio_tools/models/transformer.py:406-441) together with the backward of apply_rotary_pos_emb (:154-174) on
// dQ / dK, so that the gradient buffers land directly in the layout of the fused qkv projection.
//
//   delta[q]  = sum_d O[q,d] dO[q,d]
//   P = exp(S*scale - lse),  dP = dO V^T,  dS = P o (dP - delta)
//   dV = P^T dO,   dK = scale * dS^T Q,   dQ = scale * dS K
//
// Two kernels, each the forward kernel's structure (TMA producer warp, one MMA-issuing thread, four softmax warps whose
// threads own one row of the score tile in TMEM):
//   attention_bwd_dkv : CTA = (128 keys, kv head, batch); loops over the q heads of the GQA group x q tiles; S^T and dP^T are
//                       produced transposed (A = K / V) so P^T, dS^T are written row-wise as A operands; dV, dK accumulate
//                       in TMEM across the whole loop.  Q and dO tiles are consumed twice from the same shared-memory bytes:
//                       as K-major B operands (S^T, dP^T) and as MN-major B operands (dV += P^T dO, dK += dS^T Q).
//   attention_bwd_dq  : CTA = (128 queries, head, batch); loops over key tiles; dQ accumulates in TMEM.
#include "common.cuh"
#include <cstring>
#include <cstdlib>

namespace b200sat {

__device__ __forceinline__ float bw_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

constexpr int BW_T = 128 * 64 * 2;  // one 128 x 64 bf16 tile = 16 KB

template <int CW>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&v)[CW]);
template <>
__device__ __forceinline__ void tmem_ld_cols<32>(uint32_t taddr, uint32_t (&v)[32]) { tmem_ld_32x32(taddr, v); }
template <>
__device__ __forceinline__ void tmem_ld_cols<16>(uint32_t taddr, uint32_t (&v)[16]) { tmem_ld_32x16(taddr, v); }

struct AttnBwdParams {
  CUtensorMap tmQ, tmK, tmV, tmdO;
  const float* lse;    // [B, Hq, Nq]
  const float* delta;  // [B, Hq, Nq]
  const float* nstat;  // v3 kernels: [2][B, Hq, Npad] = -lse * log2(e) (padding -inf) and -delta (padding 0), Npad = roundup(Nq, 64)
  int Npad;
  __nv_bfloat16* dQ;
  __nv_bfloat16* dK;
  __nv_bfloat16* dV;
  long dq_bs, dq_ss, dq_hs, dk_bs, dk_ss, dk_hs, dv_bs, dv_ss, dv_hs;
  const float* rope_cos;  // [N, 16] or null: inverse rotation applied to dQ / dK rows (position = sequence index)
  const float* rope_sin;
  int B, Hq, Hkv, Nq, Nk;
  float scale, scale_log2;
};

// write 32 fp32 values (already scaled) as bf16 into the swizzled 128-column A-operand tile: row r, columns [c*32, c*32+32)
__device__ __forceinline__ void store_operand_chunk(uint8_t* tile_row, int sw, int c, const float (&v)[32]) {
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int chunk = c * 4 + t;
    const int kb = chunk >> 3, cc = chunk & 7;
    uint4 u;
    u.x = pack_bf16(v[8 * t + 0], v[8 * t + 1]);
    u.y = pack_bf16(v[8 * t + 2], v[8 * t + 3]);
    u.z = pack_bf16(v[8 * t + 4], v[8 * t + 5]);
    u.w = pack_bf16(v[8 * t + 6], v[8 * t + 7]);
    *reinterpret_cast<uint4*>(tile_row + kb * BW_T + ((cc ^ sw) << 4)) = u;
  }
}

// rows of a [*, 64] gradient: optional inverse RoPE on dims [0,32), scale, bf16 store (128 bytes)
__device__ __forceinline__ void store_grad_row(__nv_bfloat16* dst, float (&g)[64], float scale, const float* cs, const float* sn) {
#pragma unroll
  for (int i = 0; i < 64; ++i) g[i] *= scale;
  if (cs) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float c_ = __ldg(cs + i), s_ = __ldg(sn + i);
      const float a = g[i], b = g[i + 16];
      g[i] = a * c_ + b * s_;
      g[i + 16] = b * c_ - a * s_;
    }
  }
  uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    uint4 u;
    u.x = pack_bf16(g[8 * i + 0], g[8 * i + 1]);
    u.y = pack_bf16(g[8 * i + 2], g[8 * i + 3]);
    u.z = pack_bf16(g[8 * i + 4], g[8 * i + 5]);
    u.w = pack_bf16(g[8 * i + 6], g[8 * i + 7]);
    d4[i] = u;
  }
}

// 32 consecutive dims of a gradient row (dims [0,32) carry the rotary pairs (i, i+16) when cs != null)
__device__ __forceinline__ void store_grad_half(__nv_bfloat16* dst, float (&g)[32], float scale, const float* cs, const float* sn) {
#pragma unroll
  for (int i = 0; i < 32; ++i) g[i] *= scale;
  if (cs) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float c_ = __ldg(cs + i), s_ = __ldg(sn + i);
      const float a = g[i], b = g[i + 16];
      g[i] = a * c_ + b * s_;
      g[i + 16] = b * c_ - a * s_;
    }
  }
  uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 u;
    u.x = pack_bf16(g[8 * i + 0], g[8 * i + 1]);
    u.y = pack_bf16(g[8 * i + 2], g[8 * i + 3]);
    u.z = pack_bf16(g[8 * i + 4], g[8 * i + 5]);
    u.w = pack_bf16(g[8 * i + 6], g[8 * i + 7]);
    d4[i] = u;
  }
}

// ------------------------------------------------------------------------------------------------------------
__global__ void attn_delta_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ d_o, const float* __restrict__ lse,
                                  float* __restrict__ delta, float* __restrict__ nstat, int B, int H, int N, int Npad, long o_bs, long o_ss,
                                  long o_hs, long d_bs, long d_ss, long d_hs) {
  griddep_launch();
  griddep_wait();
  const long ip = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;   // index into the padded [B, H, Npad] layout
  const long total = static_cast<long>(B) * H * Npad;
  if (ip >= total) return;
  const int n = ip % Npad;
  const int h = (ip / Npad) % H;
  const int b = ip / (static_cast<long>(Npad) * H);
  if (n >= N) {
    nstat[ip] = -INFINITY;
    nstat[total + ip] = 0.f;
    return;
  }
  const long i = (static_cast<long>(b) * H + h) * N + n;
  const uint4* po = reinterpret_cast<const uint4*>(o + b * o_bs + n * o_ss + h * o_hs);
  const uint4* pd = reinterpret_cast<const uint4*>(d_o + b * d_bs + n * d_ss + h * d_hs);
  float acc = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const uint4 a = __ldg(po + k), c = __ldg(pd + k);
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, cw[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 x = unpack_bf16(aw[j]), y = unpack_bf16(cw[j]);
      acc += x.x * y.x + x.y * y.y;
    }
  }
  delta[i] = acc;  // layout [B, H, N]
  nstat[ip] = -lse[i] * 1.4426950408889634f;
  nstat[total + ip] = -acc;
}

// ------------------------------------------------------------------------------------------------------------
// Pipelined v2 of both kernels (after the ncu/launch-list pass on v1: one score tile in flight per CTA left the tensor pipe and
// the softmax warps waiting on each other; 327 + 258 us for the B=8 self-attention backward): 64-wide inner tiles, score tiles
// (S, dP) DOUBLE-BUFFERED in TMEM so the MMA thread runs two tiles ahead, operand tiles (P^T / dS^T / dS) double-buffered in
// shared memory so the accumulate MMAs of tile j overlap the softmax math of tile j+1.  Accumulators never leave TMEM.
constexpr int BW_HT = 64 * 64 * 2;  // a 64 x 64 bf16 tile = 8 KB
// Inner-tile TMA rings.  A stage is held from the issue of a tile's score MMAs until its accumulate MMAs retire, i.e. for BW_NB = 3
// iterations, so a ring of S stages gives the NEXT load only S - 3 iterations of lead.  Round 2 (ncu source view: the softmax warps'
// top stall was the spin on s_full, the MMA warp's the spin on the ring's full barrier): with 4 stages the lead was one iteration
// ~ 1 us ~ the L2 -> shared-memory latency of a TMA tile, so every iteration waited for its load.  Now 6 (dK/dV kernel, 224 KB of
// shared memory) and 8 (dQ kernel) stages.
constexpr int DKV_STAGES = 6;
constexpr int DQ_STAGES = 8;
constexpr int BW_NB = 3;       // score-tile buffers in TMEM and operand-tile buffers in shared memory (MMA thread runs BW_NB tiles ahead)
constexpr int DKV_SMEM = 2 * BW_T /*K,V*/ + DKV_STAGES * 2 * BW_HT /*(Q,dO) ring*/ + BW_NB * BW_T /*P^T*/ + BW_NB * BW_T /*dS^T*/ + 2 * BW_NB * 64 * 4 + 256;

// SW = number of softmax warps: 8 (two threads per score row, 32 columns each) or 16 (four threads per row, 16 columns each).
// Round 2: the kernels hold ONE CTA per SM (192 KB of tiles, all 512 TMEM columns), so 8 warps = 2 per scheduler could not hide the
// tcgen05.ld / MUFU / shared-memory latencies of their own tile (measured ~1500 clk per 128x64 tile against a 512-clk MUFU bound).
template <int SW>
__global__ void __launch_bounds__(64 + SW * 32, 1) attention_bwd_dkv_tcgen05(const __grid_constant__ AttnBwdParams p) {
  constexpr int CW = 256 / SW;   // score columns per thread
  constexpr int NT = SW * 32;    // softmax threads
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sK = smem;
  uint8_t* sV = smem + BW_T;
  uint8_t* sRing = smem + 2 * BW_T;                         // stage s: Q tile (64 rows) at + s*2*BW_HT, dO tile at + BW_HT
  uint8_t* sPT = sRing + DKV_STAGES * 2 * BW_HT;             // BW_NB buffers of [128 keys x 64 queries]
  uint8_t* sdST = sPT + BW_NB * BW_T;
  float* s_lse = reinterpret_cast<float*>(sdST + BW_NB * BW_T); // [BW_NB][64] (pre-multiplied by log2 e)
  float* s_delta = s_lse + BW_NB * 64;                          // [BW_NB][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_delta + BW_NB * 64);
  uint64_t* kv_full = bars + 0;
  uint64_t* qdo_full = bars + 1;                      // [DKV_STAGES]
  uint64_t* qdo_empty = qdo_full + DKV_STAGES;        // [DKV_STAGES]
  uint64_t* s_full = qdo_empty + DKV_STAGES;          // [BW_NB]
  uint64_t* p_full = s_full + BW_NB;                  // [BW_NB]
  uint64_t* acc_free = p_full + BW_NB;                // [BW_NB] accumulate MMAs of a tile retired: its P^T / dS^T buffer is reusable
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(acc_free + BW_NB);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k0 = blockIdx.x * 128;
  const int hk = blockIdx.y;
  const int b = blockIdx.z;
  const int G = p.Hq / p.Hkv;
  const int nq = (p.Nq + 63) / 64;
  const int iters = G * nq;

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) __trap();
    mbar_init(kv_full, 1);
    for (int i = 0; i < DKV_STAGES; ++i) { mbar_init(&qdo_full[i], 1); mbar_init(&qdo_empty[i], 1); }
    for (int i = 0; i < BW_NB; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_full[i], NT); mbar_init(&acc_free[i], 1); }
    fence_barrier_init();
  }
  griddep_launch();
  if (warp == 1) { tmem_alloc(tmem_ptr_smem, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  griddep_wait();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // columns: 3 S^T buffers [0,192); 3 dP^T buffers [192,384); dV [384,448); dK [448,512)
  const uint32_t tm_dV = tmem_base + 384, tm_dK = tmem_base + 448;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(kv_full, 2 * BW_T);
      tma_load_4d(sK, &p.tmK, kv_full, 0, hk, k0, b);
      tma_load_4d(sV, &p.tmV, kv_full, 0, hk, k0, b);
      for (int it = 0; it < iters; ++it) {
        const int s = it % DKV_STAGES;
        const int h = hk * G + it / nq, qt = it % nq;
        mbar_wait(&qdo_empty[s], ((it / DKV_STAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(&qdo_full[s], 2 * BW_HT);
        tma_load_4d(sRing + s * 2 * BW_HT, &p.tmQ, &qdo_full[s], 0, h, qt * 64, b);
        tma_load_4d(sRing + s * 2 * BW_HT + BW_HT, &p.tmdO, &qdo_full[s], 0, h, qt * 64, b);
      }
    }
  } else if (warp == 1) {
    // MMA issue by the whole warp in uniform control flow, one elected lane issues (umma_bf16_lo, common.cuh): ~2 instructions per MMA
    // instead of ~18 - the 32-cycle 128x64x16 MMAs were starved by descriptor arithmetic in the issuing thread
    const uint32_t leader = elect_one() ? 1u : 0u;
    constexpr uint32_t id_s = make_idesc_bf16(128, 64, 0, 0);
    constexpr uint32_t id_acc = make_idesc_bf16(128, 64, 0, 1);
    const uint32_t loK = desc_lo_kmajor(smem_u32(sK)), loV = desc_lo_kmajor(smem_u32(sV));
    const uint32_t loQ0 = desc_lo_kmajor(smem_u32(sRing)), lodO0 = desc_lo_kmajor(smem_u32(sRing) + BW_HT);
    const uint32_t loQm0 = desc_lo_mnmajor(smem_u32(sRing), 1024), lodOm0 = desc_lo_mnmajor(smem_u32(sRing) + BW_HT, 1024);
    const uint32_t loPT0 = desc_lo_kmajor(smem_u32(sPT)), lodST0 = desc_lo_kmajor(smem_u32(sdST));
    constexpr uint32_t kStageStep = (2 * BW_HT) >> 4, kBufStep = BW_T >> 4;
    auto issue_scores = [&](int it) {
      const int s = it % DKV_STAGES;
      mbar_wait(&qdo_full[s], (it / DKV_STAGES) & 1);
      tc_fence_after();
      const uint32_t loQ = loQ0 + s * kStageStep, lodO = lodO0 + s * kStageStep;
      const uint32_t tS = tmem_base + (it % BW_NB) * 64, tP = tmem_base + 192 + (it % BW_NB) * 64;
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16_lo(tS, loK + 2 * k, loQ + 2 * k, id_s, k != 0, leader);     // S^T[keys, q] = K Q^T
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16_lo(tP, loV + 2 * k, lodO + 2 * k, id_s, k != 0, leader);    // dP^T[keys, q] = V dO^T
      umma_commit_if(&s_full[it % BW_NB], leader);
    };
    mbar_wait(kv_full, 0);
    for (int i = 0; i < BW_NB && i < iters; ++i) issue_scores(i);
    for (int it = 0; it < iters; ++it) {
      const int s = it % DKV_STAGES;
      mbar_wait(&p_full[it % BW_NB], (it / BW_NB) & 1);
      tc_fence_after();
      const uint32_t loQm = loQm0 + s * kStageStep, lodOm = lodOm0 + s * kStageStep;
      const uint32_t loPT = loPT0 + (it % BW_NB) * kBufStep, lodST = lodST0 + (it % BW_NB) * kBufStep;
#pragma unroll
      for (int k = 0; k < 4; ++k)   // dV[keys, d] += P^T dO   (dO tile re-read MN-major: rows = queries = MMA K)
        umma_bf16_lo(tm_dV, loPT + 2 * k, lodOm + 128 * k, id_acc, (it | k) != 0, leader);
#pragma unroll
      for (int k = 0; k < 4; ++k)   // dK[keys, d] += dS^T Q
        umma_bf16_lo(tm_dK, lodST + 2 * k, loQm + 128 * k, id_acc, (it | k) != 0, leader);
      umma_commit_if(&qdo_empty[s], leader);
      umma_commit_if(&acc_free[it % BW_NB], leader);
      if (it + BW_NB < iters) issue_scores(it + BW_NB);   // its TMEM buffers were drained before p_full(it) completed
    }
  } else {
    // SW warps: SW/4 threads per key row, each owning CW of the 64 query columns of a tile
    const int quarter = warp & 3;
    const int sub = (warp - 2) >> 2;            // which CW-column slice
    const int r = quarter * 32 + lane;          // key row of this thread
    const int tid = (warp - 2) * 32 + lane;     // 0..NT-1 among the softmax threads
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const bool key_ok = (k0 + r) < p.Nk;
    const int sw = r & 7;
    // column statistics (lse, delta) of the NEXT q tile are fetched one iteration ahead so their global-memory latency is off the
    // critical path (ncu v2: the whole CTA sat at the staging barrier waiting for this load)
    auto fetch_stat = [&](int it) -> float {
      if (tid >= 128 || it >= iters) return 0.f;
      const int hh = hk * G + it / nq, qq = (it % nq) * 64 + (tid & 63);
      const long idx = (static_cast<long>(b) * p.Hq + hh) * p.Nq + qq;
      if (tid < 64) return (qq < p.Nq) ? p.lse[idx] * 1.4426950408889634f : INFINITY;   // out-of-range query => P = 0
      return (qq < p.Nq) ? p.delta[idx] : 0.f;
    };
    float stat_next = fetch_stat(0);
    for (int it = 0; it < iters; ++it) {
      const int buf = it % BW_NB;
      if (tid < 64) s_lse[buf * 64 + tid] = stat_next;
      else if (tid < 128) s_delta[buf * 64 + (tid & 63)] = stat_next;
      stat_next = fetch_stat(it + 1);
      asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
      mbar_wait(&s_full[buf], (it / BW_NB) & 1);
      tc_fence_after();
      uint32_t rs[CW], rp[CW];
      tmem_ld_cols<CW>(tmem_base + lane_off + buf * 64 + sub * CW, rs);
      tmem_ld_cols<CW>(tmem_base + 192 + lane_off + buf * 64 + sub * CW, rp);
      tmem_ld_wait();
      uint32_t pk[CW / 2], dk_[CW / 2];
#pragma unroll
      for (int i = 0; i < CW; i += 2) {
        const float2 l2 = *reinterpret_cast<const float2*>(&s_lse[buf * 64 + sub * CW + i]);
        const float2 d2 = *reinterpret_cast<const float2*>(&s_delta[buf * 64 + sub * CW + i]);
        const float p0 = key_ok ? bw_exp2(fmaf(__uint_as_float(rs[i]), p.scale_log2, -l2.x)) : 0.f;
        const float p1 = key_ok ? bw_exp2(fmaf(__uint_as_float(rs[i + 1]), p.scale_log2, -l2.y)) : 0.f;
        pk[i >> 1] = pack_bf16(p0, p1);
        dk_[i >> 1] = pack_bf16(p0 * (__uint_as_float(rp[i]) - d2.x), p1 * (__uint_as_float(rp[i + 1]) - d2.y));
      }
      if (it >= BW_NB) mbar_wait(&acc_free[buf], ((it / BW_NB) - 1) & 1);   // accumulate MMAs of tile it-BW_NB finished reading this buffer
      uint8_t* pt_row = sPT + buf * BW_T + r * 128;
      uint8_t* ds_row = sdST + buf * BW_T + r * 128;
#pragma unroll
      for (int t = 0; t < CW / 8; ++t) {
        const int ch = sub * (CW / 8) + t;
        *reinterpret_cast<uint4*>(pt_row + ((ch ^ sw) << 4)) = make_uint4(pk[4 * t], pk[4 * t + 1], pk[4 * t + 2], pk[4 * t + 3]);
        *reinterpret_cast<uint4*>(ds_row + ((ch ^ sw) << 4)) = make_uint4(dk_[4 * t], dk_[4 * t + 1], dk_[4 * t + 2], dk_[4 * t + 3]);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(&p_full[buf]);
    }
    mbar_wait(&acc_free[(iters - 1) % BW_NB], ((iters - 1) / BW_NB) & 1);
    tc_fence_after();
    if (sub < 2) {   // read-out: two threads per key row, 32 dims each (the rotary pairs (i, i+16) stay inside one thread)
      const int half = sub;
      float g[32];
      const int krow = k0 + r;
      {
        uint32_t raw[32];
        tmem_ld_32x32(tm_dV + lane_off + half * 32, raw);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) g[i] = __uint_as_float(raw[i]);
      }
      if (key_ok) store_grad_half(p.dV + b * p.dv_bs + krow * p.dv_ss + hk * p.dv_hs + half * 32, g, 1.0f, nullptr, nullptr);
      {
        uint32_t raw[32];
        tmem_ld_32x32(tm_dK + lane_off + half * 32, raw);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) g[i] = __uint_as_float(raw[i]);
      }
      if (key_ok) {
        const bool rot = p.rope_cos != nullptr && half == 0;
        store_grad_half(p.dK + b * p.dk_bs + krow * p.dk_ss + hk * p.dk_hs + half * 32, g, p.scale,
                        rot ? p.rope_cos + krow * 16 : nullptr, rot ? p.rope_sin + krow * 16 : nullptr);
      }
    }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ------------------------------------------------------------------------------------------------------------
constexpr int DQ_SMEM = 2 * BW_T /*Q,dO*/ + DQ_STAGES * 2 * BW_HT /*(K,V) ring*/ + BW_NB * BW_T /*dS*/ + 256;

template <int SW>
__global__ void __launch_bounds__(64 + SW * 32, 1) attention_bwd_dq_tcgen05(const __grid_constant__ AttnBwdParams p) {
  constexpr int CW = 256 / SW;
  constexpr int NT = SW * 32;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sdO = smem + BW_T;
  uint8_t* sRing = smem + 2 * BW_T;                       // stage s: K tile (64 keys) at + s*2*BW_HT, V tile at + BW_HT
  uint8_t* sdS = sRing + DQ_STAGES * 2 * BW_HT;           // 2 buffers of [128 queries x 64 keys]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdS + BW_NB * BW_T);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;                       // [DQ_STAGES]
  uint64_t* kv_empty = kv_full + DQ_STAGES;           // [DQ_STAGES]
  uint64_t* s_full = kv_empty + DQ_STAGES;            // [BW_NB]
  uint64_t* p_full = s_full + BW_NB;                  // [BW_NB]
  uint64_t* acc_free = p_full + BW_NB;                // [BW_NB]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(acc_free + BW_NB);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int hk = h / (p.Hq / p.Hkv);
  const int nkv = (p.Nk + 63) / 64;

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) __trap();
    mbar_init(q_full, 1);
    for (int i = 0; i < DQ_STAGES; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    for (int i = 0; i < BW_NB; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_full[i], NT); mbar_init(&acc_free[i], 1); }
    fence_barrier_init();
  }
  griddep_launch();
  if (warp == 1) { tmem_alloc(tmem_ptr_smem, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  griddep_wait();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // columns: 3 S buffers [0,192); 3 dP buffers [192,384); dQ [384,448)
  const uint32_t tm_dQ = tmem_base + 384;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, 2 * BW_T);
      tma_load_4d(sQ, &p.tmQ, q_full, 0, h, q0, b);
      tma_load_4d(sdO, &p.tmdO, q_full, 0, h, q0, b);
      for (int j = 0; j < nkv; ++j) {
        const int s = j % DQ_STAGES;
        mbar_wait(&kv_empty[s], ((j / DQ_STAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(&kv_full[s], 2 * BW_HT);
        tma_load_4d(sRing + s * 2 * BW_HT, &p.tmK, &kv_full[s], 0, hk, j * 64, b);
        tma_load_4d(sRing + s * 2 * BW_HT + BW_HT, &p.tmV, &kv_full[s], 0, hk, j * 64, b);
      }
    }
  } else if (warp == 1) {
    const uint32_t leader = elect_one() ? 1u : 0u;      // see the dK/dV kernel
    constexpr uint32_t id_s = make_idesc_bf16(128, 64, 0, 0);
    constexpr uint32_t id_acc = make_idesc_bf16(128, 64, 0, 1);
    const uint32_t loQ = desc_lo_kmajor(smem_u32(sQ)), lodO = desc_lo_kmajor(smem_u32(sdO));
    const uint32_t loK0 = desc_lo_kmajor(smem_u32(sRing)), loV0 = desc_lo_kmajor(smem_u32(sRing) + BW_HT);
    const uint32_t loKm0 = desc_lo_mnmajor(smem_u32(sRing), 1024);
    const uint32_t lodS0 = desc_lo_kmajor(smem_u32(sdS));
    constexpr uint32_t kStageStep = (2 * BW_HT) >> 4, kBufStep = BW_T >> 4;
    auto issue_scores = [&](int j) {
      const int s = j % DQ_STAGES;
      mbar_wait(&kv_full[s], (j / DQ_STAGES) & 1);
      tc_fence_after();
      const uint32_t loK = loK0 + s * kStageStep, loV = loV0 + s * kStageStep;
      const uint32_t tS = tmem_base + (j % BW_NB) * 64, tP = tmem_base + 192 + (j % BW_NB) * 64;
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16_lo(tS, loQ + 2 * k, loK + 2 * k, id_s, k != 0, leader);     // S[q, keys] = Q K^T
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16_lo(tP, lodO + 2 * k, loV + 2 * k, id_s, k != 0, leader);    // dP[q, keys] = dO V^T
      umma_commit_if(&s_full[j % BW_NB], leader);
    };
    mbar_wait(q_full, 0);
    for (int i = 0; i < BW_NB && i < nkv; ++i) issue_scores(i);
    for (int j = 0; j < nkv; ++j) {
      const int s = j % DQ_STAGES;
      mbar_wait(&p_full[j % BW_NB], (j / BW_NB) & 1);
      tc_fence_after();
      const uint32_t loKm = loKm0 + s * kStageStep;
      const uint32_t lodS = lodS0 + (j % BW_NB) * kBufStep;
#pragma unroll
      for (int k = 0; k < 4; ++k)   // dQ[q, d] += dS K   (K tile re-read MN-major: rows = keys = MMA K)
        umma_bf16_lo(tm_dQ, lodS + 2 * k, loKm + 128 * k, id_acc, (j | k) != 0, leader);
      umma_commit_if(&kv_empty[s], leader);
      umma_commit_if(&acc_free[j % BW_NB], leader);
      if (j + BW_NB < nkv) issue_scores(j + BW_NB);
    }
  } else {
    const int quarter = warp & 3;
    const int sub = (warp - 2) >> 2;
    const int r = quarter * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const int qrow = q0 + r;
    const bool q_ok = qrow < p.Nq;
    const long sidx = (static_cast<long>(b) * p.Hq + h) * p.Nq + qrow;
    const float lse2 = q_ok ? p.lse[sidx] * 1.4426950408889634f : INFINITY;
    const float dl = q_ok ? p.delta[sidx] : 0.f;
    const int sw = r & 7;
    for (int j = 0; j < nkv; ++j) {
      const int buf = j % BW_NB;
      const int nvalid = p.Nk - j * 64 - sub * CW;
      mbar_wait(&s_full[buf], (j / BW_NB) & 1);
      tc_fence_after();
      uint32_t rs[CW], rp[CW];
      tmem_ld_cols<CW>(tmem_base + lane_off + buf * 64 + sub * CW, rs);
      tmem_ld_cols<CW>(tmem_base + 192 + lane_off + buf * 64 + sub * CW, rp);
      tmem_ld_wait();
      uint32_t dk_[CW / 2];
#pragma unroll
      for (int i = 0; i < CW; i += 2) {
        const float p0 = (i < nvalid) ? bw_exp2(fmaf(__uint_as_float(rs[i]), p.scale_log2, -lse2)) : 0.f;
        const float p1 = (i + 1 < nvalid) ? bw_exp2(fmaf(__uint_as_float(rs[i + 1]), p.scale_log2, -lse2)) : 0.f;
        dk_[i >> 1] = pack_bf16(p0 * (__uint_as_float(rp[i]) - dl), p1 * (__uint_as_float(rp[i + 1]) - dl));
      }
      if (j >= BW_NB) mbar_wait(&acc_free[buf], ((j / BW_NB) - 1) & 1);
      uint8_t* ds_row = sdS + buf * BW_T + r * 128;
#pragma unroll
      for (int t = 0; t < CW / 8; ++t) {
        const int ch = sub * (CW / 8) + t;
        *reinterpret_cast<uint4*>(ds_row + ((ch ^ sw) << 4)) = make_uint4(dk_[4 * t], dk_[4 * t + 1], dk_[4 * t + 2], dk_[4 * t + 3]);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(&p_full[buf]);
    }
    mbar_wait(&acc_free[(nkv - 1) % BW_NB], ((nkv - 1) / BW_NB) & 1);
    tc_fence_after();
    if (sub < 2) {
      const int half = sub;
      float g[32];
      {
        uint32_t raw[32];
        tmem_ld_32x32(tm_dQ + lane_off + half * 32, raw);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) g[i] = __uint_as_float(raw[i]);
      }
      if (q_ok) {
        const bool rot = p.rope_cos != nullptr && half == 0;
        store_grad_half(p.dQ + b * p.dq_bs + qrow * p.dq_ss + h * p.dq_hs + half * 32, g, p.scale,
                        rot ? p.rope_cos + qrow * 16 : nullptr, rot ? p.rope_sin + qrow * 16 : nullptr);
      }
    }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}


// ============================================================================================================
// Round-2 rewrite of both kernels ("v3") along the lines of the forward kernel (attention.cu): the per-tile timeline showed a chain
// of serial barrier / TMEM / MUFU latencies in the softmax warps with too few independent instruction streams per scheduler to hide it
// (the v2 kernels held ONE CTA per SM and took ~3000 cycles per 128 x 64 tile against ~600 of MMA or MUFU work).  v3:
//   * one thread per score row (4 softmax warps, 64 columns per tile), no cross-thread exchange, no CTA-wide named barrier;
//   * two CTAs per SM (256 TMEM columns, <= 113 KB of shared memory each);
//   * score tiles (S, dP) single-buffered and handed back to the MMA warp as soon as they are in registers (s_free), so the next
//     tile's score MMAs run under this tile's exp / dS arithmetic; operand tiles (P^T, dS^T / dS) single-buffered, guarded by acc_done;
//   * one mbarrier arrival per warp; packed fp32 arithmetic; a quarter of the exp2 on the FMA pipe (poly_exp2_x2);
//   * the per-query statistics arrive pre-negated and padded (nstat), so the dK/dV kernel reads them as aligned float4 broadcasts.
constexpr int BW3_POLY = 2;   // of every 8 column pairs, this many take the FMA-pipe exp2

// one pair of score columns: p = exp2(s * scale_log2 - lse2), ds = p * (dp - delta); nl2 / nd2 carry the NEGATED statistics
template <bool POLY>
__device__ __forceinline__ void bw3_pair(uint32_t s0, uint32_t s1, uint32_t dp0, uint32_t dp1, uint64_t sc2, uint64_t nl2, uint64_t nd2, float& p0,
                                         float& p1, float& ds0, float& ds1) {
  const uint64_t x2 = fma_f32x2(pack_f32x2(__uint_as_float(s0), __uint_as_float(s1)), sc2, nl2);
  if (POLY) {
    poly_exp2_x2(x2, p0, p1);
  } else {
    unpack_f32x2(x2, p0, p1);
    p0 = fast_exp2(p0);
    p1 = fast_exp2(p1);
  }
  const uint64_t t2 = add_f32x2(pack_f32x2(__uint_as_float(dp0), __uint_as_float(dp1)), nd2);
  unpack_f32x2(mul_f32x2(pack_f32x2(p0, p1), t2), ds0, ds1);
}

constexpr int DQ3_STAGES = 3;
constexpr int DQ3_SMEM = 2 * BW_T /*Q,dO*/ + DQ3_STAGES * 2 * BW_HT /*(K,V) ring*/ + BW_T /*dS*/ + 256;

__global__ void __launch_bounds__(192, 2) attention_bwd_dq_v3(const __grid_constant__ AttnBwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sdO = smem + BW_T;
  uint8_t* sRing = smem + 2 * BW_T;                       // stage s: K tile (64 keys) at + s*2*BW_HT, V tile at + BW_HT
  uint8_t* sdS = sRing + DQ3_STAGES * 2 * BW_HT;          // [128 queries x 64 keys]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdS + BW_T);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;                       // [DQ3_STAGES]
  uint64_t* kv_empty = kv_full + DQ3_STAGES;          // [DQ3_STAGES]
  uint64_t* s_full = kv_empty + DQ3_STAGES;
  uint64_t* s_free = s_full + 1;
  uint64_t* p_full = s_free + 1;
  uint64_t* acc_done = p_full + 1;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(acc_done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int hk = h / (p.Hq / p.Hkv);
  const int nkv = (p.Nk + 63) / 64;

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) __trap();
    mbar_init(q_full, 1);
    for (int i = 0; i < DQ3_STAGES; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    mbar_init(s_full, 1); mbar_init(s_free, 4); mbar_init(p_full, 4); mbar_init(acc_done, 1);
    fence_barrier_init();
  }
  griddep_launch();
  if (warp == 1) { tmem_alloc(tmem_ptr_smem, 256); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  griddep_wait();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t tS = tmem_base, tP = tmem_base + 64, tm_dQ = tmem_base + 128;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, 2 * BW_T);
      tma_load_4d(sQ, &p.tmQ, q_full, 0, h, q0, b);
      tma_load_4d(sdO, &p.tmdO, q_full, 0, h, q0, b);
      for (int j = 0; j < nkv; ++j) {
        const int s = j % DQ3_STAGES;
        mbar_wait(&kv_empty[s], ((j / DQ3_STAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(&kv_full[s], 2 * BW_HT);
        tma_load_4d(sRing + s * 2 * BW_HT, &p.tmK, &kv_full[s], 0, hk, j * 64, b);
        tma_load_4d(sRing + s * 2 * BW_HT + BW_HT, &p.tmV, &kv_full[s], 0, hk, j * 64, b);
      }
    }
  } else if (warp == 1) {
    const uint32_t leader = elect_one() ? 1u : 0u;
    constexpr uint32_t id_s = make_idesc_bf16(128, 64, 0, 0);
    constexpr uint32_t id_acc = make_idesc_bf16(128, 64, 0, 1);
    const uint32_t loQ = desc_lo_kmajor(smem_u32(sQ)), lodO = desc_lo_kmajor(smem_u32(sdO));
    const uint32_t loK0 = desc_lo_kmajor(smem_u32(sRing)), loV0 = desc_lo_kmajor(smem_u32(sRing) + BW_HT);
    const uint32_t loKm0 = desc_lo_mnmajor(smem_u32(sRing), 1024);
    const uint32_t lodS = desc_lo_kmajor(smem_u32(sdS));
    constexpr uint32_t kStageStep = (2 * BW_HT) >> 4;
    auto issue_scores = [&](int j) {
      const int s = j % DQ3_STAGES;
      mbar_wait(&kv_full[s], (j / DQ3_STAGES) & 1);
      tc_fence_after();
      const uint32_t loK = loK0 + s * kStageStep, loV = loV0 + s * kStageStep;
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16_lo(tS, loQ + 2 * k, loK + 2 * k, id_s, k != 0, leader);     // S[q, keys] = Q K^T
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16_lo(tP, lodO + 2 * k, loV + 2 * k, id_s, k != 0, leader);    // dP[q, keys] = dO V^T
      umma_commit_if(s_full, leader);
    };
    mbar_wait(q_full, 0);
    issue_scores(0);
    for (int j = 0; j < nkv; ++j) {
      if (j + 1 < nkv) {
        mbar_wait(s_free, j & 1);          // S(j), dP(j) are in registers
        issue_scores(j + 1);
      }
      const int s = j % DQ3_STAGES;
      mbar_wait(p_full, j & 1);
      tc_fence_after();
      const uint32_t loKm = loKm0 + s * kStageStep;
#pragma unroll
      for (int k = 0; k < 4; ++k)   // dQ[q, d] += dS K   (K tile re-read MN-major: rows = keys = MMA K)
        umma_bf16_lo(tm_dQ, lodS + 2 * k, loKm + 128 * k, id_acc, (j | k) != 0, leader);
      umma_commit_if(&kv_empty[s], leader);
      umma_commit_if(acc_done, leader);
    }
  } else {
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const int qrow = q0 + r;
    const bool q_ok = qrow < p.Nq;
    const long sidx = (static_cast<long>(b) * p.Hq + h) * p.Npad + qrow;      // q0 + r < Npad always (Npad is a multiple of 64... of 128 rows: guarded)
    const long stot = static_cast<long>(p.B) * p.Hq * p.Npad;
    const float nl = (qrow < p.Npad) ? p.nstat[sidx] : -INFINITY;
    const float nd = (qrow < p.Npad) ? p.nstat[stot + sidx] : 0.f;
    const uint64_t sc2 = pack_f32x2(p.scale_log2, p.scale_log2), nl2 = pack_f32x2(nl, nl), nd2 = pack_f32x2(nd, nd);
    const int sw = r & 7;
    uint8_t* ds_row = sdS + r * 128;
    for (int j = 0; j < nkv; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t rs[32], rp[32];
        tmem_ld_32x32(tS + lane_off + c * 32, rs);
        tmem_ld_32x32(tP + lane_off + c * 32, rp);
        tmem_ld_wait();
        if (c == 1) {
          tc_fence_before();
          __syncwarp();
          mbar_arrive_if(s_free, lane == 0);
        }
        const int nv = p.Nk - j * 64 - c * 32;     // valid keys among these 32 columns (may be <= 0)
        if (nv < 32) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i >= nv) rs[i] = 0xff800000u;      // -inf: p = 0
        }
        uint32_t dk_[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const int qp = c * 16 + (i >> 1);
          float p0, p1, d0, d1;
          if (((qp * BW3_POLY) & 7) < BW3_POLY) bw3_pair<true>(rs[i], rs[i + 1], rp[i], rp[i + 1], sc2, nl2, nd2, p0, p1, d0, d1);
          else bw3_pair<false>(rs[i], rs[i + 1], rp[i], rp[i + 1], sc2, nl2, nd2, p0, p1, d0, d1);
          dk_[i >> 1] = pack_bf16(d0, d1);
        }
        if (c == 0 && j >= 1) mbar_wait(acc_done, (j - 1) & 1);   // dQ MMAs of tile j-1 have finished reading the dS buffer
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int ch = c * 4 + t;
          *reinterpret_cast<uint4*>(ds_row + ((ch ^ sw) << 4)) = make_uint4(dk_[4 * t], dk_[4 * t + 1], dk_[4 * t + 2], dk_[4 * t + 3]);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      mbar_arrive_if(p_full, lane == 0);
    }
    mbar_wait(acc_done, (nkv - 1) & 1);
    tc_fence_after();
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float g[32];
      uint32_t raw[32];
      tmem_ld_32x32(tm_dQ + lane_off + half * 32, raw);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) g[i] = __uint_as_float(raw[i]);
      if (q_ok) {
        const bool rot = p.rope_cos != nullptr && half == 0;
        store_grad_half(p.dQ + b * p.dq_bs + qrow * p.dq_ss + h * p.dq_hs + half * 32, g, p.scale,
                        rot ? p.rope_cos + qrow * 16 : nullptr, rot ? p.rope_sin + qrow * 16 : nullptr);
      }
    }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 256); }
}

// ------------------------------------------------------------------------------------------------------------
constexpr int DKV3_STAGES = 3;
constexpr int DKV3_SMEM = 2 * BW_T /*K,V*/ + DKV3_STAGES * 2 * BW_HT /*(Q,dO) ring*/ + 2 * BW_T /*P^T, dS^T*/ + 256;

__global__ void __launch_bounds__(192, 2) attention_bwd_dkv_v3(const __grid_constant__ AttnBwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sK = smem;
  uint8_t* sV = smem + BW_T;
  uint8_t* sRing = smem + 2 * BW_T;                         // stage s: Q tile (64 rows) at + s*2*BW_HT, dO tile at + BW_HT
  uint8_t* sPT = sRing + DKV3_STAGES * 2 * BW_HT;            // [128 keys x 64 queries]
  uint8_t* sdST = sPT + BW_T;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdST + BW_T);
  uint64_t* kv_full = bars + 0;
  uint64_t* qdo_full = bars + 1;                      // [DKV3_STAGES]
  uint64_t* qdo_empty = qdo_full + DKV3_STAGES;       // [DKV3_STAGES]
  uint64_t* s_full = qdo_empty + DKV3_STAGES;
  uint64_t* s_free = s_full + 1;
  uint64_t* p_full = s_free + 1;
  uint64_t* acc_done = p_full + 1;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(acc_done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k0 = blockIdx.x * 128;
  const int hk = blockIdx.y;
  const int b = blockIdx.z;
  const int G = p.Hq / p.Hkv;
  const int nq = (p.Nq + 63) / 64;
  const int iters = G * nq;

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) __trap();
    mbar_init(kv_full, 1);
    for (int i = 0; i < DKV3_STAGES; ++i) { mbar_init(&qdo_full[i], 1); mbar_init(&qdo_empty[i], 1); }
    mbar_init(s_full, 1); mbar_init(s_free, 4); mbar_init(p_full, 4); mbar_init(acc_done, 1);
    fence_barrier_init();
  }
  griddep_launch();
  if (warp == 1) { tmem_alloc(tmem_ptr_smem, 256); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  griddep_wait();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t tS = tmem_base, tP = tmem_base + 64, tm_dV = tmem_base + 128, tm_dK = tmem_base + 192;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(kv_full, 2 * BW_T);
      tma_load_4d(sK, &p.tmK, kv_full, 0, hk, k0, b);
      tma_load_4d(sV, &p.tmV, kv_full, 0, hk, k0, b);
      for (int it = 0; it < iters; ++it) {
        const int s = it % DKV3_STAGES;
        const int h = hk * G + it / nq, qt = it % nq;
        mbar_wait(&qdo_empty[s], ((it / DKV3_STAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(&qdo_full[s], 2 * BW_HT);
        tma_load_4d(sRing + s * 2 * BW_HT, &p.tmQ, &qdo_full[s], 0, h, qt * 64, b);
        tma_load_4d(sRing + s * 2 * BW_HT + BW_HT, &p.tmdO, &qdo_full[s], 0, h, qt * 64, b);
      }
    }
  } else if (warp == 1) {
    const uint32_t leader = elect_one() ? 1u : 0u;
    constexpr uint32_t id_s = make_idesc_bf16(128, 64, 0, 0);
    constexpr uint32_t id_acc = make_idesc_bf16(128, 64, 0, 1);
    const uint32_t loK = desc_lo_kmajor(smem_u32(sK)), loV = desc_lo_kmajor(smem_u32(sV));
    const uint32_t loQ0 = desc_lo_kmajor(smem_u32(sRing)), lodO0 = desc_lo_kmajor(smem_u32(sRing) + BW_HT);
    const uint32_t loQm0 = desc_lo_mnmajor(smem_u32(sRing), 1024), lodOm0 = desc_lo_mnmajor(smem_u32(sRing) + BW_HT, 1024);
    const uint32_t loPT = desc_lo_kmajor(smem_u32(sPT)), lodST = desc_lo_kmajor(smem_u32(sdST));
    constexpr uint32_t kStageStep = (2 * BW_HT) >> 4;
    auto issue_scores = [&](int it) {
      const int s = it % DKV3_STAGES;
      mbar_wait(&qdo_full[s], (it / DKV3_STAGES) & 1);
      tc_fence_after();
      const uint32_t loQ = loQ0 + s * kStageStep, lodO = lodO0 + s * kStageStep;
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16_lo(tS, loK + 2 * k, loQ + 2 * k, id_s, k != 0, leader);     // S^T[keys, q] = K Q^T
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16_lo(tP, loV + 2 * k, lodO + 2 * k, id_s, k != 0, leader);    // dP^T[keys, q] = V dO^T
      umma_commit_if(s_full, leader);
    };
    mbar_wait(kv_full, 0);
    issue_scores(0);
    for (int it = 0; it < iters; ++it) {
      if (it + 1 < iters) {
        mbar_wait(s_free, it & 1);
        issue_scores(it + 1);
      }
      const int s = it % DKV3_STAGES;
      mbar_wait(p_full, it & 1);
      tc_fence_after();
      const uint32_t loQm = loQm0 + s * kStageStep, lodOm = lodOm0 + s * kStageStep;
#pragma unroll
      for (int k = 0; k < 4; ++k)   // dV[keys, d] += P^T dO   (dO tile re-read MN-major: rows = queries = MMA K)
        umma_bf16_lo(tm_dV, loPT + 2 * k, lodOm + 128 * k, id_acc, (it | k) != 0, leader);
#pragma unroll
      for (int k = 0; k < 4; ++k)   // dK[keys, d] += dS^T Q
        umma_bf16_lo(tm_dK, lodST + 2 * k, loQm + 128 * k, id_acc, (it | k) != 0, leader);
      umma_commit_if(&qdo_empty[s], leader);
      umma_commit_if(acc_done, leader);
    }
  } else {
    // one thread per key row; the 64 columns of a tile are 64 queries whose (negated) statistics are warp-uniform float4 loads.
    // Rows beyond Nk carry zero K / V rows: their P^T / dS^T rows are finite garbage that only reaches dV / dK rows never stored.
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const bool key_ok = (k0 + r) < p.Nk;
    const int sw = r & 7;
    const long stot = static_cast<long>(p.B) * p.Hq * p.Npad;
    const uint64_t sc2 = pack_f32x2(p.scale_log2, p.scale_log2);
    uint8_t* pt_row = sPT + r * 128;
    uint8_t* ds_row = sdST + r * 128;
    for (int it = 0; it < iters; ++it) {
      const int hh = hk * G + it / nq, qt = it % nq;
      const float* nls = p.nstat + (static_cast<long>(b) * p.Hq + hh) * p.Npad + qt * 64;
      const float* nds = nls + stot;
      if (it + 1 < iters && lane < 4) {       // next tile's statistics into L1 (2 x 256 bytes)
        const int h2 = hk * G + (it + 1) / nq, q2 = (it + 1) % nq;
        const float* nx = p.nstat + (static_cast<long>(b) * p.Hq + h2) * p.Npad + q2 * 64 + (lane & 1) * 32 + ((lane & 2) ? stot : 0);
        asm volatile("prefetch.global.L1 [%0];" ::"l"(nx));
      }
      mbar_wait(s_full, it & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 4; ++c) {          // 16 query columns at a time
        float4 l4[4], d4[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          l4[t] = __ldg(reinterpret_cast<const float4*>(nls + c * 16) + t);
          d4[t] = __ldg(reinterpret_cast<const float4*>(nds + c * 16) + t);
        }
        uint32_t rs[16], rp[16];
        tmem_ld_32x16(tS + lane_off + c * 16, rs);
        tmem_ld_32x16(tP + lane_off + c * 16, rp);
        tmem_ld_wait();
        if (c == 3) {
          tc_fence_before();
          __syncwarp();
          mbar_arrive_if(s_free, lane == 0);
        }
        uint32_t pk[8], dk_[8];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          float p0, p1, p2, p3, e0, e1, e2, e3;
          const int qp = c * 8 + t * 2;
          if (((qp * BW3_POLY) & 7) < BW3_POLY)
            bw3_pair<true>(rs[4 * t], rs[4 * t + 1], rp[4 * t], rp[4 * t + 1], sc2, pack_f32x2(l4[t].x, l4[t].y), pack_f32x2(d4[t].x, d4[t].y), p0, p1, e0, e1);
          else
            bw3_pair<false>(rs[4 * t], rs[4 * t + 1], rp[4 * t], rp[4 * t + 1], sc2, pack_f32x2(l4[t].x, l4[t].y), pack_f32x2(d4[t].x, d4[t].y), p0, p1, e0, e1);
          if ((((qp + 1) * BW3_POLY) & 7) < BW3_POLY)
            bw3_pair<true>(rs[4 * t + 2], rs[4 * t + 3], rp[4 * t + 2], rp[4 * t + 3], sc2, pack_f32x2(l4[t].z, l4[t].w), pack_f32x2(d4[t].z, d4[t].w), p2, p3, e2, e3);
          else
            bw3_pair<false>(rs[4 * t + 2], rs[4 * t + 3], rp[4 * t + 2], rp[4 * t + 3], sc2, pack_f32x2(l4[t].z, l4[t].w), pack_f32x2(d4[t].z, d4[t].w), p2, p3, e2, e3);
          pk[2 * t] = pack_bf16(p0, p1); pk[2 * t + 1] = pack_bf16(p2, p3);
          dk_[2 * t] = pack_bf16(e0, e1); dk_[2 * t + 1] = pack_bf16(e2, e3);
        }
        if (c == 0 && it >= 1) mbar_wait(acc_done, (it - 1) & 1);   // accumulate MMAs of tile it-1 have finished reading P^T / dS^T
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const int ch = c * 2 + t;
          *reinterpret_cast<uint4*>(pt_row + ((ch ^ sw) << 4)) = make_uint4(pk[4 * t], pk[4 * t + 1], pk[4 * t + 2], pk[4 * t + 3]);
          *reinterpret_cast<uint4*>(ds_row + ((ch ^ sw) << 4)) = make_uint4(dk_[4 * t], dk_[4 * t + 1], dk_[4 * t + 2], dk_[4 * t + 3]);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      mbar_arrive_if(p_full, lane == 0);
    }
    mbar_wait(acc_done, (iters - 1) & 1);
    tc_fence_after();
    const int krow = k0 + r;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float g[32];
      uint32_t raw[32];
      tmem_ld_32x32(tm_dV + lane_off + half * 32, raw);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) g[i] = __uint_as_float(raw[i]);
      if (key_ok) store_grad_half(p.dV + b * p.dv_bs + krow * p.dv_ss + hk * p.dv_hs + half * 32, g, 1.0f, nullptr, nullptr);
      tmem_ld_32x32(tm_dK + lane_off + half * 32, raw);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) g[i] = __uint_as_float(raw[i]);
      if (key_ok) {
        const bool rot = p.rope_cos != nullptr && half == 0;
        store_grad_half(p.dK + b * p.dk_bs + krow * p.dk_ss + hk * p.dk_hs + half * 32, g, p.scale,
                        rot ? p.rope_cos + krow * 16 : nullptr, rot ? p.rope_sin + krow * 16 : nullptr);
      }
    }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 256); }
}

static int bw_head_map(CUtensorMap* tm, const void* base, int B, int H, int N, long bs, long ss, long hs, int rows) {
  uint64_t dims[4] = {64, static_cast<uint64_t>(H), static_cast<uint64_t>(N), static_cast<uint64_t>(B)};
  uint64_t strides[3] = {static_cast<uint64_t>(hs) * 2, static_cast<uint64_t>(ss) * 2, static_cast<uint64_t>(bs) * 2};
  uint32_t box[4] = {64, 1, static_cast<uint32_t>(rows), 1};
  return encode_tmap_bf16(tm, base, 4, dims, strides, box, 1);
}

}  // namespace b200sat

using namespace b200sat;

extern "C" int b200sat_attention_bwd(const void* q, const void* k, const void* v, const void* o, const void* d_o, const float* lse,
                                     float* delta_scratch, void* dq, void* dk, void* dv, int B, int Hq, int Hkv, int Nq, int Nk,
                                     const long* strides /* 8 x (batch, seq, head): q k v o do dq dk dv */, int head_dim, float scale,
                                     const float* rope_cos, const float* rope_sin, void* stream) {
  if (!q || !k || !v || !o || !d_o || !lse || !delta_scratch || !dq || !dk || !dv || !strides) { set_last_error("attention_bwd: null argument"); return B200SAT_EINVAL; }
  if (head_dim != 64) { set_last_error("attention_bwd: only head_dim 64 is implemented"); return B200SAT_EUNSUPPORTED; }
  if (B <= 0 || Hq <= 0 || Hkv <= 0 || Hq % Hkv || Nq <= 0 || Nk <= 0) { set_last_error("attention_bwd: bad shape"); return B200SAT_EINVAL; }
  if ((rope_cos == nullptr) != (rope_sin == nullptr)) { set_last_error("attention_bwd: rope tables go together"); return B200SAT_EINVAL; }
  const long* sq = strides, *sk = strides + 3, *sv = strides + 6, *so = strides + 9, *sdo = strides + 12, *sdq = strides + 15,
              *sdk = strides + 18, *sdv = strides + 21;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int Npad = (Nq + 63) / 64 * 64;
  long npad_total = 0;
  {
    const long n = static_cast<long>(B) * Hq * Nq;
    npad_total = static_cast<long>(B) * Hq * Npad;
    B200SAT_CHECK_CUDA(launch_k(attn_delta_kernel, dim3(static_cast<int>((npad_total + 127) / 128)), dim3(128), 0, s, 1, static_cast<const __nv_bfloat16*>(o),
                                static_cast<const __nv_bfloat16*>(d_o), lse, delta_scratch + 2 * npad_total, delta_scratch, B, Hq, Nq, Npad, so[0], so[1], so[2], sdo[0], sdo[1], sdo[2]));
  }
  AttnBwdParams p, pq;   // p: dK/dV kernel (128-key rows, 64-query inner tiles); pq: dQ kernel (128-query rows, 64-key inner tiles)
  memset(&p, 0, sizeof(p));
  int rc;
  if ((rc = bw_head_map(&p.tmQ, q, B, Hq, Nq, sq[0], sq[1], sq[2], 64))) return rc;
  if ((rc = bw_head_map(&p.tmK, k, B, Hkv, Nk, sk[0], sk[1], sk[2], 128))) return rc;
  if ((rc = bw_head_map(&p.tmV, v, B, Hkv, Nk, sv[0], sv[1], sv[2], 128))) return rc;
  if ((rc = bw_head_map(&p.tmdO, d_o, B, Hq, Nq, sdo[0], sdo[1], sdo[2], 64))) return rc;
  p.nstat = delta_scratch; p.Npad = Npad;                       // [2][B, Hq, Npad] first (16-byte aligned rows), then delta [B, Hq, Nq]
  p.lse = lse; p.delta = delta_scratch + 2 * npad_total;
  p.dQ = static_cast<__nv_bfloat16*>(dq); p.dK = static_cast<__nv_bfloat16*>(dk); p.dV = static_cast<__nv_bfloat16*>(dv);
  p.dq_bs = sdq[0]; p.dq_ss = sdq[1]; p.dq_hs = sdq[2];
  p.dk_bs = sdk[0]; p.dk_ss = sdk[1]; p.dk_hs = sdk[2];
  p.dv_bs = sdv[0]; p.dv_ss = sdv[1]; p.dv_hs = sdv[2];
  p.rope_cos = rope_cos; p.rope_sin = rope_sin;
  p.B = B; p.Hq = Hq; p.Hkv = Hkv; p.Nq = Nq; p.Nk = Nk;
  p.scale = scale; p.scale_log2 = scale * 1.4426950408889634f;
  static bool attr_set = false;
  if (!attr_set) {
    B200SAT_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_dkv_tcgen05<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, DKV_SMEM));
    B200SAT_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_dq_tcgen05<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, DQ_SMEM));
    B200SAT_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_dkv_v3, cudaFuncAttributeMaxDynamicSharedMemorySize, DKV3_SMEM));
    B200SAT_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_dq_v3, cudaFuncAttributeMaxDynamicSharedMemorySize, DQ3_SMEM));
    B200SAT_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_dkv_v3, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    B200SAT_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_dq_v3, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    attr_set = true;
  }
  pq = p;
  if ((rc = bw_head_map(&pq.tmQ, q, B, Hq, Nq, sq[0], sq[1], sq[2], 128))) return rc;
  if ((rc = bw_head_map(&pq.tmK, k, B, Hkv, Nk, sk[0], sk[1], sk[2], 64))) return rc;
  if ((rc = bw_head_map(&pq.tmV, v, B, Hkv, Nk, sv[0], sv[1], sv[2], 64))) return rc;
  if ((rc = bw_head_map(&pq.tmdO, d_o, B, Hq, Nq, sdo[0], sdo[1], sdo[2], 128))) return rc;
  // B200SAT_ATTN_BWD_V3 (default 3): bit 0 = v3 dK/dV kernel, bit 1 = v3 dQ kernel; a cleared bit runs the round-1 (v2) kernel of that half
  // (read per call so one process can compare them)
  const char* ver_env = getenv("B200SAT_ATTN_BWD_V3");
  const int v3mask = ver_env ? atoi(ver_env) : 3;
  if (v3mask & 1) B200SAT_CHECK_CUDA(launch_k(attention_bwd_dkv_v3, dim3((Nk + 127) / 128, Hkv, B), dim3(192), DKV3_SMEM, s, 1, p));
  else B200SAT_CHECK_CUDA(launch_k(attention_bwd_dkv_tcgen05<8>, dim3((Nk + 127) / 128, Hkv, B), dim3(320), DKV_SMEM, s, 1, p));
  if (v3mask & 2) B200SAT_CHECK_CUDA(launch_k(attention_bwd_dq_v3, dim3((Nq + 127) / 128, Hq, B), dim3(192), DQ3_SMEM, s, 1, pq));
  else B200SAT_CHECK_CUDA(launch_k(attention_bwd_dq_tcgen05<8>, dim3((Nq + 127) / 128, Hq, B), dim3(320), DQ_SMEM, s, 1, pq));
  B200SAT_CHECK_CUDA(cudaGetLastError());
  return B200SAT_OK;
}

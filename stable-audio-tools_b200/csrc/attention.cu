// b200sat — flash attention forward on tcgen05 (non-causal, dh = 64, GQA-aware, ragged sequence tails).
//
// Replaces flash_attn_func / F.scaled_dot_product_attention in the reference
// (stable_audio_tools/models/transformer.py:406-441) for both the self-attention (N = 1025) and the grouped-query
// cross-attention over the conditioning tokens (Nk = 130, 24 q heads / 12 kv heads, :408-411).  Heads are read in place
// from the projection outputs ([B, N, heads, 64] views with arbitrary strides), so the 'b n (h d) -> b h n d' rearranges
// and the repeat_interleave of K/V never touch HBM.
//
// One CTA = one 128-query tile of one (batch, head); two CTAs are resident per SM so one CTA's softmax overlaps the
// other's MMAs.  320 threads:
//   warp 0 lane 0 : TMA producer (Q once; K,V tiles of 128 keys through a 2-stage ring; 4-D tensor maps, 128B swizzle)
//   warp 1        : TMEM allocation; lane 0 issues S = Q K^T (128x128x64) and O_j = P V (128x64x128) on tcgen05
//   warps 2..9    : online softmax — two threads per query row (64 keys each; tcgen05.ld 32x32b), P written to smem as the
//                   bf16 A operand, O (32 dims per thread) accumulated in registers with the running-max rescale.
//                   (ncu r1: with 4 softmax warps the kernel was latency-bound: issue-active 31 %, tensor pipe 17 %.)
#include "common.cuh"
#include <cstring>

namespace b200sat {

__device__ __forceinline__ float fast_exp2(float x) {
#if defined(AT_VARIANT) && (AT_VARIANT & 1)
  return x * 0.001f + 1.0f;
#else
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
#endif
}

// packed fp32 pairs (sm_100: FFMA2 / FADD2 issue two lanes per slot)
__device__ __forceinline__ uint64_t pack_f32x2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpack_f32x2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma_f32x2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.ftz.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ uint64_t add_f32x2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.ftz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

struct AttnParams {
  CUtensorMap tmQ, tmK, tmV;
  __nv_bfloat16* O;
  float* lse;  // optional [B, Hq, Nq]
  long o_bs, o_ss, o_hs;
  int B, Hq, Hkv, Nq, Nk;
  float scale_log2;  // softmax scale * log2(e)
};

constexpr int AT_BM = 128;   // queries per CTA
constexpr int AT_BN = 64;    // keys per tile
constexpr int AT_D = 64;
constexpr int AT_QT = AT_BM * AT_D * 2;   // 16 KB: Q tile, P tile (128 x 64 bf16)
constexpr int AT_KT = AT_BN * AT_D * 2;   //  8 KB: K tile, V tile
constexpr int AT_KV_STAGES = 3;
constexpr int AT_SMEM = AT_QT /*Q*/ + AT_KV_STAGES * 2 * AT_KT /*K,V ring*/ + 2 * AT_QT /*P x2*/ + 256 /*barriers*/ + 1024 /*row-max exchange, two parities*/;
constexpr float AT_RESCALE_THRESHOLD = 8.0f;
// AT_VARIANT bit 7: per-tile timeline of one early and one late CTA (SM clock, low 32 bits) written over the start of p.lse:
// region r (CTA (0,0,0) -> 0, CTA (4,12,4) -> 1), record [r][j][16]: slots 0-7 softmax warp 2, 8-12 MMA warp, 13-15 CTA entry / loop start / end.
#define AT_TRACE(j, slot)                                                                                                     \
  do {                                                                                                                        \
    if ((AT_VARIANT & 128) && trace_region >= 0 && lane == 0)                                                                 \
      reinterpret_cast<uint32_t*>(p.lse)[(trace_region * 64 + (j)) * 16 + (slot)] = static_cast<uint32_t>(clock64());         \
  } while (0)
#ifndef AT_VARIANT
#define AT_VARIANT 0   // timing experiments only (tools/attn_variants.py): bit 0 no MUFU, 1 no pair exchange, 2 no max, 3 no P store, 4 no TMEM load
#endif  // log2 units: O/l are only rescaled when the row max grows by more than 2^8

// Pipeline (round-1 redesign after the ncu capture of the first version: tensor pipe 17 %, issue slots 31 %, every tile paid
// four serial barrier hops S -> softmax -> P -> PV -> O read):
//   * 64-key tiles, the score tile S double-buffered in TMEM: the MMA thread keeps S(j+1), S(j+2) ahead of the softmax warps;
//   * O accumulates in TMEM across tiles (PV issued with accumulate) — the softmax warps never wait for PV on the common path;
//   * lazy rescale: the running max only moves when it grows by more than 2^8 (any upper bound is a valid softmax shift), and
//     only then a thread waits for the previous PV and rescales its half row of O in TMEM (tcgen05.ld -> scale -> tcgen05.st);
//   * P double-buffered in shared memory so the softmax of tile j+1 overlaps the PV MMA of tile j.
__global__ void __launch_bounds__(320, 2) attention_fwd_tcgen05(const __grid_constant__ AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sKV = smem + AT_QT;                                   // stage s: K at + s*2*AT_KT, V at + AT_KT
  uint8_t* sP = sKV + AT_KV_STAGES * 2 * AT_KT;                  // 2 x [128 rows x 64 keys]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * AT_QT);
  // K and V tiles travel through SEPARATE rings (round 2): a K stage is free again as soon as its S = Q K^T MMAs retire, a V stage
  // only after P V - with one shared ring every stage was held for ~3 iterations, the next load got a lead of less than one
  // iteration (~1 us ~ the L2 -> shared-memory latency of a TMA tile), and each CTA's iteration time sat at that latency.
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;                       // [AT_KV_STAGES]
  uint64_t* k_empty = k_full + AT_KV_STAGES;
  uint64_t* v_full = k_empty + AT_KV_STAGES;
  uint64_t* v_empty = v_full + AT_KV_STAGES;
  uint64_t* s_full = v_empty + AT_KV_STAGES;         // [2]
  uint64_t* p_full = s_full + 2;                     // [2]
  uint64_t* pv_done = p_full + 2;                    // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(pv_done + 2);
  __nv_bfloat16* s_max = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(bars) + 256);  // [2 halves][128 rows]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * AT_BM;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int hkv = h / (p.Hq / p.Hkv);
  const int num_kv = (p.Nk + AT_BN - 1) / AT_BN;
  const int trace_region = (AT_VARIANT & 128) ? ((blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) ? 0 : (blockIdx.x == 4 && blockIdx.y == 12 && blockIdx.z == 4) ? 1 : -1) : -1;
  if (warp == 2) AT_TRACE(0, 13);

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) __trap();  // the swizzled layouts need a 1024-byte aligned base
    tma_prefetch_desc(&p.tmQ);
    tma_prefetch_desc(&p.tmK);
    tma_prefetch_desc(&p.tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < AT_KV_STAGES; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_full[i], (AT_VARIANT & 256) ? 8 : 256); mbar_init(&pv_done[i], 1); }
    fence_barrier_init();
  }
  griddep_launch();
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  griddep_wait();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t tmem_O = tmem_base + 128;   // 64 fp32 columns; S buffers at +0 and +64

  if (warp == 0) {
    if (lane == 0) {          // Q, then the K tiles
      mbar_arrive_expect_tx(q_full, AT_QT);
      tma_load_4d(sQ, &p.tmQ, q_full, 0, h, q0, b);
      for (int j = 0; j < num_kv; ++j) {
        const int s = j % AT_KV_STAGES;
        mbar_wait(&k_empty[s], ((j / AT_KV_STAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(&k_full[s], AT_KT);
        tma_load_4d(sKV + s * 2 * AT_KT, &p.tmK, &k_full[s], 0, hkv, j * AT_BN, b);
      }
    } else if (lane == 1) {   // the V tiles, on their own ring so they never hold a K stage back
      for (int j = 0; j < num_kv; ++j) {
        const int s = j % AT_KV_STAGES;
        mbar_wait(&v_empty[s], ((j / AT_KV_STAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(&v_full[s], AT_KT);
        tma_load_4d(sKV + s * 2 * AT_KT + AT_KT, &p.tmV, &v_full[s], 0, hkv, j * AT_BN, b);
      }
    }
  } else if (warp == 1) {
    // MMA issue: the whole warp walks the loop (uniform control flow), one elected lane issues (see umma_bf16_lo in common.cuh)
    const uint32_t leader = elect_one() ? 1u : 0u;
    constexpr uint32_t idesc_s = make_idesc_bf16(128, 64, 0, 0);
    constexpr uint32_t idesc_o = make_idesc_bf16(128, 64, 0, 1);  // B = V is MN-major (keys along rows)
    const uint32_t loQ = desc_lo_kmajor(smem_u32(sQ));
    const uint32_t loK0 = desc_lo_kmajor(smem_u32(sKV));
    const uint32_t loV0 = desc_lo_mnmajor(smem_u32(sKV) + AT_KT, 1024);
    const uint32_t loP0 = desc_lo_kmajor(smem_u32(sP));
    auto issue_s = [&](int j) {
      const int s = j % AT_KV_STAGES;
      mbar_wait(&k_full[s], (j / AT_KV_STAGES) & 1);
      tc_fence_after();
      const uint32_t loK = loK0 + s * ((2 * AT_KT) >> 4);
      const uint32_t tS = tmem_base + (j & 1) * 64;
#pragma unroll
      for (int k = 0; k < AT_D / 16; ++k) umma_bf16_lo(tS, loQ + 2 * k, loK + 2 * k, idesc_s, k != 0, leader);
      umma_commit_if(&s_full[j & 1], leader);
      umma_commit_if(&k_empty[s], leader);       // the K stage is reusable as soon as these MMAs retire
    };
    mbar_wait(q_full, 0);
    issue_s(0);
    if (num_kv > 1) issue_s(1);
    for (int j = 0; j < num_kv; ++j) {
      const int s = j % AT_KV_STAGES;
      AT_TRACE(j, 8);
      mbar_wait(&p_full[j & 1], (j >> 1) & 1);
      AT_TRACE(j, 9);
      mbar_wait(&v_full[s], (j / AT_KV_STAGES) & 1);
      AT_TRACE(j, 10);
      tc_fence_after();
      const uint32_t loP = loP0 + (j & 1) * (AT_QT >> 4);
      const uint32_t loV = loV0 + s * ((2 * AT_KT) >> 4);
#pragma unroll
      for (int k = 0; k < AT_BN / 16; ++k) umma_bf16_lo(tmem_O, loP + 2 * k, loV + 128 * k, idesc_o, (j | k) != 0, leader);
      umma_commit_if(&pv_done[j & 1], leader);
      umma_commit_if(&v_empty[s], leader);
      AT_TRACE(j, 11);
      if (j + 2 < num_kv) issue_s(j + 2);   // its S buffer was consumed before p_full(j) completed
      AT_TRACE(j, 12);
    }
  } else {
    // ===================== softmax / output warps: two threads per query row, 32 keys each per tile =====================
    const int quarter = warp & 3;          // TMEM lane quarter accessible to this warp
    const int half = (warp - 2) >> 2;      // which 32 keys of the 64-key tile / which 32 output dims
    const int r = quarter * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    float m_run = -INFINITY, l_run = 0.f;
    const int sw = r & 7;
    if (warp == 2) AT_TRACE(0, 14);
    __nv_bfloat16* my_max = s_max + half * 128 + r;
    const __nv_bfloat16* other_max = s_max + (half ^ 1) * 128 + r;

    for (int j = 0; j < num_kv; ++j) {
      const int buf = j & 1;
      if (warp == 2) AT_TRACE(j, 0);
      mbar_wait(&s_full[buf], (j >> 1) & 1);
      if (warp == 2) AT_TRACE(j, 1);
      tc_fence_after();
      const int nvalid = p.Nk - j * AT_BN - half * 32;   // valid keys among this thread's 32 columns (may be <= 0)
      uint32_t raw[32];
      if (AT_VARIANT & 16) {
#pragma unroll
        for (int i = 0; i < 32; ++i) raw[i] = __float_as_uint(static_cast<float>((lane + i + j) & 7));
      } else {
        tmem_ld_32x32(tmem_base + lane_off + buf * 64 + half * 32, raw);
        tmem_ld_wait();
      }
      if (warp == 2) AT_TRACE(j, 2);
      float mx = -INFINITY;
      if (AT_VARIANT & 4) {
        mx = 8.0f / p.scale_log2;
      } else if (nvalid >= 32) {
#pragma unroll
        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(raw[i]));
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, (i < nvalid) ? __uint_as_float(raw[i]) : -INFINITY);
      }
      // the two threads of a row agree on a shift: max of the two half-row maxima, rounded UP to bf16
#if AT_VARIANT & 32
      my_max[(j & 1) * 256] = __float2bfloat16_ru(mx * p.scale_log2);
      asm volatile("bar.sync %0, 64;" ::"r"(quarter + 1) : "memory");
      const float m_pair = fmaxf(__bfloat162float(__float2bfloat16_ru(mx * p.scale_log2)), __bfloat162float(other_max[(j & 1) * 256]));
#elif AT_VARIANT & 2
      const float m_pair = __bfloat162float(__float2bfloat16_ru(mx * p.scale_log2));
#else
      *my_max = __float2bfloat16_ru(mx * p.scale_log2);
      asm volatile("bar.sync %0, 64;" ::"r"(quarter + 1) : "memory");
      const float m_pair = fmaxf(__bfloat162float(*my_max), __bfloat162float(*other_max));
      asm volatile("bar.sync %0, 64;" ::"r"(quarter + 1) : "memory");   // partner has read my slot before I overwrite it next tile
#endif
      // rare path: move the shift.  l and the TMEM-resident O row are rescaled; all earlier PV MMAs must have retired.
      // tcgen05.ld/st are warp-collective, so the branch is taken by the whole warp when ANY of its rows needs it (rows
      // that do not move use alpha = 1); the partner warp owns the same rows and takes the same decision.
      if (warp == 2) AT_TRACE(j, 3);
      const bool need = m_pair > m_run + AT_RESCALE_THRESHOLD;
      if (__any_sync(0xffffffffu, need)) {
        const float alpha = need ? fast_exp2(m_run - m_pair) : 1.0f;
        if (j > 0) {
          mbar_wait(&pv_done[(j - 1) & 1], ((j - 1) >> 1) & 1);
          tc_fence_after();
          uint32_t ov[32];
          tmem_ld_32x32(tmem_O + lane_off + half * 32, ov);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
          tmem_st_32x32(tmem_O + lane_off + half * 32, ov);
          tmem_st_wait();
        }
        l_run *= alpha;
        if (need) m_run = m_pair;
      }
      // p = exp2(s*scale - m), partial row sum, bf16 P -> swizzled smem (A operand of the PV MMA)
      float lsum = 0.f;
      uint32_t pk[16];
      if (nvalid >= 32 && (AT_VARIANT & 64)) {
        const uint64_t sc2 = pack_f32x2(p.scale_log2, p.scale_log2), nm2 = pack_f32x2(-m_run, -m_run);
        uint64_t acc0 = pack_f32x2(0.f, 0.f), acc1 = acc0;
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          float a0, a1, b0, b1;
          unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(raw[i]), __uint_as_float(raw[i + 1])), sc2, nm2), a0, a1);
          unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(raw[i + 2]), __uint_as_float(raw[i + 3])), sc2, nm2), b0, b1);
          a0 = fast_exp2(a0); a1 = fast_exp2(a1); b0 = fast_exp2(b0); b1 = fast_exp2(b1);
          acc0 = add_f32x2(acc0, pack_f32x2(a0, a1));
          acc1 = add_f32x2(acc1, pack_f32x2(b0, b1));
          pk[i >> 1] = pack_bf16(a0, a1);
          pk[(i >> 1) + 1] = pack_bf16(b0, b1);
        }
        float s0, s1;
        unpack_f32x2(add_f32x2(acc0, acc1), s0, s1);
        lsum = s0 + s1;
      } else if (nvalid >= 32) {
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float p0 = fast_exp2(fmaf(__uint_as_float(raw[i]), p.scale_log2, -m_run));
          const float p1 = fast_exp2(fmaf(__uint_as_float(raw[i + 1]), p.scale_log2, -m_run));
          lsum += p0 + p1;
          pk[i >> 1] = pack_bf16(p0, p1);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float p0 = (i < nvalid) ? fast_exp2(fmaf(__uint_as_float(raw[i]), p.scale_log2, -m_run)) : 0.f;
          const float p1 = (i + 1 < nvalid) ? fast_exp2(fmaf(__uint_as_float(raw[i + 1]), p.scale_log2, -m_run)) : 0.f;
          lsum += p0 + p1;
          pk[i >> 1] = pack_bf16(p0, p1);
        }
      }
      l_run += lsum;
      if (warp == 2) AT_TRACE(j, 4);
      if (j >= 2) mbar_wait(&pv_done[buf], ((j >> 1) - 1) & 1);   // PV(j-2) has finished reading this P buffer
      if (warp == 2) AT_TRACE(j, 5);
      uint8_t* prow = sP + buf * AT_QT + r * 128;
#pragma unroll
      for (int t = 0; t < ((AT_VARIANT & 8) ? 0 : 4); ++t) {
        const int ch = half * 4 + t;              // 16-byte chunk index along the 64 keys
        *reinterpret_cast<uint4*>(prow + ((ch ^ sw) << 4)) = make_uint4(pk[4 * t], pk[4 * t + 1], pk[4 * t + 2], pk[4 * t + 3]);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      if (warp == 2) AT_TRACE(j, 6);
      if (AT_VARIANT & 256) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[buf]);
      } else {
        mbar_arrive(&p_full[buf]);
      }
      if (warp == 2) AT_TRACE(j, 7);
    }
    // epilogue: all PV MMAs retired -> read this thread's 32 output dims, combine the two partial row sums, normalise
    mbar_wait(&pv_done[(num_kv - 1) & 1], ((num_kv - 1) >> 1) & 1);
    tc_fence_after();
    uint32_t ov[32];
    tmem_ld_32x32(tmem_O + lane_off + half * 32, ov);
    tmem_ld_wait();
    float* s_l = reinterpret_cast<float*>(sP);   // P buffers are free now
    s_l[half * 128 + r] = l_run;
    asm volatile("bar.sync %0, 64;" ::"r"(quarter + 1) : "memory");
    const float l_tot = l_run + s_l[(half ^ 1) * 128 + r];
    const int qrow = q0 + r;
    if (qrow < p.Nq) {
      const float inv = 1.0f / l_tot;
      __nv_bfloat16* dst = p.O + static_cast<long>(b) * p.o_bs + static_cast<long>(qrow) * p.o_ss + static_cast<long>(h) * p.o_hs + half * 32;
      uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 u;
        u.x = pack_bf16(__uint_as_float(ov[8 * i + 0]) * inv, __uint_as_float(ov[8 * i + 1]) * inv);
        u.y = pack_bf16(__uint_as_float(ov[8 * i + 2]) * inv, __uint_as_float(ov[8 * i + 3]) * inv);
        u.z = pack_bf16(__uint_as_float(ov[8 * i + 4]) * inv, __uint_as_float(ov[8 * i + 5]) * inv);
        u.w = pack_bf16(__uint_as_float(ov[8 * i + 6]) * inv, __uint_as_float(ov[8 * i + 7]) * inv);
        d4[i] = u;
      }
      if (p.lse && half == 0 && !(AT_VARIANT & 128)) p.lse[(static_cast<long>(b) * p.Hq + h) * p.Nq + qrow] = m_run * 0.6931471805599453f + logf(l_tot);
    }
    tc_fence_before();
    if (warp == 2) AT_TRACE(0, 15);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

static int make_head_map(CUtensorMap* tm, const void* base, int B, int H, int N, long bs, long ss, long hs, int box_rows) {
  uint64_t dims[4] = {AT_D, static_cast<uint64_t>(H), static_cast<uint64_t>(N), static_cast<uint64_t>(B)};
  uint64_t strides[3] = {static_cast<uint64_t>(hs) * 2, static_cast<uint64_t>(ss) * 2, static_cast<uint64_t>(bs) * 2};
  uint32_t box[4] = {AT_D, 1, static_cast<uint32_t>(box_rows), 1};
  return encode_tmap_bf16(tm, base, 4, dims, strides, box, 1);
}

}  // namespace b200sat

using namespace b200sat;

extern "C" int b200sat_attention_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int B, int Hq, int Hkv,
                                     int Nq, int Nk, long q_bs, long q_ss, long q_hs, long k_bs, long k_ss, long k_hs,
                                     long v_bs, long v_ss, long v_hs, long o_bs, long o_ss, long o_hs, int head_dim,
                                     float scale, void* stream) {
  if (!q || !k || !v || !o || B <= 0 || Hq <= 0 || Hkv <= 0 || Nq <= 0 || Nk <= 0) { set_last_error("attention: bad arguments"); return B200SAT_EINVAL; }
  if (head_dim != AT_D) { set_last_error("attention: only head_dim 64 is implemented"); return B200SAT_EUNSUPPORTED; }
  if (Hq % Hkv) { set_last_error("attention: Hq must be a multiple of Hkv"); return B200SAT_EINVAL; }
  if ((o_ss % 8) || (o_hs % 8) || (o_bs % 8)) { set_last_error("attention: output strides must be multiples of 8 elements"); return B200SAT_EINVAL; }
  AttnParams p;
  memset(&p, 0, sizeof(p));
  int rc;
  if ((rc = make_head_map(&p.tmQ, q, B, Hq, Nq, q_bs, q_ss, q_hs, AT_BM))) return rc;
  if ((rc = make_head_map(&p.tmK, k, B, Hkv, Nk, k_bs, k_ss, k_hs, AT_BN))) return rc;
  if ((rc = make_head_map(&p.tmV, v, B, Hkv, Nk, v_bs, v_ss, v_hs, AT_BN))) return rc;
  p.O = static_cast<__nv_bfloat16*>(o); p.lse = lse;
  p.o_bs = o_bs; p.o_ss = o_ss; p.o_hs = o_hs;
  p.B = B; p.Hq = Hq; p.Hkv = Hkv; p.Nq = Nq; p.Nk = Nk;
  p.scale_log2 = scale * 1.4426950408889634f;
  static bool attr_set = false;
  if (!attr_set) {
    B200SAT_CHECK_CUDA(cudaFuncSetAttribute(attention_fwd_tcgen05, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM));
    attr_set = true;
  }
  dim3 grid((Nq + AT_BM - 1) / AT_BM, Hq, B);
  B200SAT_CHECK_CUDA(launch_k(attention_fwd_tcgen05, dim3(grid), dim3(320), AT_SMEM, static_cast<cudaStream_t>(stream), 1, p));
  B200SAT_CHECK_CUDA(cudaGetLastError());
  return B200SAT_OK;
}

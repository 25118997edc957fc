// b200sat — flash attention forward on tcgen05 (non-causal, dh = 64, GQA-aware, ragged sequence tails).
//
// Replaces flash_attn_func / F.scaled_dot_product_attention in the reference
// (stable_audio_tools/models/transformer.py:406-441) for both the self-attention (N = 1025) and the grouped-query
// cross-attention over the conditioning tokens (Nk = 130, 24 q heads / 12 kv heads, :408-411).  Heads are read in place
// from the projection outputs ([B, N, heads, 64] views with arbitrary strides), so the 'b n (h d) -> b h n d' rearranges
// and the repeat_interleave of K/V never touch HBM.
//
// One CTA = one 128-query tile of one (batch, head); THREE CTAs are resident per SM so one CTA's softmax runs under the others' MMAs and
// latencies.  192 threads (the round-2 redesign and what its in-kernel timeline showed are described above the kernel):
//   warp 0 lane 0 : TMA producer (Q once; K and V tiles of 64 keys through separate rings; 4-D tensor maps, 128B swizzle)
//   warp 1        : TMEM allocation; one elected lane issues S = Q K^T (128x64x64) and O += P V (128x64x64) on tcgen05
//   warps 2..5    : online softmax, ONE thread per query row (64 keys per tile in registers from one tcgen05.ld pair), P written to smem
//                   as the bf16 A operand, O accumulated in TMEM with a lazy rescale; a quarter of the exp2 on the FMA pipe.
#include "common.cuh"
#include <cstring>

namespace b200sat {


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
constexpr int AT_STAGES = 2;              // K ring and V ring depth
constexpr int AT_THREADS = 192;
constexpr int AT_SMEM = AT_QT /*Q*/ + 2 * AT_STAGES * AT_KT /*K ring, V ring*/ + AT_QT /*P*/ + 256 /*barriers*/;
constexpr float AT_RESCALE_THRESHOLD = 8.0f;  // log2 units: O/l are only rescaled when the row max grows by more than 2^8
#ifndef AT_VARIANT
#define AT_VARIANT 0   // debug builds only (tools/attn_variants.py): bit 7 = per-tile timeline; bits 0-3 override AT_POLY_EIGHTHS; bit 4 = branchy arrive
#endif
#ifndef AT_POLY_EIGHTHS
#define AT_POLY_EIGHTHS 2   // of every 8 (column pairs), this many take the FMA-pipe exp2 instead of the MUFU
#endif
constexpr int AT_POLY = (AT_VARIANT & 15) ? ((AT_VARIANT & 15) == 15 ? 0 : (AT_VARIANT & 15)) : AT_POLY_EIGHTHS;
// AT_VARIANT bit 7: per-tile timeline of one early and one late CTA (SM clock, low 32 bits) written over the start of p.lse:
// region r (CTA (0,0,0) -> 0, CTA (4,12,4) -> 1), record [r][j][16]: slots 0-7 softmax warp 2, 8-12 MMA warp, 13-15 CTA entry / loop start / end.
#define AT_TRACE(j, slot)                                                                                                     \
  do {                                                                                                                        \
    if ((AT_VARIANT & 128) && trace_region >= 0 && lane == 0)                                                                 \
      reinterpret_cast<uint32_t*>(p.lse)[(trace_region * 64 + (j)) * 16 + (slot)] = static_cast<uint32_t>(clock64());         \
  } while (0)

__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]),
               "r"(v[13]), "r"(v[14]), "r"(v[15])
               : "memory");
}

// Round-2 redesign.  The in-kernel timeline of the previous version (profiles/r2_attention_fwd_timeline_before.txt: two threads per
// row, 8 softmax warps, 2 CTAs / SM) showed every tile paying ~1850 cycles of SERIAL softmax-warp latency - barrier hops, the
// row-max exchange between the two threads of a row, 800 cycles in the exp phase because the two warps of a pair hit the MUFU at the
// same time - with only two independent instruction streams per scheduler to hide it; knocking out any one phase shortened the
// kernel by exactly that phase (profiles/r2_attention_fwd_phase_variants.txt).  So:
//   * ONE thread per query row (4 softmax warps, 64 keys each per tile): no cross-thread exchange, no named barriers;
//   * THREE CTAs per SM (192 threads, 128 TMEM columns, 65 KB of shared memory each): three independent softmax streams per
//     scheduler, so one CTA's MUFU phase (64 ex2 per row and tile = 512 cycles per warp) runs under the others' latencies;
//   * S is single-buffered: a row's 64 scores go to registers in one tcgen05.ld pair and the buffer is handed back at once
//     (s_free), so S(j+1) = Q K(j+1)^T is issued while the softmax of tile j is still in its max / exp phase;
//   * one mbarrier arrival per WARP (after __syncwarp), not per thread;
//   * O accumulates in TMEM across tiles with the lazy rescale (the running max only moves when it grows by more than 2^8);
//   * packed fp32 pairs (FFMA2 / FADD2) for the scale-shift and the row sum.
// Per tile and row: 64 ex2 on the MUFU (16 / clk / SM -> 512 cycles per 128 x 64 tile and SM) is the bound this layout aims at.
__global__ void __launch_bounds__(AT_THREADS, 3) attention_fwd_tcgen05(const __grid_constant__ AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + AT_QT;                      // [AT_STAGES][64 keys x 64]
  uint8_t* sV = sK + AT_STAGES * AT_KT;          // [AT_STAGES][64 keys x 64]
  uint8_t* sP = sV + AT_STAGES * AT_KT;          // [128 rows x 64 keys]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + AT_QT);
  // K and V tiles travel through SEPARATE rings: a K stage is free again as soon as its S = Q K^T MMAs retire, a V stage only after P V.
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;                       // [AT_STAGES]
  uint64_t* k_empty = k_full + AT_STAGES;
  uint64_t* v_full = k_empty + AT_STAGES;
  uint64_t* v_empty = v_full + AT_STAGES;
  uint64_t* s_full = v_empty + AT_STAGES;            // S(j) is in TMEM                      (MMA -> softmax)
  uint64_t* s_free = s_full + 1;                     // S(j) has been read into registers    (softmax -> MMA, one arrival per warp)
  uint64_t* p_full = s_free + 1;                     // P(j) is in shared memory             (softmax -> MMA, one arrival per warp)
  uint64_t* pv_done = p_full + 1;                    // P V(j) retired: P buffer free, O current
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(pv_done + 1);

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
    for (int i = 0; i < AT_STAGES; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
    mbar_init(s_full, 1); mbar_init(s_free, 4); mbar_init(p_full, 4); mbar_init(pv_done, 1);
    fence_barrier_init();
  }
  griddep_launch();
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, 128);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  griddep_wait();
  const uint32_t tmem_S = *tmem_ptr_smem;    // 64 fp32 columns
  const uint32_t tmem_O = tmem_S + 64;       // 64 fp32 columns

  if (warp == 0) {
    if (lane == 0) {          // Q, then the K tiles
      mbar_arrive_expect_tx(q_full, AT_QT);
      tma_load_4d(sQ, &p.tmQ, q_full, 0, h, q0, b);
      for (int j = 0; j < num_kv; ++j) {
        const int s = j % AT_STAGES;
        mbar_wait(&k_empty[s], ((j / AT_STAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(&k_full[s], AT_KT);
        tma_load_4d(sK + s * AT_KT, &p.tmK, &k_full[s], 0, hkv, j * AT_BN, b);
      }
    } else if (lane == 1) {   // the V tiles
      for (int j = 0; j < num_kv; ++j) {
        const int s = j % AT_STAGES;
        mbar_wait(&v_empty[s], ((j / AT_STAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(&v_full[s], AT_KT);
        tma_load_4d(sV + s * AT_KT, &p.tmV, &v_full[s], 0, hkv, j * AT_BN, b);
      }
    }
  } else if (warp == 1) {
    // MMA issue: the whole warp walks the loop (uniform control flow), one elected lane issues (see umma_bf16_lo in common.cuh)
    const uint32_t leader = elect_one() ? 1u : 0u;
    constexpr uint32_t idesc_s = make_idesc_bf16(128, 64, 0, 0);
    constexpr uint32_t idesc_o = make_idesc_bf16(128, 64, 0, 1);  // B = V is MN-major (keys along rows)
    const uint32_t loQ = desc_lo_kmajor(smem_u32(sQ));
    const uint32_t loK0 = desc_lo_kmajor(smem_u32(sK));
    const uint32_t loV0 = desc_lo_mnmajor(smem_u32(sV), 1024);
    const uint32_t loP = desc_lo_kmajor(smem_u32(sP));
    auto issue_s = [&](int j) {
      const int s = j % AT_STAGES;
      mbar_wait(&k_full[s], (j / AT_STAGES) & 1);
      tc_fence_after();
      const uint32_t loK = loK0 + s * (AT_KT >> 4);
#pragma unroll
      for (int k = 0; k < AT_D / 16; ++k) umma_bf16_lo(tmem_S, loQ + 2 * k, loK + 2 * k, idesc_s, k != 0, leader);
      umma_commit_if(s_full, leader);
      umma_commit_if(&k_empty[s], leader);       // the K stage is reusable as soon as these MMAs retire
    };
    mbar_wait(q_full, 0);
    issue_s(0);
    for (int j = 0; j < num_kv; ++j) {
      AT_TRACE(j, 8);
      if (j + 1 < num_kv) {
        mbar_wait(s_free, j & 1);              // every softmax warp holds S(j) in registers
        issue_s(j + 1);
      }
      AT_TRACE(j, 9);
      const int s = j % AT_STAGES;
      mbar_wait(p_full, j & 1);
      AT_TRACE(j, 10);
      mbar_wait(&v_full[s], (j / AT_STAGES) & 1);
      AT_TRACE(j, 11);
      tc_fence_after();
      const uint32_t loV = loV0 + s * (AT_KT >> 4);
#pragma unroll
      for (int k = 0; k < AT_BN / 16; ++k) umma_bf16_lo(tmem_O, loP + 2 * k, loV + 128 * k, idesc_o, (j | k) != 0, leader);
      umma_commit_if(pv_done, leader);
      umma_commit_if(&v_empty[s], leader);
      AT_TRACE(j, 12);
    }
  } else {
    // ===================== softmax / output warps: one thread per query row =====================
    const int quarter = warp & 3;          // TMEM lane quarter accessible to this warp
    const int r = quarter * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    float m_run = -INFINITY, l_run = 0.f;
    const int sw = r & 7;
    uint8_t* prow = sP + r * 128;
    if (warp == 2) AT_TRACE(0, 14);

    for (int j = 0; j < num_kv; ++j) {
      if (warp == 2) AT_TRACE(j, 0);
      mbar_wait(s_full, j & 1);
      if (warp == 2) AT_TRACE(j, 1);
      tc_fence_after();
      uint32_t lo[32], hi[32];
      tmem_ld_32x32(tmem_S + lane_off, lo);
      tmem_ld_32x32(tmem_S + lane_off + 32, hi);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (AT_VARIANT & 16) { if (lane == 0) mbar_arrive(s_free); } else mbar_arrive_if(s_free, lane == 0);
      if (warp == 2) AT_TRACE(j, 2);
      const int nvalid = p.Nk - j * AT_BN;   // valid keys in this tile (>= 1)
      if (nvalid < AT_BN) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (i >= nvalid) lo[i] = 0xff800000u;        // -inf: exp2 gives exactly 0
          if (i + 32 >= nvalid) hi[i] = 0xff800000u;
        }
      }
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        mx0 = fmaxf(mx0, fmaxf(__uint_as_float(lo[i]), __uint_as_float(lo[i + 1])));
        mx1 = fmaxf(mx1, fmaxf(__uint_as_float(lo[i + 2]), __uint_as_float(lo[i + 3])));
        mx2 = fmaxf(mx2, fmaxf(__uint_as_float(hi[i]), __uint_as_float(hi[i + 1])));
        mx3 = fmaxf(mx3, fmaxf(__uint_as_float(hi[i + 2]), __uint_as_float(hi[i + 3])));
      }
      const float m_new = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * p.scale_log2;
      // rare path: move the shift.  l and the TMEM-resident O row are rescaled; all earlier PV MMAs must have retired.
      // tcgen05.ld/st are warp-collective, so the branch is taken by the whole warp when ANY of its rows needs it (rows
      // that do not move use alpha = 1).
      const bool need = m_new > m_run + AT_RESCALE_THRESHOLD;
      if (__any_sync(0xffffffffu, need)) {
        const float alpha = need ? fast_exp2(m_run - m_new) : 1.0f;
        if (j > 0) {
          mbar_wait(pv_done, (j - 1) & 1);
          tc_fence_after();
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t ov[16];
            tmem_ld_32x16(tmem_O + lane_off + c * 16, ov);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
            tmem_st_32x16(tmem_O + lane_off + c * 16, ov);
          }
          tmem_st_wait();
        }
        l_run *= alpha;
        if (need) m_run = m_new;
      }
      if (warp == 2) AT_TRACE(j, 3);
      if (j >= 1) mbar_wait(pv_done, (j - 1) & 1);   // PV(j-1) has finished reading the P buffer (normally long retired by now)
      if (warp == 2) AT_TRACE(j, 4);
      // p = exp2(s*scale - m), row sum, bf16 P (A operand of the PV MMA) stored 16 bytes (8 keys) at a time, 128B-swizzled
      const uint64_t sc2 = pack_f32x2(p.scale_log2, p.scale_log2), nm2 = pack_f32x2(-m_run, -m_run);
      uint64_t acc0 = pack_f32x2(0.f, 0.f), acc1 = acc0;
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        const uint32_t* src = ch < 4 ? &lo[8 * ch] : &hi[8 * (ch - 4)];
        float e[8];
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
          const int q = ch * 4 + (i >> 1);                  // column pair 0..31 of the row
          const uint64_t x2 = fma_f32x2(pack_f32x2(__uint_as_float(src[i]), __uint_as_float(src[i + 1])), sc2, nm2);
          if (((q * AT_POLY) & 7) < AT_POLY) {
            poly_exp2_x2(x2, e[i], e[i + 1]);
          } else {
            unpack_f32x2(x2, e[i], e[i + 1]);
            e[i] = fast_exp2(e[i]);
            e[i + 1] = fast_exp2(e[i + 1]);
          }
        }
        acc0 = add_f32x2(acc0, add_f32x2(pack_f32x2(e[0], e[1]), pack_f32x2(e[2], e[3])));
        acc1 = add_f32x2(acc1, add_f32x2(pack_f32x2(e[4], e[5]), pack_f32x2(e[6], e[7])));
        *reinterpret_cast<uint4*>(prow + ((ch ^ sw) << 4)) = make_uint4(pack_bf16(e[0], e[1]), pack_bf16(e[2], e[3]), pack_bf16(e[4], e[5]), pack_bf16(e[6], e[7]));
      }
      float s0, s1;
      unpack_f32x2(add_f32x2(acc0, acc1), s0, s1);
      l_run += s0 + s1;
      if (warp == 2) AT_TRACE(j, 5);
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (warp == 2) AT_TRACE(j, 6);
      if (AT_VARIANT & 16) { if (lane == 0) mbar_arrive(p_full); } else mbar_arrive_if(p_full, lane == 0);
      if (warp == 2) AT_TRACE(j, 7);
    }
    // epilogue: all PV MMAs retired -> read the row's 64 output dims, normalise, store
    mbar_wait(pv_done, (num_kv - 1) & 1);
    tc_fence_after();
    const int qrow = q0 + r;
    const float inv = 1.0f / l_run;
    __nv_bfloat16* dst = p.O + static_cast<long>(b) * p.o_bs + static_cast<long>(qrow) * p.o_ss + static_cast<long>(h) * p.o_hs;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t ov[32];
      tmem_ld_32x32(tmem_O + lane_off + c * 32, ov);
      tmem_ld_wait();
      if (qrow < p.Nq) {
        uint4* d4 = reinterpret_cast<uint4*>(dst + c * 32);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 u;
          u.x = pack_bf16(__uint_as_float(ov[8 * i + 0]) * inv, __uint_as_float(ov[8 * i + 1]) * inv);
          u.y = pack_bf16(__uint_as_float(ov[8 * i + 2]) * inv, __uint_as_float(ov[8 * i + 3]) * inv);
          u.z = pack_bf16(__uint_as_float(ov[8 * i + 4]) * inv, __uint_as_float(ov[8 * i + 5]) * inv);
          u.w = pack_bf16(__uint_as_float(ov[8 * i + 6]) * inv, __uint_as_float(ov[8 * i + 7]) * inv);
          d4[i] = u;
        }
      }
    }
    if (qrow < p.Nq && p.lse && !(AT_VARIANT & 128)) p.lse[(static_cast<long>(b) * p.Hq + h) * p.Nq + qrow] = m_run * 0.6931471805599453f + logf(l_run);
    tc_fence_before();
    if (warp == 2) AT_TRACE(0, 15);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_S, 128);
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
    B200SAT_CHECK_CUDA(cudaFuncSetAttribute(attention_fwd_tcgen05, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    attr_set = true;
  }
  dim3 grid((Nq + AT_BM - 1) / AT_BM, Hq, B);
  B200SAT_CHECK_CUDA(launch_k(attention_fwd_tcgen05, dim3(grid), dim3(AT_THREADS), AT_SMEM, static_cast<cudaStream_t>(stream), 1, p));
  B200SAT_CHECK_CUDA(cudaGetLastError());
  return B200SAT_OK;
}

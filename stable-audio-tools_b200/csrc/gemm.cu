// b200sat — tcgen05/TMA GEMM with fused epilogues (the DiT linear layers).
//
//   D[M,N] = epilogue( A[M,K] (bf16, K contiguous)  x  B[N,K]^T (bf16, K contiguous; an nn.Linear weight) )
//
// Replaces the cuBLASLt calls behind nn.Linear in the reference
// (stable_audio_tools/models/transformer.py:263,308,356-364,481,534,747-748) and the eager elementwise kernels
// that follow them: bias add, SwiGLU (transformer.py:272-275), residual add (:704-712), the partial NeoX RoPE on
// q/k (:154-174, :491-507) and SiLU (dit.py:41-76).
//
// Structure (one persistent CTA per SM, 384 threads):
//   warp 0  lane 0 : TMA producer  — cp.async.bulk.tensor tiles of A (128x64) and B (BNx64), 128B swizzle, kStages ring
//   warp 1  lane 0 : MMA issuer    — tcgen05.mma.cta_group::1.kind::f16, 128 x BN x 16, fp32 accumulators in TMEM
//   warp 2         : TMEM allocator (2 accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1)
//   (CTAS == 2: the same roles in both CTAs of a cluster pair; only the leader CTA issues the 256-row MMAs)
//   warps 4..11    : epilogue      — tcgen05.ld 32x32b (thread == output row), fused math, 16-byte global stores;
//                    two warps per SM sub-partition because a lone warp cannot hide its own ALU latency (ncu r1: the
//                    4-warp epilogue ran at IPC 0.1 and starved the MMA issuer on the tmem_empty barrier)
#include "common.cuh"
#include <cstring>

namespace b200sat {

enum GemmFlags : int {
  GEMM_BIAS = 1,        // + bias[n] (fp32)
  GEMM_RESIDUAL = 2,    // out = residual + bf16(acc + bias)
  GEMM_SILU = 4,        // out = silu(acc + bias)
  GEMM_SWIGLU = 8,      // out[:, j] = u[:, j] * silu(u[:, j + n_half]);  B rows j and j+n_half share one tile
  GEMM_ROPE = 16,       // q/k halves of a fused qkv projection get the partial rotary embedding
  GEMM_OUT_F32 = 32,    // fp32 output instead of bf16
  GEMM_ROW_REMAP = 64,  // out_row = (r / seg_in) * seg_out + seg_off + r % seg_in
  GEMM_GATE = 128,      // out = residual + bf16(acc+bias) * gate[b, n]   (adaLN: gate = sigmoid(1 - g), fp32 [B, N])
  GEMM_A_MN = 256,      // A is stored [K, M] (M contiguous): the transposed operand of a weight-gradient GEMM
  GEMM_B_MN = 512,      // B is stored [K, N] (N contiguous): data-gradient GEMM through an nn.Linear weight [N_out=K, N]
  GEMM_ACCUM = 1024,    // fp32 output accumulates: D += acc (gradient accumulation into .grad)
  GEMM_SWIGLU_BWD = 2048,  // D[M, 2*N]: [dact*silu(g) | dact*a*silu'(g)] with (a|g) read from aux [M, 2*N] (n_half = N)
  GEMM_ATOMIC = 4096,
  GEMM_LN_A = 8192,        // A is the RAW LayerNorm input and B has gamma folded in: out = rstd_m*(acc - mean_m*colsum_n) (+bias ...)
  GEMM_ROWSTATS = 16384,   // also accumulate per-row (sum, sum of squares) of the bf16 output into out_stats (next LayerNorm)      // fp32 D += acc with red.global.add (split-K weight gradients: several CTAs own one output tile)
};

struct GemmParams {
  CUtensorMap tmA;
  CUtensorMap tmB;
  void* D;
  const float* bias;
  const __nv_bfloat16* residual;
  const float* rope_cos;  // [rope_seq, 16]
  const float* rope_sin;
  const float* gate;      // [B, N] fp32 (adaLN)
  const float* ln_stats;   // [M, 2] (sum x, sum x^2) over the K features of each A row
  const float* ln_colsum;  // [N] (SwiGLU: [n_half + N]) sum_k B[n, k] of the gamma-folded weight
  float* out_stats;        // [rows, 2]
  float ln_eps;
  __nv_bfloat16* aux;     // SwiGLU fwd: optional copy of the pre-activation u [M, 2*n_half]; SwiGLU bwd: the saved u
  int ld_aux;
  int M, N, K;
  int ldd, ldr;
  int flags;
  int seg_in, seg_out, seg_off;
  int rope_seq, rope_dmodel, rope_dh;
  int n_half;
  int num_m_tiles, num_n_tiles;
  int ksplit, kb_per_split;  // split-K: tile index -> (k-slice, n, m); each slice covers kb_per_split k-blocks
  // conv weight-gradient addressing (both operands are time-major activation planes behind 4-D maps {C, s, T/s, B}): the K loop walks
  // (item, 64-step time block); each operand is read at its own (phase r, row offset) = one tap of the convolution
  int wg_kb_per_item, wg_rA, wg_offA, wg_rB, wg_offB;
  // all taps of a flattened 2-D conv weight gradient in ONE launch: the N extent is (tap, 64 channels) and every 64-column block of a B
  // tile is the same 64-channel plane read at that tap's row shift, so the dY tile is staged once per BN/64 taps instead of once per tap
  int wg_ntaps;
  int wg_tap_off[32];
};

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;   // 64 bf16 = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;

// CTAS == 2: a CTA pair (cluster of 2 on one TPC) computes a 256 x BN tile with tcgen05.mma.cta_group::2 — each CTA
// stages its own 128 rows of A and HALF of the B tile, so the L2 -> SMEM bytes per MMA flop drop by 1.5x (the 1-CTA kernel
// is L2-bandwidth bound: 96 B/clk/SM needed vs ~43 B/clk/SM available chip-wide).
template <int BN, int CTAS>
struct GemmCfg {
  static constexpr int kABytes = BLOCK_M * BLOCK_K * 2;
  static constexpr int kBBytes = (BN / CTAS) * BLOCK_K * 2;   // per CTA
  static constexpr int kStageBytes = kABytes + kBBytes;
  // ring depth: as many stages as 224 KB hold (<= 8).  A stage feeds 4 MMAs = 384-512 tensor-pipe cycles, so the previous 192 KB budget
  // (6 stages at BN = 192 / 256) gave a load ~2k cycles of lead - about one L2 -> shared-memory TMA latency
  static constexpr int kStages = (224 * 1024) / kStageBytes > 8 ? 8 : (224 * 1024) / kStageBytes;
  static constexpr int kAccStride = BN <= 64 ? 64 : (BN <= 128 ? 128 : 256);  // TMEM columns per accumulator stage
  static constexpr int kTmemCols = 2 * kAccStride;                            // two stages, power of two
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

__device__ __forceinline__ void store_chunk_bf16(__nv_bfloat16* dst, const float (&o)[32], int ncols_valid) {
  if (ncols_valid >= 32) {
    uint4* p = reinterpret_cast<uint4*>(dst);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 u;
      u.x = pack_bf16(o[8 * i + 0], o[8 * i + 1]);
      u.y = pack_bf16(o[8 * i + 2], o[8 * i + 3]);
      u.z = pack_bf16(o[8 * i + 4], o[8 * i + 5]);
      u.w = pack_bf16(o[8 * i + 6], o[8 * i + 7]);
      p[i] = u;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) if (i < ncols_valid) dst[i] = __float2bfloat16_rn(o[i]);
  }
}
__device__ __forceinline__ void store_chunk_f32(float* dst, const float (&o)[32], int ncols_valid) {
  if (ncols_valid >= 32) {
    float4* p = reinterpret_cast<float4*>(dst);
#pragma unroll
    for (int i = 0; i < 8; ++i) p[i] = make_float4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) if (i < ncols_valid) dst[i] = o[i];
  }
}

template <int BN, int CTAS>
__global__ void __launch_bounds__(384, 1) gemm_bf16_tcgen05(const __grid_constant__ GemmParams p) {
  using Cfg = GemmCfg<BN, CTAS>;
  const uint32_t cta_rank = (CTAS == 2) ? cluster_ctarank() : 0u;
  const int unit = (CTAS == 2) ? (blockIdx.x >> 1) : blockIdx.x;          // scheduling unit = CTA or CTA pair
  const int num_units = (CTAS == 2) ? (gridDim.x >> 1) : gridDim.x;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + Cfg::kStages;
  uint64_t* tmem_full = bars + 2 * Cfg::kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int mn_tiles = p.num_m_tiles * p.num_n_tiles;
  const int num_tiles = mn_tiles * p.ksplit;
  const int total_kb = (p.K + BLOCK_K - 1) / BLOCK_K;
  const bool swiglu = (p.flags & GEMM_SWIGLU) != 0;
  const bool a_mn = (p.flags & GEMM_A_MN) != 0;
  const bool b_mn = (p.flags & GEMM_B_MN) != 0;
  // Output columns handled per tile (SwiGLU folds value|gate halves of the tile into BN/2 outputs).
  const int out_bn = swiglu ? BN / 2 : BN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < Cfg::kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 256 * CTAS);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if (CTAS == 2) { tmem_alloc_pair(tmem_ptr_smem, Cfg::kTmemCols); tmem_relinquish_pair(); }
    else { tmem_alloc(tmem_ptr_smem, Cfg::kTmemCols); tmem_relinquish(); }
  }
  griddep_launch();
  tc_fence_before();
  if (CTAS == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  griddep_wait();   // everything above overlapped the previous kernel's tail; global memory is touched only below
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = unit; tile < num_tiles; tile += num_units) {
        const int mn = tile % mn_tiles, ks = tile / mn_tiles;
        const int m_blk = mn % p.num_m_tiles;
        const int n_blk = mn / p.num_m_tiles;
        const int m0 = m_blk * (BLOCK_M * CTAS) + cta_rank * BLOCK_M;
        const int kb_begin = ks * p.kb_per_split, kb_end = min(total_kb, kb_begin + p.kb_per_split);
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          if (p.wg_kb_per_item > 0) {
            if (CTAS == 1 || cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], CTAS * Cfg::kStageBytes);
            const int item = kb / p.wg_kb_per_item, t0 = (kb % p.wg_kb_per_item) * BLOCK_K;
            const int nb0 = n_blk * BN + static_cast<int>(cta_rank) * (BN / CTAS);
#pragma unroll
            for (int i = 0; i < BLOCK_M / 64; ++i) {
              if (CTAS == 2) tma_load_4d_pair(sa + i * 8192, &p.tmA, &full_bar[stage], m0 + 64 * i, p.wg_rA, t0 + p.wg_offA, item);
              else tma_load_4d(sa + i * 8192, &p.tmA, &full_bar[stage], m0 + 64 * i, p.wg_rA, t0 + p.wg_offA, item);
            }
#pragma unroll
            for (int i = 0; i < BN / CTAS / 64; ++i) {
              int chan = nb0 + 64 * i, offB = p.wg_offB;
              if (p.wg_ntaps > 0) {   // column block -> tap (blocks past the last tap re-read it; the epilogue never stores them)
                offB = p.wg_tap_off[min(chan >> 6, p.wg_ntaps - 1)];
                chan = 0;
              }
              if (CTAS == 2) tma_load_4d_pair(sb + i * 8192, &p.tmB, &full_bar[stage], chan, p.wg_rB, t0 + offB, item);
              else tma_load_4d(sb + i * 8192, &p.tmB, &full_bar[stage], chan, p.wg_rB, t0 + offB, item);
            }
          } else if (a_mn || b_mn) {
            // MN-major operands: 64(mn) x 64(k) boxes, one per 64-wide block of the M / N extent (8 KB each)
            if (CTAS == 1 || cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], CTAS * Cfg::kStageBytes);
            if (a_mn) {
#pragma unroll
              for (int i = 0; i < BLOCK_M / 64; ++i) {
                if (CTAS == 2) tma_load_2d_pair(sa + i * 8192, &p.tmA, &full_bar[stage], m0 + 64 * i, kb * BLOCK_K);
                else tma_load_2d(sa + i * 8192, &p.tmA, &full_bar[stage], m0 + 64 * i, kb * BLOCK_K);
              }
            } else {
              if (CTAS == 2) tma_load_2d_pair(sa, &p.tmA, &full_bar[stage], kb * BLOCK_K, m0);
              else tma_load_2d(sa, &p.tmA, &full_bar[stage], kb * BLOCK_K, m0);
            }
            const int nb0 = n_blk * BN + static_cast<int>(cta_rank) * (BN / CTAS);
            if (b_mn) {
#pragma unroll
              for (int i = 0; i < BN / CTAS / 64; ++i) {
                if (CTAS == 2) tma_load_2d_pair(sb + i * 8192, &p.tmB, &full_bar[stage], nb0 + 64 * i, kb * BLOCK_K);
                else tma_load_2d(sb + i * 8192, &p.tmB, &full_bar[stage], nb0 + 64 * i, kb * BLOCK_K);
              }
            } else {
              if (CTAS == 2) tma_load_2d_pair(sb, &p.tmB, &full_bar[stage], kb * BLOCK_K, nb0);
              else tma_load_2d(sb, &p.tmB, &full_bar[stage], kb * BLOCK_K, nb0);
            }
          } else if (CTAS == 1) {
            mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
            tma_load_2d(sa, &p.tmA, &full_bar[stage], kb * BLOCK_K, m0);
            if (swiglu) {
              tma_load_2d(sb, &p.tmB, &full_bar[stage], kb * BLOCK_K, n_blk * (BN / 2));
              tma_load_2d(sb + Cfg::kBBytes / 2, &p.tmB, &full_bar[stage], kb * BLOCK_K, p.n_half + n_blk * (BN / 2));
            } else {
              tma_load_2d(sb, &p.tmB, &full_bar[stage], kb * BLOCK_K, n_blk * BN);
            }
          } else {
            // both CTAs load; all bytes are credited to the leader's barrier, which expects the pair's total
            if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
            tma_load_2d_pair(sa, &p.tmA, &full_bar[stage], kb * BLOCK_K, m0);
            const int brow = swiglu ? (cta_rank == 0 ? n_blk * (BN / 2) : p.n_half + n_blk * (BN / 2))
                                    : n_blk * BN + static_cast<int>(cta_rank) * (BN / 2);
            tma_load_2d_pair(sb, &p.tmB, &full_bar[stage], kb * BLOCK_K, brow);
          }
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (cta_rank == 0) {
      // ===================== MMA issuer (leader CTA only in pair mode) =====================
      // The whole warp walks the loop in uniform control flow and one elected lane issues (umma_bf16_lo in common.cuh): the
      // descriptor low words live in uniform registers and advance by 32-bit adds.  Before, ~25 SASS instructions of 64-bit
      // descriptor arithmetic + a waterfall loop separated consecutive MMAs - more than the 96-cycle 256x192x16 pair MMA itself.
      const uint32_t leader = elect_one() ? 1u : 0u;
      const uint32_t idesc = make_idesc_bf16(BLOCK_M * CTAS, BN, a_mn ? 1u : 0u, b_mn ? 1u : 0u);
      // per UMMA_K step: K-major advances 32 B inside the swizzle row; MN-major advances 16 k-rows = 2048 B
      const uint32_t a_step = a_mn ? (2048u >> 4) : 2u, b_step = b_mn ? (2048u >> 4) : 2u;
      const uint32_t a_hi = (a_mn ? (8192u >> 4) : 1u) << 16, b_hi = (b_mn ? (8192u >> 4) : 1u) << 16;
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = unit; tile < num_tiles; tile += num_units) {
        mbar_wait(&tmem_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * Cfg::kAccStride;
        const int ks = tile / mn_tiles;
        const int kb_begin = ks * p.kb_per_split, kb_end = min(total_kb, kb_begin + p.kb_per_split);
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint32_t la = (sa >> 4) | a_hi;
          const uint32_t lb = ((sa + Cfg::kABytes) >> 4) | b_hi;
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            if (CTAS == 2) umma_bf16_pair_lo(tmem_d, la + a_step * k, lb + b_step * k, idesc, (kb != kb_begin) || (k != 0), leader);
            else umma_bf16_lo(tmem_d, la + a_step * k, lb + b_step * k, idesc, (kb != kb_begin) || (k != 0), leader);
          }
          if (CTAS == 2) umma_commit_pair_if(&empty_bar[stage], leader); else umma_commit_if(&empty_bar[stage], leader);
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        if (CTAS == 2) umma_commit_pair_if(&tmem_full[as], leader); else umma_commit_if(&tmem_full[as], leader);
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (8 warps: 2 per SM sub-partition, each owns half of the tile's columns) ===========
    const int ew = warp - 4;
    const int q = ew & 3;      // TMEM lane quarter this warp may access (warp % 4)
    const int half = ew >> 2;  // column half
    const int chunks_per_warp = (out_bn / 32 + 1) / 2;
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = unit; tile < num_tiles; tile += num_units) {
      const int mn = tile % mn_tiles;
      const int m_blk = mn % p.num_m_tiles;
      const int n_blk = mn / p.num_m_tiles;
      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after();
      const int row = m_blk * (BLOCK_M * CTAS) + static_cast<int>(cta_rank) * BLOCK_M + q * 32 + lane;
      const bool row_ok = row < p.M;
      int out_row = row;
      if (p.flags & GEMM_ROW_REMAP) out_row = (row / p.seg_in) * p.seg_out + p.seg_off + (row % p.seg_in);
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * Cfg::kAccStride;
      const int n0 = n_blk * out_bn;
      float ln_mean = 0.f, ln_rstd = 1.f, st1 = 0.f, st2 = 0.f;
      if ((p.flags & GEMM_LN_A) && row_ok) {
        const float2 st = __ldg(reinterpret_cast<const float2*>(p.ln_stats) + row);
        ln_mean = st.x / p.K;
        ln_rstd = rsqrtf(fmaxf(st.y / p.K - ln_mean * ln_mean, 0.f) + p.ln_eps);
      }
      for (int cc = 0; cc < chunks_per_warp; ++cc) {
        const int c = half * chunks_per_warp + cc;
        const int col = n0 + c * 32;
        if (c * 32 >= out_bn || col >= p.N) break;  // warp-uniform
        uint32_t raw[32];
        float v[32];
        tmem_ld_32x32(taddr + c * 32, raw);
        const int ncols = min(32, p.N - col);
        const bool full = ncols == 32;
        float bv[32];
        if (p.flags & GEMM_BIAS) {
          if (full) {
            const float4* bp = reinterpret_cast<const float4*>(p.bias + col);
#pragma unroll
            for (int i = 0; i < 8; ++i) { const float4 t = __ldg(bp + i); bv[4 * i] = t.x; bv[4 * i + 1] = t.y; bv[4 * i + 2] = t.z; bv[4 * i + 3] = t.w; }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) bv[i] = (i < ncols) ? __ldg(p.bias + col + i) : 0.f;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) bv[i] = 0.f;
        }
        tmem_ld_wait();
        if (p.flags & GEMM_LN_A) {
          // LayerNorm folded into the GEMM: acc = x_raw . (gamma o W)^T  =>  LN(x).W^T = rstd*(acc - mean*colsum)
          const float4* cp = reinterpret_cast<const float4*>(p.ln_colsum + col);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 c4 = full ? __ldg(cp + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            raw[4 * i + 0] = __float_as_uint(ln_rstd * (__uint_as_float(raw[4 * i + 0]) - ln_mean * c4.x));
            raw[4 * i + 1] = __float_as_uint(ln_rstd * (__uint_as_float(raw[4 * i + 1]) - ln_mean * c4.y));
            raw[4 * i + 2] = __float_as_uint(ln_rstd * (__uint_as_float(raw[4 * i + 2]) - ln_mean * c4.z));
            raw[4 * i + 3] = __float_as_uint(ln_rstd * (__uint_as_float(raw[4 * i + 3]) - ln_mean * c4.w));
          }
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]) + bv[i];
        if (swiglu) {
          // value * silu(gate); the gate half of the tile sits BN/2 accumulator columns further
          tmem_ld_32x32(taddr + BN / 2 + c * 32, raw);
          if (p.flags & GEMM_BIAS) {
            const float4* bp = reinterpret_cast<const float4*>(p.bias + p.n_half + col);
#pragma unroll
            for (int i = 0; i < 8; ++i) { const float4 t = __ldg(bp + i); bv[4 * i] = t.x; bv[4 * i + 1] = t.y; bv[4 * i + 2] = t.z; bv[4 * i + 3] = t.w; }
          }
          tmem_ld_wait();
          if (p.flags & GEMM_LN_A) {
            const float4* cp = reinterpret_cast<const float4*>(p.ln_colsum + p.n_half + col);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 c4 = __ldg(cp + i);
              raw[4 * i + 0] = __float_as_uint(ln_rstd * (__uint_as_float(raw[4 * i + 0]) - ln_mean * c4.x));
              raw[4 * i + 1] = __float_as_uint(ln_rstd * (__uint_as_float(raw[4 * i + 1]) - ln_mean * c4.y));
              raw[4 * i + 2] = __float_as_uint(ln_rstd * (__uint_as_float(raw[4 * i + 2]) - ln_mean * c4.z));
              raw[4 * i + 3] = __float_as_uint(ln_rstd * (__uint_as_float(raw[4 * i + 3]) - ln_mean * c4.w));
            }
          }
          float gv[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) gv[i] = __uint_as_float(raw[i]) + bv[i];
          if (p.aux && row_ok) {  // training forward: keep the pre-activation (value | gate) for the backward
            __nv_bfloat16* ua = p.aux + static_cast<size_t>(row) * p.ld_aux + col;
            store_chunk_bf16(ua, v, 32);
            store_chunk_bf16(ua + p.n_half, gv, 32);
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = v[i] * silu_f(gv[i]);
        }
        if ((p.flags & GEMM_SWIGLU_BWD) && row_ok && full) {
          // v = d(act); saved u = (a | g):  d a = v*silu(g),  d g = v*a*sigmoid(g)*(1 + g*(1 - sigmoid(g)))
          const __nv_bfloat16* ua = p.aux + static_cast<size_t>(row) * p.ld_aux + col;
          float da[32], dg[32];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint4 au = __ldg(reinterpret_cast<const uint4*>(ua) + i);
            const uint4 gu = __ldg(reinterpret_cast<const uint4*>(ua + p.n_half) + i);
            const uint32_t aw[4] = {au.x, au.y, au.z, au.w}, gw[4] = {gu.x, gu.y, gu.z, gu.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 af = unpack_bf16(aw[j]), gf = unpack_bf16(gw[j]);
              const float a2[2] = {af.x, af.y}, g2[2] = {gf.x, gf.y};
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int idx = 8 * i + 2 * j + e;
                const float sg = __fdividef(1.0f, 1.0f + __expf(-g2[e]));
                da[idx] = v[idx] * g2[e] * sg;
                dg[idx] = v[idx] * a2[e] * sg * (1.0f + g2[e] * (1.0f - sg));
              }
            }
          }
          __nv_bfloat16* dd = reinterpret_cast<__nv_bfloat16*>(p.D) + static_cast<size_t>(out_row) * p.ldd + col;
          store_chunk_bf16(dd, da, 32);
          store_chunk_bf16(dd + p.n_half, dg, 32);
          continue;
        }
        if (p.flags & GEMM_SILU) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = silu_f(v[i]);
        }
        if ((p.flags & GEMM_ROPE) && row_ok) {
          // column -> (which in {q,k,v}, head, dim); rotate dims [0,32) of q and k heads (NeoX half-split, 16 freqs).
          // The reference rotates the bf16 projection output in fp32 (transformer.py:491-507): round first.
          const int which = col / p.rope_dmodel;
          const int dim0 = (col % p.rope_dmodel) % p.rope_dh;
          if (which < 2 && dim0 == 0) {
            const int pos = row % p.rope_seq;
            const float4* cs = reinterpret_cast<const float4*>(p.rope_cos + pos * 16);
            const float4* sn = reinterpret_cast<const float4*>(p.rope_sin + pos * 16);
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4) {
              const float4 c4 = __ldg(cs + i4), s4 = __ldg(sn + i4);
              const float cc_[4] = {c4.x, c4.y, c4.z, c4.w}, ss_[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int i = 4 * i4 + j;
                const float x1 = bf16_round(v[i]), x2 = bf16_round(v[i + 16]);
                v[i] = x1 * cc_[j] - x2 * ss_[j];
                v[i + 16] = x2 * cc_[j] + x1 * ss_[j];
              }
            }
          }
        }
        if (row_ok) {
          if (p.flags & GEMM_RESIDUAL) {
            const __nv_bfloat16* r = p.residual + static_cast<size_t>(out_row) * p.ldr + col;
            if (p.flags & GEMM_GATE) {
              const float* gp = p.gate + static_cast<size_t>(row / p.seg_in) * p.N + col;
              if (p.aux) {   // training forward: keep the un-gated branch output (its product with dh is the gate gradient)
                __nv_bfloat16* ua = p.aux + static_cast<size_t>(row) * p.ld_aux + col;
#pragma unroll
                for (int i = 0; i < 32; ++i) if (i < ncols) ua[i] = __float2bfloat16_rn(v[i]);
              }
#pragma unroll
              for (int i = 0; i < 32; ++i) if (i < ncols) v[i] = bf16_round(v[i]) * __ldg(gp + i);
            }
            if (full) {
              const uint4* rp = reinterpret_cast<const uint4*>(r);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const uint4 u = __ldg(rp + i);
                const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float2 f = unpack_bf16(w[j]);
                  // the reference rounds the branch output to bf16 before the residual add (transformer.py:704-712)
                  v[8 * i + 2 * j] = bf16_round(v[8 * i + 2 * j]) + f.x;
                  v[8 * i + 2 * j + 1] = bf16_round(v[8 * i + 2 * j + 1]) + f.y;
                }
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) if (i < ncols) v[i] = bf16_round(v[i]) + __bfloat162float(r[i]);
            }
          }
          if (p.flags & GEMM_ROWSTATS) {
#pragma unroll
            for (int i = 0; i < 32; ++i) if (i < ncols) { const float r_ = bf16_round(v[i]); st1 += r_; st2 += r_ * r_; }
          }
          if (p.flags & GEMM_OUT_F32) {
            float* dp = reinterpret_cast<float*>(p.D) + static_cast<size_t>(out_row) * p.ldd + col;
            if (p.flags & GEMM_ATOMIC) {
              if (full) {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dp + 4 * i), "f"(v[4 * i]), "f"(v[4 * i + 1]), "f"(v[4 * i + 2]), "f"(v[4 * i + 3]) : "memory");
              } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) if (i < ncols) atomicAdd(dp + i, v[i]);
              }
              continue;
            }
            if (p.flags & GEMM_ACCUM) {
#pragma unroll
              for (int i = 0; i < 32; ++i) if (i < ncols) v[i] += dp[i];
            }
            store_chunk_f32(dp, v, ncols);
          } else {
            store_chunk_bf16(reinterpret_cast<__nv_bfloat16*>(p.D) + static_cast<size_t>(out_row) * p.ldd + col, v, ncols);
          }
        }
      }
      if ((p.flags & GEMM_ROWSTATS) && row_ok) {
        atomicAdd(p.out_stats + 2 * static_cast<size_t>(out_row), st1);
        atomicAdd(p.out_stats + 2 * static_cast<size_t>(out_row) + 1, st2);
      }
      tc_fence_before();
      if (CTAS == 2 && cta_rank != 0) mbar_arrive_remote(&tmem_empty[as], 0);
      else mbar_arrive(&tmem_empty[as]);
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  }

  tc_fence_before();
  if (CTAS == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if (CTAS == 2) tmem_dealloc_pair(tmem_base, Cfg::kTmemCols); else tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

template <int BN, int CTAS>
static int launch_gemm(GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, CTAS>;
  static bool attr_set = false;
  if (!attr_set) {
    B200SAT_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_tcgen05<BN, CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  const int out_bn = (p.flags & GEMM_SWIGLU) ? BN / 2 : BN;
  p.num_m_tiles = (p.M + BLOCK_M * CTAS - 1) / (BLOCK_M * CTAS);
  p.num_n_tiles = (p.N + out_bn - 1) / out_bn;
  const int units = num_sms() / CTAS;
  const int total_kb = (p.K + BLOCK_K - 1) / BLOCK_K;
  p.ksplit = 1;
  if ((p.flags & GEMM_ACCUM) && (p.flags & GEMM_OUT_F32)) {
    // gradient accumulation with few output tiles and a long K (tokens): split K across idle SMs, combine with fp32 reds
    const int mn = p.num_m_tiles * p.num_n_tiles;
    int ks = units / mn;
    // dense-layer wgrads: at most 8 slices (reduction traffic); conv wgrads have tiny outputs ([C,C] per tap) and a K extent of
    // batch x time, so every SM takes a slice
    if (ks > 8 && p.wg_kb_per_item == 0) ks = 8;
    while (ks > 1 && total_kb / ks < 16) --ks;
    if (ks > 1) { p.ksplit = ks; p.flags = (p.flags & ~GEMM_ACCUM) | GEMM_ATOMIC; }
  }
  p.kb_per_split = (total_kb + p.ksplit - 1) / p.ksplit;
  p.ksplit = (total_kb + p.kb_per_split - 1) / p.kb_per_split;  // no empty slices
  const int tiles = p.num_m_tiles * p.num_n_tiles * p.ksplit;
  const int grid = (tiles < units ? tiles : units) * CTAS;
  B200SAT_CHECK_CUDA(launch_k(gemm_bf16_tcgen05<BN, CTAS>, dim3(grid), dim3(384), Cfg::kSmemBytes, stream, CTAS, p));
  return B200SAT_OK;
}

// Tile-shape heuristic: minimise waves x per-tile time over {CTA-pair 256/192/128, single-CTA 128/64}.
// Efficiencies are relative MMA-pipe rates measured with tools/gemm_sweep.py on B200 (profiles/).
static void pick_config(int M, int N, bool swiglu, bool mn_major, int* bn_out, int* ctas_out) {
  if (swiglu) { *bn_out = 256; *ctas_out = M > 128 ? 2 : 1; return; }
  // transposed-operand (gradient) GEMMs: the 256x256 pair tile is the only shape that stays MMA-bound with MN-major boxes
  // (measured: 256x128 pair tiles reach 0.6 PF, 256x256 1.24 PF); low tile counts are handled by split-K instead
  if (mn_major && M > 128 && N > 128) { *bn_out = 256; *ctas_out = 2; return; }
  const int sms = num_sms();
  struct Cand { int bn, ctas; double eff; };
  const Cand cands[5] = {{256, 2, 1.00}, {192, 2, 0.85}, {128, 2, 0.62}, {128, 1, 0.55}, {64, 1, 0.32}};
  double best = 1e30;
  *bn_out = 128; *ctas_out = 1;
  for (int i = 0; i < 5; ++i) {
    const Cand& c = cands[i];
    if (c.ctas == 2 && M <= 128) continue;
    if (c.bn == 192 && mn_major) continue;  // 96-column halves are not a whole number of 64-wide MN-major boxes
    if (c.bn > 64 && N <= c.bn / 2) continue;
    const long mt = (M + 128 * c.ctas - 1) / (128 * c.ctas);
    const long nt = (N + c.bn - 1) / c.bn;
    const long units = sms / c.ctas;
    const long waves = (mt * nt + units - 1) / units;
    const double cost = waves * (c.bn / c.eff);
    if (cost < best) { best = cost; *bn_out = c.bn; *ctas_out = c.ctas; }
  }
}

}  // namespace b200sat

using namespace b200sat;

// C-ABI — see include/b200sat.h
extern "C" int b200sat_gemm_bf16(const void* A, int lda, const void* B, int ldb, void* D, int ldd, int M, int N, int K,
                                 int flags, const float* bias, const void* residual, int ldr, const float* rope_cos,
                                 const float* rope_sin, int rope_seq, int rope_dmodel, int rope_dh, int n_half, int seg_in,
                                 int seg_out, int seg_off, const float* gate, void* aux, int ld_aux, const float* ln_stats,
                                 const float* ln_colsum, float ln_eps, float* out_stats, int force_bn, void* stream) {
  if (!A || !B || !D || M <= 0 || N <= 0 || K <= 0) { set_last_error("gemm: null pointer or empty shape"); return B200SAT_EINVAL; }
  if ((lda % 8) || (ldb % 8)) { set_last_error("gemm: lda/ldb must be multiples of 8 (16-byte TMA strides)"); return B200SAT_EINVAL; }
  if ((K % 8) && !(flags & (GEMM_A_MN | GEMM_B_MN))) { set_last_error("gemm: K must be a multiple of 8"); return B200SAT_EINVAL; }
  if ((flags & GEMM_LN_A) && (!ln_stats || !ln_colsum || (N % 32))) { set_last_error("gemm: LN_A needs row statistics, column sums and N % 32 == 0"); return B200SAT_EINVAL; }
  if ((flags & GEMM_ROWSTATS) && !out_stats) { set_last_error("gemm: ROWSTATS needs out_stats"); return B200SAT_EINVAL; }
  if ((flags & GEMM_SWIGLU_BWD) && (!aux || n_half != N || (N % 32))) { set_last_error("gemm: swiglu_bwd needs aux and n_half == N"); return B200SAT_EINVAL; }
  if ((flags & (GEMM_A_MN | GEMM_B_MN)) && (flags & GEMM_SWIGLU)) { set_last_error("gemm: swiglu with MN-major operands"); return B200SAT_EUNSUPPORTED; }
  if ((ldd % 8) || ((flags & GEMM_RESIDUAL) && (ldr % 8))) { set_last_error("gemm: ldd/ldr must be multiples of 8"); return B200SAT_EINVAL; }
  if ((flags & GEMM_SWIGLU) && (N % 128 || n_half <= 0)) { set_last_error("gemm: swiglu needs N % 128 == 0 and n_half"); return B200SAT_EINVAL; }
  if ((flags & GEMM_ROPE) && (!rope_cos || !rope_sin || rope_dh != 64 || rope_seq <= 0)) { set_last_error("gemm: rope needs tables, dh == 64"); return B200SAT_EINVAL; }
  if ((flags & GEMM_GATE) && (!gate || !(flags & GEMM_RESIDUAL) || seg_in <= 0)) { set_last_error("gemm: gate needs residual + seg_in"); return B200SAT_EINVAL; }
  GemmParams p;
  memset(&p, 0, sizeof(p));
  const bool swiglu = flags & GEMM_SWIGLU;
  // force_bn: 0 = heuristic; 64/128/256 = 1-CTA tile width; 2128/2256 = CTA-pair (cta_group::2) with BN 128/256
  int ctas = 1, bn;
  if (force_bn >= 2000) { ctas = 2; bn = force_bn - 2000; }
  else if (force_bn) bn = force_bn;
  else pick_config(M, N, swiglu, (flags & (GEMM_A_MN | GEMM_B_MN)) != 0, &bn, &ctas);
  if (bn != 64 && bn != 128 && bn != 256 && !(bn == 192 && ctas == 2)) { set_last_error("gemm: force_bn must be 64/128/256/2128/2192/2256"); return B200SAT_EINVAL; }
  if (ctas == 2 && bn == 64) { set_last_error("gemm: pair mode needs BN >= 128"); return B200SAT_EINVAL; }
  if (bn == 192 && (flags & (GEMM_A_MN | GEMM_B_MN))) { set_last_error("gemm: BN=192 is K-major only"); return B200SAT_EUNSUPPORTED; }
  if (swiglu && bn != 256) { set_last_error("gemm: swiglu requires BN=256"); return B200SAT_EINVAL; }
  const int b_rows_total = swiglu ? 2 * N : N;  // value rows [0,N) and gate rows [n_half, n_half+N)
  if (flags & GEMM_A_MN) {  // A stored [K, M]
    uint64_t dims[2] = {static_cast<uint64_t>(M), static_cast<uint64_t>(K)};
    uint64_t strides[1] = {static_cast<uint64_t>(lda) * 2};
    uint32_t box[2] = {64, BLOCK_K};
    int rc = encode_tmap_bf16(&p.tmA, A, 2, dims, strides, box, 1);
    if (rc) return rc;
  } else {
    uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(M)};
    uint64_t strides[1] = {static_cast<uint64_t>(lda) * 2};
    uint32_t box[2] = {BLOCK_K, BLOCK_M};
    int rc = encode_tmap_bf16(&p.tmA, A, 2, dims, strides, box, 1);
    if (rc) return rc;
  }
  if (flags & GEMM_B_MN) {  // B stored [K, N]
    uint64_t dims[2] = {static_cast<uint64_t>(N), static_cast<uint64_t>(K)};
    uint64_t strides[1] = {static_cast<uint64_t>(ldb) * 2};
    uint32_t box[2] = {64, BLOCK_K};
    int rc = encode_tmap_bf16(&p.tmB, B, 2, dims, strides, box, 1);
    if (rc) return rc;
  } else {
    uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(swiglu ? n_half + N : b_rows_total)};
    uint64_t strides[1] = {static_cast<uint64_t>(ldb) * 2};
    uint32_t box[2] = {BLOCK_K, static_cast<uint32_t>((swiglu || ctas == 2) ? bn / 2 : bn)};
    int rc = encode_tmap_bf16(&p.tmB, B, 2, dims, strides, box, 1);
    if (rc) return rc;
  }
  p.D = D; p.bias = bias; p.residual = static_cast<const __nv_bfloat16*>(residual);
  p.rope_cos = rope_cos; p.rope_sin = rope_sin; p.gate = gate;
  p.aux = static_cast<__nv_bfloat16*>(aux); p.ld_aux = ld_aux;
  p.ln_stats = ln_stats; p.ln_colsum = ln_colsum; p.ln_eps = ln_eps; p.out_stats = out_stats;
  p.M = M; p.N = N; p.K = K; p.ldd = ldd; p.ldr = ldr; p.flags = flags;
  p.seg_in = seg_in > 0 ? seg_in : 1; p.seg_out = seg_out; p.seg_off = seg_off;
  p.rope_seq = rope_seq > 0 ? rope_seq : 1; p.rope_dmodel = rope_dmodel > 0 ? rope_dmodel : 1; p.rope_dh = rope_dh > 0 ? rope_dh : 64;
  p.n_half = n_half;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (ctas == 2) return bn == 256 ? launch_gemm<256, 2>(p, s) : (bn == 192 ? launch_gemm<192, 2>(p, s) : launch_gemm<128, 2>(p, s));
  switch (bn) {
    case 256: return launch_gemm<256, 1>(p, s);
    case 128: return launch_gemm<128, 1>(p, s);
    default: return launch_gemm<64, 1>(p, s);
  }
}


// Weight gradient of a 1-D convolution tap: dW[Ca, Cb] += sum over items b and time steps t < T_iter of
//     A[b, (t + offA) * sA + rA, :]^T  (x)  B[b, (t + offB) * sB + rB, :]
// with A, B time-major bf16 activation planes [B, T, C] read in place through 4-D tensor maps (rows outside [0, T/s) are
// zero-filled = the convolution's zero padding).  One launch per tap; split-K over time with fp32 reds fills the machine.
// Backward of the Conv1d / ConvTranspose1d weights of models/autoencoders.py:58-83, :233-283 (dW = dY (*) X).
extern "C" int b200sat_conv_wgrad(const void* a_plane, int Ca, int Ta, int sA, int rA, int offA, const void* b_plane, int Cb, int Tb,
                                  int sB, int rB, int offB, float* dW, int B, int T_iter, void* stream) {
  if (!a_plane || !b_plane || !dW || B <= 0 || T_iter <= 0 || sA < 1 || sB < 1) { set_last_error("conv_wgrad: bad arguments"); return B200SAT_EINVAL; }
  if (Ca % 64 || Cb % 64 || Ta % sA || Tb % sB) { set_last_error("conv_wgrad: channels must be multiples of 64 and T of the stride"); return B200SAT_EUNSUPPORTED; }
  GemmParams p;
  memset(&p, 0, sizeof(p));
  auto plane_map = [](CUtensorMap* tm, const void* base, int Bn, int T, int C, int s) {
    uint64_t dims[4] = {static_cast<uint64_t>(C), static_cast<uint64_t>(s), static_cast<uint64_t>(T / s), static_cast<uint64_t>(Bn)};
    uint64_t strides[3] = {static_cast<uint64_t>(C) * 2, static_cast<uint64_t>(C) * s * 2, static_cast<uint64_t>(C) * T * 2};
    uint32_t box[4] = {64, 1, 64, 1};
    return encode_tmap_bf16(tm, base, 4, dims, strides, box, 1);
  };
  int rc;
  if ((rc = plane_map(&p.tmA, a_plane, B, Ta, Ca, sA))) return rc;
  if ((rc = plane_map(&p.tmB, b_plane, B, Tb, Cb, sB))) return rc;
  p.D = dW; p.M = Ca; p.N = Cb; p.ldd = Cb;
  p.wg_kb_per_item = (T_iter + BLOCK_K - 1) / BLOCK_K;
  p.K = B * p.wg_kb_per_item * BLOCK_K;
  p.wg_rA = rA; p.wg_offA = offA; p.wg_rB = rB; p.wg_offB = offB;
  p.flags = GEMM_A_MN | GEMM_B_MN | GEMM_OUT_F32 | GEMM_ACCUM;
  p.seg_in = 1; p.rope_seq = 1; p.rope_dmodel = 1; p.rope_dh = 64;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (Ca > 128 && Cb > 128) return launch_gemm<256, 2>(p, s);
  if (Cb > 128) return launch_gemm<256, 1>(p, s);
  return launch_gemm<128, 1>(p, s);
}

// All taps of a 64 -> 64 channel flattened 2-D conv weight gradient in ONE launch (see GemmParams::wg_ntaps):
//     dWc[ca][tap][cb] += sum_{b,t} A[b,t,ca] * B[b,t + tap_off[tap],cb]          (note the [Ca][ntaps][Cb] output layout)
// The per-tap entry below reads both planes from HBM once per tap (27 x for a 3x9 kernel; the planes are far larger than L2); here a
// 128 x 256 tile covers four taps, so the planes stream 7 x instead of 27 x and the zero-padded upper half of the M = 64 accumulator is
// amortised over four taps.  Backward (dW = dY (*) X) of the Conv2d stacks of models/encodec.py:94-138.
extern "C" int b200sat_conv_wgrad_taps_cat(const void* a_plane, const void* b_plane, int T, const int* tap_off, int ntaps, float* dWc, int B,
                                           void* stream) {
  if (!a_plane || !b_plane || !dWc || !tap_off || ntaps <= 0 || ntaps > 32 || B <= 0 || T <= 0) { set_last_error("conv_wgrad_taps_cat: bad arguments"); return B200SAT_EINVAL; }
  GemmParams p;
  memset(&p, 0, sizeof(p));
  auto plane_map = [](CUtensorMap* tm, const void* base, int Bn, int Tn) {
    uint64_t dims[4] = {64, 1, static_cast<uint64_t>(Tn), static_cast<uint64_t>(Bn)};
    uint64_t strides[3] = {128, 128, static_cast<uint64_t>(Tn) * 128};
    uint32_t box[4] = {64, 1, 64, 1};
    return encode_tmap_bf16(tm, base, 4, dims, strides, box, 1);
  };
  int rc;
  if ((rc = plane_map(&p.tmA, a_plane, B, T))) return rc;
  if ((rc = plane_map(&p.tmB, b_plane, B, T))) return rc;
  p.D = dWc; p.M = 64; p.N = ntaps * 64; p.ldd = ntaps * 64;
  p.wg_kb_per_item = (T + BLOCK_K - 1) / BLOCK_K;
  p.K = B * p.wg_kb_per_item * BLOCK_K;
  p.wg_ntaps = ntaps;
  for (int k = 0; k < ntaps; ++k) p.wg_tap_off[k] = tap_off[k];
  p.flags = GEMM_A_MN | GEMM_B_MN | GEMM_OUT_F32 | GEMM_ACCUM;
  p.seg_in = 1; p.rope_seq = 1; p.rope_dmodel = 1; p.rope_dh = 64;
  return launch_gemm<256, 1>(p, static_cast<cudaStream_t>(stream));
}

// All taps of one conv weight gradient with a per-tap row shift table on the B operand (flattened 2-D convs): dW[tap][Ca][Cb] +=
// sum_{b,t} A[b,t,:]^T (x) B[b,t + tap_off[tap],:].  One call = ntaps launches of the kernel above (saves the host round trips).
extern "C" int b200sat_conv_wgrad_taps(const void* a_plane, int Ca, const void* b_plane, int Cb, int T, const int* tap_off, int ntaps, float* dW,
                                       int B, void* stream) {
  if (!tap_off || ntaps <= 0) { set_last_error("conv_wgrad_taps: bad arguments"); return B200SAT_EINVAL; }
  for (int k = 0; k < ntaps; ++k) {
    const int rc = b200sat_conv_wgrad(a_plane, Ca, T, 1, 0, 0, b_plane, Cb, T, 1, 0, tap_off[k], dW + static_cast<size_t>(k) * Ca * Cb, B, T, stream);
    if (rc) return rc;
  }
  return B200SAT_OK;
}

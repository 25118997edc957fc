// b200sat — Oobleck Conv1d / ConvTranspose1d stacks as im2col-free implicit GEMMs on tcgen05, TMA-staged.
//
// Replaces weight_norm + F.conv1d / F.conv_transpose1d (cuDNN) and the eager SnakeBeta kernels of the reference
// (stable_audio_tools/models/autoencoders.py:23-27, :58-83, :233-283, :285-362; models/blocks.py:291-329).
//
// Data layout (DESIGN.md "Oobleck activations"): activations are TIME-MAJOR / channels-last [B, T, C] so that a
// convolution tap is a row shift: the A operand of tap k is the plain 2-D box rows [t0 + k*dil - pad, +128) x 64 input
// channels of the SAME tensor, fetched by TMA with out-of-range rows zero-filled (= the conv's zero padding; the batch
// index is its own tensor-map dimension so padding never leaks between items).  No im2col buffer exists anywhere.
// Strided (down-sampling) convs view the input as [B, T/s, s, C]; transposed convs are s independent 2-tap convs, one
// per output phase.  Weights are weight-normalised and packed once per step to [Cout, taps*Cin] (K contiguous).
//
// Precision: every activation / weight is a pair of bf16 planes (hi, lo) with hi + lo ~= the fp32 value (2^-17 rel).
//   passes = 1 : hi x hi only (bf16 GEMM, what autocast training uses)
//   passes = 3 : hi*hi + hi*lo + lo*hi accumulated in the same fp32 TMEM tile -> fp32-class results on bf16 tensor cores.
//
// Epilogue (thread == output time step): + bias, + residual, store raw planes, SnakeBeta with the NEXT layer's
// (alpha, beta) and store the activated planes — so Snake never makes its own pass over HBM.
#include "common.cuh"
#include <cstring>
#include <cstdlib>

namespace b200sat {

struct ConvParams {
  CUtensorMap tmA[2];  // input planes hi, lo: dims {Cin, s_in, T_in/s_in, B}
  CUtensorMap tmB[2];  // packed weight planes hi, lo: dims {taps*Cin, rows}
  CUtensorMap tmRes, tmOut, tmAct;  // hi planes of the residual / raw / activated outputs: dims {Cout, s_o, T_out/s_o, B}, box 32 x 32, 64-byte swizzle
  int tma_epi;         // 1: the epilogue moves the hi planes with TMA (residual load, raw / activated stores)
  const float* bias;
  const float* snake_a;     // exp(alpha) per output channel (of the consumer's SnakeBeta), or null
  const float* snake_invb;  // 1/(exp(beta)+1e-9)
  const __nv_bfloat16* res_hi;
  const __nv_bfloat16* res_lo;
  __nv_bfloat16* out_hi;
  __nv_bfloat16* out_lo;
  __nv_bfloat16* act_hi;
  __nv_bfloat16* act_lo;
  int B, T_in, T_out, Cin, Cout;
  int taps, dil, pad, stride;
  int mode;    // 0 conv (stride 1), 1 strided conv, 2 transposed conv (grid covers `stride` phases)
  int passes;  // 1 or 3
  int m_rows;  // GEMM rows per batch item
  int m_tiles, n_tiles, phases;
  // flattened 2-D convolutions (discriminator, models/encodec.py:76-92): the activation plane is [B, frames * fp, C] with zero pad columns
  // materialised, a 2-D tap is the row shift tap_off[tap]; rows whose (m % fp) falls outside [mask_f0, mask_f1) are stored as zeros
  int tap_off[32];
  int use_tap_table, mask_fp, mask_f0, mask_f1;
  // window groups (2-D convs): `win_groups` windows per channel block, window g starts at row group_off[g] - pad and serves
  // taps [g * taps_per_group, (g+1) * taps_per_group) at consecutive (dil-spaced) row shifts: the 9 frequency taps of one time offset
  int win_groups, taps_per_group, group_off[4];
  float leaky;       // > 0: LeakyReLU slope applied to the raw output (nn.LeakyReLU(0.2), encodec.py:68)
  int window;        // 1: stride-1 conv whose taps share one (128 + (taps-1)*dil)-row A window per channel block
  int win_rows;      // rows of one A item: 128 + (taps-1)*dil in window mode, 128 otherwise
  // shared-memory layout chosen per launch (launch_conv): A items of exactly the window's size and one 2 KB transposition buffer per
  // epilogue warp when only one bf16 plane is written leave room for a deeper weight ring (round 2: a ring stage covers ~512 MMA
  // cycles, so 4 stages gave the weight loads ~1.5k cycles of lead - less than a TMA tile's L2 latency; the tensor pipe sat at 55-60 %)
  int a_item_bytes, b_stages, stg_warp_bytes;
};

constexpr int CV_BM = 128;
constexpr int CV_BK = 64;

template <int BN, int MSUB>
struct ConvCfg {
  // Two rings.  A ring: one item = the input rows one 64-channel block of one 128-row sub-tile needs.  For stride-1 convs that is
  // the WINDOW [t0 - pad, t0 - pad + 128 + (taps-1)*dil) fetched ONCE and read by every tap through a row-shifted UMMA descriptor
  // (the im2col reuse a per-tap fetch would throw away: 7x less shared-memory fill for the k7 convs); strided / transposed convs
  // use one 128-row item per (tap, block).  B ring: one 64-deep K slice of the packed weights per MMA group.
  // MSUB = 2 (stride-1 convs): a CTA tile is 256 time steps = two 128-row accumulators that share every B slice, which halves
  // the weight bytes streamed per flop - with the window the weights are what fills shared memory.
  static constexpr int kAItemBytes = 24 * 1024;            // up to 192 rows x 128 B
  static constexpr int kAItems = (MSUB == 2) ? 4 : ((BN == 256) ? 2 : 3);
  static constexpr int kBBytes = BN * CV_BK * 2;
  static constexpr int kBStagesMin = (MSUB == 2) ? 4 : ((BN == 256) ? 3 : 5);
  static constexpr int kBStagesMax = 8;
  static constexpr int kTmemCols = 2 * MSUB * BN;
  static constexpr int kStagingBytes = 16 * 4096;  // two 32-row x 64-byte transposition buffers (raw, activated) per epilogue warp
  static constexpr int kSmemMax = 227 * 1024;
  static constexpr int kFixedBytes = 1024 /*align slack*/ + 512 /*barriers*/;
  static_assert(kTmemCols <= 512, "TMEM has 512 columns");
};



__device__ __forceinline__ void split_store8(__nv_bfloat16* hi, __nv_bfloat16* lo, const float* v) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float a0 = bf16_round(v[2 * i]), a1 = bf16_round(v[2 * i + 1]);
    h[i] = pack_bf16(a0, a1);
    l[i] = pack_bf16(v[2 * i] - a0, v[2 * i + 1] - a1);
  }
  *reinterpret_cast<uint4*>(hi) = make_uint4(h[0], h[1], h[2], h[3]);
  if (lo) *reinterpret_cast<uint4*>(lo) = make_uint4(l[0], l[1], l[2], l[3]);
}

// Epilogue geometry.  TMEM hands every thread one output ROW (32 fp32 columns per tcgen05.ld); global memory wants whole lines.
// Sixteen epilogue warps (four per TMEM lane quarter, one 32-column block each per 128 columns of tile) own a pair of 32-row x
// 64-byte transposition buffers in shared memory (raw values, activated values); 16-byte chunks are XOR-swizzled with
// (row >> 1) & 3 so that both the row-per-thread side and the line side (8 rows x 4 chunks per instruction) are bank-conflict
// free.  Every global load / store of the bf16 planes then moves eight complete 64-byte row segments per instruction.
constexpr int CV_EPI_WARPS = 16;
constexpr int CV_THREADS = 128 + CV_EPI_WARPS * 32;

__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
// this thread's 32 values -> bf16 -> its row of the transposition buffer; LO: the rounding remainder goes straight to the lo plane
template <bool LO>
__device__ __forceinline__ void stage_row(uint32_t stg_row, int swz, const float* v, __nv_bfloat16* lo_row) {
#pragma unroll
  for (int jx = 0; jx < 4; ++jx) {
    uint32_t hw[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) hw[e] = pack_bf16(v[8 * jx + 2 * e], v[8 * jx + 2 * e + 1]);
    sts128(stg_row + ((jx ^ swz) << 4), make_uint4(hw[0], hw[1], hw[2], hw[3]));
    if (LO) {
      if (lo_row) {
        uint32_t lw[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 h = unpack_bf16(hw[e]);
          lw[e] = pack_bf16(v[8 * jx + 2 * e] - h.x, v[8 * jx + 2 * e + 1] - h.y);
        }
        *reinterpret_cast<uint4*>(lo_row + 8 * jx) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
      }
    }
  }
}

template <int BN, bool LO, int MSUB>
__global__ void __launch_bounds__(CV_THREADS, 1) conv1d_tcgen05(const __grid_constant__ ConvParams p) {
  using Cfg = ConvCfg<BN, MSUB>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  const int a_item_bytes = p.a_item_bytes, nb = p.b_stages;
  uint8_t* smem_b = smem + Cfg::kAItems * a_item_bytes;
  uint8_t* stage_base = smem_b + nb * Cfg::kBBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage_base + CV_EPI_WARPS * p.stg_warp_bytes);
  uint64_t* full_a = bars;
  uint64_t* empty_a = full_a + Cfg::kAItems;
  uint64_t* full_b = empty_a + Cfg::kAItems;
  uint64_t* empty_b = full_b + nb;
  uint64_t* tmem_full = empty_b + nb;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* res_bar = tmem_empty + 2;                       // one per epilogue warp (TMA residual loads)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(res_bar + CV_EPI_WARPS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_per_phase = p.m_tiles * p.B * p.n_tiles;
  const int num_tiles = tiles_per_phase * p.phases;
  const int cin_blocks = p.Cin / CV_BK;
  const bool window = p.window != 0;                         // one A item per channel block, shared by all taps
  const int groups = (window && p.win_groups > 0) ? p.win_groups : 1;
  const int taps_per_item = window ? (p.win_groups > 0 ? p.taps_per_group : p.taps) : 1;
  const int items_per_pass = window ? groups * cin_blocks : p.taps * cin_blocks;
  const int num_items = p.passes * items_per_pass;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA[0]);
    tma_prefetch_desc(&p.tmB[0]);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < Cfg::kAItems; ++i) { mbar_init(&full_a[i], 1); mbar_init(&empty_a[i], 1); }
    for (int i = 0; i < nb; ++i) { mbar_init(&full_b[i], 1); mbar_init(&empty_b[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], CV_EPI_WARPS * 32); }
    for (int i = 0; i < CV_EPI_WARPS; ++i) mbar_init(&res_bar[i], 1);
    fence_barrier_init();
  }
  if (warp == 2) { tmem_alloc(tmem_ptr_smem, Cfg::kTmemCols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  // tile -> (phase, n_blk, b, m_blk): m fastest so neighbouring CTAs share the weight tile in L2
  auto decode = [&](int tile, int& ph, int& n_blk, int& b, int& m_blk) {
    ph = tile / tiles_per_phase;
    int r = tile % tiles_per_phase;
    n_blk = r / (p.m_tiles * p.B);
    r = r % (p.m_tiles * p.B);
    b = r / p.m_tiles;
    m_blk = r % p.m_tiles;
  };

  // item index -> (pass, tap (non-window modes only), channel block); K order: pass, [tap,] block, then the taps of a window
  auto item_decode = [&](int it, int& pass, int& tap, int& cib) {
    pass = it / items_per_pass;
    const int rem = it % items_per_pass;
    tap = window ? (rem / cin_blocks) * taps_per_item : rem / cin_blocks;   // window: first tap of the item's group
    cib = rem % cin_blocks;
  };

  if (warp == 0) {
    if (lane == 0) {   // B producer: one weight slice per MMA group
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int ph, n_blk, b, m_blk;
        decode(tile, ph, n_blk, b, m_blk);
        for (int it = 0; it < num_items; ++it) {
          int pass, tap0, cib;
          item_decode(it, pass, tap0, cib);
          const int b_plane = (pass == 1) ? 1 : 0;            // passes: hi*hi, hi*lo, lo*hi
          for (int tt = 0; tt < taps_per_item; ++tt) {
            const int tap = tap0 + tt;
            mbar_wait(&empty_b[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full_b[stage], Cfg::kBBytes);
            tma_load_2d(smem_b + stage * Cfg::kBBytes, &p.tmB[b_plane], &full_b[stage], tap * p.Cin + cib * CV_BK, ph * p.Cout + n_blk * BN);
            if (++stage == nb) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 3) {
    if (lane == 0) {   // A producer: input rows of one channel block (window or per-tap tile)
      int slot = 0; uint32_t phase = 0;
      const uint32_t item_bytes = static_cast<uint32_t>(p.win_rows) * 128u;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int ph, n_blk, b, m_blk;
        decode(tile, ph, n_blk, b, m_blk);
        const int m0 = m_blk * (CV_BM * MSUB);
        for (int it = 0; it < num_items; ++it) {
          int pass, tap, cib;
          item_decode(it, pass, tap, cib);
          int r = 0, row_off;
          if (p.mode == 0) {
            // window: tap k of the group reads rows [k*dil, k*dil + 128) of the item
            row_off = window ? (p.win_groups > 0 ? p.group_off[tap / taps_per_item] : 0) - p.pad
                             : (p.use_tap_table ? p.tap_off[tap] : tap * p.dil - p.pad);
          } else if (p.mode == 1) {
            const int d = tap - p.pad;                      // input time = t_out*s + d
            const int j = (d >= 0) ? d / p.stride : -((-d + p.stride - 1) / p.stride);
            r = d - j * p.stride;
            row_off = j;
          } else {
            row_off = -tap;                                 // transposed conv phase: x[q - j]
          }
          const int a_plane = (pass == 2) ? 1 : 0;
#pragma unroll
          for (int sub = 0; sub < MSUB; ++sub) {
            mbar_wait(&empty_a[slot], phase ^ 1);
            mbar_arrive_expect_tx(&full_a[slot], item_bytes);
            tma_load_4d(smem_a + slot * a_item_bytes, &p.tmA[a_plane], &full_a[slot], cib * CV_BK, r, m0 + sub * 128 + row_off, b);
            if (++slot == Cfg::kAItems) { slot = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    {   // MMA issue: whole warp in uniform control flow, one elected lane issues (umma_bf16_lo: ~2 instructions per MMA instead of ~9)
      const uint32_t leader = elect_one() ? 1u : 0u;
      constexpr uint32_t idesc = make_idesc_bf16(CV_BM, BN, 0, 0);
      int stage = 0; uint32_t phase = 0; int slot = 0; uint32_t sphase = 0; int as = 0; uint32_t aphase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * (MSUB * BN);
        for (int it = 0; it < num_items; ++it) {
          uint32_t la0[MSUB];
          int slots[MSUB];
#pragma unroll
          for (int sub = 0; sub < MSUB; ++sub) {
            mbar_wait(&full_a[slot], sphase);
            la0[sub] = desc_lo_kmajor(smem_u32(smem_a + slot * a_item_bytes));
            slots[sub] = slot;
            if (++slot == Cfg::kAItems) { slot = 0; sphase ^= 1; }
          }
          for (int tt = 0; tt < taps_per_item; ++tt) {
            mbar_wait(&full_b[stage], phase);
            tc_fence_after();
            // tap tt reads the window shifted down by tt*dil rows: rows sit at a 128-byte pitch (SBO = 8 rows = 1024 B) and the
            // 128-byte swizzle is a function of the shared-memory ADDRESS bits [7,10), so a shift by any number of rows is a plain
            // start-address offset with the descriptor's base-offset field left 0 (measured: setting it to shift%8 gives wrong sums)
            const uint32_t shift = static_cast<uint32_t>(tt * p.dil) * (128u >> 4);
            const uint32_t lb = desc_lo_kmajor(smem_u32(smem_b + stage * Cfg::kBBytes));
#pragma unroll
            for (int sub = 0; sub < MSUB; ++sub) {
#pragma unroll
              for (int k = 0; k < CV_BK / 16; ++k)
                umma_bf16_lo(tmem_d + sub * BN, la0[sub] + shift + 2 * k, lb + 2 * k, idesc, (it | tt | k) != 0, leader);
            }
            umma_commit_if(&empty_b[stage], leader);
            if (++stage == nb) { stage = 0; phase ^= 1; }
          }
#pragma unroll
          for (int sub = 0; sub < MSUB; ++sub) umma_commit_if(&empty_a[slots[sub]], leader);
        }
        umma_commit_if(&tmem_full[as], leader);
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    const int ew = warp - 4;
    const int q = ew & 3;                  // TMEM lane quarter = 32 rows of the tile
    const int cg = ew >> 2;                // column group: BN/4 consecutive columns
    constexpr int NB = BN / 128;           // 32-column blocks per warp per tile
    uint8_t* stg_ptr = stage_base + ew * p.stg_warp_bytes;
    const uint32_t stg_o = smem_u32(stg_ptr);                  // raw values (the residual is staged here first, overwritten in place)
    const uint32_t stg_a = stg_o + (p.stg_warp_bytes - 2048);  // activated values (the same 2 KB when only one plane is produced)
    const int lrow = lane >> 2, lchunk = lane & 3;             // line side: 8 rows x 4 chunks per instruction
    const int swz = (lane >> 1) & 3;
    const uint32_t my_o = stg_o + lane * 64, my_a = stg_a + lane * 64;
    const bool tma_epi = p.tma_epi != 0;
    uint32_t rphase = 0;
    int as = 0; uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int ph, n_blk, b, m_blk;
      decode(tile, ph, n_blk, b, m_blk);
      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after();
#pragma unroll 1
      for (int sub = 0; sub < MSUB; ++sub) {
      const int m_base = m_blk * (CV_BM * MSUB) + sub * CV_BM;
      int grow[4];   // line side: b * T_out + t of rows lrow + 8 i, or -1 when the row does not exist
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int m = m_base + q * 32 + lrow + 8 * i;
        int t = m;
        bool ok = m < p.m_rows;
        if (p.mode == 2) { t = m * p.stride + ph - p.pad; ok = ok && t >= 0 && t < p.T_out; }
        grow[i] = ok ? b * p.T_out + t : -1;
      }
      size_t my_off = 0;
      bool my_ok = false;
      if (LO) {
        int my_t = m_base + q * 32 + lane;
        my_ok = my_t < p.m_rows;
        if (p.mode == 2) { my_t = my_t * p.stride + ph - p.pad; my_ok = my_ok && my_t >= 0 && my_t < p.T_out; }
        my_off = (static_cast<size_t>(b) * p.T_out + (my_ok ? my_t : 0)) * p.Cout;
      }
      const int col0 = n_blk * BN + cg * (BN / 4);
      // TMA coordinates of this warp's 32 rows in the {Cout, s_o, T_out/s_o, B} planes (mode 2 interleaves `stride` phases)
      int trow = m_base + q * 32, tph = 0;
      if (p.mode == 2) {
        const int d = ph - p.pad;
        const int off = (d >= 0) ? d / p.stride : -((-d + p.stride - 1) / p.stride);
        tph = d - off * p.stride;
        trow += off;
      }
      // residual of the first block: in flight while the accumulator is still being produced
      uint4 rres[4];
      if (!tma_epi && p.res_hi && col0 < p.Cout) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          rres[i] = grow[i] >= 0 ? __ldg(reinterpret_cast<const uint4*>(p.res_hi + static_cast<size_t>(grow[i]) * p.Cout + col0 + lchunk * 8)) : make_uint4(0, 0, 0, 0);
      }
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + (as * MSUB + sub) * BN + cg * (BN / 4);
#pragma unroll
      for (int cc = 0; cc < NB; ++cc) {
        const int col = col0 + cc * 32;
        if (col >= p.Cout) break;
        uint32_t raw[32];
        tmem_ld_32x32(taddr + cc * 32, raw);
        if (tma_epi) {
          if (lane == 0) {
            tma_store_wait_read<0>();                   // the previous block's stores have read the buffers
            if (p.res_hi) {
              mbar_arrive_expect_tx(&res_bar[ew], 2048);
              tma_load_4d(stg_ptr, &p.tmRes, &res_bar[ew], col, tph, trow, b);
            }
          }
          __syncwarp();
          if (p.res_hi) { mbar_wait(&res_bar[ew], rphase); rphase ^= 1; }
        } else {
          __syncwarp();                                 // the previous block's line-side stores have read the buffers
          if (p.res_hi) {
#pragma unroll
            for (int i = 0; i < 4; ++i) { const int r = lrow + 8 * i; sts128(stg_o + r * 64 + ((lchunk ^ ((r >> 1) & 3)) << 4), rres[i]); }
            if (cc + 1 < NB && col + 32 < p.Cout) {
#pragma unroll
              for (int i = 0; i < 4; ++i)
                rres[i] = grow[i] >= 0 ? __ldg(reinterpret_cast<const uint4*>(p.res_hi + static_cast<size_t>(grow[i]) * p.Cout + col + 32 + lchunk * 8)) : make_uint4(0, 0, 0, 0);
            }
            __syncwarp();
          }
        }
        tmem_ld_wait_regs(raw);
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
        if (p.bias) {
          const float4* bp = reinterpret_cast<const float4*>(p.bias + col);
#pragma unroll
          for (int i = 0; i < 8; ++i) { const float4 t4 = __ldg(bp + i); v[4 * i] += t4.x; v[4 * i + 1] += t4.y; v[4 * i + 2] += t4.z; v[4 * i + 3] += t4.w; }
        }
        if (p.res_hi) {
#pragma unroll
          for (int jx = 0; jx < 4; ++jx) {
            const uint4 u = lds128(my_o + ((jx ^ swz) << 4));
            const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) { const float2 f = unpack_bf16(w[e]); v[8 * jx + 2 * e] += f.x; v[8 * jx + 2 * e + 1] += f.y; }
          }
          if (LO) {
            if (p.res_lo && my_ok) {   // fp32-class residual: the lo plane is read row-wise (inference-only path)
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const uint4 u = __ldg(reinterpret_cast<const uint4*>(p.res_lo + my_off + col) + i);
                const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) { const float2 f = unpack_bf16(w[e]); v[8 * i + 2 * e] += f.x; v[8 * i + 2 * e + 1] += f.y; }
              }
            }
          }
        }
        if (p.leaky > 0.f) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = v[i] > 0.f ? v[i] : v[i] * p.leaky;
        }
        if (p.mask_fp > 0) {
          const int fcol = (m_base + q * 32 + lane) % p.mask_fp;
          if (fcol < p.mask_f0 || fcol >= p.mask_f1) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = 0.f;
          }
        }
        if (p.out_hi) stage_row<LO>(my_o, swz, v, (LO && p.out_lo && my_ok) ? p.out_lo + my_off + col : nullptr);
        if (p.act_hi) {
          const float4* ap = reinterpret_cast<const float4*>(p.snake_a + col);
          const float4* ip = reinterpret_cast<const float4*>(p.snake_invb + col);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 a4 = __ldg(ap + i), b4 = __ldg(ip + i);
            const float aa[4] = {a4.x, a4.y, a4.z, a4.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              // single-pass bf16 mode: the activated value is rounded to bf16 (2^-9 relative) right below, so the SFU sine without the
              // Cody-Waite reduction is enough (abs error ~2^-21 |x|) and saves three instructions per element of an epilogue-bound kernel
              const float sn = LO ? fast_sin(v[4 * i + e] * aa[e]) : __sinf(v[4 * i + e] * aa[e]);
              v[4 * i + e] += bb[e] * sn * sn;
            }
          }
          stage_row<LO>(my_a, swz, v, (LO && p.act_lo && my_ok) ? p.act_lo + my_off + col : nullptr);
        }
        if (tma_epi) {
          fence_proxy_async_smem();                     // generic-proxy writes of the buffers -> visible to the TMA engine
          __syncwarp();
          if (lane == 0) {
            if (p.out_hi) tma_store_4d(&p.tmOut, stg_ptr, col, tph, trow, b);
            if (p.act_hi) tma_store_4d(&p.tmAct, stg_ptr + (p.stg_warp_bytes - 2048), col, tph, trow, b);
            tma_store_commit();
          }
        } else {
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = lrow + 8 * i;
            if (grow[i] >= 0) {
              const size_t o = static_cast<size_t>(grow[i]) * p.Cout + col + lchunk * 8;
              const uint32_t so = r * 64 + ((lchunk ^ ((r >> 1) & 3)) << 4);
              if (p.out_hi) *reinterpret_cast<uint4*>(p.out_hi + o) = lds128(stg_o + so);
              if (p.act_hi) *reinterpret_cast<uint4*>(p.act_hi + o) = lds128(stg_a + so);
            }
          }
        }
      }
      }   // sub-tile
      tc_fence_before();
      mbar_arrive(&tmem_empty[as]);
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
    if (tma_epi && lane == 0) tma_store_wait_read<0>();   // shared memory must outlive the last stores' reads
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, Cfg::kTmemCols); }
}

template <int BN, bool LO, int MSUB>
static int launch_conv(ConvParams& p, cudaStream_t stream) {
  using Cfg = ConvCfg<BN, MSUB>;
  static bool attr_set = false;
  if (!attr_set) {
    B200SAT_CHECK_CUDA(cudaFuncSetAttribute(conv1d_tcgen05<BN, LO, MSUB>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemMax));
    attr_set = true;
  }
  // shared-memory layout: A items sized to the rows they hold, one transposition buffer per epilogue warp unless both a raw and an
  // activated plane (or a staged residual next to an activated plane) are needed, the rest goes to the weight ring (<= 8 stages)
  p.a_item_bytes = ((p.win_rows * 128 + 1023) / 1024) * 1024;
  const bool two_bufs = p.act_hi && (p.out_hi || p.res_hi);
  p.stg_warp_bytes = two_bufs ? 4096 : 2048;
  {
    static const int ring_env = [] { const char* e = getenv("B200SAT_CONV_BSTAGES"); return e ? atoi(e) : 0; }();   // A/B measurement
    const int room = Cfg::kSmemMax - Cfg::kFixedBytes - Cfg::kAItems * p.a_item_bytes - CV_EPI_WARPS * p.stg_warp_bytes;
    int nb = room / Cfg::kBBytes;
    if (nb > Cfg::kBStagesMax) nb = Cfg::kBStagesMax;
    if (ring_env > 0 && ring_env < nb) nb = ring_env;
    if (nb < 2) { set_last_error("conv1d: shared memory does not fit two weight stages"); return B200SAT_EUNSUPPORTED; }
    p.b_stages = nb;
  }
  const int smem_bytes = Cfg::kFixedBytes + Cfg::kAItems * p.a_item_bytes + p.b_stages * Cfg::kBBytes + CV_EPI_WARPS * p.stg_warp_bytes;
  p.m_tiles = (p.m_rows + CV_BM * MSUB - 1) / (CV_BM * MSUB);
  p.n_tiles = (p.Cout + BN - 1) / BN;
  const long tiles = static_cast<long>(p.m_tiles) * p.B * p.n_tiles * p.phases;
  const int grid = static_cast<int>(tiles < num_sms() ? tiles : num_sms());
  conv1d_tcgen05<BN, LO, MSUB><<<grid, CV_THREADS, smem_bytes, stream>>>(p);
  B200SAT_CHECK_CUDA(cudaGetLastError());
  return B200SAT_OK;
}

// ---------------------------------------------------------------------------------------------------------
// weight_norm (old-style, dim 0: w = g * v / ||v||, autoencoders.py:23-27) + packing into the GEMM layout, hi/lo planes.
//   conv:        v [Cout, Cin, K]  -> W[co][k*Cin + ci]
//   transposed:  v [Cin, Cout, K]  -> W[(ph*Cout + co)][j*Cin + ci] = w[ci][co][ph + s*j]     (K = 2s)
__global__ void wn_norm_kernel(const float* __restrict__ v, float* __restrict__ inv_norm, int inner) {
  const int r = blockIdx.x;
  const float* p = v + static_cast<size_t>(r) * inner;
  float s = 0.f;
  for (int i = threadIdx.x; i < inner; i += blockDim.x) { const float x = p[i]; s += x * x; }
  __shared__ float red[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) inv_norm[r] = rsqrtf(s);
  }
}
__global__ void wn_pack_kernel(const float* __restrict__ v, const float* __restrict__ g, const float* __restrict__ inv_norm,
                               __nv_bfloat16* __restrict__ w_hi, __nv_bfloat16* __restrict__ w_lo, int Cout, int Cin, int K,
                               int transposed, int stride) {
  const long n = static_cast<long>(Cout) * Cin * K;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * blockDim.x) {
    // i indexes the packed layout (coalesced writes)
    float w;
    if (!transposed) {
      const int co = i / (static_cast<long>(K) * Cin);
      const int rem = i % (static_cast<long>(K) * Cin);
      const int k = rem / Cin, ci = rem % Cin;
      const float vv = v[(static_cast<long>(co) * Cin + ci) * K + k];
      w = g ? vv * g[co] * inv_norm[co] : vv;
    } else {
      const int kp = 2 * Cin;  // packed K per phase row
      const long row = i / kp;
      const int rem = i % kp;
      const int ph = row / Cout, co = row % Cout;
      const int j = rem / Cin, ci = rem % Cin;
      const float vv = v[(static_cast<long>(ci) * Cout + co) * K + ph + stride * j];
      w = g ? vv * g[ci] * inv_norm[ci] : vv;
    }
    const float h = bf16_round(w);
    w_hi[i] = __float2bfloat16_rn(h);
    if (w_lo) w_lo[i] = __float2bfloat16_rn(w - h);
  }
}

// SnakeBeta parameters -> exp(alpha), 1/(exp(beta)+1e-9)   (blocks.py:321-329, :291-292)
__global__ void snake_prep_kernel(const float* __restrict__ alpha, const float* __restrict__ beta, float* __restrict__ a,
                                  float* __restrict__ invb, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < C) { a[i] = expf(alpha[i]); invb[i] = 1.0f / (expf(beta[i]) + 1e-9f); }
}

// ---------------------------------------------------------------------------------------------------------
// Edge layers with 2 audio channels (tensor cores would idle): SIMT kernels.
// conv_in: x fp32 [B, Cin<=4, T] -> channels-last planes [B, T, Cout] (+ snake planes).  autoencoders.py:303
__global__ void __launch_bounds__(256) conv_in_kernel(const float* __restrict__ x, const float* __restrict__ w /*[Cout,Cin,K]*/,
                                                      const float* __restrict__ bias, const float* __restrict__ sa,
                                                      const float* __restrict__ sib, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo,
                                                      __nv_bfloat16* act_hi, __nv_bfloat16* act_lo, int B, int Cin, int T, int Cout,
                                                      int K, int pad) {
  extern __shared__ float sm[];
  float* sw = sm;                       // Cout*Cin*K
  float* sx = sm + Cout * Cin * K;      // Cin * (64 + K - 1)
  const int b = blockIdx.y, t0 = blockIdx.x * 64;
  for (int i = threadIdx.x; i < Cout * Cin * K; i += 256) sw[i] = w[i];
  const int span = 64 + K - 1;
  for (int i = threadIdx.x; i < Cin * span; i += 256) {
    const int ci = i / span, tt = i % span;
    const int t = t0 + tt - pad;
    sx[i] = (t >= 0 && t < T) ? x[(static_cast<long>(b) * Cin + ci) * T + t] : 0.f;
  }
  __syncthreads();
  // thread -> 8 consecutive output channels of one time step per iteration
  const int groups = Cout / 8;
  for (int it = threadIdx.x; it < 64 * groups; it += 256) {
    const int tt = it / groups, cg = it % groups;
    const int t = t0 + tt;
    if (t >= T) continue;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int co = cg * 8 + j;
      float acc = bias ? bias[co] : 0.f;
      for (int ci = 0; ci < Cin; ++ci)
        for (int k = 0; k < K; ++k) acc += sw[(co * Cin + ci) * K + k] * sx[ci * span + tt + k];
      v[j] = acc;
    }
    const size_t off = (static_cast<size_t>(b) * T + t) * Cout + cg * 8;
    if (out_hi) split_store8(out_hi + off, out_lo ? out_lo + off : nullptr, v);
    if (act_hi) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float s = fast_sin(v[j] * sa[cg * 8 + j]); v[j] += sib[cg * 8 + j] * s * s; }
      split_store8(act_hi + off, act_lo ? act_lo + off : nullptr, v);
    }
  }
}

// conv_out: activated planes [B, T, Cin] -> y fp32 [B, Cout<=4, T], K taps, zero padding, optional tanh.  autoencoders.py:355-357
__global__ void __launch_bounds__(256) conv_out_kernel(const __nv_bfloat16* __restrict__ in_hi, const __nv_bfloat16* __restrict__ in_lo,
                                                       const float* __restrict__ w /*[Cout,Cin,K]*/, const float* __restrict__ bias,
                                                       float* __restrict__ y, int B, int Cin, int T, int Cout, int K, int pad, int tanh_out) {
  extern __shared__ float sm[];
  float* sw = sm;                         // Cout*K*Cin laid out [co][k][ci]
  float* sx = sm + Cout * K * Cin;        // (32 + K - 1) * (Cin + 1)
  const int b = blockIdx.y, t0 = blockIdx.x * 32;
  for (int i = threadIdx.x; i < Cout * Cin * K; i += 256) {
    const int co = i / (Cin * K), rem = i % (Cin * K), ci = rem / K, k = rem % K;
    sw[(co * K + k) * Cin + ci] = w[i];
  }
  const int span = 32 + K - 1;
  for (int i = threadIdx.x; i < span * Cin; i += 256) {
    const int tt = i / Cin, ci = i % Cin;
    const int t = t0 + tt - pad;
    float v = 0.f;
    if (t >= 0 && t < T) {
      const size_t off = (static_cast<size_t>(b) * T + t) * Cin + ci;
      v = __bfloat162float(in_hi[off]) + (in_lo ? __bfloat162float(in_lo[off]) : 0.f);
    }
    sx[tt * (Cin + 1) + ci] = v;
  }
  __syncthreads();
  // warp w handles time steps w, w+8, ...; lanes split Cin
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int tt = warp; tt < 32; tt += 8) {
    const int t = t0 + tt;
    if (t >= T) break;
    for (int co = 0; co < Cout; ++co) {
      float acc = 0.f;
      for (int k = 0; k < K; ++k)
        for (int ci = lane; ci < Cin; ci += 32) acc += sw[(co * K + k) * Cin + ci] * sx[(tt + k) * (Cin + 1) + ci];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) {
        if (bias) acc += bias[co];
        y[(static_cast<size_t>(b) * Cout + co) * T + t] = tanh_out ? tanhf(acc) : acc;
      }
    }
  }
}

// Fast paths of the two edge layers (stereo audio, 7 taps - every Oobleck config the reference ships):
//
// conv_in_fast: lane == CPT consecutive output channels with their CPT*CIN*K weights in registers; a warp walks 32 consecutive
// time steps and reads each audio sample as a shared-memory broadcast, so a step costs CIN*K broadcast loads + CPT*CIN*K FMAs
// and every store instruction writes one complete [Cout] row (256 B at Cout = 128).  ~25x faster than the generic kernel above,
// which spent its time on two shared-memory loads per FMA.
template <int CPT, int CIN, int K>
__global__ void __launch_bounds__(256) conv_in_fast_kernel(const float* __restrict__ x, const float* __restrict__ w /*[Cout,CIN,K]*/,
                                                           const float* __restrict__ bias, const float* __restrict__ sa,
                                                           const float* __restrict__ sib, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo,
                                                           __nv_bfloat16* act_hi, __nv_bfloat16* act_lo, int T, int Cout, int pad) {
  constexpr int TT = 256, TS = 32, SPAN = TT + K - 1;
  __shared__ float sx[CIN][SPAN];
  const int b = blockIdx.y, t0 = blockIdx.x * TT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < CIN * SPAN; i += 256) {
    const int ci = i / SPAN, tt = i % SPAN;
    const int t = t0 + tt - pad;
    sx[ci][tt] = (t >= 0 && t < T) ? x[(static_cast<long>(b) * CIN + ci) * T + t] : 0.f;
  }
  const int co0 = lane * CPT;
  float wr[CPT][CIN * K], br[CPT], ar[CPT], ir[CPT];
#pragma unroll
  for (int c = 0; c < CPT; ++c) {
#pragma unroll
    for (int j = 0; j < CIN * K; ++j) wr[c][j] = w[(co0 + c) * CIN * K + j];
    br[c] = bias ? bias[co0 + c] : 0.f;
    ar[c] = sa ? sa[co0 + c] : 0.f;
    ir[c] = sib ? sib[co0 + c] : 0.f;
  }
  __syncthreads();
  for (int s2 = 0; s2 < TS; ++s2) {
    const int tl = warp * TS + s2;
    const int t = t0 + tl;
    if (t >= T) break;
    float acc[CPT];
#pragma unroll
    for (int c = 0; c < CPT; ++c) acc[c] = br[c];
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) {
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const float xv = sx[ci][tl + k];
#pragma unroll
        for (int c = 0; c < CPT; ++c) acc[c] = fmaf(wr[c][ci * K + k], xv, acc[c]);
      }
    }
    const size_t off = (static_cast<size_t>(b) * T + t) * Cout + co0;
    auto store = [&](__nv_bfloat16* hi, __nv_bfloat16* lo) {
      uint32_t hw[CPT / 2], lw[CPT / 2];
#pragma unroll
      for (int c = 0; c < CPT / 2; ++c) {
        hw[c] = pack_bf16(acc[2 * c], acc[2 * c + 1]);
        const float2 h = unpack_bf16(hw[c]);
        lw[c] = pack_bf16(acc[2 * c] - h.x, acc[2 * c + 1] - h.y);
      }
      if (CPT == 4) {
        *reinterpret_cast<uint2*>(hi + off) = make_uint2(hw[0], hw[CPT / 2 - 1]);
        if (lo) *reinterpret_cast<uint2*>(lo + off) = make_uint2(lw[0], lw[CPT / 2 - 1]);
      } else {
        *reinterpret_cast<uint32_t*>(hi + off) = hw[0];
        if (lo) *reinterpret_cast<uint32_t*>(lo + off) = lw[0];
      }
    };
    if (out_hi) store(out_hi, out_lo);
    if (act_hi) {
#pragma unroll
      for (int c = 0; c < CPT; ++c) { const float sn = fast_sin(acc[c] * ar[c]); acc[c] += ir[c] * sn * sn; }
      store(act_hi, act_lo);
    }
  }
}

// conv_out_fast: planes [B, T, Cin] -> y fp32 [B, COUT, T].  A block owns 64 time steps; lane = (time group of 4 steps) x (quarter of
// the input channels, interleaved ci = 4 j + cq so that the fp32 shared-memory tile with a (Cin + 1)-word pitch is read
// conflict-free); each thread keeps 4 x COUT accumulators, reads 10 tile values + K weight pairs per channel for 4*K*COUT FMAs,
// and the four channel quarters are combined with two shuffles.  Output stores are 128-byte coalesced along time.
template <int COUT, int K>
__global__ void __launch_bounds__(64) conv_out_fast_kernel(const __nv_bfloat16* __restrict__ in_hi, const __nv_bfloat16* __restrict__ in_lo,
                                                           const float* __restrict__ w /*[COUT,Cin,K]*/, const float* __restrict__ bias,
                                                           float* __restrict__ y, int Cin, int T, int pad, int tanh_out) {
  constexpr int TT = 64, ROWS = TT + K - 1;
  extern __shared__ float sm[];
  const int pitch = Cin + 1;
  float* sx = sm;                     // [ROWS][pitch]
  float* sw = sm + ROWS * pitch;      // [K][Cin][COUT]
  const int b = blockIdx.y, t0 = blockIdx.x * TT;
  for (int i = threadIdx.x; i < COUT * Cin * K; i += 64) {
    const int co = i / (Cin * K), rem = i % (Cin * K), ci = rem / K, k = rem % K;
    sw[(k * Cin + ci) * COUT + co] = w[i];
  }
  const int chunks = Cin / 8;
  for (int i = threadIdx.x; i < ROWS * chunks; i += 64) {
    const int r = i / chunks, c8 = (i % chunks) * 8;
    const int t = t0 + r - pad;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = 0.f;
    if (t >= 0 && t < T) {
      const size_t off = (static_cast<size_t>(b) * T + t) * Cin + c8;
      const uint4 u = *reinterpret_cast<const uint4*>(in_hi + off);
      const uint32_t uw[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) { const float2 f = unpack_bf16(uw[e]); v[2 * e] = f.x; v[2 * e + 1] = f.y; }
      if (in_lo) {
        const uint4 u2 = *reinterpret_cast<const uint4*>(in_lo + off);
        const uint32_t lw[4] = {u2.x, u2.y, u2.z, u2.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) { const float2 f = unpack_bf16(lw[e]); v[2 * e] += f.x; v[2 * e + 1] += f.y; }
      }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) sx[r * pitch + c8 + e] = v[e];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tq = lane >> 2, cq = lane & 3;
  const int tl = (warp * 8 + tq) * 4;
  float acc[4][COUT];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int co = 0; co < COUT; ++co) acc[a][co] = 0.f;
  for (int j = 0; j < Cin / 4; ++j) {
    const int ci = 4 * j + cq;
    float xr[4 + K - 1];
#pragma unroll
    for (int r = 0; r < 4 + K - 1; ++r) xr[r] = sx[(tl + r) * pitch + ci];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      float wv[COUT];
#pragma unroll
      for (int co = 0; co < COUT; ++co) wv[co] = sw[(k * Cin + ci) * COUT + co];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int co = 0; co < COUT; ++co) acc[a][co] = fmaf(xr[a + k], wv[co], acc[a][co]);
    }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int co = 0; co < COUT; ++co) {
      acc[a][co] += __shfl_xor_sync(0xffffffffu, acc[a][co], 1);
      acc[a][co] += __shfl_xor_sync(0xffffffffu, acc[a][co], 2);
    }
  const int t = t0 + tl + cq;      // lane cq of a time group writes step cq: 32 consecutive steps per warp
  if (t < T) {
#pragma unroll
    for (int co = 0; co < COUT; ++co) {
      float v = (cq == 0) ? acc[0][co] : (cq == 1) ? acc[1][co] : (cq == 2) ? acc[2][co] : acc[3][co];
      if (bias) v += bias[co];
      y[(static_cast<size_t>(b) * COUT + co) * T + t] = tanh_out ? tanhf(v) : v;
    }
  }
}

// fp32 [B, C, T] -> channels-last hi/lo planes [B, T, C] (decoder input latents)
__global__ void to_planes_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                 int B, int C, int T) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, c0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, t = t0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && t < T) ? x[(static_cast<size_t>(b) * C + c) * T + t] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int t = t0 + i, c = c0 + threadIdx.x;
    if (t < T && c < C) {
      const float v = tile[threadIdx.x][i];
      const float h = bf16_round(v);
      const size_t off = (static_cast<size_t>(b) * T + t) * C + c;
      hi[off] = __float2bfloat16_rn(h);
      if (lo) lo[off] = __float2bfloat16_rn(v - h);
    }
  }
}

// Encoder tail: planes [B, T, 2*L] (mean | scale) -> VAE sample z = noise*(softplus(scale)+1e-4) + mean, fp32 [B, L, T],
// and KL partial sums (bottleneck.py:105-113).  kl_out accumulates sum over (b, t, c) of (mean^2 + var - log var - 1).
__global__ void vae_sample_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo,
                                  const float* __restrict__ noise, float* __restrict__ z, float* __restrict__ mean_scale_out,
                                  float* __restrict__ kl_out, int B, int L, int T) {
  const long n = static_cast<long>(B) * L * T;
  float klacc = 0.f;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int t = i % T;
    const int c = (i / T) % L;
    const int b = i / (static_cast<long>(T) * L);
    const size_t off = (static_cast<size_t>(b) * T + t) * (2 * L);
    const float mean = __bfloat162float(hi[off + c]) + (lo ? __bfloat162float(lo[off + c]) : 0.f);
    const float sc = __bfloat162float(hi[off + L + c]) + (lo ? __bfloat162float(lo[off + L + c]) : 0.f);
    const float sp = (sc > 20.f) ? sc : log1pf(expf(sc));  // F.softplus (threshold 20)
    const float stdev = sp + 1e-4f;
    const float var = stdev * stdev;
    if (z) z[i] = (noise ? noise[i] : 0.f) * stdev + mean;
    if (mean_scale_out) {
      mean_scale_out[(static_cast<size_t>(b) * 2 * L + c) * T + t] = mean;
      mean_scale_out[(static_cast<size_t>(b) * 2 * L + L + c) * T + t] = sc;
    }
    klacc += mean * mean + var - logf(var) - 1.f;
  }
  if (kl_out) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) klacc += __shfl_xor_sync(0xffffffffu, klacc, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(kl_out, klacc);
  }
}

}  // namespace b200sat

using namespace b200sat;

static int make_plane_map(CUtensorMap* tm, const void* base, int B, int T, int C, int s_in, int box_rows) {
  uint64_t dims[4] = {static_cast<uint64_t>(C), static_cast<uint64_t>(s_in), static_cast<uint64_t>(T / s_in), static_cast<uint64_t>(B)};
  uint64_t strides[3] = {static_cast<uint64_t>(C) * 2, static_cast<uint64_t>(C) * s_in * 2, static_cast<uint64_t>(C) * T * 2};
  uint32_t box[4] = {CV_BK, 1, static_cast<uint32_t>(box_rows), 1};
  return encode_tmap_bf16(tm, base, 4, dims, strides, box, 1);
}

extern "C" int b200sat_conv1d_fwd(const void* in_hi, const void* in_lo, const void* w_hi, const void* w_lo, const float* bias,
                                  const void* res_hi, const void* res_lo, void* out_hi, void* out_lo, void* act_hi, void* act_lo,
                                  const float* snake_a, const float* snake_invb, int B, int T_in, int Cin, int Cout, int taps,
                                  int dil, int pad, int stride, int mode, int passes, void* stream) {
  if (!in_hi || !w_hi || (!out_hi && !act_hi) || B <= 0 || T_in <= 0) { set_last_error("conv1d: bad arguments"); return B200SAT_EINVAL; }
  if (Cin % 64 || Cout % 32) { set_last_error("conv1d: Cin must be a multiple of 64 and Cout of 32 (edge layers use conv_in/conv_out)"); return B200SAT_EUNSUPPORTED; }
  if (passes != 1 && passes != 3) { set_last_error("conv1d: passes must be 1 or 3"); return B200SAT_EINVAL; }
  if (passes == 3 && (!in_lo || !w_lo)) { set_last_error("conv1d: passes=3 needs lo planes"); return B200SAT_EINVAL; }
  if (act_hi && (!snake_a || !snake_invb)) { set_last_error("conv1d: activated output needs snake parameters"); return B200SAT_EINVAL; }
  if (mode < 0 || mode > 2 || stride < 1 || (mode == 1 && T_in % stride)) { set_last_error("conv1d: bad mode/stride"); return B200SAT_EINVAL; }
  if (mode == 2 && taps != 2 * stride) { set_last_error("conv1d: transposed conv needs kernel_size == 2*stride"); return B200SAT_EUNSUPPORTED; }
  ConvParams p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.T_in = T_in; p.Cin = Cin; p.Cout = Cout; p.dil = dil; p.pad = pad; p.stride = stride; p.mode = mode; p.passes = passes;
  int wrows, wk;
  if (mode == 0) { p.T_out = T_in + 2 * pad - dil * (taps - 1); p.taps = taps; p.m_rows = p.T_out; p.phases = 1; wrows = Cout; wk = taps * Cin; }
  else if (mode == 1) { p.T_out = (T_in + 2 * pad - taps) / stride + 1; p.taps = taps; p.m_rows = p.T_out; p.phases = 1; wrows = Cout; wk = taps * Cin; }
  else { p.T_out = (T_in - 1) * stride - 2 * pad + taps; p.taps = 2; p.m_rows = T_in + 1; p.phases = stride; wrows = stride * Cout; wk = 2 * Cin; }
  if (p.T_out <= 0) { set_last_error("conv1d: empty output"); return B200SAT_EINVAL; }
  const int s_in = (mode == 1) ? stride : 1;
  int bn = (Cout >= 256) ? 256 : 128;
  bool tall_env = true;
  int rc;
  {
    // B200SAT_CONV_WINDOW=0 falls back to per-tap A tiles (A/B measurement of the shared window)
    static const int win_env = [] { const char* e = getenv("B200SAT_CONV_WINDOW"); return e ? atoi(e) : 1; }();
    tall_env = win_env != 3;   // B200SAT_CONV_WINDOW=3: window but 128-row tiles (A/B measurement)
    const int wr = CV_BM + (p.taps - 1) * dil;
    p.window = (mode == 0 && win_env != 0 && wr * 128 <= 24 * 1024) ? 1 : 0;
    p.win_rows = p.window ? wr : CV_BM;
    if (p.window && tall_env) bn = 128;   // 256 x 128 tiles (the weight box must match the kernel's BN)
  }
  if ((rc = make_plane_map(&p.tmA[0], in_hi, B, T_in, Cin, s_in, p.win_rows))) return rc;
  if (in_lo && (rc = make_plane_map(&p.tmA[1], in_lo, B, T_in, Cin, s_in, p.win_rows))) return rc;
  {
    uint64_t dims[2] = {static_cast<uint64_t>(wk), static_cast<uint64_t>(wrows)};
    uint64_t strides[1] = {static_cast<uint64_t>(wk) * 2};
    uint32_t box[2] = {CV_BK, static_cast<uint32_t>(bn)};
    if ((rc = encode_tmap_bf16(&p.tmB[0], w_hi, 2, dims, strides, box, 1))) return rc;
    if (w_lo && (rc = encode_tmap_bf16(&p.tmB[1], w_lo, 2, dims, strides, box, 1))) return rc;
  }
  {
    // epilogue planes through TMA: {Cout, s_o, T_out/s_o, B}, 32 x 32 boxes, 64-byte swizzle (= the staging buffers' layout)
    static const int epi_env = [] { const char* e = getenv("B200SAT_CONV_TMA_EPI"); return e ? atoi(e) : 1; }();
    const int s_o = (mode == 2) ? stride : 1;
    // transposed convs keep the line-side path: their phase-interleaved rows start at negative coordinates, which the TMA store
    // unit rejects (illegal instruction on sm_100a; loads accept them)
    p.tma_epi = (epi_env && mode != 2) ? 1 : 0;
    if (p.tma_epi) {
      auto omap = [&](CUtensorMap* tm, const void* base) {
        uint64_t dims[4] = {static_cast<uint64_t>(Cout), static_cast<uint64_t>(s_o), static_cast<uint64_t>(p.T_out / s_o), static_cast<uint64_t>(B)};
        uint64_t strides[3] = {static_cast<uint64_t>(Cout) * 2, static_cast<uint64_t>(Cout) * s_o * 2, static_cast<uint64_t>(Cout) * p.T_out * 2};
        uint32_t box[4] = {32, 1, 32, 1};
        return encode_tmap_bf16(tm, base, 4, dims, strides, box, 2);
      };
      if (res_hi && (rc = omap(&p.tmRes, res_hi))) return rc;
      if (out_hi && (rc = omap(&p.tmOut, out_hi))) return rc;
      if (act_hi && (rc = omap(&p.tmAct, act_hi))) return rc;
    }
  }
  p.bias = bias; p.snake_a = snake_a; p.snake_invb = snake_invb;
  p.res_hi = static_cast<const __nv_bfloat16*>(res_hi); p.res_lo = static_cast<const __nv_bfloat16*>(res_lo);
  p.out_hi = static_cast<__nv_bfloat16*>(out_hi); p.out_lo = static_cast<__nv_bfloat16*>(out_lo);
  p.act_hi = static_cast<__nv_bfloat16*>(act_hi); p.act_lo = static_cast<__nv_bfloat16*>(act_lo);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool lo = p.res_lo || p.out_lo || p.act_lo;
  // stride-1 convs with a shared window: 256 x 128 tiles (two accumulators per weight slice) for every width - the window makes A
  // cheap, so tall tiles minimise the bytes of shared-memory fill per flop
  if (p.window && tall_env) return lo ? launch_conv<128, true, 2>(p, s) : launch_conv<128, false, 2>(p, s);
  if (bn == 256) return lo ? launch_conv<256, true, 1>(p, s) : launch_conv<256, false, 1>(p, s);
  return lo ? launch_conv<128, true, 1>(p, s) : launch_conv<128, false, 1>(p, s);
}

extern "C" int b200sat_wn_pack(const float* v, const float* g, float* inv_norm_scratch, void* w_hi, void* w_lo, int Cout, int Cin,
                               int K, int transposed, int stride, void* stream) {
  if (!v || !w_hi || Cout <= 0 || Cin <= 0 || K <= 0) { set_last_error("wn_pack: bad arguments"); return B200SAT_EINVAL; }
  if (g && !inv_norm_scratch) { set_last_error("wn_pack: weight-norm needs a scratch buffer of dim-0 floats"); return B200SAT_EINVAL; }
  if (transposed && K != 2 * stride) { set_last_error("wn_pack: transposed conv needs K == 2*stride"); return B200SAT_EUNSUPPORTED; }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int dim0 = transposed ? Cin : Cout;
  const int inner = (transposed ? Cout : Cin) * K;
  if (g) wn_norm_kernel<<<dim0, 256, 0, s>>>(v, inv_norm_scratch, inner);
  const long n = static_cast<long>(Cout) * Cin * K;
  const int grid = static_cast<int>((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096);
  wn_pack_kernel<<<grid, 256, 0, s>>>(v, g, inv_norm_scratch, static_cast<__nv_bfloat16*>(w_hi), static_cast<__nv_bfloat16*>(w_lo),
                                      Cout, Cin, K, transposed, stride);
  B200SAT_CHECK_CUDA(cudaGetLastError());
  return B200SAT_OK;
}

extern "C" int b200sat_snake_prep(const float* alpha, const float* beta, float* a, float* invb, int C, void* stream) {
  if (!alpha || !beta || !a || !invb || C <= 0) { set_last_error("snake_prep: bad arguments"); return B200SAT_EINVAL; }
  snake_prep_kernel<<<(C + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(alpha, beta, a, invb, C);
  B200SAT_CHECK_CUDA(cudaGetLastError());
  return B200SAT_OK;
}

extern "C" int b200sat_conv_in(const float* x, const float* w, const float* bias, const float* snake_a, const float* snake_invb,
                               void* out_hi, void* out_lo, void* act_hi, void* act_lo, int B, int Cin, int T, int Cout, int K, int pad,
                               void* stream) {
  if (!x || !w || B <= 0 || Cin <= 0 || Cin > 8 || Cout % 8 || K <= 0) { set_last_error("conv_in: bad arguments (Cin <= 8, Cout % 8 == 0)"); return B200SAT_EINVAL; }
  if (Cin == 2 && K == 7 && (Cout == 128 || Cout == 64)) {
    dim3 fgrid((T + 255) / 256, B);
    auto* oh = static_cast<__nv_bfloat16*>(out_hi); auto* ol = static_cast<__nv_bfloat16*>(out_lo);
    auto* ah = static_cast<__nv_bfloat16*>(act_hi); auto* al = static_cast<__nv_bfloat16*>(act_lo);
    if (Cout == 128) conv_in_fast_kernel<4, 2, 7><<<fgrid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, w, bias, snake_a, snake_invb, oh, ol, ah, al, T, Cout, pad);
    else conv_in_fast_kernel<2, 2, 7><<<fgrid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, w, bias, snake_a, snake_invb, oh, ol, ah, al, T, Cout, pad);
    B200SAT_CHECK_CUDA(cudaGetLastError());
    return B200SAT_OK;
  }
  const int smem = (Cout * Cin * K + Cin * (64 + K - 1)) * 4;
  if (smem > 48 * 1024) { set_last_error("conv_in: weights do not fit in 48 KB of shared memory"); return B200SAT_EUNSUPPORTED; }
  dim3 grid((T + 63) / 64, B);
  conv_in_kernel<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(
      x, w, bias, snake_a, snake_invb, static_cast<__nv_bfloat16*>(out_hi), static_cast<__nv_bfloat16*>(out_lo),
      static_cast<__nv_bfloat16*>(act_hi), static_cast<__nv_bfloat16*>(act_lo), B, Cin, T, Cout, K, pad);
  B200SAT_CHECK_CUDA(cudaGetLastError());
  return B200SAT_OK;
}

extern "C" int b200sat_conv_out(const void* in_hi, const void* in_lo, const float* w, const float* bias, float* y, int B, int Cin,
                                int T, int Cout, int K, int pad, int tanh_out, void* stream) {
  if (!in_hi || !w || !y || B <= 0 || Cout <= 0 || Cout > 8) { set_last_error("conv_out: bad arguments (Cout <= 8)"); return B200SAT_EINVAL; }
  if (Cout == 2 && K == 7 && Cin % 8 == 0 && ((64 + 6) * (Cin + 1) + 2 * 7 * Cin) * 4 <= 48 * 1024) {
    const int fsmem = ((64 + 6) * (Cin + 1) + 2 * 7 * Cin) * 4;
    dim3 fgrid((T + 63) / 64, B);
    conv_out_fast_kernel<2, 7><<<fgrid, 64, fsmem, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(in_hi), static_cast<const __nv_bfloat16*>(in_lo), w, bias, y, Cin, T, pad, tanh_out);
    B200SAT_CHECK_CUDA(cudaGetLastError());
    return B200SAT_OK;
  }
  const int smem = (Cout * K * Cin + (32 + K - 1) * (Cin + 1)) * 4;
  if (smem > 48 * 1024) { set_last_error("conv_out: tile does not fit in 48 KB of shared memory"); return B200SAT_EUNSUPPORTED; }
  dim3 grid((T + 31) / 32, B);
  conv_out_kernel<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(in_hi), static_cast<const __nv_bfloat16*>(in_lo), w, bias, y, B, Cin, T, Cout, K, pad, tanh_out);
  B200SAT_CHECK_CUDA(cudaGetLastError());
  return B200SAT_OK;
}

extern "C" int b200sat_to_planes(const float* x, void* hi, void* lo, int B, int C, int T, void* stream) {
  if (!x || !hi || B <= 0 || C <= 0 || T <= 0) { set_last_error("to_planes: bad arguments"); return B200SAT_EINVAL; }
  dim3 grid((T + 31) / 32, (C + 31) / 32, B), block(32, 8);
  to_planes_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(x, static_cast<__nv_bfloat16*>(hi), static_cast<__nv_bfloat16*>(lo), B, C, T);
  B200SAT_CHECK_CUDA(cudaGetLastError());
  return B200SAT_OK;
}

extern "C" int b200sat_vae_sample(const void* hi, const void* lo, const float* noise, float* z, float* mean_scale_out, float* kl_sum,
                                  int B, int L, int T, void* stream) {
  if (!hi || B <= 0 || L <= 0 || T <= 0) { set_last_error("vae_sample: bad arguments"); return B200SAT_EINVAL; }
  const long n = static_cast<long>(B) * L * T;
  const int grid = static_cast<int>((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184);
  vae_sample_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(hi), static_cast<const __nv_bfloat16*>(lo),
                                                                        noise, z, mean_scale_out, kl_sum, B, L, T);
  B200SAT_CHECK_CUDA(cudaGetLastError());
  return B200SAT_OK;
}

// Flattened 2-D convolution on the tcgen05 conv kernel (see ConvParams::tap_off): in / out bf16 planes [B, P, C], w packed [Cout][ntaps*Cin]
// (b200sat_wn_pack with K = ntaps), tap_off[ntaps] = row shift of each tap, rows with (row % fp) outside [f0, f1) are written as zeros
// (fp = 0: no mask), optional LeakyReLU on the output.  Forward and data gradient of the Conv2d stacks of models/encodec.py:76-92.
extern "C" int b200sat_conv2d_flat(const void* in, const void* w, const float* bias, void* out, int B, int P, int Cin, int Cout, int ntaps,
                                   const int* tap_off, int fp, int f0, int f1, float leaky, void* stream) {
  if (!in || !w || !out || !tap_off || B <= 0 || P <= 0 || ntaps <= 0 || ntaps > 32) { set_last_error("conv2d_flat: bad arguments (ntaps <= 32)"); return B200SAT_EINVAL; }
  if (Cin % 64 || Cout % 32) { set_last_error("conv2d_flat: Cin must be a multiple of 64 and Cout of 32"); return B200SAT_EUNSUPPORTED; }
  ConvParams p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.T_in = P; p.T_out = P; p.Cin = Cin; p.Cout = Cout; p.dil = 1; p.pad = 0; p.stride = 1; p.mode = 0; p.passes = 1;
  p.taps = ntaps; p.m_rows = P; p.phases = 1;
  p.use_tap_table = 1;
  for (int i = 0; i < ntaps; ++i) p.tap_off[i] = tap_off[i];
  p.mask_fp = fp; p.mask_f0 = f0; p.mask_f1 = f1; p.leaky = leaky;
  p.window = 0; p.win_rows = CV_BM;
  // taps that come in runs of consecutive rows (the frequency taps of one time offset) share one window per run:
  // 3 windows of 128 + 8 rows instead of 27 per-tap tiles (B200SAT_DISC_WINDOW=0 keeps the per-tap path)
  static const int win_env = [] { const char* e = getenv("B200SAT_DISC_WINDOW"); return e ? atoi(e) : 1; }();
  bool tall = false;
  if (win_env) {
    int run = 1;
    while (run < ntaps && tap_off[run] == tap_off[run - 1] + 1) ++run;
    bool ok = run > 1 && ntaps % run == 0 && ntaps / run <= 4;
    for (int g = 0; ok && g < ntaps / run; ++g)
      for (int k = 1; k < run; ++k) ok = ok && tap_off[g * run + k] == tap_off[g * run] + k;
    if (ok) {
      p.window = 1; p.use_tap_table = 0;
      p.win_groups = ntaps / run; p.taps_per_group = run;
      for (int g = 0; g < p.win_groups; ++g) p.group_off[g] = tap_off[g * run];
      p.dil = 1; p.pad = 0;
      p.win_rows = CV_BM + run - 1;
      tall = true;
    }
  }
  const int bn = (Cout >= 256 && !tall) ? 256 : 128;
  int rc;
  if ((rc = make_plane_map(&p.tmA[0], in, B, P, Cin, 1, p.win_rows))) return rc;
  {
    uint64_t dims[2] = {static_cast<uint64_t>(ntaps) * Cin, static_cast<uint64_t>(Cout)};
    uint64_t strides[1] = {static_cast<uint64_t>(ntaps) * Cin * 2};
    uint32_t box[2] = {CV_BK, static_cast<uint32_t>(bn)};
    if ((rc = encode_tmap_bf16(&p.tmB[0], w, 2, dims, strides, box, 1))) return rc;
  }
  {
    uint64_t dims[4] = {static_cast<uint64_t>(Cout), 1, static_cast<uint64_t>(P), static_cast<uint64_t>(B)};
    uint64_t strides[3] = {static_cast<uint64_t>(Cout) * 2, static_cast<uint64_t>(Cout) * 2, static_cast<uint64_t>(Cout) * P * 2};
    uint32_t box[4] = {32, 1, 32, 1};
    if ((rc = encode_tmap_bf16(&p.tmOut, out, 4, dims, strides, box, 2))) return rc;
    p.tma_epi = 1;
  }
  p.bias = bias;
  p.out_hi = static_cast<__nv_bfloat16*>(out);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (tall) return launch_conv<128, false, 2>(p, s);
  return bn == 256 ? launch_conv<256, false, 1>(p, s) : launch_conv<128, false, 1>(p, s);
}

// b200sat — one Oobleck ResidualUnit in ONE kernel (C = 128, bf16 single pass):
//
//     y = x + conv1( snake1( conv7_dil( snake0(x) ) ) )          stable_audio_tools/models/autoencoders.py:58-83
//
// inputs : x_act = snake0(x) and x (raw) as time-major bf16 planes [B, T, 128] (the previous layer's epilogue wrote both)
// outputs: y (raw plane, the next unit's skip input) and snake_next(y) (the next layer's conv input)
//
// Why: as two launches (conv1d.cu) the k = 1 conv is HBM-bound — it reads the k = 7 conv's activated output and the skip plane and
// writes two planes: 2.1 GB per launch at T = 2 097 152, 540 us against 447 us for the k = 7 conv that does 7x its flops (ncu,
// profiles/r1_launches_ae_bf16_47s_v2.csv).  Fused, the intermediate h = snake1(conv7(.)) never leaves the SM: the k = 7
// accumulator is read out of TMEM, activated, written to shared memory as the A operand of a second tcgen05 GEMM against the
// 128 x 128 k = 1 weights, and only x_act, x in and y, snake(y) out touch HBM (2.0 GB instead of 3.1 GB per unit).
//
// Structure = conv1d_tcgen05<128, false, 2> (256-step tiles, shared A window for the 7 taps, TMA rings) plus:
//   * TMEM: accumulator 1 (conv7, 2 x 128 columns) and accumulator 2 (conv1, 2 x 128 columns) = all 512 columns;
//   * the B ring carries the 14 conv7 weight slices of a tile followed by the 2 conv1 slices (re-streamed from L2 per tile: +14 %
//     weight traffic, and no 32 KB of shared memory pinned for W1);
//   * epilogue 1 (all 16 epilogue warps): acc1 -> + b7 -> SnakeBeta(s1) -> bf16 -> H operand tiles (128-byte swizzle, K-major) in
//     shared memory -> h_full; the MMA thread then issues conv1 (8 MMAs per sub-tile) and goes straight on to the next tile's conv7,
//     which overlaps epilogue 2: acc2 -> + b1 -> + x (skip, prefetched line-wise) -> raw plane, SnakeBeta(next) -> activated plane,
//     both through per-warp transposition buffers and TMA stores.  The H tiles and the transposition buffers share the same 64 KB.
#include "common.cuh"
#include <cstring>
#include <cstdlib>

namespace b200sat {

struct RuParams {
  CUtensorMap tmA;     // x_act plane: dims {128, 1, T, B}, box {64, 1, win_rows, 1}, 128-byte swizzle
  CUtensorMap tmW7;    // packed conv7 weights [128][7*128]: box {64, 128}
  CUtensorMap tmW1;    // packed conv1 weights [128][128]:   box {64, 128}
  CUtensorMap tmOut, tmAct;   // raw / activated output planes: dims {128, 1, T, B}, box {32, 1, 32, 1}, 64-byte swizzle
  const __nv_bfloat16* res;   // x (raw) plane [B, T, 128]
  const float* b7; const float* s1_a; const float* s1_invb;
  const float* b1; const float* nx_a; const float* nx_invb;   // nx_* null: no activated output
  int B, T, dil, pad, win_rows, m_tiles;
  int has_out, has_act;
};

constexpr int RU_C = 128;
constexpr int RU_TAPS = 7;
constexpr int RU_AITEMS = 4, RU_AITEM_BYTES = 24 * 1024;
constexpr int RU_BSTAGES = 4, RU_BBYTES = RU_C * 64 * 2;        // one 128 x 64 weight slice = 16 KB
constexpr int RU_HBYTES = 2 * 2 * 16384;                        // [sub][k-block] 128 rows x 128 B; aliased by the 16 x 4 KB transposition buffers
constexpr int RU_EPI_WARPS = 16;
constexpr int RU_THREADS = 128 + RU_EPI_WARPS * 32;
constexpr int RU_SMEM = RU_AITEMS * RU_AITEM_BYTES + RU_BSTAGES * RU_BBYTES + RU_HBYTES + 1024 + 512;

__device__ __forceinline__ void ru_sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ru_lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
// this thread's 32 values -> bf16 -> its row of a 32-row x 64-byte transposition buffer (16-byte chunks XOR-swizzled with (row>>1)&3,
// which is exactly TMA's SWIZZLE_64B)
__device__ __forceinline__ void ru_stage_row(uint32_t stg_row, int swz, const float* v) {
#pragma unroll
  for (int jx = 0; jx < 4; ++jx) {
    uint32_t hw[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) hw[e] = pack_bf16(v[8 * jx + 2 * e], v[8 * jx + 2 * e + 1]);
    ru_sts128(stg_row + ((jx ^ swz) << 4), make_uint4(hw[0], hw[1], hw[2], hw[3]));
  }
}
__device__ __forceinline__ void ru_snake32(float* v, const float* a, const float* invb) {
  const float4* ap = reinterpret_cast<const float4*>(a);
  const float4* ip = reinterpret_cast<const float4*>(invb);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 a4 = __ldg(ap + i), b4 = __ldg(ip + i);
    const float aa[4] = {a4.x, a4.y, a4.z, a4.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float sn = __sinf(v[4 * i + e] * aa[e]);   // bf16 output: the SFU sine's own range handling is accurate enough
      v[4 * i + e] += bb[e] * sn * sn;
    }
  }
}
// an opaque zero: parameter loads indexed with it cannot be hoisted out of the sub-tile loops (96 hoisted floats would spill)
__device__ __forceinline__ int ru_zero() {
  int z;
  asm volatile("mov.u32 %0, 0;" : "=r"(z));
  return z;
}
__device__ __forceinline__ void ru_add_bias32(float* v, const float* bias) {
  const float4* bp = reinterpret_cast<const float4*>(bias);
#pragma unroll
  for (int i = 0; i < 8; ++i) { const float4 t4 = __ldg(bp + i); v[4 * i] += t4.x; v[4 * i + 1] += t4.y; v[4 * i + 2] += t4.z; v[4 * i + 3] += t4.w; }
}

__global__ void __launch_bounds__(RU_THREADS, 1) residual_unit_tcgen05(const __grid_constant__ RuParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + RU_AITEMS * RU_AITEM_BYTES;
  uint8_t* smem_h = smem_b + RU_BSTAGES * RU_BBYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_h + RU_HBYTES);
  uint64_t* full_a = bars;
  uint64_t* empty_a = full_a + RU_AITEMS;
  uint64_t* full_b = empty_a + RU_AITEMS;
  uint64_t* empty_b = full_b + RU_BSTAGES;
  uint64_t* acc1_full = empty_b + RU_BSTAGES;
  uint64_t* h_full = acc1_full + 1;      // all epilogue threads: H written (and accumulator 1 drained)
  uint64_t* acc2_full = h_full + 1;
  uint64_t* acc2_empty = acc2_full + 1;  // all epilogue threads: accumulator 2 drained
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(acc2_empty + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.m_tiles * p.B;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmW7);
    tma_prefetch_desc(&p.tmW1);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < RU_AITEMS; ++i) { mbar_init(&full_a[i], 1); mbar_init(&empty_a[i], 1); }
    for (int i = 0; i < RU_BSTAGES; ++i) { mbar_init(&full_b[i], 1); mbar_init(&empty_b[i], 1); }
    mbar_init(acc1_full, 1);
    mbar_init(h_full, RU_EPI_WARPS * 32);
    mbar_init(acc2_full, 1);
    mbar_init(acc2_empty, RU_EPI_WARPS * 32);
    fence_barrier_init();
  }
  if (warp == 2) { tmem_alloc(tmem_ptr_smem, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t tmem_acc1 = tmem_base, tmem_acc2 = tmem_base + 256;

  if (warp == 0) {
    if (lane == 0) {   // weight slices: 2 channel blocks x 7 taps of conv7, then the 2 K-blocks of conv1
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        for (int s = 0; s < 2 * RU_TAPS + 2; ++s) {
          mbar_wait(&empty_b[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_b[stage], RU_BBYTES);
          if (s < 2 * RU_TAPS) {
            const int cib = s / RU_TAPS, tap = s % RU_TAPS;
            tma_load_2d(smem_b + stage * RU_BBYTES, &p.tmW7, &full_b[stage], tap * RU_C + cib * 64, 0);
          } else {
            tma_load_2d(smem_b + stage * RU_BBYTES, &p.tmW1, &full_b[stage], (s - 2 * RU_TAPS) * 64, 0);
          }
          if (++stage == RU_BSTAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 3) {
    if (lane == 0) {   // input windows: rows [t0 - pad, t0 - pad + 128 + 6 dil) of one 64-channel block, one per 128-step sub-tile
      int slot = 0; uint32_t phase = 0;
      const uint32_t item_bytes = static_cast<uint32_t>(p.win_rows) * 128u;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int b = tile / p.m_tiles, m0 = (tile % p.m_tiles) * 256;
        for (int cib = 0; cib < 2; ++cib) {
#pragma unroll
          for (int sub = 0; sub < 2; ++sub) {
            mbar_wait(&empty_a[slot], phase ^ 1);
            mbar_arrive_expect_tx(&full_a[slot], item_bytes);
            tma_load_4d(smem_a + slot * RU_AITEM_BYTES, &p.tmA, &full_a[slot], cib * 64, 0, m0 + sub * 128 - p.pad, b);
            if (++slot == RU_AITEMS) { slot = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    {   // MMA issue: whole warp in uniform control flow, one elected lane issues (umma_bf16_lo, common.cuh)
      const uint32_t leader = elect_one() ? 1u : 0u;
      constexpr uint32_t idesc = make_idesc_bf16(128, RU_C, 0, 0);
      int stage = 0; uint32_t phase = 0; int slot = 0; uint32_t sphase = 0;
      uint32_t tphase = 0;   // per-tile parity of the single-stage accumulator barriers
      const uint32_t lh0 = desc_lo_kmajor(smem_u32(smem_h));
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        // ---- conv7: accumulator 1 is free (h_full of the previous tile has been waited on below)
        for (int cib = 0; cib < 2; ++cib) {
          uint32_t la0[2];
          int slots[2];
#pragma unroll
          for (int sub = 0; sub < 2; ++sub) {
            mbar_wait(&full_a[slot], sphase);
            la0[sub] = desc_lo_kmajor(smem_u32(smem_a + slot * RU_AITEM_BYTES));
            slots[sub] = slot;
            if (++slot == RU_AITEMS) { slot = 0; sphase ^= 1; }
          }
          for (int tt = 0; tt < RU_TAPS; ++tt) {
            mbar_wait(&full_b[stage], phase);
            tc_fence_after();
            const uint32_t shift = static_cast<uint32_t>(tt * p.dil) * (128u >> 4);   // tap tt = the window shifted down tt*dil rows
            const uint32_t lb = desc_lo_kmajor(smem_u32(smem_b + stage * RU_BBYTES));
#pragma unroll
            for (int sub = 0; sub < 2; ++sub) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16_lo(tmem_acc1 + sub * RU_C, la0[sub] + shift + 2 * k, lb + 2 * k, idesc, (cib | tt | k) != 0, leader);
            }
            umma_commit_if(&empty_b[stage], leader);
            if (++stage == RU_BSTAGES) { stage = 0; phase ^= 1; }
          }
#pragma unroll
          for (int sub = 0; sub < 2; ++sub) umma_commit_if(&empty_a[slots[sub]], leader);
        }
        umma_commit_if(acc1_full, leader);
        // ---- conv1 on the activated intermediate (H tiles written by the epilogue warps)
        mbar_wait(h_full, tphase);
        mbar_wait(acc2_empty, tphase ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < 2; ++kb) {
          mbar_wait(&full_b[stage], phase);
          tc_fence_after();
          const uint32_t lb = desc_lo_kmajor(smem_u32(smem_b + stage * RU_BBYTES));
#pragma unroll
          for (int sub = 0; sub < 2; ++sub) {
            const uint32_t lh = lh0 + (sub * 2 + kb) * (16384 >> 4);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16_lo(tmem_acc2 + sub * RU_C, lh + 2 * k, lb + 2 * k, idesc, (kb | k) != 0, leader);
          }
          umma_commit_if(&empty_b[stage], leader);
          if (++stage == RU_BSTAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_if(acc2_full, leader);
        tphase ^= 1;
      }
    }
  } else if (warp >= 4) {
    const int ew = warp - 4;
    const int q = ew & 3;                  // TMEM lane quarter = 32 rows of a sub-tile
    const int cg = ew >> 2;                // 32-column group
    const int r = q * 32 + lane;           // this thread's row within a sub-tile
    const uint32_t stg_o = smem_u32(smem_h) + ew * 4096;   // per-warp transposition buffers (alias the H tiles)
    const uint32_t stg_a = stg_o + 2048;
    const int lrow = lane >> 2, lchunk = lane & 3;         // line side: 8 rows x 4 chunks per instruction
    const int swz = (lane >> 1) & 3;
    const uint32_t my_o = stg_o + lane * 64, my_a = stg_a + lane * 64;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const int col = cg * 32;
    // H operand address of this thread's row / 32-column slice: k-block cg>>1, 16-byte chunks (cg&1)*4 .. +3, XOR (row & 7)
    const uint32_t h_row = smem_u32(smem_h) + (cg >> 1) * 16384 + r * 128;
    const int h_c0 = (cg & 1) * 4, h_sw = r & 7;
    uint32_t tphase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int b = tile / p.m_tiles, m0 = (tile % p.m_tiles) * 256;
      // skip rows (line side: rows lrow + 8 i of this warp's 32, chunk lchunk of its 64 bytes)
      uint4 rres[4];
      auto load_res = [&](int sub) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int t = m0 + sub * 128 + q * 32 + lrow + 8 * i;
          rres[i] = (t < p.T) ? __ldg(reinterpret_cast<const uint4*>(p.res + (static_cast<size_t>(b) * p.T + t) * RU_C + col + lchunk * 8)) : make_uint4(0, 0, 0, 0);
        }
      };
      // ---- epilogue 1: accumulator 1 -> + b7 -> SnakeBeta(s1) -> H operand tiles
      if (lane == 0) tma_store_wait_read<0>();     // this warp's stores of the previous tile have read its transposition buffers ...
      __syncwarp();
      asm volatile("bar.sync 1, %0;" ::"n"(RU_EPI_WARPS * 32) : "memory");   // ... and so have everybody else's: H may be overwritten
      mbar_wait(acc1_full, tphase);
      tc_fence_after();
#pragma unroll 1
      for (int sub = 0; sub < 2; ++sub) {
        uint32_t raw[32];
        tmem_ld_32x32(tmem_acc1 + lane_off + sub * RU_C + col, raw);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
        const int z = ru_zero();
        ru_add_bias32(v, p.b7 + col + z);
        ru_snake32(v, p.s1_a + col + z, p.s1_invb + col + z);
        const uint32_t hr = h_row + sub * 32768;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          uint32_t hw[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) hw[e] = pack_bf16(v[8 * t + 2 * e], v[8 * t + 2 * e + 1]);
          ru_sts128(hr + (((h_c0 + t) ^ h_sw) << 4), make_uint4(hw[0], hw[1], hw[2], hw[3]));
        }
      }
      fence_proxy_async_smem();      // generic-proxy writes of H -> visible to the tensor core's async proxy
      tc_fence_before();
      mbar_arrive(h_full);
      load_res(0);                    // in flight while conv1 runs
      // ---- epilogue 2: accumulator 2 -> + b1 -> + skip -> raw plane; SnakeBeta(next) -> activated plane
      mbar_wait(acc2_full, tphase);   // conv1 has finished reading H: the transposition buffers may be written
      tc_fence_after();
#pragma unroll 1
      for (int sub = 0; sub < 2; ++sub) {
        if (sub == 1) {
          if (lane == 0) tma_store_wait_read<0>();   // sub-tile 0's stores have read the buffers
          __syncwarp();
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) { const int rr = lrow + 8 * i; ru_sts128(stg_o + rr * 64 + ((lchunk ^ ((rr >> 1) & 3)) << 4), rres[i]); }
        if (sub == 0) load_res(1);
        __syncwarp();
        uint32_t raw[32];
        tmem_ld_32x32(tmem_acc2 + lane_off + sub * RU_C + col, raw);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
        const int z = ru_zero();
        ru_add_bias32(v, p.b1 + col + z);
#pragma unroll
        for (int jx = 0; jx < 4; ++jx) {
          const uint4 u = ru_lds128(my_o + ((jx ^ swz) << 4));
          const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) { const float2 f = unpack_bf16(w[e]); v[8 * jx + 2 * e] += f.x; v[8 * jx + 2 * e + 1] += f.y; }
        }
        __syncwarp();                               // every lane has read its skip row before the buffer takes the raw output
        if (p.has_out) ru_stage_row(my_o, swz, v);
        if (p.has_act) {
          ru_snake32(v, p.nx_a + col + z, p.nx_invb + col + z);
          ru_stage_row(my_a, swz, v);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          const int trow = m0 + sub * 128 + q * 32;
          if (p.has_out) tma_store_4d(&p.tmOut, smem_h + ew * 4096, col, 0, trow, b);
          if (p.has_act) tma_store_4d(&p.tmAct, smem_h + ew * 4096 + 2048, col, 0, trow, b);
          tma_store_commit();
        }
      }
      tc_fence_before();
      mbar_arrive(acc2_empty);
      tphase ^= 1;
    }
    if (lane == 0) tma_store_wait_read<0>();   // shared memory must outlive the last stores' reads
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

}  // namespace b200sat

using namespace b200sat;

// One ResidualUnit forward (models/autoencoders.py:58-83) on bf16 time-major planes, C = 128:
//   out_raw = x_raw + conv1(snake1(conv7_dil(x_act) + b7)) + b1;   out_act = snake_next(out_raw)   (either output may be NULL)
// w7: packed [128][7*128] (b200sat_wn_pack layout), w1: packed [128][128]; snake parameters as produced by b200sat_snake_prep.
extern "C" int b200sat_residual_unit_fwd(const void* x_act, const void* x_raw, const void* w7, const float* b7, const float* s1_a,
                                         const float* s1_invb, const void* w1, const float* b1, const float* next_a, const float* next_invb,
                                         void* out_raw, void* out_act, int B, int T, int C, int dil, void* stream) {
  if (!x_act || !x_raw || !w7 || !w1 || !b7 || !b1 || !s1_a || !s1_invb || (!out_raw && !out_act) || B <= 0 || T <= 0 || dil < 1) {
    set_last_error("residual_unit: bad arguments"); return B200SAT_EINVAL;
  }
  if (C != RU_C) { set_last_error("residual_unit: the fused kernel is built for C == 128 (other widths: conv1d_fwd x2)"); return B200SAT_EUNSUPPORTED; }
  if (out_act && (!next_a || !next_invb)) { set_last_error("residual_unit: activated output needs the next SnakeBeta's parameters"); return B200SAT_EINVAL; }
  const int win_rows = 128 + (RU_TAPS - 1) * dil;
  if (win_rows * 128 > RU_AITEM_BYTES) { set_last_error("residual_unit: dilation too large for the shared window"); return B200SAT_EUNSUPPORTED; }
  RuParams p;
  memset(&p, 0, sizeof(p));
  int rc;
  {
    uint64_t dims[4] = {RU_C, 1, static_cast<uint64_t>(T), static_cast<uint64_t>(B)};
    uint64_t strides[3] = {RU_C * 2, RU_C * 2, static_cast<uint64_t>(RU_C) * T * 2};
    uint32_t box[4] = {64, 1, static_cast<uint32_t>(win_rows), 1};
    if ((rc = encode_tmap_bf16(&p.tmA, x_act, 4, dims, strides, box, 1))) return rc;
    uint32_t obox[4] = {32, 1, 32, 1};
    if (out_raw && (rc = encode_tmap_bf16(&p.tmOut, out_raw, 4, dims, strides, obox, 2))) return rc;
    if (out_act && (rc = encode_tmap_bf16(&p.tmAct, out_act, 4, dims, strides, obox, 2))) return rc;
  }
  {
    uint64_t dims[2] = {RU_TAPS * RU_C, RU_C};
    uint64_t strides[1] = {RU_TAPS * RU_C * 2};
    uint32_t box[2] = {64, RU_C};
    if ((rc = encode_tmap_bf16(&p.tmW7, w7, 2, dims, strides, box, 1))) return rc;
    uint64_t dims1[2] = {RU_C, RU_C};
    uint64_t strides1[1] = {RU_C * 2};
    if ((rc = encode_tmap_bf16(&p.tmW1, w1, 2, dims1, strides1, box, 1))) return rc;
  }
  p.res = static_cast<const __nv_bfloat16*>(x_raw);
  p.b7 = b7; p.s1_a = s1_a; p.s1_invb = s1_invb; p.b1 = b1; p.nx_a = next_a; p.nx_invb = next_invb;
  p.B = B; p.T = T; p.dil = dil; p.pad = 3 * dil; p.win_rows = win_rows;
  p.m_tiles = (T + 255) / 256;
  p.has_out = out_raw ? 1 : 0; p.has_act = out_act ? 1 : 0;
  static bool attr_set = false;
  if (!attr_set) {
    B200SAT_CHECK_CUDA(cudaFuncSetAttribute(residual_unit_tcgen05, cudaFuncAttributeMaxDynamicSharedMemorySize, RU_SMEM));
    attr_set = true;
  }
  const long tiles = static_cast<long>(p.m_tiles) * B;
  const int grid = static_cast<int>(tiles < num_sms() ? tiles : num_sms());
  residual_unit_tcgen05<<<grid, RU_THREADS, RU_SMEM, static_cast<cudaStream_t>(stream)>>>(p);
  B200SAT_CHECK_CUDA(cudaGetLastError());
  return B200SAT_OK;
}

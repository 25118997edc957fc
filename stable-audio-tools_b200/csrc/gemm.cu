// b200sat — tcgen05/TMA GEMM with fused epilogues (the DiT linear layers).
//
//   D[M,N] = epilogue( A[M,K] (bf16, K contiguous)  x  B[N,K]^T (bf16, K contiguous; an nn.Linear weight) )
//
// Replaces the cuBLASLt calls behind nn.Linear in the reference
// (stable_audio_tools/models/transformer.py:263,308,356-364,481,534,747-748) and the eager elementwise kernels
// that follow them: bias add, SwiGLU (transformer.py:272-275), residual add (:704-712), the partial NeoX RoPE on
// q/k (:154-174, :491-507) and SiLU (dit.py:41-76).
//
// Structure (one persistent CTA per SM, 256 threads):
//   warp 0  lane 0 : TMA producer  — cp.async.bulk.tensor tiles of A (128x64) and B (BNx64), 128B swizzle, kStages ring
//   warp 1  lane 0 : MMA issuer    — tcgen05.mma.cta_group::1.kind::f16, 128 x BN x 16, fp32 accumulators in TMEM
//   warp 2         : TMEM allocator (2 accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1)
//   warps 4..7     : epilogue      — tcgen05.ld 32x32b (thread == output row), fused math, 16-byte global stores
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
  int M, N, K;
  int ldd, ldr;
  int flags;
  int seg_in, seg_out, seg_off;
  int rope_seq, rope_dmodel, rope_dh;
  int n_half;
  int num_m_tiles, num_n_tiles;
};

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;   // 64 bf16 = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;

template <int BN>
struct GemmCfg {
  static constexpr int kStages = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int kABytes = BLOCK_M * BLOCK_K * 2;
  static constexpr int kBBytes = BN * BLOCK_K * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = 2 * BN;  // two accumulator stages
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

template <int BN>
__global__ void __launch_bounds__(256, 1) gemm_bf16_tcgen05(const __grid_constant__ GemmParams p) {
  using Cfg = GemmCfg<BN>;
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
  const int num_tiles = p.num_m_tiles * p.num_n_tiles;
  const int num_kb = (p.K + BLOCK_K - 1) / BLOCK_K;
  const bool swiglu = (p.flags & GEMM_SWIGLU) != 0;
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
      mbar_init(&tmem_empty[i], 128);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr_smem, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile % p.num_m_tiles;
        const int n_blk = tile / p.num_m_tiles;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          tma_load_2d(sa, &p.tmA, &full_bar[stage], kb * BLOCK_K, m_blk * BLOCK_M);
          if (swiglu) {
            tma_load_2d(sb, &p.tmB, &full_bar[stage], kb * BLOCK_K, n_blk * (BN / 2));
            tma_load_2d(sb + Cfg::kBBytes / 2, &p.tmB, &full_bar[stage], kb * BLOCK_K, p.n_half + n_blk * (BN / 2));
          } else {
            tma_load_2d(sb, &p.tmB, &full_bar[stage], kb * BLOCK_K, n_blk * BN);
          }
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc = make_idesc_bf16(BLOCK_M, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint32_t sb = sa + Cfg::kABytes;
          const uint64_t da = make_smem_desc_sw128(sa, 16, 1024);
          const uint64_t db = make_smem_desc_sw128(sb, 16, 1024);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // advance K inside the 128B swizzle row: 16 bf16 = 32 bytes = +2 in the (addr >> 4) field
            umma_bf16(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[as]);
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int q = warp - 4;  // == warp % 4: the TMEM lane quarter this warp may access
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile % p.num_m_tiles;
      const int n_blk = tile / p.num_m_tiles;
      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after();
      const int row = m_blk * BLOCK_M + q * 32 + lane;
      const bool row_ok = row < p.M;
      int out_row = row;
      if (p.flags & GEMM_ROW_REMAP) out_row = (row / p.seg_in) * p.seg_out + p.seg_off + (row % p.seg_in);
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN;
      const int n0 = n_blk * out_bn;
      for (int c = 0; c < out_bn / 32; ++c) {
        const int col = n0 + c * 32;
        if (col >= p.N) break;  // warp-uniform
        uint32_t raw[32];
        float v[32];
        tmem_ld_32x32(taddr + c * 32, raw);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
        const int ncols = min(32, p.N - col);
        if (p.flags & GEMM_BIAS) {
#pragma unroll
          for (int i = 0; i < 32; ++i) if (i < ncols) v[i] += __ldg(p.bias + col + i);
        }
        if (swiglu) {
          uint32_t graw[32];
          tmem_ld_32x32(taddr + BN / 2 + c * 32, graw);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float g = __uint_as_float(graw[i]);
            if ((p.flags & GEMM_BIAS) && i < ncols) g += __ldg(p.bias + p.n_half + col + i);
            // bf16 rounding points follow the reference's bf16 eager path: linear -> silu -> mul
            g = bf16_round(g);
            const float a = bf16_round(v[i]);
            v[i] = a * bf16_round(silu_f(g));
          }
        }
        if (p.flags & GEMM_SILU) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = silu_f(bf16_round(v[i]));
        }
        if ((p.flags & GEMM_ROPE) && row_ok) {
          // column -> (which in {q,k,v}, head, dim); rotate dims [0,32) of q and k heads (NeoX half-split, 16 freqs)
          const int which = col / p.rope_dmodel;
          const int dim0 = (col % p.rope_dmodel) % p.rope_dh;
          if (which < 2 && dim0 == 0) {
            const int pos = row % p.rope_seq;
            const float* cs = p.rope_cos + pos * 16;
            const float* sn = p.rope_sin + pos * 16;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float c_ = __ldg(cs + i), s_ = __ldg(sn + i);
              const float x1 = bf16_round(v[i]), x2 = bf16_round(v[i + 16]);
              v[i] = x1 * c_ - x2 * s_;
              v[i + 16] = x2 * c_ + x1 * s_;
            }
          }
        }
        if (row_ok) {
          if (p.flags & GEMM_RESIDUAL) {
            const __nv_bfloat16* r = p.residual + static_cast<size_t>(out_row) * p.ldr + col;
            float gate[32];
            if (p.flags & GEMM_GATE) {
              const float* gp = p.gate + static_cast<size_t>(row / p.seg_in) * p.N + col;
#pragma unroll
              for (int i = 0; i < 32; ++i) gate[i] = (i < ncols) ? __ldg(gp + i) : 0.f;
            }
            if (ncols >= 32) {
              const uint4* rp = reinterpret_cast<const uint4*>(r);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const uint4 u = __ldg(rp + i);
                const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float2 f = unpack_bf16(w[j]);
                  float y0 = bf16_round(v[8 * i + 2 * j]), y1 = bf16_round(v[8 * i + 2 * j + 1]);
                  if (p.flags & GEMM_GATE) { y0 = bf16_round(y0 * gate[8 * i + 2 * j]); y1 = bf16_round(y1 * gate[8 * i + 2 * j + 1]); }
                  v[8 * i + 2 * j] = y0 + f.x;
                  v[8 * i + 2 * j + 1] = y1 + f.y;
                }
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                if (i < ncols) {
                  float y = bf16_round(v[i]);
                  if (p.flags & GEMM_GATE) y = bf16_round(y * gate[i]);
                  v[i] = y + __bfloat162float(r[i]);
                }
              }
            }
          }
          if (p.flags & GEMM_OUT_F32) {
            store_chunk_f32(reinterpret_cast<float*>(p.D) + static_cast<size_t>(out_row) * p.ldd + col, v, ncols);
          } else {
            store_chunk_bf16(reinterpret_cast<__nv_bfloat16*>(p.D) + static_cast<size_t>(out_row) * p.ldd + col, v, ncols);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty[as]);
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

template <int BN>
static int launch_gemm(GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    B200SAT_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_tcgen05<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  const int out_bn = (p.flags & GEMM_SWIGLU) ? BN / 2 : BN;
  p.num_m_tiles = (p.M + BLOCK_M - 1) / BLOCK_M;
  p.num_n_tiles = (p.N + out_bn - 1) / out_bn;
  const int tiles = p.num_m_tiles * p.num_n_tiles;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  gemm_bf16_tcgen05<BN><<<grid, 256, Cfg::kSmemBytes, stream>>>(p);
  B200SAT_CHECK_CUDA(cudaGetLastError());
  return B200SAT_OK;
}

static int pick_bn(int M, int N, bool swiglu) {
  if (swiglu) return 256;
  const int sms = num_sms();
  const int mt = (M + BLOCK_M - 1) / BLOCK_M;
  int best = 256;
  double best_cost = 1e30;
  const int cands[3] = {256, 128, 64};
  // cost model: waves x per-tile time; narrow tiles re-read A from shared memory more often (lower MMA efficiency)
  const double eff[3] = {1.0, 0.92, 0.62};
  for (int i = 0; i < 3; ++i) {
    const int bn = cands[i];
    const int nt = (N + bn - 1) / bn;
    const long tiles = static_cast<long>(mt) * nt;
    const long waves = (tiles + sms - 1) / sms;
    const double cost = waves * (bn / eff[i]);
    if (cost < best_cost) { best_cost = cost; best = bn; }
  }
  return best;
}

}  // namespace b200sat

using namespace b200sat;

// C-ABI — see include/b200sat.h
extern "C" int b200sat_gemm_bf16(const void* A, int lda, const void* B, int ldb, void* D, int ldd, int M, int N, int K,
                                 int flags, const float* bias, const void* residual, int ldr, const float* rope_cos,
                                 const float* rope_sin, int rope_seq, int rope_dmodel, int rope_dh, int n_half, int seg_in,
                                 int seg_out, int seg_off, const float* gate, int force_bn, void* stream) {
  if (!A || !B || !D || M <= 0 || N <= 0 || K <= 0) { set_last_error("gemm: null pointer or empty shape"); return B200SAT_EINVAL; }
  if ((lda % 8) || (ldb % 8) || (K % 8)) { set_last_error("gemm: lda/ldb/K must be multiples of 8 (16-byte TMA strides)"); return B200SAT_EINVAL; }
  if ((ldd % 8) || ((flags & GEMM_RESIDUAL) && (ldr % 8))) { set_last_error("gemm: ldd/ldr must be multiples of 8"); return B200SAT_EINVAL; }
  if ((flags & GEMM_SWIGLU) && (N % 128 || n_half <= 0)) { set_last_error("gemm: swiglu needs N % 128 == 0 and n_half"); return B200SAT_EINVAL; }
  if ((flags & GEMM_ROPE) && (!rope_cos || !rope_sin || rope_dh != 64 || rope_seq <= 0)) { set_last_error("gemm: rope needs tables, dh == 64"); return B200SAT_EINVAL; }
  if ((flags & GEMM_GATE) && (!gate || !(flags & GEMM_RESIDUAL) || seg_in <= 0)) { set_last_error("gemm: gate needs residual + seg_in"); return B200SAT_EINVAL; }
  GemmParams p;
  memset(&p, 0, sizeof(p));
  const bool swiglu = flags & GEMM_SWIGLU;
  const int bn = force_bn ? force_bn : pick_bn(M, N, swiglu);
  if (bn != 64 && bn != 128 && bn != 256) { set_last_error("gemm: force_bn must be 64/128/256"); return B200SAT_EINVAL; }
  if (swiglu && bn != 256) { set_last_error("gemm: swiglu requires BN=256"); return B200SAT_EINVAL; }
  const int b_rows_total = swiglu ? 2 * N : N;  // value rows [0,N) and gate rows [n_half, n_half+N)
  {
    uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(M)};
    uint64_t strides[1] = {static_cast<uint64_t>(lda) * 2};
    uint32_t box[2] = {BLOCK_K, BLOCK_M};
    int rc = encode_tmap_bf16(&p.tmA, A, 2, dims, strides, box, 1);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(swiglu ? n_half + N : b_rows_total)};
    uint64_t strides[1] = {static_cast<uint64_t>(ldb) * 2};
    uint32_t box[2] = {BLOCK_K, static_cast<uint32_t>(swiglu ? bn / 2 : bn)};
    int rc = encode_tmap_bf16(&p.tmB, B, 2, dims, strides, box, 1);
    if (rc) return rc;
  }
  p.D = D; p.bias = bias; p.residual = static_cast<const __nv_bfloat16*>(residual);
  p.rope_cos = rope_cos; p.rope_sin = rope_sin; p.gate = gate;
  p.M = M; p.N = N; p.K = K; p.ldd = ldd; p.ldr = ldr; p.flags = flags;
  p.seg_in = seg_in > 0 ? seg_in : 1; p.seg_out = seg_out; p.seg_off = seg_off;
  p.rope_seq = rope_seq > 0 ? rope_seq : 1; p.rope_dmodel = rope_dmodel > 0 ? rope_dmodel : 1; p.rope_dh = rope_dh > 0 ? rope_dh : 64;
  p.n_half = n_half;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (bn) {
    case 256: return launch_gemm<256>(p, s);
    case 128: return launch_gemm<128>(p, s);
    default: return launch_gemm<64>(p, s);
  }
}

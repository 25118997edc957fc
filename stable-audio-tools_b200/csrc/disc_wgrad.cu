// b200sat — weight gradient of the 64 -> 64 channel flattened 2-D convs of the Encodec discriminator, all taps in ONE pass over the planes.
//
//     dW[tap][ca][cb] += sum_{b, t} dY[b, t, ca] * X[b, t + off[tap], cb]            (models/encodec.py:94-138 under autograd: dW = dY (*) X)
//
// Why a kernel of its own.  The taps of a 3 x 9 (or 3 x 3) kernel are row shifts dt*d*Fp + df of the SAME plane: three bands (dt) of nine
// (three) consecutive shifts.  `b200sat_conv_wgrad_taps` reads both planes from HBM once per tap (27 x; 180 us per tap at batch 32, the planes
// are 0.5 GB each); `b200sat_conv_wgrad_taps_cat` covers four taps per 128 x 256 tile but is bound by L2 -> shared-memory delivery (48 KB per
// 64 time steps and four taps = 46 B/clk/SM, profiles/r2_launches_ae_adv_b32_after.csv: 1.7 ms per 27-tap layer).  Here one pipeline stage
// holds 64 time steps of dY (8 KB) and, per band, ONE 72-row window of X (9 KB) that serves all of the band's taps: a tap is a start-address
// offset of df rows into the window (both operands are MN-major, rows = time, so a time shift moves the descriptor start by 128 B per row; the
// 128-byte swizzle is a function of the shared-memory address bits, as the conv kernel's row-shifted windows already rely on).
// Two taps share one MMA: M = 128 = (tap of the pair, cb), N = 64 = ca, the second tap's window is the first's plus the descriptor's leading
// byte offset.  Seven pairs of 128 x 64 fp32 accumulators fill 448 TMEM columns, so a 27-tap layer takes two passes (split over the grid) and
// 35 KB of shared-memory fill feed 28 MMAs: the planes stream twice instead of 27 (7) times.
//
// Roles (256 threads): warp 0 TMA producer, warp 1 MMA issue (lean path: whole warp, one elected lane), warp 2 TMEM allocator, warps 4-7
// epilogue (TMEM -> coalesced fp32 reds into dW; thread = (tap of the pair, cb), consecutive lanes = consecutive cb).
#include "common.cuh"

namespace b200sat {

constexpr int WG_ROWS = 64;                  // time steps per pipeline stage
constexpr int WG_WIN = 72;                   // rows of one band window (64 + 8 shifts)
constexpr int WG_A_BYTES = WG_ROWS * 128;    // dY tile
constexpr int WG_BAND_BYTES = WG_WIN * 128;  // 9216 = 9 swizzle atoms
constexpr int WG_STAGE_BYTES = 36864;        // 8192 + 3 * 9216 = 35840, rounded to a multiple of 1024
constexpr int WG_STAGES = 6;
constexpr int WG_MAX_PAIRS = 14;
constexpr int WG_PAIRS_PER_PASS = 7;
constexpr int WG_SMEM = WG_STAGES * WG_STAGE_BYTES + 1024 + 256;

struct WgWinParams {
  CUtensorMap tmA;   // dY plane  {64, 1, T, B}, box {64, 1, 64, 1}
  CUtensorMap tmX;   // X plane   {64, 1, T, B}, box {64, 1, 72, 1}
  float* dW;         // [ntaps][64 ca][64 cb]
  int kb_per_item, total_kb;
  int nbands, band_off[3];
  int npairs, npass, ctas_per_pass, kb_per_cta;
  int pair_addr[WG_MAX_PAIRS];      // byte offset of the pair's first tap inside the stage's window area
  int pair_lbo[WG_MAX_PAIRS];       // byte distance to its second tap (> 0)
  int pair_tap[WG_MAX_PAIRS][2];    // tap index stored from each half, or -1
};

__global__ void __launch_bounds__(256, 1) disc_wgrad_window_tcgen05(const __grid_constant__ WgWinParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + WG_STAGES * WG_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + WG_STAGES;
  uint64_t* acc_full = empty_bar + WG_STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(acc_full + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int pass = blockIdx.x / p.ctas_per_pass;
  const int slice = blockIdx.x % p.ctas_per_pass;
  const int kb_begin = slice * p.kb_per_cta;
  const int kb_end = min(p.total_kb, kb_begin + p.kb_per_cta);
  const int pair0 = pass * WG_PAIRS_PER_PASS;
  const int npairs = min(WG_PAIRS_PER_PASS, p.npairs - pair0);
  const bool active = pass < p.npass && kb_begin < kb_end && npairs > 0;

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&p.tmA); tma_prefetch_desc(&p.tmX); }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < WG_STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) { tmem_alloc(tmem_ptr_smem, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (active) {
    if (warp == 0) {
      if (lane == 0) {
        // ===================== TMA producer =====================
        int stage = 0; uint32_t phase = 0;
        const uint32_t bytes = WG_A_BYTES + p.nbands * WG_BAND_BYTES;
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          const int item = kb / p.kb_per_item, t0 = (kb % p.kb_per_item) * WG_ROWS;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], bytes);
          uint8_t* s = smem + stage * WG_STAGE_BYTES;
          tma_load_4d(s, &p.tmA, &full_bar[stage], 0, 0, t0, item);
          for (int b = 0; b < p.nbands; ++b)
            tma_load_4d(s + WG_A_BYTES + b * WG_BAND_BYTES, &p.tmX, &full_bar[stage], 0, 0, t0 + p.band_off[b], item);
          if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    } else if (warp == 1) {
      // ===================== MMA issue =====================
      const uint32_t leader = elect_one() ? 1u : 0u;
      constexpr uint32_t idesc = make_idesc_bf16(128, 64, 1, 1);
      int stage = 0; uint32_t phase = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * WG_STAGE_BYTES);
        const uint32_t lb = desc_lo_mnmajor(sa, 8192u);
        for (int j = 0; j < npairs; ++j) {
          const uint32_t la = desc_lo_mnmajor(sa + WG_A_BYTES + static_cast<uint32_t>(p.pair_addr[pair0 + j]), static_cast<uint32_t>(p.pair_lbo[pair0 + j]));
#pragma unroll
          for (int k = 0; k < WG_ROWS / 16; ++k)
            umma_bf16_lo(tmem_base + j * 64, la + k * (2048u >> 4), lb + k * (2048u >> 4), idesc, (kb != kb_begin) || (k != 0), leader);
        }
        umma_commit_if(&empty_bar[stage], leader);
        if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit_if(acc_full, leader);
    } else if (warp >= 4) {
      // ===================== epilogue: thread = accumulator row (half of the pair, cb), 64 columns = ca =====================
      const int q = warp & 3;
      const int row = q * 32 + lane;
      const int half = row >> 6, cb = row & 63;
      mbar_wait(acc_full, 0);
      tc_fence_after();
      for (int j = 0; j < npairs; ++j) {
        const int tap = p.pair_tap[pair0 + j][half];     // warp-uniform (a warp lies inside one half)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + j * 64 + c * 32, v);
          tmem_ld_wait();
          if (tap >= 0) {
            float* dst = p.dW + (static_cast<size_t>(tap) * 64 + c * 32) * 64 + cb;
#pragma unroll
            for (int i = 0; i < 32; ++i) atomicAdd(dst + i * 64, __uint_as_float(v[i]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

}  // namespace b200sat

using namespace b200sat;

// See include/b200sat.h.  Returns B200SAT_EUNSUPPORTED when the tap table is not "at most three bands of at most nine consecutive row shifts"
// (the caller then uses b200sat_conv_wgrad_taps_cat).
extern "C" int b200sat_conv_wgrad_taps_win(const void* a_plane, const void* b_plane, int T, const int* tap_off, int ntaps, float* dW, int B,
                                           void* stream) {
  if (!a_plane || !b_plane || !dW || !tap_off || ntaps <= 0 || B <= 0 || T <= 0) { set_last_error("conv_wgrad_taps_win: bad arguments"); return B200SAT_EINVAL; }
  if (ntaps > 2 * WG_MAX_PAIRS) { set_last_error("conv_wgrad_taps_win: more than 28 taps"); return B200SAT_EUNSUPPORTED; }
  WgWinParams p;
  memset(&p, 0, sizeof(p));
  int tap_addr[2 * WG_MAX_PAIRS];
  int nb = 0, start = 0;
  for (int k = 0; k < ntaps; ++k) {
    if (k > 0 && tap_off[k] <= tap_off[k - 1]) { set_last_error("conv_wgrad_taps_win: tap offsets must ascend"); return B200SAT_EUNSUPPORTED; }
    if (nb == 0 || tap_off[k] - start > WG_WIN - WG_ROWS) {
      if (nb == 3) { set_last_error("conv_wgrad_taps_win: more than three bands of row shifts"); return B200SAT_EUNSUPPORTED; }
      start = tap_off[k];
      p.band_off[nb++] = start;
    }
    tap_addr[k] = (nb - 1) * WG_BAND_BYTES + (tap_off[k] - start) * 128;
  }
  p.nbands = nb;
  // pairs of taps: (0,1), (2,3), ...; an odd count pairs the last tap with its predecessor again and stores only the last
  int np = 0;
  for (int k = 0; k + 1 < ntaps; k += 2, ++np) {
    p.pair_addr[np] = tap_addr[k]; p.pair_lbo[np] = tap_addr[k + 1] - tap_addr[k];
    p.pair_tap[np][0] = k; p.pair_tap[np][1] = k + 1;
  }
  if (ntaps & 1) {
    if (ntaps == 1) { set_last_error("conv_wgrad_taps_win: needs at least two taps"); return B200SAT_EUNSUPPORTED; }
    p.pair_addr[np] = tap_addr[ntaps - 2]; p.pair_lbo[np] = tap_addr[ntaps - 1] - tap_addr[ntaps - 2];
    p.pair_tap[np][0] = -1; p.pair_tap[np][1] = ntaps - 1;
    ++np;
  }
  p.npairs = np;
  p.npass = (np + WG_PAIRS_PER_PASS - 1) / WG_PAIRS_PER_PASS;
  auto plane_map = [](CUtensorMap* tm, const void* base, int Bn, int Tn, int rows) {
    uint64_t dims[4] = {64, 1, static_cast<uint64_t>(Tn), static_cast<uint64_t>(Bn)};
    uint64_t strides[3] = {128, 128, static_cast<uint64_t>(Tn) * 128};
    uint32_t box[4] = {64, 1, static_cast<uint32_t>(rows), 1};
    return encode_tmap_bf16(tm, base, 4, dims, strides, box, 1);
  };
  int rc;
  if ((rc = plane_map(&p.tmA, a_plane, B, T, WG_ROWS))) return rc;
  if ((rc = plane_map(&p.tmX, b_plane, B, T, WG_WIN))) return rc;
  p.dW = dW;
  p.kb_per_item = (T + WG_ROWS - 1) / WG_ROWS;
  p.total_kb = B * p.kb_per_item;
  int ctas = num_sms() / p.npass;
  if (ctas > p.total_kb) ctas = p.total_kb;
  if (ctas < 1) ctas = 1;
  p.kb_per_cta = (p.total_kb + ctas - 1) / ctas;
  ctas = (p.total_kb + p.kb_per_cta - 1) / p.kb_per_cta;    // no empty slices
  p.ctas_per_pass = ctas;
  static bool attr = false;
  if (!attr) {
    B200SAT_CHECK_CUDA(cudaFuncSetAttribute(disc_wgrad_window_tcgen05, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM));
    attr = true;
  }
  disc_wgrad_window_tcgen05<<<ctas * p.npass, 256, WG_SMEM, static_cast<cudaStream_t>(stream)>>>(p);
  B200SAT_CHECK_CUDA(cudaGetLastError());
  return B200SAT_OK;
}

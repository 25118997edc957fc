// b200sat — fused AdamW + EMA + bf16 weight refresh over a flat fp32 parameter buffer: one pass over HBM instead of three
// (torch.optim.AdamW step; ema_pytorch.EMA.update, training/diffusion.py:239-247, 489-491; the fp32 -> bf16 working-copy cast).
// Reads p, g, m, v, ema (20 B/param), writes p, m, v, ema, bf16 w (18 B/param).
#include "common.cuh"

namespace b200sat {

struct AdamArgs {
  float lr, beta1, beta2, eps, weight_decay, bc1, bc2, ema_decay, grad_scale;
  int ema_before_step;   // 1: the EMA takes the weights BEFORE this update (AutoencoderTrainingWrapper calls ema.update() ahead of
                         //    opt_gen.step(), training/autoencoders.py:499-506); 0: after it (DiffusionCondTrainingWrapper.on_before_zero_grad)
};

__global__ void __launch_bounds__(256) adamw_ema_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                        float* __restrict__ v, float* __restrict__ ema, __nv_bfloat16* __restrict__ w16,
                                                        long n, long n16, AdamArgs a) {
  const long n4 = n >> 2;
  const float inv_bc1 = 1.f / a.bc1, inv_sqrt_bc2 = rsqrtf(a.bc2);
  const float decay_w = 1.f - a.lr * a.weight_decay;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n4; i += static_cast<long>(gridDim.x) * blockDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    float pe[4] = {pp.x, pp.y, pp.z, pp.w};
    const float po[4] = {pp.x, pp.y, pp.z, pp.w};
    const float ge[4] = {gg.x, gg.y, gg.z, gg.w};
    float me[4] = {mm.x, mm.y, mm.z, mm.w}, ve[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gr = ge[j] * a.grad_scale;
      me[j] = a.beta1 * me[j] + (1.f - a.beta1) * gr;
      ve[j] = a.beta2 * ve[j] + (1.f - a.beta2) * gr * gr;
      const float denom = sqrtf(ve[j]) * inv_sqrt_bc2 + a.eps;
      pe[j] = pe[j] * decay_w - a.lr * inv_bc1 * (me[j] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = make_float4(pe[0], pe[1], pe[2], pe[3]);
    reinterpret_cast<float4*>(m)[i] = make_float4(me[0], me[1], me[2], me[3]);
    reinterpret_cast<float4*>(v)[i] = make_float4(ve[0], ve[1], ve[2], ve[3]);
    if (ema) {
      float4 ee = reinterpret_cast<float4*>(ema)[i];
      const float* src = a.ema_before_step ? po : pe;
      ee.x = ee.x * a.ema_decay + src[0] * (1.f - a.ema_decay);
      ee.y = ee.y * a.ema_decay + src[1] * (1.f - a.ema_decay);
      ee.z = ee.z * a.ema_decay + src[2] * (1.f - a.ema_decay);
      ee.w = ee.w * a.ema_decay + src[3] * (1.f - a.ema_decay);
      reinterpret_cast<float4*>(ema)[i] = ee;
    }
    if (w16 && (i << 2) < n16) {
      reinterpret_cast<uint2*>(w16)[i] = make_uint2(pack_bf16(pe[0], pe[1]), pack_bf16(pe[2], pe[3]));
    }
  }
}

}  // namespace b200sat

using namespace b200sat;

extern "C" int b200sat_adamw_ema_step(float* p, const float* g, float* m, float* v, float* ema, void* w_bf16, long n, long n_bf16, float lr,
                                      float beta1, float beta2, float eps, float weight_decay, int step, float ema_decay, float grad_scale,
                                      int ema_before_step, void* stream) {
  if (!p || !g || !m || !v || n <= 0 || step < 1) { set_last_error("adamw_ema_step: bad arguments"); return B200SAT_EINVAL; }
  if ((n & 3) || (n_bf16 & 3) || n_bf16 > n) { set_last_error("adamw_ema_step: element counts must be multiples of 4 (pad the flat buffer)"); return B200SAT_EINVAL; }
  // step >= (1 << 24) marks a per-layer slice launched on a side stream NEXT TO the backward pass (b200sat/ddp.py): many short blocks
  // that leave the SM quickly instead of a grid-stride loop that would hold registers the persistent GEMM CTAs are waiting for
  const bool slice = step >= (1 << 24);
  if (slice) step -= (1 << 24);
  AdamArgs a;
  a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay;
  a.bc1 = 1.f - powf(beta1, static_cast<float>(step));
  a.bc2 = 1.f - powf(beta2, static_cast<float>(step));
  a.ema_decay = ema_decay; a.grad_scale = grad_scale; a.ema_before_step = ema_before_step ? 1 : 0;
  const long n4 = n >> 2;
  long blocks = (n4 + 255) / 256;
  const long cap = static_cast<long>(num_sms()) * 16;
  if (!slice && blocks > cap) blocks = cap;
  adamw_ema_kernel<<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(p, g, m, v, ema, static_cast<__nv_bfloat16*>(w_bf16), n,
                                                                                             w_bf16 ? n_bf16 : 0, a);
  B200SAT_CHECK_CUDA(cudaGetLastError());
  return B200SAT_OK;
}

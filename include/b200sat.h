/* b200sat — C ABI of the B200-native Stable Audio hot path.
 *
 * The reference (Stability-AI/stable-audio-tools) has no FFI: its "operators" are PyTorch nn.Module forward bodies.
 * Each entry point below replaces the ATen/cuDNN/cuBLAS/cuFFT calls behind one of those bodies; the reference
 * location is cited per function (paths relative to /root/reference/stable_audio_tools).  The Python shim
 * (stable-audio-tools_b200/b200sat) binds these with ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless stated otherwise; the library never allocates and never synchronises,
 *     so every entry is CUDA-graph capturable; `stream` is a cudaStream_t passed as void*.
 *   - return value: 0 ok; <0 invalid argument / unsupported shape (b200sat_last_error() has the text);
 *     >0 a cudaError_t from the launch.
 *   - bf16 tensors are row-major with the stated leading dimensions (elements).
 */
#ifndef B200SAT_H
#define B200SAT_H
#ifdef __cplusplus
extern "C" {
#endif

const char* b200sat_last_error(void);
int b200sat_version(void);
int b200sat_num_sms(void);
/* Grid budget of the persistent kernels while a collective shares the GPU (the reference's DDP all-reduce overlaps its backward,
 * train.py:124-164): limit > 0 caps the SM count the persistent GEMM / conv kernels size their grids from (rounded down to an even
 * number), limit <= 0 restores the whole device.  Returns the previous limit. */
int b200sat_set_sm_limit(int limit);
unsigned long long b200sat_launch_count(void);

/* GEMM flags (bitmask) */
#define B200SAT_GEMM_BIAS 1
#define B200SAT_GEMM_RESIDUAL 2
#define B200SAT_GEMM_SILU 4
#define B200SAT_GEMM_SWIGLU 8
#define B200SAT_GEMM_ROPE 16
#define B200SAT_GEMM_OUT_F32 32
#define B200SAT_GEMM_ROW_REMAP 64
#define B200SAT_GEMM_GATE 128
#define B200SAT_GEMM_A_MN 256        /* A stored [K,M]: weight-gradient GEMMs (dW = dY^T X) */
#define B200SAT_GEMM_B_MN 512        /* B stored [K,N]: data-gradient GEMMs (dX = dY W) */
#define B200SAT_GEMM_ACCUM 1024      /* fp32 D += acc */
#define B200SAT_GEMM_LN_A 8192       /* A = raw LayerNorm input, gamma folded into B: out = rstd*(acc - mean*colsum) (transformer.py:236-238 fused) */
#define B200SAT_GEMM_ROWSTATS 16384  /* accumulate (sum, sumsq) of each output row into out_stats for the next fused LayerNorm */
#define B200SAT_GEMM_SWIGLU_BWD 2048 /* D[M,2N] = SwiGLU backward of acc against saved pre-activation aux */

/* D[M,N] = epilogue(A[M,K] x B[N,K]^T), bf16 in, fp32 accumulate (tcgen05 + TMA).
 * Replaces nn.Linear (cuBLASLt) + the eager epilogues of models/transformer.py:263-275 (GLU/SwiGLU), :308 (ff out),
 * :356-364,:481 (to_qkv/to_q/to_kv), :534 (to_out), :154-174,:491-507 (partial RoPE on q,k), :704-712 (residual adds),
 * :677-701 (adaLN gate), :747-748 (project_in/out) and models/dit.py:41-76 (SiLU MLPs). */
int b200sat_gemm_bf16(const void* A, int lda, const void* B, int ldb, void* D, int ldd, int M, int N, int K, int flags,
                      const float* bias, const void* residual, int ldr, const float* rope_cos, const float* rope_sin,
                      int rope_seq, int rope_dmodel, int rope_dh, int n_half, int seg_in, int seg_out, int seg_off,
                      const float* gate, void* aux, int ld_aux, const float* ln_stats, const float* ln_colsum, float ln_eps,
                      float* out_stats, int force_bn, void* stream);

/* Flash attention forward (tcgen05; non-causal; head_dim 64; Hq % Hkv == 0 grouped-query).  q/k/v/o are bf16 views
 * [B, N, heads, 64] addressed by element strides (batch, sequence, head); lse (optional, fp32 [B,Hq,Nq]) is the natural-log
 * row normaliser kept for the backward pass.  Replaces models/transformer.py:406-441 (flash_attn_func / SDPA dispatch,
 * GQA repeat_interleave :408-411) and the head rearranges of :469-482, :526. */
int b200sat_attention_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int B, int Hq, int Hkv, int Nq,
                          int Nk, long q_bs, long q_ss, long q_hs, long k_bs, long k_ss, long k_hs, long v_bs, long v_ss,
                          long v_hs, long o_bs, long o_ss, long o_hs, int head_dim, float scale, void* stream);

/* LayerNorm over the last dim, bf16 in/out, fp32 statistics; gamma (and optional beta) fp32 [D].  With scale/shift
 * (fp32 [B, ld_mod] rows, rows_per_batch rows of x per batch entry) applies the adaLN modulate y*(1+scale)+shift.
 * Replaces models/transformer.py:236-241 (LayerNorm.forward) and :680-682, :695-697. */
int b200sat_layernorm_fwd(const void* x, long ldx, const float* gamma, const float* beta, const float* scale,
                          const float* shift, long ld_mod, int rows_per_batch, void* y, long ldy, int rows, int D, float eps,
                          void* stream);

/* y[M,N] = act(x[M,K] w[N,K]^T + bias) (+ add), 1 <= M <= 8: the conditioning MLPs whose M is the batch size.
 * Replaces models/dit.py:41-76,:140-168 (to_timestep_embed / to_global_embed) and models/transformer.py:767-773,:677-684. */
int b200sat_small_linear(const void* x, long ldx, const void* w, long ldw, const float* bias, const void* add, long ldadd,
                         void* y, long ldy, int M, int N, int K, int act_silu, int out_f32, int act_sigmoid_1m, float* stats,
                         long stats_stride, void* stream);

/* out[B, 2*half] = [cos(2 pi t w) | sin(2 pi t w)] (bf16).  t is read at t[*step * t_stride + b] when step != NULL so a
 * captured CUDA graph can walk a per-step table.  Replaces models/blocks.py:85-94 (FourierFeatures.forward). */
int b200sat_fourier_features(const float* t, const void* w, void* out, int B, int half, const int* step, int t_stride,
                             void* stream);

/* DiT input stage: bf16(x*c_in) -> 1x1 conv + residual -> [reps*B*T, C] bf16 rows (token-major).
 * Replaces models/dit.py:193-195 (+ the CFG batch duplication :330-331 via reps).  x fp32 [B,C,T]; c_in = cin_table[*step]. */
int b200sat_dit_pre(const float* x, const void* wconv, void* out, int B, int C, int T, int reps, const float* cin_table,
                    const int* step, void* stream);

/* DiT input stage with channel-concatenated conditioning (inpainting; dit.py:160-165 `torch.cat([x, input_concat_cond], dim=1)`, CFG duplication
 * :336-337): out rows [(rep*B+b)*T + t][Cp] bf16 = (x[b,:,t]*c_in | cond[b,:,t] | 0).  x fp32 [B,C,T], cond fp32 [B,Dc,T]; the 1x1
 * preprocess_conv + residual (dit.py:193) then run as b200sat_gemm_bf16 with the residual epilogue on a [Cp,Cp] zero-padded weight. */
int b200sat_dit_concat(const float* x, const float* cond, void* out, int B, int C, int Dc, int Cp, int T, int reps, const float* cin_table,
                       const int* step, void* stream);

/* DiT output stage: drop `prepend` tokens, transpose to [B,C,T], 1x1 conv + residual, optional classifier-free guidance
 * over the (cond|uncond) batch halves with std-rescale.  Replaces models/dit.py:219-224 and :398-408.  out fp32 [B,C,T]. */
int b200sat_dit_post(const void* h, long ld_batch, int prepend, const void* wconv, float* out, int B, int C, int T, int cfg,
                     float cfg_scale, float scale_phi, void* stream);

/* Sampler state update for one step s = *step: den = v*c0 + x*c1; x = c2*x + c3*den + c4*d1 + c5*d2 + c6*noise[s];
 * coef is fp32 [steps, 8]; hist is a ring of 3 denoised tensors.  Covers k-diffusion's VDenoiser + sample_dpmpp_3m_sde
 * (called from inference/sampling.py:351-387) and the in-repo v-DDIM update (inference/sampling.py:281-300). */
int b200sat_sampler_update(float* x, const float* v, float* hist, const float* noise, const float* coef, int* step, long n,
                           int advance, void* stream);
int b200sat_step_set(int* step, int value, void* stream);

/* ---- Oobleck autoencoder (models/autoencoders.py) ------------------------------------------------------------------
 * Activations are time-major bf16 plane pairs (hi, lo), [B, T, C] each, hi + lo ~ fp32 value; lo pointers may be NULL
 * when passes == 1 (plain bf16). */

/* Implicit-GEMM Conv1d / ConvTranspose1d on tcgen05 with fused bias, residual add and SnakeBeta epilogue.
 *   mode 0: stride-1 (dilated) conv, zero padding `pad`        — ResidualUnit convs, autoencoders.py:68-72
 *   mode 1: strided conv, kernel `taps`, stride `stride`        — EncoderBlock down-sampler, :245-246
 *   mode 2: transposed conv, kernel 2*stride, stride `stride`   — DecoderBlock up-sampler, :267-269
 * w_hi/w_lo: packed by b200sat_wn_pack.  out_*: raw result planes (optional); act_*: SnakeBeta(result) planes with the
 * consumer layer's snake_a = exp(alpha), snake_invb = 1/(exp(beta)+1e-9) (optional).  res_*: residual planes added
 * before both (ResidualUnit skip, :83).  passes: 1 (bf16) or 3 (hi*hi + hi*lo + lo*hi, fp32-class). */
int b200sat_conv1d_fwd(const void* in_hi, const void* in_lo, const void* w_hi, const void* w_lo, const float* bias,
                       const void* res_hi, const void* res_lo, void* out_hi, void* out_lo, void* act_hi, void* act_lo,
                       const float* snake_a, const float* snake_invb, int B, int T_in, int Cin, int Cout, int taps, int dil,
                       int pad, int stride, int mode, int passes, void* stream);

/* One ResidualUnit (models/autoencoders.py:58-83: Snake -> WNConv1d k7 dilated -> Snake -> WNConv1d k1 -> + x) in ONE launch, bf16 planes,
 * C == 128 (other widths: two b200sat_conv1d_fwd calls).  x_act = snake0(x) and x_raw = x are the [B, T, 128] planes the previous layer's
 * epilogue wrote; w7 / w1 are b200sat_wn_pack outputs ([128][7*128], [128][128]); s1_* / next_* are b200sat_snake_prep outputs.
 *   out_raw = x_raw + conv1(snake1(conv7_dil(x_act) + b7)) + b1,   out_act = snake_next(out_raw)      (either may be NULL)
 * The k7 output never reaches HBM: 2.0 GB instead of 3.1 GB of traffic per unit at T = 2 097 152. */
int b200sat_residual_unit_fwd(const void* x_act, const void* x_raw, const void* w7, const float* b7, const float* s1_a, const float* s1_invb,
                              const void* w1, const float* b1, const float* next_a, const float* next_invb, void* out_raw, void* out_act,
                              int B, int T, int C, int dil, void* stream);

/* weight_norm (w = g*v/||v||, norm over all dims but 0; autoencoders.py:23-27) + packing to the GEMM layout as hi/lo bf16
 * planes.  v fp32 [Cout,Cin,K] (conv) or [Cin,Cout,K] (transposed); g NULL = plain weight. */
int b200sat_wn_pack(const float* v, const float* g, float* inv_norm_scratch, void* w_hi, void* w_lo, int Cout, int Cin, int K,
                    int transposed, int stride, void* stream);

/* SnakeBeta parameters (log-scale alpha, beta; blocks.py:321-329) -> exp(alpha), 1/(exp(beta)+1e-9). */
int b200sat_snake_prep(const float* alpha, const float* beta, float* a, float* invb, int C, void* stream);

/* First conv of the encoder (audio channels -> C; autoencoders.py:303): x fp32 [B,Cin,T] -> planes [B,T,Cout] (+snake). */
int b200sat_conv_in(const float* x, const float* w, const float* bias, const float* snake_a, const float* snake_invb,
                    void* out_hi, void* out_lo, void* act_hi, void* act_lo, int B, int Cin, int T, int Cout, int K, int pad,
                    void* stream);

/* Last conv of the decoder (C -> audio channels; autoencoders.py:355-357): planes [B,T,Cin] -> y fp32 [B,Cout,T]. */
int b200sat_conv_out(const void* in_hi, const void* in_lo, const float* w, const float* bias, float* y, int B, int Cin, int T,
                     int Cout, int K, int pad, int tanh_out, void* stream);

/* fp32 [B,C,T] -> hi/lo planes [B,T,C] (decoder input latents). */
int b200sat_to_planes(const float* x, void* hi, void* lo, int B, int C, int T, void* stream);

/* VAE bottleneck (models/bottleneck.py:105-134): planes [B,T,2L] (mean|scale) -> z = noise*(softplus(scale)+1e-4)+mean
 * fp32 [B,L,T]; optional fp32 [B,2L,T] copy of (mean|scale); kl_sum += sum(mean^2 + var - log var - 1). */
int b200sat_vae_sample(const void* hi, const void* lo, const float* noise, float* z, float* mean_scale_out, float* kl_sum,
                       int B, int L, int T, void* stream);

/* ---- backward pass of the DiT blocks --------------------------------------------------------------------------------- */

/* Flash-attention backward (tcgen05): dQ, dK, dV from dO with the forward's O and LSE; optional inverse RoPE on dQ/dK rows
 * (rope_cos/sin [N,16], position = sequence index) so gradients land in the pre-rotation layout of the qkv projection.
 * strides: HOST array of 8 x (batch, seq, head) element strides for q, k, v, o, dO, dQ, dK, dV.  delta_scratch: fp32 workspace of B*Hq*(Nq + 2*roundup(Nq,64)) elements
 * (-lse*log2(e) and -delta padded to [2][B,Hq,roundup(Nq,64)] for the aligned broadcast loads of the dK/dV kernel, then delta [B,Hq,Nq];
 * the pointer must be 16-byte aligned).
 * Backward of models/transformer.py:406-441 (+ :154-174). */
int b200sat_attention_bwd(const void* q, const void* k, const void* v, const void* o, const void* d_o, const float* lse,
                          float* delta_scratch, void* dq, void* dk, void* dv, int B, int Hq, int Hkv, int Nq, int Nk,
                          const long* strides, int head_dim, float scale, const float* rope_cos, const float* rope_sin,
                          void* stream);

/* LayerNorm backward: dx_out = dres + dLN(dy; x, gamma); dgamma (fp32 [D], optional) accumulates.  Backward of
 * models/transformer.py:236-238 fused with the residual-stream gradient add of :703-712. */
int b200sat_layernorm_bwd(const void* x, long ldx, const void* dy, long lddy, const float* gamma, const void* dres, long ldr,
                          void* dx_out, long ldo, float* dgamma, int rows, int D, float eps, void* stream);

/* adaLN variant (transformer.py:680-697, y = LN(x; gamma) * (1 + scale_b) + shift_b): gain gamma * (1 + mod_scale[b, :]) (fp32 [B, ld_mod],
 * rows_per_batch rows of x per batch entry); dp (fp32 [B, D], +=) receives sum_n dy * xhat per batch entry: dgamma = sum_b (1 + scale_b) dp_b,
 * dscale_b = gamma * dp_b, dshift_b = per-batch column sums of dy. */
int b200sat_layernorm_mod_bwd(const void* x, long ldx, const void* dy, long lddy, const float* gamma, const float* mod_scale, long ld_mod,
                              int rows_per_batch, const void* dres, long ldr, void* dx_out, long ldo, float* dp, int rows, int D, float eps,
                              void* stream);

/* adaLN gate backward (transformer.py:690-701, x = x + branch * g_b with g_b = sigmoid(1 - gate_b), fp32 [B, D]):
 * dbranch = dh * g_b (bf16), dgate[b, :] += sum_n dh * branch (fp32; the host applies d sigmoid). */
int b200sat_gate_bwd(const void* dh, long ldh, const void* branch, long ldb, const float* gate, void* dbranch, long ldo, float* dgate,
                     int rows_per_batch, int batches, int D, void* stream);

/* out[n] += sum_m dY[m,n]: bias gradients of nn.Linear. */
int b200sat_colsum(const void* dy, long ld, float* out, int M, int N, void* stream);

/* ---- multi-resolution STFT loss (training/losses/auraloss.py) -------------------------------------------------------- */

/* out[b,r,t] = FIR_taps( sum_c mix[r,c] x[b,c,t] ), zero padded: FIRFilter.forward (A-weighting, auraloss.py:155-169) fused with
 * SumAndDifference (:44-73) / channel selection via the R x C mixing matrix.  x fp32 [B,C,T] -> out fp32 [B,R,T], R <= 4. */
int b200sat_stft_prefilter(const float* x, float* out, const float* mix, const float* taps, int B, int C, int T, int R, int ntaps,
                           void* stream);

/* One STFT resolution over `rows` mono signal pairs (input xf, target yf; fp32 [rows, T]): per row acc[row] += { sum (|Y|-|X|)^2,
 * sum |Y|^2, sum |log|X| - log|Y|| } over all frames and bins, |.| = sqrt(max(re^2+im^2, eps)); torch.stft semantics
 * (center, reflect pad, periodic hann `window` [n_fft], onesided).  twiddle: fp32 pairs exp(-2 pi i k / n_fft), k < n_fft/2.
 * Replaces STFTLoss.stft + SpectralConvergenceLoss + STFTMagnitudeLoss (auraloss.py:171-223, :368-449); nothing but the three
 * sums leaves the SM. */
int b200sat_stft_loss_accumulate(const float* xf, const float* yf, double* acc, const float* window, const float* twiddle, int rows,
                                 int T, int n_fft, int hop, float eps, void* stream);

/* Backward of b200sat_stft_loss_accumulate for one resolution: recomputes each frame's transform, forms the magnitude gradients from
 * the per-row coefficients coef[row] = {a1, a2, a3} (a1 = c_sc/sqrt(S1 S2), a2 = c_sc sqrt(S1)/S2^1.5, a3 = c_log), runs ONE
 * inverse-direction complex FFT per frame for both signals and scatter-adds d loss / d xf, d yf (fp32 [rows,T], accumulated).
 * Backward of auraloss.py:368-449. */
int b200sat_stft_loss_backward(const float* xf, const float* yf, float* dxf, float* dyf, const float* coef, const float* window,
                               const float* twiddle, int rows, int T, int n_fft, int hop, float eps, void* stream);

/* Backward of b200sat_stft_prefilter: dx[B,C,T] = mix^T . FIR^T(dout[B,R,T]). */
int b200sat_stft_prefilter_backward(const float* dout, float* dx, const float* mix, const float* taps, int B, int C, int T, int R,
                                    int ntaps, void* stream);

/* ---- backward of the Oobleck convolutions -------------------------------------------------------------------------------- */

/* One tap of a conv weight gradient: dW[Ca,Cb] (fp32, +=) = sum_{b, t < T_iter} A[b,(t+offA)*sA+rA,:]^T (x) B[b,(t+offB)*sB+rB,:],
 * A/B = time-major bf16 planes [B,T,C] read in place (out-of-range rows = zero padding).  tcgen05 GEMM with both operands
 * MN-major, split-K over time.  Backward (dW = dY (*) X) of models/autoencoders.py:58-83, :233-283. */
int b200sat_conv_wgrad(const void* a_plane, int Ca, int Ta, int sA, int rA, int offA, const void* b_plane, int Cb, int Tb, int sB,
                       int rB, int offB, float* dW, int B, int T_iter, void* stream);

/* All taps of a flattened 2-D conv weight gradient: dW[tap][Ca][Cb] += sum_{b,t} A[b,t,:]^T (x) B[b,t+tap_off[tap],:] (ntaps launches of the above). */
int b200sat_conv_wgrad_taps(const void* a_plane, int Ca, const void* b_plane, int Cb, int T, const int* tap_off, int ntaps, float* dW, int B,
                            void* stream);

/* The same weight gradient for 64 -> 64 channel convs in ONE launch: a 128 x 256 accumulator tile covers four taps (column block -> tap row
 * shift), so both planes stream from HBM ceil(ntaps/4) times instead of ntaps times.  Output layout dWc[Ca=64][ntaps][Cb=64] (fp32, +=);
 * ntaps <= 32.  Backward of the Conv2d stacks of models/encodec.py:94-138 (dW = dY (*) X). */
int b200sat_conv_wgrad_taps_cat(const void* a_plane, const void* b_plane, int T, const int* tap_off, int ntaps, float* dWc, int B, void* stream);

/* The same weight gradient in ONE pass over the planes (csrc/disc_wgrad.cu): per 64 time steps one dY tile and, per band of consecutive row
 * shifts (the frequency taps of one time offset), one 72-row window of X in shared memory; every tap is a row-shifted MN-major descriptor
 * into its window, two taps share a 128 x 64 MMA.  dW[ntaps][64][64] (fp32, +=).  tap_off must ascend and form at most three bands of at
 * most nine consecutive shifts (else B200SAT_EUNSUPPORTED: use the _cat entry); ntaps in [2, 28]. */
int b200sat_conv_wgrad_taps_win(const void* a_plane, const void* b_plane, int T, const int* tap_off, int ntaps, float* dW, int B, void* stream);

/* SnakeBeta backward fused with the skip-connection add and the parameter reductions (models/blocks.py:291-329 under autograd):
 * d_raw = d_skip + d_act * (1 + invb*a*sin(2 a x)); dalpha/dbeta [C] (log-scale parameters) and dbias [C] (= column sums of d_raw, the
 * bias gradient of the conv that produced x; optional) are accumulated with fp32 atomics.  Planes are bf16 [rows, C]. */
int b200sat_snake_bwd(const void* d_act, const void* x_raw, const void* d_skip, const float* snake_a, const float* snake_invb, void* d_raw,
                      float* dalpha, float* dbeta, float* dbias, long rows, int C, void* stream);

/* Weight-normalised weights packed for the data-gradient convolution (run through b200sat_conv1d_fwd on the output gradient):
 * mode 0 -> flipped taps, run as mode 0 with pad' = dil*(K-1)-pad; mode 1 (strided) -> run as mode 2; mode 2 (transposed) -> run
 * as mode 1; channels swap roles.  inv_norm = the 1/||v|| rows b200sat_wn_pack left in its scratch.  autoencoders.py:23-27. */
int b200sat_wn_pack_dgrad(const float* v, const float* g, const float* inv_norm, void* out, int Cout, int Cin, int K, int mode, int stride,
                          void* stream);

/* weight_norm backward: dw_taps fp32 [K][R][Cc] (filled by b200sat_conv_wgrad, one tap per launch) -> dv [R][Cc][K], dg [R]
 * (g NULL: plain weight, dv = dw re-laid out).  torch.nn.utils.weight_norm under autograd, autoencoders.py:23-27. */
int b200sat_wn_bwd(const float* dw_taps, const float* v, const float* g, const float* inv_norm, float* dv, float* dg, int R, int Cc, int K,
                   void* stream);

/* Weight gradient of the audio-channel edge convs: dW[c*stride_c + a*stride_a + k] += sum_{b,t} plane[b,t,c]*sig[b,a,t+sign*(k-pad)];
 * plane bf16 [B,T,C], sig fp32 [B,A,T].  Encoder conv_in (autoencoders.py:303): sign +1; decoder conv_out (:355-357): sign -1. */
int b200sat_edge_wgrad(const void* plane, const float* sig, float* dW, int B, int T, int C, int A, int K, int pad, int sign, long stride_c,
                       long stride_a, void* stream);

/* VAE bottleneck backward (bottleneck.py:105-113): dz bf16 plane [B,T,L] (NULL = 0), ms forward plane [B,T,2L], noise fp32 [B,L,T];
 * the KL term contributes kl_scale * (*kl_grad) * d(sum_c(mean^2 + var - log var - 1)); d_ms bf16 plane [B,T,2L]. */
int b200sat_vae_sample_bwd(const void* dz, const void* ms, const float* noise, const float* kl_grad, float kl_scale, void* d_ms, int B,
                           int L, int T, void* stream);

/* ---- optimizer ------------------------------------------------------------------------------------------------------- */

/* Fused AdamW (decoupled weight decay, bias correction at `step` >= 1) + EMA + bf16 working-copy refresh over a flat fp32 buffer:
 * g' = g*grad_scale; m, v updated; p = p(1 - lr wd) - lr/(1-b1^t) * m / (sqrt(v/(1-b2^t)) + eps); ema = ema*ema_decay + p*(1-ema_decay)
 * (ema NULL = off); the first n_bf16 elements of p are also written as bf16 to w_bf16 (NULL = off).  One HBM pass for
 * torch.optim.AdamW + ema_pytorch.EMA.update (training/diffusion.py:239-247, 489-491; training/utils.py:60-79) + the weight cast.
 * step | (1 << 24): "slice" launch - the call covers one layer's slice of the flat buffers and runs next to the backward pass on a side
 * stream (b200sat/ddp.py): many short blocks instead of a grid-stride loop, so SM resources return to the backward's kernels quickly.
 * ema_before_step != 0: the EMA averages the weights as they were BEFORE this update — AutoencoderTrainingWrapper.training_step calls
 * autoencoder_ema.update() ahead of opt_gen.step() (training/autoencoders.py:499-506); 0 = after it (DiffusionCondTrainingWrapper). */
int b200sat_adamw_ema_step(float* p, const float* g, float* m, float* v, float* ema, void* w_bf16, long n, long n_bf16, float lr, float beta1,
                           float beta2, float eps, float weight_decay, int step, float ema_decay, float grad_scale, int ema_before_step,
                           void* stream);

/* ---- Encodec multi-scale STFT discriminator (models/encodec.py:38-138, models/discriminators.py:13-58) ---------------------------
 * One scale's activations are flattened planes [B, P, C], P = frames * (F + 8), F = n_fft/2 + 1: bins in columns [4, 4+F) of each frame's
 * row group, zero pad columns either side, so a 2-D conv tap (dt, df) with dilation (d, 1) is the row shift dt*d*(F+8) + df. */

/* The 64 -> 64 channel Conv2d layers (3x9 with dilation (d,1), 3x3) and their data gradients on the tcgen05 conv kernel: w packed by
 * b200sat_wn_pack (K = ntaps; b200sat_wn_pack_dgrad mode 0 for the gradient), tap_off[ntaps] (host array) = row shift per tap, rows whose
 * (row % fp) is outside [f0, f1) are written as zeros, optional LeakyReLU (encodec.py:80-87, :102-103). */
int b200sat_conv2d_flat(const void* in, const void* w, const float* bias, void* out, int B, int P, int Cin, int Cout, int ntaps,
                        const int* tap_off, int fp, int f0, int f1, float leaky, void* stream);

/* First conv (4 -> 64 channels, 3 x 9; encodec.py:77-79) on the tensor cores: the nine frequency taps are folded into channels,
 * S9[row][ci*9 + df] = spec[row + df - 4][ci] (bf16 [B, P, 64], 36 channels used), which makes the layer a 3-tap (row shifts -Fp, 0, +Fp)
 * 64 -> 64 flattened conv served by b200sat_conv2d_flat / b200sat_conv_wgrad_taps_cat / b200sat_wn_pack / b200sat_wn_bwd.
 * backward == 0: s9 = pack(spec fp32 [B, P, 4]);  backward != 0: spec (d spec) = pack^T(s9 (d S9)). */
int b200sat_disc_spec_pack(float* spec, void* s9, int B, int frames, int F, int backward, void* stream);

/* Complex STFT front end (torchaudio Spectrogram normalized=True, center=False, power=None, hann, win = n_fft; encodec.py:72-74, :96-99):
 * x fp32 [B,2,T] -> spec fp32 [B,P,4] = (re ch0, re ch1, im ch0, im ch1); pad columns are NOT written (pre-zero the buffer).
 * backward != 0: spec holds d spec and x (fp32 [B,2,T]) ACCUMULATES d audio. */
int b200sat_disc_stft(const float* x, float* spec, const float* window, const float* twiddle, int B, int T, int n_fft, int hop, int backward,
                      void* stream);

/* First conv (4 -> 64, 3x9) + LeakyReLU: spec -> out bf16 [B,P,64] (w fp32 [64,4,27] weight-normalised, encodec.py:77-79); and/or its data
 * gradient dpre bf16 [B,P,64] -> dspec fp32 [B,P,4].  Either pair may be NULL. */
int b200sat_disc_conv0(const float* spec, const float* w, const float* bias, void* out, const void* dpre, float* dspec, int B, int frames, int F,
                       float leaky, void* stream);

/* conv_post (64 -> 1, 3x3, encodec.py:88-90): act bf16 [B,P,64] -> logits fp32 [B,P] (pad columns 0). */
int b200sat_disc_convpost(const void* act, const float* w, const float* bias, float* logits, int B, int frames, int F, void* stream);

/* Hinge sums over the valid bins (discriminators.py:13-16): sums[0] += sum relu(1-lt), sums[1] += sum relu(1+lf), sums[2] += sum lf. */
int b200sat_disc_hinge_sums(const float* lt, const float* lf, double* sums, int B, int frames, int F, void* stream);

/* out[0] += sum |a - b| over two bf16 planes of n elements (feature matching, discriminators.py:24, :41-47). */
int b200sat_disc_l1_sum(const void* a, const void* b, double* out, long n, void* stream);

/* d logits: mode 0 generator (-scale), 1 discriminator/reals (-scale where 1-l>0), 2 discriminator/fakes (+scale where 1+l>0). */
int b200sat_disc_logit_grad(const float* logits, float* g, int B, int frames, int F, int mode, float scale, void* stream);

/* Backward through one feature map: d_pre = (d_in + conv_post^T(d_logit) + fm_coef*sign(post-other)) * (post>0 ? 1 : leaky); any of the
 * three sources may be NULL; pad columns -> 0. */
int b200sat_disc_act_bwd(const void* d_in, const float* d_logit, const float* w_post, const void* post, const void* other, float fm_coef,
                         float leaky, void* d_pre, int B, int frames, int F, void* stream);

/* Weight gradients of the first conv (dW fp32 [64,4,27] += sum_p d_pre[p] (x) spec[p + tap]) and of conv_post (dW fp32 [64,9] +=
 * sum_p g[p] * act[p + tap], dbias += sum g) for the discriminator step (training/autoencoders.py:476-489). */
int b200sat_disc_conv0_wgrad(const void* dpre, const float* spec, float* dW, int B, int frames, int F, void* stream);
int b200sat_disc_convpost_wgrad(const float* g, const void* act, float* dW, float* dbias, int B, int frames, int F, void* stream);

#ifdef __cplusplus
}
#endif
#endif

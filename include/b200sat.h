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

/* D[M,N] = epilogue(A[M,K] x B[N,K]^T), bf16 in, fp32 accumulate (tcgen05 + TMA).
 * Replaces nn.Linear (cuBLASLt) + the eager epilogues of models/transformer.py:263-275 (GLU/SwiGLU), :308 (ff out),
 * :356-364,:481 (to_qkv/to_q/to_kv), :534 (to_out), :154-174,:491-507 (partial RoPE on q,k), :704-712 (residual adds),
 * :677-701 (adaLN gate), :747-748 (project_in/out) and models/dit.py:41-76 (SiLU MLPs). */
int b200sat_gemm_bf16(const void* A, int lda, const void* B, int ldb, void* D, int ldd, int M, int N, int K, int flags,
                      const float* bias, const void* residual, int ldr, const float* rope_cos, const float* rope_sin,
                      int rope_seq, int rope_dmodel, int rope_dh, int n_half, int seg_in, int seg_out, int seg_off,
                      const float* gate, int force_bn, void* stream);

#ifdef __cplusplus
}
#endif
#endif

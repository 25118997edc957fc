// b200sat — library plumbing: error reporting, device query, TMA descriptor encoding.
#include "common.cuh"
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <mutex>

namespace b200sat {

static thread_local char g_err[512] = "";
void set_last_error(const char* msg) {
  strncpy(g_err, msg ? msg : "", sizeof(g_err) - 1);
  g_err[sizeof(g_err) - 1] = 0;
}

int pdl_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("B200SAT_PDL"); v = (e && e[0] == '0') ? 0 : 1; }
  return v;
}

static int g_sm_limit = 0;   // 0 = all SMs; set while a collective kernel needs SMs of its own (b200sat_set_sm_limit)
int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return (g_sm_limit > 0 && g_sm_limit < n) ? g_sm_limit : n;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    // resolved from the installed driver at run time: no link-time dependency on libcuda.so
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box, int swizzle128) {
  EncodeTiledFn fn = get_encode();
  if (!fn) { set_last_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)"); return B200SAT_EDRIVER; }
  if (reinterpret_cast<uintptr_t>(base) & 15) { set_last_error("tensor map base must be 16-byte aligned"); return B200SAT_EINVAL; }
  cuuint64_t d[5]; cuuint64_t s[4]; cuuint32_t b[5]; cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) {
    s[i] = strides_bytes[i];
    if (s[i] & 15) { set_last_error("tensor map strides must be multiples of 16 bytes"); return B200SAT_EINVAL; }
  }
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), d, s, b, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle128 == 1 ? CU_TENSOR_MAP_SWIZZLE_128B : (swizzle128 == 2 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE),
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled failed: CUresult %d (rank %d dims %llu,%llu box %u,%u)", (int)r, rank,
             (unsigned long long)d[0], (unsigned long long)(rank > 1 ? d[1] : 0), b[0], rank > 1 ? b[1] : 0);
    set_last_error(buf);
    return B200SAT_EDRIVER;
  }
  return B200SAT_OK;
}

}  // namespace b200sat

extern "C" const char* b200sat_last_error() { return b200sat::g_err; }
extern "C" int b200sat_version() { return 100; }
extern "C" int b200sat_num_sms() { return b200sat::num_sms(); }
// Persistent kernels size their grids from num_sms(): while NCCL's all-reduce kernel shares the GPU with the backward pass
// (b200sat/ddp.py), leaving it `reserve` SMs keeps every persistent CTA resident in ONE wave instead of queueing a second wave behind
// the collective.  limit <= 0 restores the full device; the value is rounded down to an even count (CTA pairs).  Returns the old limit.
extern "C" int b200sat_set_sm_limit(int limit) {
  const int old = b200sat::g_sm_limit;
  b200sat::g_sm_limit = limit > 0 ? (limit & ~1) : 0;
  return old;
}
// Number of kernel launches issued by this library since load (claimed in bench.py's gpu_launches).
namespace b200sat { unsigned long long g_launches = 0; }
extern "C" unsigned long long b200sat_launch_count() { return b200sat::g_launches; }

// common.cuh -- shared helpers for libpio_sm100 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <atomic>

#include "../../include/pio.h"

namespace pio {

extern thread_local char g_err[512];
extern std::atomic<long long> g_launches;

inline int fail(int code, const char* fmt, ...) __attribute__((format(printf, 2, 3)));
}  // namespace pio

#include <stdarg.h>
namespace pio {
inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define PIO_CUDA(expr)                                                                          \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess)                                                                      \
      return pio::fail(PIO_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

#define PIO_CHECK(cond, ...)                                   \
  do {                                                         \
    if (!(cond)) return pio::fail(PIO_EINVAL, __VA_ARGS__);    \
  } while (0)

// after every kernel launch: count it and pick up launch-configuration errors
#define PIO_LAUNCHED()                                                                          \
  do {                                                                                          \
    pio::g_launches.fetch_add(1, std::memory_order_relaxed);                                    \
    cudaError_t _e = cudaGetLastError();                                                        \
    if (_e != cudaSuccess)                                                                      \
      return pio::fail(PIO_ECUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

#define PIO_TRY(expr)          \
  do {                         \
    int _r = (expr);           \
    if (_r != PIO_OK) return _r; \
  } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

constexpr int kNumSMs = 148;  // B200

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device setting: remember it per (call site, device), so that a process
// that drives several GPUs still configures every one of them (one process per GPU is the deployment model, this is the guard).
struct SmemAttrOnce {
  bool done[64] = {};
  template <typename K>
  cudaError_t ensure(K kernel, int bytes) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64 && done[dev]) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess && dev >= 0 && dev < 64) done[dev] = true;
    return e;
  }
};

// ---- programmatic dependent launch (PDL): a kernel launched through launch_pdl() may start while its predecessor in
// the stream is still draining; it must execute pdl_wait() before it touches anything a predecessor wrote (or before it
// overwrites anything a predecessor reads).  Everything ahead of pdl_wait() -- barrier init, tensor-memory allocation,
// tensor-map prefetch -- then overlaps the predecessor's tail.  PIO_PDL=0 restores plain stream order.
bool pdl_enabled();  // elementwise.cu
bool pdl_kind_enabled(int kind);  // PIO_PDL_OFF=<bit mask of kinds> launches those kernels fully serialised; see elementwise.cu
// RAII: every kernel launched by this thread while the object lives is fully serialised (the ViT forward: see elementwise.cu)
struct PdlScopeOff {
  PdlScopeOff();
  ~PdlScopeOff();
};
enum { PDL_KIND_OTHER = 0, PDL_KIND_LN = 1, PDL_KIND_GEMM = 2, PDL_KIND_GEMM2 = 3, PDL_KIND_ATTN = 4 };
// griddepcontrol.wait makes the predecessor's writes visible to this grid's ordinary (generic-proxy) accesses; the TMA engine reads
// through the async proxy, which needs its own fence.  Without it a kernel that TMA-loads what its predecessor wrote with
// st.global (LayerNorm rows, the 256-wide GEMM tile's direct stores) raced under programmatic dependent launch: the bf16 ViT
// at 4 x 224 px differed from run to run in 12 of 29 forwards (max |diff| 0.06), 0 of 29 with the attention kernel launched
// fully serialised (profiles/r02y_*, r02z_*) -- and 0 of 29 with this fence.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;\n\tfence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_k(int kind, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl_enabled() && pdl_kind_enabled(kind)) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// -------------------------------------------------------------------------------- device helpers
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_new(float x) {
  const float k = 0.79788456080286535588f;  // sqrt(2/pi)
  return 0.5f * x * (1.0f + tanhf(k * (x + 0.044715f * x * x * x)));
}
__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == PIO_ACT_GELU_ERF) return gelu_erf(v);
  if (act == PIO_ACT_GELU_NEW) return gelu_new(v);
  if (act == PIO_ACT_TANH) return tanhf(v);
  if (act == PIO_ACT_RELU) return fmaxf(v, 0.f);
  return v;
}

// The epilogue shared by the SIMT and the tcgen05 GEMMs (see PioLinear in pio.h).
struct Epilogue {
  const float* bias;
  const float* colscale;
  const float* gamma;
  const float* residual;
  const float* res_rowscale;
  int ldres;
  float alpha;
  int act;
  int rows_per_group, group_stride, group_offset;
  float* argmax_val;
  int* argmax_idx;
  float* argmax_sumexp;
  int argmax_ld;
  const float* exp_ref;
  float* exp_psum;
  float* exp_pmax;
  int exp_ld;
  float* split_ws;  // deterministic split-K: partial tiles [work item][128][BN] fp32 (gemm_sm100.cu)
  int* split_cnt;   //                        arrival counters per output tile, zero between launches
  __device__ __forceinline__ long long out_row(int m) const {
    if (rows_per_group == 0) return m;
    return (long long)(m / rows_per_group) * group_stride + group_offset + (m % rows_per_group);
  }
  // value for output element (m, n); orow = out_row(m)
  __device__ __forceinline__ float apply(float acc, int m, long long orow, int n) const {
    float v = acc * alpha;
    if (colscale) v *= __ldg(colscale + n);
    if (bias) v += __ldg(bias + n);
    v = apply_act(v, act);
    if (gamma) v *= __ldg(gamma + n);
    if (residual) {
      float r = residual[orow * ldres + n];
      if (res_rowscale) r *= __ldg(res_rowscale + m);
      v += r;
    }
    return v;
  }
};

inline Epilogue make_epilogue(const PioLinear& p) {
  Epilogue e;
  e.bias = p.bias;
  e.colscale = p.colscale;
  e.gamma = p.gamma;
  e.residual = p.residual;
  e.res_rowscale = p.res_rowscale;
  e.ldres = p.ldres;
  e.alpha = p.alpha;
  e.act = p.act;
  e.rows_per_group = p.rows_per_group;
  e.group_stride = p.group_stride;
  e.group_offset = p.group_offset;
  e.argmax_val = p.argmax_val;
  e.argmax_idx = p.argmax_idx;
  e.argmax_sumexp = p.argmax_sumexp;
  e.argmax_ld = p.argmax_ld;
  e.exp_ref = p.exp_ref;
  e.exp_psum = p.exp_psum;
  e.exp_pmax = p.exp_pmax;
  e.exp_ld = p.exp_ld;
  e.split_ws = nullptr;
  e.split_cnt = nullptr;
  return e;
}

// kernels implemented in other translation units
int linear_simt(const PioLinear& p, cudaStream_t st);    // gemm_simt.cu  (fp32 FFMA)
int linear_tc(const PioLinear& p, cudaStream_t st);      // gemm_sm100.cu (tcgen05 / TMA / TMEM)
int argmax_slabs_tc(int M, int N);
void release_split_scratch();                             // gemm_sm100.cu
bool linear_tc2_eligible(const PioLinear& p);           // gemm2_sm100.cu: 2-CTA (cta_group::2) 256 x 256 pair tiles
int linear_tc2(const PioLinear& p, cudaStream_t st);                       // column slabs a fused arg-max call writes per row
int f32_to_bf16(const float* in, __nv_bfloat16* out, long long n, cudaStream_t st);
int transpose_f32(const float* in, float* out, int rows, int cols, cudaStream_t st);           // out[c][r] = in[r][c]
int transpose_to_bf16(const float* in, __nv_bfloat16* out, int rows, int cols, cudaStream_t st);
int layernorm(const float* x, int ldx, const float* w, const float* b, void* out, int out_dt, int ldo, int rows, int dim,
              float eps, cudaStream_t st);
int im2col14(const float* imgs, void* cols, int cols_dt, int B, int S, int g, int Kp, cudaStream_t st);
int init_global_tokens(float* x, const float* cls, const float* reg, const float* pos, int B, int N, int D, cudaStream_t st);
int l2norm_rows(float* x, int rows, int dim, cudaStream_t st);
int softmax_rows(const float* in, float* out, int rows, int cols, float scale, cudaStream_t st);
int cls_attention(const void* qkv, int dt, int B, int N, int D, int ng, float* out, float* logits_ws, cudaStream_t st);
int vit_attention(const void* qkv, void* out, int dt, int B, int N, int H, cudaStream_t st);
size_t vit_attention_tc_workspace(int B, int N, int H);
int vit_attention_tc(const void* qkv, void* out, void* vt_ws, int B, int N, int H, cudaStream_t st);  // attention_sm100.cu
int decode_attention(const void* qkv, void* kc, void* vc, void* out, int dt, int R, int H, int T, int t, cudaStream_t st,
                     const int* anc = nullptr);  // anc: per-(row, position) cache-row table of the beam search (12 x 64 kernel only)
// bidirectional (or causal) attention over n <= 64 tokens; kc / vc != null also writes positions 0..n-1 of a KV cache [R][H][Tmax][hd]
int small_attention(const void* q, long long ldq, const void* kv, long long ldkv, void* out, long long ldo, int dt, int R, int n, int H,
                    int hd, bool causal, void* kc, void* vc, int Tmax, cudaStream_t st);
size_t small_attention_smem(int n, int hd);

}  // namespace pio

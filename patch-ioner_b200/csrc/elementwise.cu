// elementwise.cu -- HBM-bound helpers: LayerNorm, dtype/layout repacks, im2col, token assembly,
// row L2-normalisation, row softmax, row arg-max.  All vectorised 16-byte accesses, one warp per
// row where a row reduction is needed (warp-shuffle reductions, no shared memory).
#include "common.cuh"
#include <stdlib.h>

namespace pio {

thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("PIO_PDL"); return !(e && e[0] == '0'); }();
  return on;
}
// Programmatic dependent launch and reproducibility (round 2).  Launched as programmatic dependents, the kernels of the bf16 ViT
// forward made it differ from run to run (4 x 224 px: 12 of 29 forwards, max |diff| 0.06; profiles/r02y_*, r02z_*, r02ab_*).  Two
// causes were separated: (1) TMA loads of what the predecessor wrote with st.global need an async-proxy fence after
// griddepcontrol.wait -- now inside pdl_wait() (29/29 -> 0/29 differing forwards with direct-store epilogues); (2) a remainder
// tied to the tcgen05 attention kernel overlapping the tail of the qkv GEMM, which in a process that had used several streams
// survived serialising the attention launch alone.  It is not root-caused; with LayerNorm, both GEMM kinds and the attention
// serialised the forward is bit-reproducible in every context tested.  Hence: inside pio_vit_forward / pio_vit_block_rows
// (PdlScopeOff) every launch is fully serialised -- the kernels there run 50-700 us each, the measured cost is 0.4 % of a
// bench step -- while the decode / projection path, whose kernels are a few us long and whose repeatability tests are green,
// keeps the overlap (+31 % at 32 rows without it).  PIO_PDL_OFF=<bit mask over PDL_KIND_*> serialises kinds everywhere.
static thread_local int g_pdl_scope_off = 0;
// PIO_VIT_PDL=1 makes the scope inert (A/B runs: the ViT forward with programmatic overlap again)
static bool pdl_scope_inert() {
  static const bool v = [] { const char* e = getenv("PIO_VIT_PDL"); return e && e[0] == '1'; }();
  return v;
}
PdlScopeOff::PdlScopeOff() { if (!pdl_scope_inert()) ++g_pdl_scope_off; }
PdlScopeOff::~PdlScopeOff() { if (!pdl_scope_inert()) --g_pdl_scope_off; }
bool pdl_kind_enabled(int kind) {
  static const int off = [] { const char* e = getenv("PIO_PDL_OFF"); return e ? atoi(e) : 0; }();
  return g_pdl_scope_off == 0 && ((off >> kind) & 1) == 0;
}
}  // namespace pio

namespace pio {
namespace {

// ------------------------------------------------------------------------------ LayerNorm
// One warp per row; dim % 128 == 0, dim <= 1024 -> each lane owns dim/128 float4.
template <int VEC>  // float4 per lane
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ w,
                                                        const float* __restrict__ b, void* out, int out_dt, int ldo,
                                                        int rows, float eps) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  pdl_wait();
  pdl_launch_dependents();
  if (warp >= rows) return;
  constexpr int DIM = VEC * 128;
  const float4* xr = reinterpret_cast<const float4*>(x + (long long)warp * ldx);
  float4 v[VEC];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    v[i] = __ldcg(xr + lane + 32 * i);  // L2, not L1: under programmatic dependent launch this CTA may share its SM (and its L1) with the producer
    s += v[i].x + v[i].y + v[i].z + v[i].w;
  }
  const float mean = warp_sum(s) * (1.0f / DIM);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    float a = v[i].x - mean, c = v[i].y - mean, d = v[i].z - mean, e = v[i].w - mean;
    q += a * a + c * c + d * d + e * e;
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / DIM) + eps);
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const float4 g = __ldg(reinterpret_cast<const float4*>(w) + lane + 32 * i);
    const float4 bb = __ldg(reinterpret_cast<const float4*>(b) + lane + 32 * i);
    float4 o;
    o.x = (v[i].x - mean) * rstd * g.x + bb.x;
    o.y = (v[i].y - mean) * rstd * g.y + bb.y;
    o.z = (v[i].z - mean) * rstd * g.z + bb.z;
    o.w = (v[i].w - mean) * rstd * g.w + bb.w;
    if (out_dt == PIO_DT_F32) {
      reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + (long long)warp * ldo)[lane + 32 * i] = o;
    } else {
      __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&lo);
      pk.y = *reinterpret_cast<uint32_t*>(&hi);
      reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out) + (long long)warp * ldo)[lane + 32 * i] = pk;
    }
  }
}

__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ src, long long src_ld, const int* __restrict__ idx,
                                                          int T, int D, float* __restrict__ out) {
  const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (t >= T) return;
  const float4* s = reinterpret_cast<const float4*>(src + (long long)idx[t] * src_ld);
  float4* o = reinterpret_cast<float4*>(out + (long long)t * D);
  for (int i = lane; i < D / 4; i += 32) o[i] = __ldg(s + i);
}
__global__ void __launch_bounds__(64) segment_mean_kernel(const float* __restrict__ x, const int* __restrict__ seg_start,
                                                          const int* __restrict__ seg_len, int D, float* __restrict__ out) {
  const int s = blockIdx.x, c4 = blockIdx.y * 64 + threadIdx.x;
  if (c4 >= D / 4) return;
  const int r0 = seg_start[s], n = seg_len[s];
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int r = 0; r < n; ++r) {
    const float4 v = reinterpret_cast<const float4*>(x + (long long)(r0 + r) * D)[c4];
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  const float inv = 1.0f / (float)n;  // n == 0 -> inf -> 0 * inf = NaN
  reinterpret_cast<float4*>(out + (long long)s * D)[c4] = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
}

__global__ void f32_to_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n4) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n4; i += stride) {
    float4 v = reinterpret_cast<const float4*>(in)[i];
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&lo);
    pk.y = *reinterpret_cast<uint32_t*>(&hi);
    reinterpret_cast<uint2*>(out)[i] = pk;
  }
}
__global__ void f32_to_bf16_tail(const float* in, __nv_bfloat16* out, long long start, long long n) {
  long long i = start + threadIdx.x;
  if (i < n) out[i] = __float2bfloat16(in[i]);
}

template <typename OutT>
__global__ void transpose_kernel(const float* __restrict__ in, OutT* __restrict__ out, int rows, int cols) {
  __shared__ float tile[32][33];
  int c = blockIdx.x * 32 + threadIdx.x, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    int r = r0 + i;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? in[(long long)r * cols + c] : 0.f;
  }
  __syncthreads();
  int orow0 = blockIdx.x * 32, ocol = blockIdx.y * 32 + threadIdx.x;
  for (int i = threadIdx.y; i < 32; i += 8) {
    int orow = orow0 + i;
    if (orow < cols && ocol < rows) {
      float v = tile[threadIdx.x][i];
      if constexpr (sizeof(OutT) == 4)
        out[(long long)orow * rows + ocol] = v;
      else
        out[(long long)orow * rows + ocol] = __float2bfloat16(v);
    }
  }
}

// ------------------------------------------------------------------------------ patch-embed im2col
// imgs [B,3,S,S] -> cols [B*P, Kp] with k = c*196 + ky*14 + kx (Conv2d weight order), zero padded to Kp.
template <typename OutT>
__global__ void im2col14_kernel(const float* __restrict__ imgs, OutT* __restrict__ cols, int B, int S, int g, int Kp) {
  const long long total = (long long)B * g * g * Kp;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    int k = (int)(i % Kp);
    long long row = i / Kp;
    float v = 0.f;
    if (k < 588) {
      int p = (int)(row % (g * g));
      int b = (int)(row / (g * g));
      int c = k / 196, r = k % 196, ky = r / 14, kx = r % 14;
      int py = p / g, px = p % g;
      v = imgs[(((long long)b * 3 + c) * S + py * 14 + ky) * S + px * 14 + kx];
    }
    if constexpr (sizeof(OutT) == 4)
      cols[i] = v;
    else
      cols[i] = __float2bfloat16(v);
  }
}

// The same layout, one CTA per (image, row of patches): the 3 x 14 image rows of the strip are read fully coalesced into shared
// memory (converted on the way), then the g output rows are written fully coalesced, two bf16 per thread.  The element-per-thread
// kernel above reads 14-pixel runs 2 KB apart and spends six integer divisions per element: 272 us for 64 x 518 px (1.1 TB/s of
// 310 MB) -- kept for image sizes whose strip does not fit shared memory.
template <typename OutT>
__global__ void __launch_bounds__(256) im2col14_strip_kernel(const float* __restrict__ imgs, OutT* __restrict__ cols, int S, int g, int Kp) {
  extern __shared__ __align__(16) unsigned char strip_raw[];
  OutT* strip = reinterpret_cast<OutT*>(strip_raw);  // [3][14][S]
  const int py = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  for (int row = warp; row < 3 * 14; row += 8) {  // a warp per image row of the strip: no index arithmetic beyond constants
    const int c = row / 14, ky = row - c * 14;
    const float* src = imgs + (((long long)b * 3 + c) * S + py * 14 + ky) * S;
    OutT* dst = strip + row * S;
    for (int x = lane; x < S; x += 32) {
      const float v = __ldg(src + x);
      if constexpr (sizeof(OutT) == 4) dst[x] = v; else dst[x] = __float2bfloat16(v);
    }
  }
  __syncthreads();
  OutT* out = cols + ((long long)b * g * g + (long long)py * g) * Kp;
  for (int px = warp; px < g; px += 8) {          // a warp per output row (one patch)
    OutT* orow = out + (long long)px * Kp;
    const OutT* sp = strip + px * 14;
    if constexpr (sizeof(OutT) == 2) {
      for (int k = 2 * lane; k < Kp; k += 64) {    // Kp is even: two bf16 per thread
        __nv_bfloat16 v[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int kk = k + e;
          const int cr = kk / 14, kx = kk - cr * 14;  // cr = c * 14 + ky: the strip row
          v[e] = kk < 588 ? sp[cr * S + kx] : __float2bfloat16(0.f);
        }
        *reinterpret_cast<__nv_bfloat162*>(orow + k) = __halves2bfloat162(v[0], v[1]);
      }
    } else {
      for (int k = lane; k < Kp; k += 32) {
        const int cr = k / 14, kx = k - cr * 14;
        orow[k] = k < 588 ? sp[cr * S + kx] : OutT(0.f);
      }
    }
  }
}

// cls/register rows of the token matrix: x[b,0] = cls + pos[0]; x[b,1..4] = reg (no pos_embed)
__global__ void init_global_tokens_kernel(float* __restrict__ x, const float* __restrict__ cls, const float* __restrict__ reg,
                                          const float* __restrict__ pos, int B, int N, int D) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * 5 * D) return;
  int d = i % D, t = (i / D) % 5, b = i / (5 * D);
  float v = (t == 0) ? cls[d] + pos[d] : reg[(t - 1) * D + d];
  x[((long long)b * N + t) * D + d] = v;
}

// ------------------------------------------------------------------------------ row ops
__global__ void l2norm_kernel(float* __restrict__ x, int rows, int dim) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  float* r = x + (long long)warp * dim;
  float s = 0.f;
  for (int i = lane; i < dim; i += 32) s += r[i] * r[i];
  const float n = sqrtf(warp_sum(s));
  for (int i = lane; i < dim; i += 32) r[i] = r[i] / n;
}

// one CTA per row: out = softmax(in * scale)
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* in, float* out, int cols, float scale) {  // in == out allowed
  __shared__ float red[8];
  const float* r = in + (long long)blockIdx.x * cols;
  float* o = out + (long long)blockIdx.x * cols;
  float m = -INFINITY;
  for (int i = threadIdx.x; i < cols; i += 256) m = fmaxf(m, r[i] * scale);
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  m = red[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
  __syncthreads();
  float s = 0.f;
  for (int i = threadIdx.x; i < cols; i += 256) s += expf(r[i] * scale - m);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += red[i];
  for (int i = threadIdx.x; i < cols; i += 256) o[i] = expf(r[i] * scale - m) / s;
}

}  // namespace

int f32_to_bf16(const float* in, __nv_bfloat16* out, long long n, cudaStream_t st) {
  long long n4 = n / 4;
  if (n4 > 0) {
    int blocks = (int)std::min<long long>((n4 + 255) / 256, kNumSMs * 16);
    f32_to_bf16_kernel<<<blocks, 256, 0, st>>>(in, out, n4);
    PIO_LAUNCHED();
  }
  if (n4 * 4 < n) {
    f32_to_bf16_tail<<<1, 32, 0, st>>>(in, out, n4 * 4, n);
    PIO_LAUNCHED();
  }
  return PIO_OK;
}
int transpose_f32(const float* in, float* out, int rows, int cols, cudaStream_t st) {
  dim3 grid(cdiv(cols, 32), cdiv(rows, 32));
  transpose_kernel<float><<<grid, dim3(32, 8), 0, st>>>(in, out, rows, cols);
  PIO_LAUNCHED();
  return PIO_OK;
}
int transpose_to_bf16(const float* in, __nv_bfloat16* out, int rows, int cols, cudaStream_t st) {
  dim3 grid(cdiv(cols, 32), cdiv(rows, 32));
  transpose_kernel<__nv_bfloat16><<<grid, dim3(32, 8), 0, st>>>(in, out, rows, cols);
  PIO_LAUNCHED();
  return PIO_OK;
}

int layernorm(const float* x, int ldx, const float* w, const float* b, void* out, int out_dt, int ldo, int rows, int dim,
              float eps, cudaStream_t st) {
  PIO_CHECK(dim % 128 == 0 && dim <= 1024, "layernorm: dim %d must be a multiple of 128, <= 1024", dim);
  PIO_CHECK(ldx % 4 == 0 && ldo % 4 == 0, "layernorm: strides must be multiples of 4");
  if (rows == 0) return PIO_OK;
  const int blocks = cdiv((long long)rows * 32, 256);
#define LN_CASE(V)                                                                                   \
  case V:                                                                                            \
    launch_pdl_k(PDL_KIND_LN, layernorm_kernel<V>, dim3(blocks), dim3(256), 0, st, x, ldx, w, b, out, out_dt, ldo, rows, eps); \
    break;
  switch (dim / 128) {
    LN_CASE(1) LN_CASE(2) LN_CASE(3) LN_CASE(4) LN_CASE(5) LN_CASE(6) LN_CASE(7) LN_CASE(8)
  }
#undef LN_CASE
  PIO_LAUNCHED();
  return PIO_OK;
}

int im2col14(const float* imgs, void* cols, int cols_dt, int B, int S, int g, int Kp, cudaStream_t st) {
  static const bool strip_off = [] { const char* e = getenv("PIO_IM2COL_STRIP"); return e && e[0] == '0'; }();
  const size_t strip_bytes = (size_t)3 * 14 * S * (cols_dt == PIO_DT_F32 ? 4 : 2);
  if (!strip_off && strip_bytes <= 200 * 1024 && Kp % 2 == 0 && B <= 65535 && g * 14 <= S) {
    if (cols_dt == PIO_DT_F32) {
      static SmemAttrOnce once;
      PIO_CUDA(once.ensure(im2col14_strip_kernel<float>, 200 * 1024));
      im2col14_strip_kernel<float><<<dim3(g, B), 256, strip_bytes, st>>>(imgs, (float*)cols, S, g, Kp);
    } else {
      static SmemAttrOnce once;
      PIO_CUDA(once.ensure(im2col14_strip_kernel<__nv_bfloat16>, 200 * 1024));
      im2col14_strip_kernel<__nv_bfloat16><<<dim3(g, B), 256, strip_bytes, st>>>(imgs, (__nv_bfloat16*)cols, S, g, Kp);
    }
    PIO_LAUNCHED();
    return PIO_OK;
  }
  const long long total = (long long)B * g * g * Kp;
  const int blocks = (int)std::min<long long>((total + 255) / 256, kNumSMs * 32);
  if (cols_dt == PIO_DT_F32)
    im2col14_kernel<float><<<blocks, 256, 0, st>>>(imgs, (float*)cols, B, S, g, Kp);
  else
    im2col14_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(imgs, (__nv_bfloat16*)cols, B, S, g, Kp);
  PIO_LAUNCHED();
  return PIO_OK;
}

int init_global_tokens(float* x, const float* cls, const float* reg, const float* pos, int B, int N, int D, cudaStream_t st) {
  init_global_tokens_kernel<<<cdiv((long long)B * 5 * D, 256), 256, 0, st>>>(x, cls, reg, pos, B, N, D);
  PIO_LAUNCHED();
  return PIO_OK;
}

int l2norm_rows(float* x, int rows, int dim, cudaStream_t st) {
  if (rows == 0) return PIO_OK;
  l2norm_kernel<<<cdiv((long long)rows * 32, 256), 256, 0, st>>>(x, rows, dim);
  PIO_LAUNCHED();
  return PIO_OK;
}

int softmax_rows(const float* in, float* out, int rows, int cols, float scale, cudaStream_t st) {
  if (rows == 0) return PIO_OK;
  softmax_rows_kernel<<<rows, 256, 0, st>>>(in, out, cols, scale);
  PIO_LAUNCHED();
  return PIO_OK;
}

}  // namespace pio

extern "C" {
// out[t, :] = src[idx[t], :]   (rows of D floats, D % 4 == 0)
int pio_gather_rows(const float* src, long long src_ld, const int* idx, int T, int D, float* out, void* stream) {
  using namespace pio;
  if (T == 0) return PIO_OK;
  PIO_CHECK(src && idx && out && D % 4 == 0 && src_ld % 4 == 0, "gather_rows: bad arguments");
  gather_rows_kernel<<<cdiv((long long)T * 32, 256), 256, 0, as_stream(stream)>>>(src, src_ld, idx, T, D, out);
  PIO_LAUNCHED();
  return PIO_OK;
}
// out[s, :] = mean of rows [seg_start[s], seg_start[s] + seg_len[s]) of x   (an empty segment gives NaN, like tensor.mean())
int pio_segment_mean(const float* x, const int* seg_start, const int* seg_len, int nseg, int D, float* out, void* stream) {
  using namespace pio;
  if (nseg == 0) return PIO_OK;
  PIO_CHECK(x && seg_start && seg_len && out && D % 4 == 0, "segment_mean: bad arguments");
  segment_mean_kernel<<<dim3(nseg, cdiv(D / 4, 64)), 64, 0, as_stream(stream)>>>(x, seg_start, seg_len, D, out);
  PIO_LAUNCHED();
  return PIO_OK;
}

const char* pio_last_error(void) { return pio::g_err; }
int pio_version(void) { return 100; }
long long pio_launch_count(void) { return pio::g_launches.load(); }
void pio_reset_launch_count(void) { pio::g_launches.store(0); }
void pio_release_scratch(void) { pio::release_split_scratch(); }

int pio_layernorm(const float* x, int ldx, const float* w, const float* b, void* out, int out_dt, int ldo, int rows,
                  int dim, float eps, void* stream) {
  return pio::layernorm(x, ldx, w, b, out, out_dt, ldo, rows, dim, eps, pio::as_stream(stream));
}
int pio_l2_normalize(float* x, int rows, int dim, void* stream) {
  return pio::l2norm_rows(x, rows, dim, pio::as_stream(stream));
}
int pio_linear(const PioLinear* p, int mode, void* stream) {
  if (!p) return pio::fail(PIO_EINVAL, "pio_linear: null descriptor");
  if (mode == PIO_FP32) return pio::linear_simt(*p, pio::as_stream(stream));
  if (mode == PIO_BF16) return pio::linear_tc(*p, pio::as_stream(stream));
  return pio::fail(PIO_EINVAL, "pio_linear: unknown mode %d", mode);
}
}

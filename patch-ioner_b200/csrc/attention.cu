// attention.cu -- (1) ViT multi-head self-attention in fp32 arithmetic (flash-style, no N x N
// materialisation), (2) the single-query KV-cache attention of the greedy decoder.
#include "common.cuh"

namespace pio {
namespace {

template <typename T> __device__ __forceinline__ float4 load4(const T* p);
template <> __device__ __forceinline__ float4 load4<float>(const float* p) { return *reinterpret_cast<const float4*>(p); }
template <> __device__ __forceinline__ float4 load4<__nv_bfloat16>(const __nv_bfloat16* p) {
  uint2 u = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&u.x), b = *reinterpret_cast<__nv_bfloat162*>(&u.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
template <typename T> __device__ __forceinline__ void store4(T* p, float4 v);
template <> __device__ __forceinline__ void store4<float>(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
template <> __device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* p, float4 v) {
  __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
  uint2 pk;
  pk.x = *reinterpret_cast<uint32_t*>(&lo);
  pk.y = *reinterpret_cast<uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(p) = pk;
}

// ---------------------------------------------------------------------------------------------
// ViT attention, head_dim 64.  qkv [B,N,3*H*64] ([q|k|v], head-major inside each), out [B,N,H*64].
// CTA = 64 queries of one (b,h); loops over 64-key tiles with an online softmax.
// Thread (ty,tx) owns rows ty+16i and score columns tx+16j (interleaved: conflict-free LDS.128),
// output columns tx*4..tx*4+3.
constexpr int BR = 64, BC = 64, HD = 64, LDS = HD + 4;

template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256) vit_attention_kernel(const TIn* __restrict__ qkv, TOut* __restrict__ out, int N, int H,
                                                            float scale) {
  extern __shared__ __align__(16) float smem[];
  float* Qs = smem;               // [BR][LDS]
  float* Ks = Qs + BR * LDS;      // [BC][LDS]
  float* Vs = Ks + BC * LDS;      // [BC][LDS]
  float* Ps = Vs + BC * LDS;      // [BR][LDS]
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * BR;
  const int C3 = 3 * H * HD;
  const TIn* base = qkv + (long long)b * N * C3;
  const TIn* qp = base + h * HD;
  const TIn* kp = base + H * HD + h * HD;
  const TIn* vp = base + 2 * H * HD + h * HD;

  // Q tile (rows beyond N are zero)
  for (int idx = tid; idx < BR * (HD / 4); idx += 256) {
    int r = idx >> 4, d4 = (idx & 15) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q0 + r < N) v = load4<TIn>(qp + (long long)(q0 + r) * C3 + d4);
    *reinterpret_cast<float4*>(&Qs[r * LDS + d4]) = v;
  }
  float m[4], l[4], o[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m[i] = -INFINITY;
    l[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
  }

  for (int k0 = 0; k0 < N; k0 += BC) {
    __syncthreads();  // previous tile's Ks/Vs/Ps fully consumed (and Qs visible on first pass)
    for (int idx = tid; idx < BC * (HD / 4); idx += 256) {
      int r = idx >> 4, d4 = (idx & 15) * 4;
      float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
      if (k0 + r < N) {
        kv = load4<TIn>(kp + (long long)(k0 + r) * C3 + d4);
        vv = load4<TIn>(vp + (long long)(k0 + r) * C3 + d4);
      }
      *reinterpret_cast<float4*>(&Ks[r * LDS + d4]) = kv;
      *reinterpret_cast<float4*>(&Vs[r * LDS + d4]) = vv;
    }
    __syncthreads();
    float s[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll 4
    for (int d = 0; d < HD; d += 4) {
      float4 q[4], k[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) q[i] = *reinterpret_cast<const float4*>(&Qs[(ty + 16 * i) * LDS + d]);
#pragma unroll
      for (int j = 0; j < 4; ++j) k[j] = *reinterpret_cast<const float4*>(&Ks[(tx + 16 * j) * LDS + d]);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          s[i][j] += q[i].x * k[j].x + q[i].y * k[j].y + q[i].z * k[j].z + q[i].w * k[j].w;
    }
    // online softmax, rows shared by the 16 lanes with equal ty
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        s[i][j] = (k0 + tx + 16 * j < N) ? s[i][j] * scale : -INFINITY;
        mx = fmaxf(mx, s[i][j]);
      }
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
      const float mnew = fmaxf(m[i], mx);
      const float corr = __expf(m[i] - mnew);  // m = -inf on first tile -> 0
      float ps = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float p = __expf(s[i][j] - mnew);
        ps += p;
        Ps[(ty + 16 * i) * LDS + tx + 16 * j] = p;
      }
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, off);
      l[i] = l[i] * corr + ps;
      m[i] = mnew;
#pragma unroll
      for (int j = 0; j < 4; ++j) o[i][j] *= corr;
    }
    __syncthreads();
#pragma unroll 4
    for (int jj = 0; jj < BC; jj += 4) {
      float4 p[4], v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) p[i] = *reinterpret_cast<const float4*>(&Ps[(ty + 16 * i) * LDS + jj]);
#pragma unroll
      for (int e = 0; e < 4; ++e) v[e] = *reinterpret_cast<const float4*>(&Vs[(jj + e) * LDS + tx * 4]);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        o[i][0] += p[i].x * v[0].x + p[i].y * v[1].x + p[i].z * v[2].x + p[i].w * v[3].x;
        o[i][1] += p[i].x * v[0].y + p[i].y * v[1].y + p[i].z * v[2].y + p[i].w * v[3].y;
        o[i][2] += p[i].x * v[0].z + p[i].y * v[1].z + p[i].z * v[2].z + p[i].w * v[3].z;
        o[i][3] += p[i].x * v[0].w + p[i].y * v[1].w + p[i].z * v[2].w + p[i].w * v[3].w;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int r = q0 + ty + 16 * i;
    if (r < N) {
      float inv = 1.0f / l[i];
      store4<TOut>(out + ((long long)b * N + r) * (H * HD) + h * HD + tx * 4,
                   make_float4(o[i][0] * inv, o[i][1] * inv, o[i][2] * inv, o[i][3] * inv));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Decoder attention for ONE new position t: append this step's k,v to the cache, then
// out[r, h*HDIM..] = softmax(q . K[0..t] / sqrt(HDIM)) V[0..t].   One warp per (region, head).
// qkv [R, 3*H*HDIM] (this step), cache kc/vc [R][H][T][HDIM].
template <typename TIn, typename TOut, typename TC, int HDIM>
__global__ void __launch_bounds__(128) decode_attention_kernel(const TIn* __restrict__ qkv, TC* __restrict__ kc,
                                                               TC* __restrict__ vc, TOut* __restrict__ out, int R, int H,
                                                               int T, int t, float scale) {
  constexpr int PER = HDIM / 32;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= R * H) return;
  const int r = warp / H, h = warp % H;
  const TIn* row = qkv + (long long)r * 3 * H * HDIM;
  TC* kbase = kc + ((long long)(r * H + h) * T) * HDIM;
  TC* vbase = vc + ((long long)(r * H + h) * T) * HDIM;
  float q[PER];
#pragma unroll
  for (int e = 0; e < PER; ++e) {
    const int d = lane + 32 * e;
    q[e] = (float)row[h * HDIM + d];
    const float kv = (float)row[H * HDIM + h * HDIM + d], vv = (float)row[2 * H * HDIM + h * HDIM + d];
    kbase[(long long)t * HDIM + d] = (TC)kv;
    vbase[(long long)t * HDIM + d] = (TC)vv;
  }
  __syncwarp();
  float my = -INFINITY;  // lane j keeps score j
  for (int j = 0; j <= t; ++j) {
    float s = 0.f;
#pragma unroll
    for (int e = 0; e < PER; ++e) s += q[e] * (float)kbase[(long long)j * HDIM + lane + 32 * e];
    s = warp_sum(s) * scale;
    if (lane == j) my = s;
  }
  const float mx = warp_max(my);
  const float p = (lane <= t) ? __expf(my - mx) : 0.f;
  const float inv = 1.0f / warp_sum(p);
  float acc[PER];
#pragma unroll
  for (int e = 0; e < PER; ++e) acc[e] = 0.f;
  for (int j = 0; j <= t; ++j) {
    const float pj = __shfl_sync(0xffffffffu, p, j) * inv;
#pragma unroll
    for (int e = 0; e < PER; ++e) acc[e] += pj * (float)vbase[(long long)j * HDIM + lane + 32 * e];
  }
#pragma unroll
  for (int e = 0; e < PER; ++e) out[(long long)r * H * HDIM + h * HDIM + lane + 32 * e] = (TOut)acc[e];
}


// bf16 variant with 16-byte loads: head_dim 192 = 24 lanes x 8 elements (lanes 24..31 only help in the reductions).
// KV-cache rows are 384 contiguous bytes, so a warp reads each key / value row with one coalesced request.
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
// streaming 16-byte load: the KV cache is read once per step and never reused inside a launch
__device__ __forceinline__ uint4 ld_stream16(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
// scheduling fence: keeps ptxas from sinking the loads of a batch next to their uses (which leaves one load in flight)
__device__ __forceinline__ void hold8(uint4 (&v)[8]) {
#pragma unroll
  for (int u = 0; u < 8; ++u) asm volatile("" : "+r"(v[u].x), "+r"(v[u].y), "+r"(v[u].z), "+r"(v[u].w));
}
// Keys / values are walked in batches of 8 rows: the 8 loads are issued back to back (8 x 384 B in flight per warp, enough
// to cover the HBM latency at 32 resident warps per SM) and the 8 warp reductions interleave.
__global__ void __launch_bounds__(128) decode_attention_bf16_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ kc,
                                                                    __nv_bfloat16* __restrict__ vc, __nv_bfloat16* __restrict__ out,
                                                                    int R, int H, int T, int t, float scale) {
  constexpr int HDIM = 192, KB = 8;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  pdl_wait();
  pdl_launch_dependents();
  if (warp >= R * H) return;
  const int r = warp / H, h = warp % H;
  const bool act = lane < HDIM / 8;
  const __nv_bfloat16* row = qkv + (long long)r * 3 * H * HDIM + h * HDIM + lane * 8;
  __nv_bfloat16* kbase = kc + ((long long)(r * H + h) * T) * HDIM + lane * 8;
  __nv_bfloat16* vbase = vc + ((long long)(r * H + h) * T) * HDIM + lane * 8;
  float q[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const uint4 zero = make_uint4(0, 0, 0, 0);
  uint4 knew = zero, vnew = zero;
  if (act) {
    const uint4 qu = *reinterpret_cast<const uint4*>(row);
    knew = *reinterpret_cast<const uint4*>(row + H * HDIM);
    vnew = *reinterpret_cast<const uint4*>(row + 2 * H * HDIM);
    unpack8(qu, q);
    *reinterpret_cast<uint4*>(kbase + (long long)t * HDIM) = knew;   // append this step's key / value
    *reinterpret_cast<uint4*>(vbase + (long long)t * HDIM) = vnew;
  }
  const int tl = t > 0 ? t - 1 : 0;  // last row that older steps wrote (row t itself comes from registers)
  float my = -INFINITY;  // lane j keeps score j
  for (int j0 = 0; j0 <= t; j0 += KB) {
    uint4 ku[KB];
#pragma unroll
    for (int u = 0; u < KB; ++u) {
      const int j = j0 + u;
      ku[u] = ld_stream16(kbase + (long long)min(j, tl) * HDIM);  // unconditional (always inside this head's T rows)
    }
    hold8(ku);  // all 8 loads are issued before the first use
#pragma unroll
    for (int u = 0; u < KB; ++u) {
      const int j = j0 + u;
      ku[u] = (act && j < t) ? ku[u] : (j == t ? knew : zero);
    }
    float sc[KB];
#pragma unroll
    for (int u = 0; u < KB; ++u) {
      float kf[8];
      unpack8(ku[u], kf);
      float a = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) a = fmaf(q[e], kf[e], a);
      sc[u] = a;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int u = 0; u < KB; ++u) sc[u] += __shfl_xor_sync(0xffffffffu, sc[u], o);
    }
#pragma unroll
    for (int u = 0; u < KB; ++u)
      if (lane == j0 + u) my = sc[u] * scale;
  }
  my = (lane <= t) ? my : -INFINITY;
  const float mx = warp_max(my);
  const float pe = (lane <= t) ? __expf(my - mx) : 0.f;
  const float p = pe * (1.0f / warp_sum(pe));
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int j0 = 0; j0 <= t; j0 += KB) {
    uint4 vu[KB];
#pragma unroll
    for (int u = 0; u < KB; ++u) {
      const int j = j0 + u;
      vu[u] = ld_stream16(vbase + (long long)min(j, tl) * HDIM);
    }
    hold8(vu);
#pragma unroll
    for (int u = 0; u < KB; ++u) {
      const int j = j0 + u;
      vu[u] = (act && j < t) ? vu[u] : (j == t ? vnew : zero);
    }
#pragma unroll
    for (int u = 0; u < KB; ++u) {
      const float pj = __shfl_sync(0xffffffffu, p, (j0 + u) & 31);  // lanes beyond t hold p = 0
      float vf[8];
      unpack8(vu[u], vf);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = fmaf(pj, vf[e], acc[e]);
    }
  }
  if (act) {
    __nv_bfloat162 a = __floats2bfloat162_rn(acc[0], acc[1]), b = __floats2bfloat162_rn(acc[2], acc[3]);
    __nv_bfloat162 c = __floats2bfloat162_rn(acc[4], acc[5]), d = __floats2bfloat162_rn(acc[6], acc[7]);
    uint4 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&a); pk.y = *reinterpret_cast<uint32_t*>(&b);
    pk.z = *reinterpret_cast<uint32_t*>(&c); pk.w = *reinterpret_cast<uint32_t*>(&d);
    *reinterpret_cast<uint4*>(out + (long long)r * H * HDIM + h * HDIM + lane * 8) = pk;
  }
}

}  // namespace

int vit_attention(const void* qkv, void* out, int dt, int B, int N, int H, cudaStream_t st) {
  const size_t smem = (size_t)(3 * BC + BR) * LDS * sizeof(float);
  dim3 grid(cdiv(N, BR), H, B);
  const float scale = 0.125f;  // 64^-0.5
  if (dt == PIO_DT_F32) {
    static SmemAttrOnce once;
    PIO_CUDA(once.ensure(vit_attention_kernel<float, float>, (int)smem));
    vit_attention_kernel<float, float><<<grid, 256, smem, st>>>((const float*)qkv, (float*)out, N, H, scale);
  } else {
    static SmemAttrOnce once;
    PIO_CUDA(once.ensure(vit_attention_kernel<__nv_bfloat16, __nv_bfloat16>, (int)smem));
    vit_attention_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, smem, st>>>((const __nv_bfloat16*)qkv, (__nv_bfloat16*)out, N, H, scale);
  }
  PIO_LAUNCHED();
  return PIO_OK;
}

namespace {
// ---------------------------------------------------------------------------------------------
// Decoder attention for longer caches (GPT-2 small: head_dim 64, up to 128 positions).  One warp per (region, head).
// Scores: lane l owns positions l, l+32, l+64, l+96 and walks its key rows whole (a 64-element row is one or two
// 128-byte lines); output: lane l owns dims 2l, 2l+1 and the warp reads each value row with one coalesced request.
template <typename T> __device__ __forceinline__ void load8f(const T* p, float (&f)[8]);
template <> __device__ __forceinline__ void load8f<float>(const float* p, float (&f)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
template <> __device__ __forceinline__ void load8f<__nv_bfloat16>(const __nv_bfloat16* p, float (&f)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  unpack8(u, f);
}
template <typename T> __device__ __forceinline__ float2 load2f(const T* p);
template <> __device__ __forceinline__ float2 load2f<float>(const float* p) { return *reinterpret_cast<const float2*>(p); }
template <> __device__ __forceinline__ float2 load2f<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p));
}
template <typename T> __device__ __forceinline__ void store2f(T* p, float a, float b);
template <> __device__ __forceinline__ void store2f<float>(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
template <> __device__ __forceinline__ void store2f<__nv_bfloat16>(__nv_bfloat16* p, float a, float b) {
  *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b);
}

template <typename T, bool ANC>
__global__ void __launch_bounds__(128) decode_attention_long_kernel(const T* __restrict__ qkv, T* __restrict__ kc, T* __restrict__ vc,
                                                                    T* __restrict__ out, int R, int H, int Tmax, int t, float scale,
                                                                    const int* __restrict__ anc) {
  // anc (beam search, or NULL): anc[r * Tmax + j] = the cache row that holds position j of row r's history -- beams are
  // re-ordered by copying this small table, never the cache; a row appends position t to ITS OWN cache row.
  constexpr int HDIM = 64, MAXC = 4;
  __shared__ float qs[4][HDIM];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warp = blockIdx.x * 4 + wib;
  pdl_wait();               // qkv comes from the GEMM just before
  pdl_launch_dependents();  // the projection GEMM may start its prologue (it waits before touching our output)
  if (warp >= R * H) return;
  const int r = warp / H, h = warp % H;
  const T* row = qkv + (long long)r * 3 * H * HDIM + h * HDIM;
  T* kbase = kc + ((long long)(r * H + h) * Tmax) * HDIM;
  T* vbase = vc + ((long long)(r * H + h) * Tmax) * HDIM;
  {
    const float2 q2 = load2f<T>(row + lane * 2), k2 = load2f<T>(row + H * HDIM + lane * 2), v2 = load2f<T>(row + 2 * H * HDIM + lane * 2);
    qs[wib][lane * 2] = q2.x; qs[wib][lane * 2 + 1] = q2.y;
    store2f<T>(kbase + (long long)t * HDIM + lane * 2, k2.x, k2.y);   // append this step's key / value
    store2f<T>(vbase + (long long)t * HDIM + lane * 2, v2.x, v2.y);
  }
  __syncwarp();  // the appended row and q are visible to the whole warp
  float sc[MAXC];
  int arow[MAXC];
  float mx = -INFINITY;
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    const int j = lane + 32 * c;
    sc[c] = -INFINITY;
    arow[c] = r;
    if (j <= t) {
      if (ANC && j < t) arow[c] = __ldg(anc + (long long)r * Tmax + j);
      const T* kr = ANC ? kc + (((long long)arow[c] * H + h) * Tmax + j) * HDIM : kbase + (long long)j * HDIM;
      float a = 0.f;
#pragma unroll
      for (int d = 0; d < HDIM; d += 8) {
        float kf[8];
        load8f<T>(kr + d, kf);
#pragma unroll
        for (int e = 0; e < 8; ++e) a = fmaf(qs[wib][d + e], kf[e], a);
      }
      sc[c] = a * scale;
    }
    mx = fmaxf(mx, sc[c]);
  }
  mx = warp_max(mx);
  float sum = 0.f;
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    sc[c] = (lane + 32 * c <= t) ? __expf(sc[c] - mx) : 0.f;
    sum += sc[c];
  }
  const float inv = 1.0f / warp_sum(sum);
  float a0 = 0.f, a1 = 0.f;
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    if (32 * c > t) break;  // warp-uniform
    const int n = min(32, t + 1 - 32 * c);
#pragma unroll 4
    for (int jj = 0; jj < n; ++jj) {
      const float pj = __shfl_sync(0xffffffffu, sc[c], jj);
      const T* vr = vbase + (long long)(32 * c + jj) * HDIM;
      if (ANC) {
        const int rj = __shfl_sync(0xffffffffu, arow[c], jj);
        vr = vc + (((long long)rj * H + h) * Tmax + 32 * c + jj) * HDIM;
      }
      const float2 v2 = load2f<T>(vr + lane * 2);
      a0 = fmaf(pj, v2.x, a0);
      a1 = fmaf(pj, v2.y, a1);
    }
  }
  store2f<T>(out + (long long)r * H * HDIM + h * HDIM + lane * 2, a0 * inv, a1 * inv);
}

// ---------------------------------------------------------------------------------------------
// Bidirectional attention over a short sequence (ViECap mapping network: 20 tokens, 8 heads x 96; ClipCap.py:51-68).
// q [R*n, ldq] (head h at column h*hd), kv [R*n, ldkv] (keys at h*hd, values at H*hd + h*hd), out [R*n, ldo].
// One CTA per (region, head): q, k, v and the n x n probabilities live in shared memory as fp32.
template <typename T>
__global__ void __launch_bounds__(128) small_attention_kernel(const T* __restrict__ q, long long ldq, const T* __restrict__ kv,
                                                              long long ldkv, T* __restrict__ out, long long ldo, int n, int H, int hd,
                                                              float scale, int causal, T* __restrict__ kc, T* __restrict__ vc, int Tmax) {
  extern __shared__ __align__(16) float sm[];
  const int ld = hd + 1;
  float* Q = sm;
  float* K = Q + n * ld;
  float* V = K + n * ld;
  float* P = V + n * ld;  // [n][n+1]
  const int r = blockIdx.x / H, h = blockIdx.x % H, tid = threadIdx.x;
  for (int idx = tid; idx < n * hd; idx += 128) {
    const int i = idx / hd, d = idx % hd;
    const long long row = (long long)r * n + i;
    const T kk = kv[row * ldkv + h * hd + d], vv = kv[row * ldkv + (long long)H * hd + h * hd + d];
    Q[i * ld + d] = (float)q[row * ldq + h * hd + d];
    K[i * ld + d] = (float)kk;
    V[i * ld + d] = (float)vv;
    if (kc) {  // prompt prefill of a decoder: positions 0..n-1 of this head's KV cache [R][H][Tmax][hd]
      const long long c = (((long long)r * H + h) * Tmax + i) * hd + d;
      kc[c] = kk;
      vc[c] = vv;
    }
  }
  __syncthreads();
  for (int idx = tid; idx < n * n; idx += 128) {
    const int i = idx / n, j = idx % n;
    float a = 0.f;
    for (int d = 0; d < hd; ++d) a = fmaf(Q[i * ld + d], K[j * ld + d], a);
    P[i * (n + 1) + j] = (causal && j > i) ? -INFINITY : a * scale;
  }
  __syncthreads();
  if (tid < n) {
    float mx = -INFINITY, s = 0.f;
    for (int j = 0; j < n; ++j) mx = fmaxf(mx, P[tid * (n + 1) + j]);
    for (int j = 0; j < n; ++j) {
      const float e = expf(P[tid * (n + 1) + j] - mx);
      P[tid * (n + 1) + j] = e;
      s += e;
    }
    const float inv = 1.0f / s;
    for (int j = 0; j < n; ++j) P[tid * (n + 1) + j] *= inv;
  }
  __syncthreads();
  for (int idx = tid; idx < n * hd; idx += 128) {
    const int i = idx / hd, d = idx % hd;
    float a = 0.f;
    for (int j = 0; j < n; ++j) a = fmaf(P[i * (n + 1) + j], V[j * ld + d], a);
    out[((long long)r * n + i) * ldo + h * hd + d] = (T)a;
  }
}

}  // namespace

size_t small_attention_smem(int n, int hd) { return ((size_t)3 * n * (hd + 1) + (size_t)n * (n + 1)) * sizeof(float); }

int small_attention(const void* q, long long ldq, const void* kv, long long ldkv, void* out, long long ldo, int dt, int R, int n, int H,
                    int hd, bool causal, void* kc, void* vc, int Tmax, cudaStream_t st) {
  PIO_CHECK(n >= 1 && n <= 128 && hd >= 1 && hd <= 128, "small attention: n %d / head_dim %d outside the built range (<=128, <=128)", n, hd);
  const size_t smem = small_attention_smem(n, hd);
  PIO_CHECK(smem <= 200 * 1024, "small attention: %zu bytes of shared memory exceed 200 KB", smem);
  if (smem > 48 * 1024) {  // opt in to the large carve-out once per kernel
    static SmemAttrOnce once_f, once_b;
    if (dt == PIO_DT_F32) PIO_CUDA(once_f.ensure(small_attention_kernel<float>, 200 * 1024));
    else PIO_CUDA(once_b.ensure(small_attention_kernel<__nv_bfloat16>, 200 * 1024));
  }
  const float scale = 1.0f / sqrtf((float)hd);
  if (dt == PIO_DT_F32)
    small_attention_kernel<float><<<R * H, 128, smem, st>>>((const float*)q, ldq, (const float*)kv, ldkv, (float*)out, ldo, n, H, hd, scale,
                                                             causal ? 1 : 0, (float*)kc, (float*)vc, Tmax);
  else
    small_attention_kernel<__nv_bfloat16><<<R * H, 128, smem, st>>>((const __nv_bfloat16*)q, ldq, (const __nv_bfloat16*)kv, ldkv,
                                                                    (__nv_bfloat16*)out, ldo, n, H, hd, scale, causal ? 1 : 0,
                                                                    (__nv_bfloat16*)kc, (__nv_bfloat16*)vc, Tmax);
  PIO_LAUNCHED();
  return PIO_OK;
}

int decode_attention(const void* qkv, void* kc, void* vc, void* out, int dt, int R, int H, int T, int t, cudaStream_t st, const int* anc) {
  PIO_CHECK(anc == nullptr || H == 12, "decode attention: the cache-row table is only built for the 12 x 64 kernel");
  if (H == 12) {  // GPT-2 small: 12 heads x 64
    PIO_CHECK(t < T && T <= 128, "decode attention: position %d outside cache of %d (max 128)", t, T);
    const int blocks = cdiv((long long)R * H, 4);
#define PIO_DEC_ATT_LONG(TT, A)                                                                                                        \
  launch_pdl(decode_attention_long_kernel<TT, A>, dim3(blocks), dim3(128), 0, st, (const TT*)qkv, (TT*)kc, (TT*)vc, (TT*)out, R, H, T, t, \
             0.125f, anc)
    if (dt == PIO_DT_F32) { if (anc) PIO_DEC_ATT_LONG(float, true); else PIO_DEC_ATT_LONG(float, false); }
    else { if (anc) PIO_DEC_ATT_LONG(__nv_bfloat16, true); else PIO_DEC_ATT_LONG(__nv_bfloat16, false); }
#undef PIO_DEC_ATT_LONG
    PIO_LAUNCHED();
    return PIO_OK;
  }
  PIO_CHECK(H == 4, "decode attention: %d heads (built: 4 x 192 and 12 x 64)", H);

  PIO_CHECK(t < T && T <= 32, "decode attention: position %d outside cache of %d (max 32)", t, T);
  const float scale = rsqrtf(192.0f);
  const int blocks = cdiv((long long)R * H * 32, 128);
  if (dt == PIO_DT_F32)
    decode_attention_kernel<float, float, float, 192><<<blocks, 128, 0, st>>>((const float*)qkv, (float*)kc, (float*)vc, (float*)out, R, H, T, t, scale);
  else
    launch_pdl(decode_attention_bf16_kernel, dim3(blocks), dim3(128), 0, st, (const __nv_bfloat16*)qkv, (__nv_bfloat16*)kc,
               (__nv_bfloat16*)vc, (__nv_bfloat16*)out, R, H, T, t, scale);
  PIO_LAUNCHED();
  return PIO_OK;
}

}  // namespace pio

// pooling.cu -- region aggregation of patch tokens.
//
// HBM-bound byte work: each image's patch tokens are read from HBM exactly once per call.  A CTA
// stages a [P x 32-channel] slab of one image in shared memory with cp.async (128-byte row
// segments, the whole slab in flight at once -- narrow TMA boxes measured 4x slower), then its 32 warps walk the
// regions of that image: lane = channel, weights are computed
// 32 patches at a time (one per lane) and broadcast with warp shuffles, the slab is read
// conflict-free (32 consecutive floats per patch).  Index arithmetic follows the reference exactly
// (torch float floor-division, inclusive end, Python slice clamping) and is exported as int32
// bounds so that tests can compare it bit-for-bit.
#include "tc_ptx.cuh"
#include <stdlib.h>

namespace pio {
namespace {

// torch `a //= b` for float32 (c10::div_floor_floating), see bbox_utils.py:19
__device__ __forceinline__ float torch_floor_div(float a, float b) {
  if (b == 0.f) return a / b;
  float mod = fmodf(a, b);
  float div = (a - mod) / b;
  if ((mod != 0.f) && ((b < 0.f) != (mod < 0.f))) div -= 1.0f;
  float floordiv;
  if (div != 0.f) {
    floordiv = floorf(div);
    if (div - floordiv > 0.5f) floordiv += 1.0f;
  } else {
    floordiv = copysignf(0.f, a / b);
  }
  return floordiv;
}
__device__ __forceinline__ int py_floor_div(int a, int b) {
  int q = a / b, r = a % b;
  return (r != 0 && ((r < 0) != (b < 0))) ? q - 1 : q;
}
// Python a[start:stop] on a dimension of `size` -> [lo, hi)
__device__ __forceinline__ void py_slice(int start, int stop, int size, int& lo, int& hi) {
  if (start < 0) { start += size; if (start < 0) start = 0; } else if (start > size) start = size;
  if (stop < 0) { stop += size; if (stop < 0) stop = 0; } else if (stop > size) stop = size;
  lo = start;
  hi = stop < start ? start : stop;
}

// boxes [n,4] xywh px -> bounds [n,4] (y_lo,y_hi,x_lo,x_hi) and skip[n] (patch-unit sum < 0: dummy box)
__global__ void box_bounds_kernel(const void* __restrict__ boxes, int boxes_dt, int n, int patch, int grid,
                                  int* __restrict__ bounds, int* __restrict__ skip) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int pu[4];
  if (boxes_dt == PIO_DT_F32) {
    const float* bx = reinterpret_cast<const float*>(boxes) + 4 * i;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float q = torch_floor_div(bx[e], (float)patch);
      pu[e] = isfinite(q) ? (int)q : 0;
    }
  } else {
    const int* bx = reinterpret_cast<const int*>(boxes) + 4 * i;
#pragma unroll
    for (int e = 0; e < 4; ++e) pu[e] = py_floor_div(bx[e], patch);
  }
  int x1 = pu[0], y1 = pu[1], w = pu[2], h = pu[3];
  int ylo, yhi, xlo, xhi;
  py_slice(y1, y1 + h + 1, grid, ylo, yhi);
  py_slice(x1, x1 + w + 1, grid, xlo, xhi);
  bounds[4 * i + 0] = ylo; bounds[4 * i + 1] = yhi; bounds[4 * i + 2] = xlo; bounds[4 * i + 3] = xhi;
  skip[i] = (x1 + y1 + w + h) < 0 ? 1 : 0;
}

// torch.linspace(-1, 1, n)[i] in fp32
__device__ __forceinline__ float linspace_m1_1(int i, int n) {
  if (n == 1) return -1.0f;
  const float step = 2.0f / (float)(n - 1);
  return (i < n / 2) ? (-1.0f + step * (float)i) : (1.0f - step * (float)(n - 1 - i));
}
__device__ __forceinline__ float gauss_w(int iy, int hs, int ix, int ws, float variance) {
  const float y = linspace_m1_1(iy, hs), x = linspace_m1_1(ix, ws);
  return expf(-(x * x + y * y) / variance);
}

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[w] = v;
  __syncthreads();
  float s = 0.f;
  for (int i = 0; i < nw; ++i) s += red[i];
  return s;
}

// One CTA per image, boxes in order (the attention-map rescale is sequential and order dependent,
// bbox_utils.py:46-48).  Writes per-box dense weights (per_box, [B,R,P], zero outside the box) and /
// or the normalised box-set map (set_out, [B,P], bbox_utils.py:49,82,89,100).
__global__ void __launch_bounds__(256) box_weights_kernel(const int* __restrict__ bounds, const int* __restrict__ skip, int R,
                                                          int grid, int mode, float variance,
                                                          const float* __restrict__ attn_map, float* __restrict__ per_box,
                                                          float* __restrict__ set_out, int set_mode) {
  extern __shared__ float sm[];
  const int P = grid * grid;
  float* A = sm;          // private copy of the attention map
  float* total = sm + P;  // box-set accumulation
  __shared__ float red[8];
  const int b = blockIdx.x;
  for (int p = threadIdx.x; p < P; p += blockDim.x) {
    A[p] = (mode == PIO_POOL_ATTN) ? attn_map[(long long)b * P + p] : 0.f;
    total[p] = 0.f;
  }
  __syncthreads();
  for (int j = 0; j < R; ++j) {
    const int bi = b * R + j;
    if (set_mode && skip[bi]) continue;  // dummy box (bbox_utils.py:40-42)
    const int y0 = bounds[4 * bi], y1 = bounds[4 * bi + 1], x0 = bounds[4 * bi + 2], x1 = bounds[4 * bi + 3];
    const int hs = y1 - y0, ws = x1 - x0, area = hs * ws;
    float part = 0.f;
    for (int i = threadIdx.x; i < area; i += blockDim.x) {
      const int iy = i / ws, ix = i % ws;
      if (mode == PIO_POOL_ATTN) part += A[(y0 + iy) * grid + x0 + ix];
      else if (mode == PIO_POOL_GAUSS) part += gauss_w(iy, hs, ix, ws, variance);
      else part += 1.0f;
    }
    const float s = block_sum(part, red);
    if (per_box)
      for (int p = threadIdx.x; p < P; p += blockDim.x) per_box[((long long)bi) * P + p] = 0.f;
    __syncthreads();
    for (int i = threadIdx.x; i < area; i += blockDim.x) {
      const int iy = i / ws, ix = i % ws, p = (y0 + iy) * grid + x0 + ix;
      float w;
      if (mode == PIO_POOL_ATTN) { w = A[p] / s; A[p] = w; }
      else if (mode == PIO_POOL_GAUSS) w = gauss_w(iy, hs, ix, ws, variance) / s;
      else w = 1.0f / (float)area;
      total[p] += w;
      if (per_box) per_box[((long long)bi) * P + p] = w;
    }
    __syncthreads();
  }
  if (set_out) {
    float part = 0.f;
    for (int p = threadIdx.x; p < P; p += blockDim.x) part += total[p];
    const float s = block_sum(part, red);
    for (int p = threadIdx.x; p < P; p += blockDim.x) set_out[(long long)b * P + p] = total[p] / s;
  }
}

// Slab pooling.  grid = (D/32, B).  SRC 0: mean over the box, 1: gaussian over the box,
// 2: explicit weights w[b,r,P] restricted to `bounds` (or the whole grid when bounds == NULL), times `scale`.
constexpr int POOL_THREADS = 1024;
constexpr int POOL_RGROUP = 64;  // boxes accumulated per shared-memory pass

template <int SRC>
__global__ void __launch_bounds__(POOL_THREADS) pool_slab_kernel(const float* __restrict__ tokens, long long img_stride,
                                                                 long long row_stride, int grid, int D,
                                                                 const int* __restrict__ bounds, int R, float variance,
                                                                 const float* __restrict__ weights, float scale,
                                                                 float* __restrict__ out) {
  extern __shared__ __align__(128) float slab[];  // [P][32] | s_out [POOL_RGROUP][32] | s_norm [POOL_RGROUP] | s_bounds [POOL_RGROUP][4]
  const int P = grid * grid;
  const int b = blockIdx.y, c0 = blockIdx.x * 32;
  float* s_out = slab + (size_t)P * 32;
  float* s_norm = s_out + POOL_RGROUP * 32;
  int* s_bounds = reinterpret_cast<int*>(s_norm + POOL_RGROUP);
  {
    // stage the slab with cp.async: 8 lanes x 16 bytes cover one 128-byte row segment, ~11 copies in flight per
    // thread and 1024 threads -> the whole 175 KB slab is requested at once (no register staging, full MLP)
    const float* src = tokens + (long long)b * img_stride + c0;
    const uint32_t dst0 = tc::smem_u32(slab);
    for (int i = threadIdx.x; i < P * 8; i += POOL_THREADS) {
      const int p = i >> 3, q = (i & 7) * 4;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst0 + (uint32_t)(p * 32 + q) * 4),
                   "l"(src + (long long)p * row_stride + q)
                   : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
  }
  // Region walk: work unit = (box j, grid row y of the box).  Warp w takes the rows r of box j with
  // (r + j) % 32 == w, so large boxes are spread over all warps and the rotation by j balances small ones.
  // lane = channel; per row the lane-distributed x-weights are broadcast with one shuffle per patch, four
  // independent accumulators break the FMA dependency chain; partial sums meet in shared memory.
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int jg = 0; jg < R; jg += POOL_RGROUP) {
    const int rg = min(POOL_RGROUP, R - jg);
    for (int i = threadIdx.x; i < rg * 32; i += POOL_THREADS) s_out[i] = 0.f;
    for (int i = threadIdx.x; i < rg * 4; i += POOL_THREADS)  // this group's slice bounds -> shared memory
      s_bounds[i] = bounds ? bounds[4 * (b * R + jg) + i] : ((i & 1) ? grid : 0);
    __syncthreads();
    // per-box normalisation (one warp per box)
    for (int jj = warp; jj < rg; jj += POOL_THREADS / 32) {
      const int y0 = s_bounds[4 * jj], y1 = s_bounds[4 * jj + 1], x0 = s_bounds[4 * jj + 2], x1 = s_bounds[4 * jj + 3];
      const int hs = y1 - y0, ws = x1 - x0;
      float nrm;
      if (SRC == 0) nrm = 1.0f / (float)(hs * ws);  // empty box -> inf -> 0 * inf = NaN like tensor.mean()
      else if (SRC == 1) {
        float sy = 0.f, sx = 0.f;
        for (int i = lane; i < hs; i += 32) { const float y = linspace_m1_1(i, hs); sy += expf(-(y * y) / variance); }
        for (int i = lane; i < ws; i += 32) { const float x = linspace_m1_1(i, ws); sx += expf(-(x * x) / variance); }
        nrm = (hs > 0 && ws > 0) ? 1.0f / (warp_sum(sy) * warp_sum(sx)) : 0.f;  // empty box -> zeros, like the reference
      } else nrm = scale;
      if (lane == 0) s_norm[jj] = nrm;
    }
    __syncthreads();
    for (int it = 0; it < rg; ++it) {
      // every warp starts at a different box, so concurrent shared-memory accumulations rarely collide
      int jj = it + 2 * warp;
      jj -= (jj >= rg) ? rg * (jj / rg) : 0;
      const int bi = b * R + jg + jj;
      const int y0 = s_bounds[4 * jj], y1 = s_bounds[4 * jj + 1], x0 = s_bounds[4 * jj + 2], x1 = s_bounds[4 * jj + 3];
      const int hs = y1 - y0, ws = x1 - x0;
      const int r0 = (warp - jj) & 31;
      if (r0 >= hs || ws <= 0) continue;  // no row of this box for this warp
      float wx0 = 0.f, wx1 = 0.f;       // separable gaussian: x-factors for columns lane and lane + 32 (grid <= 64)
      if (SRC == 1) {
        if (lane < ws) { const float x = linspace_m1_1(lane, ws); wx0 = expf(-(x * x) / variance); }
        if (lane + 32 < ws) { const float x = linspace_m1_1(lane + 32, ws); wx1 = expf(-(x * x) / variance); }
      }
      float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
      for (int r = r0; r < hs; r += 32) {
        const int prow = (y0 + r) * grid + x0;
        float wy = 1.0f;
        if (SRC == 1) { const float y = linspace_m1_1(r, hs); wy = expf(-(y * y) / variance); }
        for (int xb = 0; xb < ws; xb += 32) {
          const int cnt = min(32, ws - xb);
          float w = 0.f;  // this lane's weight for patch (row r, column xb + lane)
          if (lane < cnt) {
            if (SRC == 0) w = 1.0f;
            else if (SRC == 1) w = wy * (xb == 0 ? wx0 : wx1);
            else w = __ldg(weights + (long long)bi * P + prow + xb + lane);
          }
          const float* sp = slab + (prow + xb) * 32 + lane;
          if (SRC == 2) {
            // explicit weights are often sparse (trace histograms): visit only the non-zero patches
            unsigned nz = __ballot_sync(0xffffffffu, w != 0.f);
            while (nz) {
              const int k = __ffs(nz) - 1;
              nz &= nz - 1;
              acc0 = fmaf(__shfl_sync(0xffffffffu, w, k), sp[k * 32], acc0);
            }
          } else {
            int k = 0;
            for (; k + 4 <= cnt; k += 4) {
              acc0 = fmaf(__shfl_sync(0xffffffffu, w, k), sp[k * 32], acc0);
              acc1 = fmaf(__shfl_sync(0xffffffffu, w, k + 1), sp[(k + 1) * 32], acc1);
              acc2 = fmaf(__shfl_sync(0xffffffffu, w, k + 2), sp[(k + 2) * 32], acc2);
              acc3 = fmaf(__shfl_sync(0xffffffffu, w, k + 3), sp[(k + 3) * 32], acc3);
            }
            for (; k < cnt; ++k) acc0 = fmaf(__shfl_sync(0xffffffffu, w, k), sp[k * 32], acc0);
          }
        }
      }
      atomicAdd(&s_out[jj * 32 + lane], (acc0 + acc1) + (acc2 + acc3));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < rg * 32; i += POOL_THREADS) {
      const int jj = i >> 5, l = i & 31;
      out[(long long)(b * R + jg + jj) * D + c0 + l] = s_out[i] * s_norm[jj];
    }
    __syncthreads();
  }
}


// ---------------------------------------------------------------------------------- box pooling (mean / gaussian)
// Warp-per-box variant for the two closed-form weightings (the dense-captioning hot path: 64 boxes per image).
// grid = (D/16, B): a CTA stages a [P x 16-channel] slab (64-byte row segments, 88 KB at 518 px) so TWO CTAs share an SM and
// one CTA's cp.async staging overlaps the other's arithmetic.  Each warp owns whole boxes (no shared-memory atomics, no
// per-box bookkeeping by the other warps): a warp visits 8 patches of a box row per step, 4 lanes x float4 per patch, so one
// LDS.128 per lane reads 512 contiguous bytes of the slab per warp (conflict free) and feeds 4 FMAs.  The separable x / y
// factors come from box_factors_kernel's table (lane-distributed, broadcast with shuffles, x-factors hoisted per box).
constexpr int PB_THREADS = 512;
// Per-box separable factors, computed ONCE per box (the pooling CTAs of an image -- 48 channel slabs -- would otherwise each
// redo the expf / linspace / normalisation work, which dominated their instruction count): fac[bi] = wx[64] | wy[64] | nrm.
// mode 0 (mean): wx = wy = 1 inside the box, nrm = 1 / (hs ws) (inf for an empty box: 0 * inf = NaN like tensor.mean());
// mode 1 (gaussian): exp(-x^2 / var) on linspace(-1, 1, span), nrm = 1 / (sum_y sum_x) (0 for an empty box, like the reference).
constexpr int FAC_STRIDE = 132;
__global__ void __launch_bounds__(256) box_factors_kernel(const int* __restrict__ bounds, int n, int mode, float variance,
                                                          float* __restrict__ fac) {
  const int bi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (bi >= n) return;
  const int4 bd = __ldg(reinterpret_cast<const int4*>(bounds) + bi);
  const int hs = bd.y - bd.x, ws = bd.w - bd.z;
  float wxa = 0.f, wxb = 0.f, wya = 0.f, wyb = 0.f, nrm;
  if (mode == 1) {
    if (lane < ws) { const float x = linspace_m1_1(lane, ws); wxa = expf(-(x * x) / variance); }
    if (lane + 32 < ws) { const float x = linspace_m1_1(lane + 32, ws); wxb = expf(-(x * x) / variance); }
    if (lane < hs) { const float y = linspace_m1_1(lane, hs); wya = expf(-(y * y) / variance); }
    if (lane + 32 < hs) { const float y = linspace_m1_1(lane + 32, hs); wyb = expf(-(y * y) / variance); }
    const float sx = warp_sum(wxa + wxb), sy = warp_sum(wya + wyb);
    nrm = (hs > 0 && ws > 0) ? 1.0f / (sy * sx) : 0.f;
  } else {
    wxa = lane < ws ? 1.f : 0.f; wxb = lane + 32 < ws ? 1.f : 0.f;
    wya = lane < hs ? 1.f : 0.f; wyb = lane + 32 < hs ? 1.f : 0.f;
    nrm = 1.0f / (float)(hs * ws);
  }
  float* f = fac + (long long)bi * FAC_STRIDE;
  f[lane] = wxa; f[lane + 32] = wxb; f[64 + lane] = wya; f[96 + lane] = wyb;
  if (lane == 0) f[128] = nrm;
}

// Rows of one box for a warp: STEPS x 8 patches per row (x = 8 k + grp), float4 = 4 channels per lane.
template <bool GAUSS, int STEPS>
__device__ __forceinline__ float4 box_rows(const float4* sp0, int grid, int hs, int ws, int grp, float wxa, float wxb, float wya,
                                           float wyb) {
  float wxk[STEPS];
  int off[STEPS];
#pragma unroll
  for (int k = 0; k < STEPS; ++k) {
    const int x = 8 * k + grp;
    wxk[k] = __shfl_sync(0xffffffffu, x < 32 ? wxa : wxb, x & 31);  // 0 beyond the box
    off[k] = min(x, ws - 1) * 4;  // float4 units; lanes beyond the row re-read its last patch with weight 0
  }
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f), acc2 = acc;
  const float4* sp = sp0;
  for (int r = 0; r < hs; ++r, sp += grid * 4) {
    float wy = 1.0f;
    if (GAUSS) wy = __shfl_sync(0xffffffffu, r < 32 ? wya : wyb, r & 31);
#pragma unroll
    for (int k = 0; k < STEPS; ++k) {
      const float w = GAUSS ? wy * wxk[k] : wxk[k];
      const float4 v = sp[off[k]];
      if (k & 1) { acc2.x = fmaf(w, v.x, acc2.x); acc2.y = fmaf(w, v.y, acc2.y); acc2.z = fmaf(w, v.z, acc2.z); acc2.w = fmaf(w, v.w, acc2.w); }
      else       { acc.x = fmaf(w, v.x, acc.x);   acc.y = fmaf(w, v.y, acc.y);   acc.z = fmaf(w, v.z, acc.z);   acc.w = fmaf(w, v.w, acc.w); }
    }
  }
  return make_float4(acc.x + acc2.x, acc.y + acc2.y, acc.z + acc2.z, acc.w + acc2.w);
}

template <bool GAUSS>  // mean mode needs no factor table: wx is the box indicator, nrm = 1 / (hs ws)
__global__ void __launch_bounds__(PB_THREADS, 2) pool_box_kernel(const float* __restrict__ tokens, long long img_stride,
                                                                long long row_stride, int grid, int D,
                                                                const int* __restrict__ bounds, const float* __restrict__ fac, int R,
                                                                float* __restrict__ out) {
  extern __shared__ __align__(128) float slab[];  // [P][16]
  const int P = grid * grid;
  const int b = blockIdx.y, c0 = blockIdx.x * 16;
  {
    // thread t copies 16-byte piece t & 3 of patches t >> 2, t >> 2 + 128, ...: two pointer bumps per copy
    const float* src = tokens + (long long)b * img_stride + c0 + (long long)(threadIdx.x >> 2) * row_stride + (threadIdx.x & 3) * 4;
    uint32_t dst = tc::smem_u32(slab) + threadIdx.x * 16;
    const long long sstep = (long long)(PB_THREADS / 4) * row_stride;
    for (int p = threadIdx.x >> 2; p < P; p += PB_THREADS / 4, src += sstep, dst += PB_THREADS * 16)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = lane >> 2, quad = lane & 3;  // 8 patches per step, each read as 4 x float4 (16 channels = 64 bytes)
  bool staged = false;
  constexpr int NW = PB_THREADS / 32;
  // lane l fetches the bounds of this warp's l-th box up front (one load under the staging copy instead of one dependent
  // global load per box); they are handed out with shuffles
  int4 myb = make_int4(0, 0, 0, 0);
  for (int k = 0; warp + k * NW < R; ++k) {
    if ((k & 31) == 0) {
      const int jm = warp + (k + lane) * NW;
      myb = jm < R ? __ldg(reinterpret_cast<const int4*>(bounds) + b * R + jm) : make_int4(0, 0, 0, 0);
    }
    const int bi = b * R + warp + k * NW;
    int4 bd;  // y0, y1, x0, x1
    bd.x = __shfl_sync(0xffffffffu, myb.x, k & 31); bd.y = __shfl_sync(0xffffffffu, myb.y, k & 31);
    bd.z = __shfl_sync(0xffffffffu, myb.z, k & 31); bd.w = __shfl_sync(0xffffffffu, myb.w, k & 31);
    const int hs = bd.y - bd.x, ws = bd.w - bd.z;
    float wxa, wxb, wya = 1.f, wyb = 1.f, nrm;
    if (GAUSS) {
      const float* f = fac + (long long)bi * FAC_STRIDE;
      wxa = __ldg(f + lane); wxb = __ldg(f + 32 + lane); wya = __ldg(f + 64 + lane); wyb = __ldg(f + 96 + lane);
      nrm = __ldg(f + 128);
    } else {
      wxa = lane < ws ? 1.f : 0.f; wxb = lane + 32 < ws ? 1.f : 0.f;
      nrm = 1.0f / (float)(hs * ws);  // empty box -> inf -> 0 * inf = NaN like tensor.mean()
    }
    if (!staged) {  // the box parameters above overlap the tail of the staging copy
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();
      staged = true;
    }
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4* sp0 = reinterpret_cast<const float4*>(slab + (bd.x * grid + bd.z) * 16) + quad;
    switch ((ws + 7) >> 3) {  // warp-uniform: a fully unrolled row loop per step count, no predicated-off issue slots
      case 1: acc = box_rows<GAUSS, 1>(sp0, grid, hs, ws, grp, wxa, wxb, wya, wyb); break;
      case 2: acc = box_rows<GAUSS, 2>(sp0, grid, hs, ws, grp, wxa, wxb, wya, wyb); break;
      case 3: acc = box_rows<GAUSS, 3>(sp0, grid, hs, ws, grp, wxa, wxb, wya, wyb); break;
      case 4: acc = box_rows<GAUSS, 4>(sp0, grid, hs, ws, grp, wxa, wxb, wya, wyb); break;
      case 5: acc = box_rows<GAUSS, 5>(sp0, grid, hs, ws, grp, wxa, wxb, wya, wyb); break;
      case 6: acc = box_rows<GAUSS, 6>(sp0, grid, hs, ws, grp, wxa, wxb, wya, wyb); break;
      case 7: acc = box_rows<GAUSS, 7>(sp0, grid, hs, ws, grp, wxa, wxb, wya, wyb); break;
      case 8: acc = box_rows<GAUSS, 8>(sp0, grid, hs, ws, grp, wxa, wxb, wya, wyb); break;
      default: break;  // empty box
    }
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) {
      acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
      acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o); acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
    }
    if (grp == 0)
      *reinterpret_cast<float4*>(out + (long long)bi * D + c0 + quad * 4) = make_float4(acc.x * nrm, acc.y * nrm, acc.z * nrm, acc.w * nrm);
  }
  if (!staged) {  // more warps than boxes: nobody may leave while copies into this CTA's shared memory are in flight
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------- traces
// One CTA per trace; Python-double binning of bbox_utils.py:158-168.
__global__ void __launch_bounds__(256) trace_bins_kernel(const double* __restrict__ pts, const int* __restrict__ offsets, int grid,
                                                         const float* __restrict__ attn, float* __restrict__ counts) {
  extern __shared__ int hist[];
  const int P = grid * grid, t = blockIdx.x;
  for (int p = threadIdx.x; p < P; p += blockDim.x) hist[p] = 0;
  __syncthreads();
  const double patch_size = 1.0 / (double)grid;
  for (int i = offsets[t] + threadIdx.x; i < offsets[t + 1]; i += blockDim.x) {
    const double x = pts[2 * i], y = pts[2 * i + 1];
    if (0.0 <= x && x <= 1.0 && 0.0 <= y && y <= 1.0) {
      int gx = (int)(x / patch_size), gy = (int)(y / patch_size);
      gx = min(gx, grid - 1);
      gy = min(gy, grid - 1);
      atomicAdd(&hist[gy * grid + gx], 1);
    }
  }
  __syncthreads();
  for (int p = threadIdx.x; p < P; p += blockDim.x) {
    float c = (float)hist[p];
    if (attn) c = attn[(long long)t * P + p] * c;  // model.py:1053
    counts[(long long)t * P + p] = c;
  }
}

// ---------------------------------------------------------------------------------- CLS attention map
// logits[b,j] = <q_cls[b], k[b,5+j]> / 128  -- one warp per (b, patch); softmax afterwards.
template <typename T>
__global__ void __launch_bounds__(256) cls_logits_kernel(const T* __restrict__ qkv, int B, int N, int D, int ng,
                                                         float* __restrict__ logits) {
  const int P = N - ng;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= (long long)B * P) return;
  const int b = (int)(warp / P), j = (int)(warp % P);
  const T* q = qkv + (long long)b * N * 3 * D;
  const T* k = qkv + ((long long)b * N + ng + j) * 3 * D + D;
  float s = 0.f;
  for (int d = lane * 4; d < D; d += 128) {
    float4 a, c;
    if constexpr (sizeof(T) == 4) {
      a = *reinterpret_cast<const float4*>(q + d);
      c = *reinterpret_cast<const float4*>(k + d);
    } else {
      uint2 ua = *reinterpret_cast<const uint2*>(q + d), uc = *reinterpret_cast<const uint2*>(k + d);
      float2 a0 = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&ua.x)), a1 = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&ua.y));
      float2 c0 = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&uc.x)), c1 = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&uc.y));
      a = make_float4(a0.x, a0.y, a1.x, a1.y);
      c = make_float4(c0.x, c0.y, c1.x, c1.y);
    }
    s += a.x * c.x + a.y * c.y + a.z * c.z + a.w * c.w;
  }
  s = warp_sum(s);
  if (lane == 0) logits[warp] = s * (0.125f / 16.0f);
}

// Per-"head" CLS logits of process_self_attention(ret_self_attn_maps=True) (dino_extraction.py:24-34): the hooked qkv is
// re-cut into `heads` = 16 groups of D / heads = 48 channels (sic, model.py:336), maps[b,h,j] = 0.125 <q_cls^h, k_j^h>.
// One warp per (b, patch): lane = (head, half), 24 contiguous channels each.
template <typename T>
__global__ void __launch_bounds__(256) cls_head_logits_kernel(const T* __restrict__ qkv, int B, int N, int D, int ng, int heads,
                                                              float scale, float* __restrict__ maps) {
  const int P = N - ng;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= (long long)B * P) return;
  const int b = (int)(warp / P), j = (int)(warp % P);
  const int hd = D / heads, per = hd / 2;  // 48, 24
  const int head = lane >> 1, c0 = head * hd + (lane & 1) * per;
  const T* q = qkv + (long long)b * N * 3 * D + c0;
  const T* k = qkv + ((long long)b * N + ng + j) * 3 * D + D + c0;
  float s = 0.f;
  for (int d = 0; d < per; ++d) s = fmaf((float)q[d], (float)k[d], s);
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  if ((lane & 1) == 0 && head < heads) maps[((long long)b * heads + head) * P + j] = s * scale;
}

// ctx_cleaner (model.py:1425-1436): rows [B*P, D] against one context vector per image.  Warp per row.
//   mode 0 orthogonal_projection: d - alpha * (<d, c> / |c|^2) * c        mode 1 contrastive_mask: d * (1 - c / (|c| + eps))
//   prenorm != 0: rows and context are L2-normalised first (the clean_after_projection=False branch, model.py:905-913).
__global__ void __launch_bounds__(256) ctx_clean_kernel(const float* __restrict__ rows, long long img_stride, long long row_stride,
                                                        const float* __restrict__ ctx, long long ctx_stride, int B, int P, int D,
                                                        int mode, float alpha, float eps, int prenorm, float* __restrict__ out) {
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= (long long)B * P) return;
  const int b = (int)(warp / P), p = (int)(warp % P);
  const float* d = rows + (long long)b * img_stride + (long long)p * row_stride;
  const float* c = ctx + (long long)b * ctx_stride;
  float* o = out + warp * D;
  float dc = 0.f, cc = 0.f, dd = 0.f;
  for (int i = lane; i < D; i += 32) { const float x = d[i], y = c[i]; dc = fmaf(x, y, dc); cc = fmaf(y, y, cc); dd = fmaf(x, x, dd); }
  dc = warp_sum(dc); cc = warp_sum(cc); dd = warp_sum(dd);
  float dn = 1.0f, cn = 1.0f;  // 1 / |d|, 1 / |c| when pre-normalising
  if (prenorm) { dn = 1.0f / sqrtf(dd); cn = 1.0f / sqrtf(cc); dc *= dn * cn; cc = 1.0f; }
  if (mode == 0) {
    const float k = alpha * dc / cc;
    for (int i = lane; i < D; i += 32) o[i] = d[i] * dn - k * (c[i] * cn);
  } else {
    const float inv = 1.0f / (sqrtf(cc) + eps);
    for (int i = lane; i < D; i += 32) o[i] = (d[i] * dn) * (1.0f - (c[i] * cn) * inv);
  }
}

__global__ void __launch_bounds__(256) region_mean_weights_kernel(int grid, float variance, float* __restrict__ w) {
  __shared__ float red[8];
  const int P = grid * grid;
  if (variance >= 100.f) {
    for (int p = threadIdx.x; p < P; p += blockDim.x) w[p] = 1.0f / (float)P;
    return;
  }
  float part = 0.f;
  for (int p = threadIdx.x; p < P; p += blockDim.x) part += gauss_w(p / grid, grid, p % grid, grid, variance);
  const float s = block_sum(part, red);
  for (int p = threadIdx.x; p < P; p += blockDim.x) w[p] = gauss_w(p / grid, grid, p % grid, grid, variance) / s;
}

template <int SRC>
int launch_slab(const float* tokens, long long img_stride, long long row_stride, int B, int grid, int D, const int* bounds,
                int R, float variance, const float* weights, float scale, float* out, cudaStream_t st) {
  const size_t smem = ((size_t)grid * grid * 32 + POOL_RGROUP * 37) * sizeof(float);
  PIO_CHECK(smem <= 227 * 1024, "pooling: grid %d too large for the shared-memory slab", grid);
  PIO_CUDA(cudaFuncSetAttribute(pool_slab_kernel<SRC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 g(D / 32, B);
  pool_slab_kernel<SRC><<<g, POOL_THREADS, smem, st>>>(tokens, img_stride, row_stride, grid, D, bounds, R, variance, weights, scale, out);
  PIO_LAUNCHED();
  return PIO_OK;
}

int launch_box(const float* tokens, long long img_stride, long long row_stride, int B, int grid, int D, const int* bounds, int R,
               int mode, float variance, float* fac, float* out, cudaStream_t st) {
  const size_t smem = (size_t)grid * grid * 16 * sizeof(float);
  static SmemAttrOnce once_g, once_m;
  PIO_CUDA(once_g.ensure(pool_box_kernel<true>, 113 * 1024));
  PIO_CUDA(once_m.ensure(pool_box_kernel<false>, 113 * 1024));
  dim3 g(D / 16, B);
  if (mode == 1) {
    box_factors_kernel<<<cdiv((long long)B * R * 32, 256), 256, 0, st>>>(bounds, B * R, mode, variance, fac);
    PIO_LAUNCHED();
    pool_box_kernel<true><<<g, PB_THREADS, smem, st>>>(tokens, img_stride, row_stride, grid, D, bounds, fac, R, out);
  } else {
    pool_box_kernel<false><<<g, PB_THREADS, smem, st>>>(tokens, img_stride, row_stride, grid, D, bounds, fac, R, out);
  }
  PIO_LAUNCHED();
  return PIO_OK;
}

}  // namespace

int cls_attention(const void* qkv, int dt, int B, int N, int D, int ng, float* out, float* logits_ws, cudaStream_t st) {
  const int P = N - ng;
  PIO_CHECK(D % 128 == 0, "cls attention: D %d must be a multiple of 128", D);
  const int blocks = cdiv((long long)B * P * 32, 256);
  if (dt == PIO_DT_F32)
    cls_logits_kernel<float><<<blocks, 256, 0, st>>>((const float*)qkv, B, N, D, ng, logits_ws);
  else
    cls_logits_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)qkv, B, N, D, ng, logits_ws);
  PIO_LAUNCHED();
  return softmax_rows(logits_ws, out, B, P, 1.0f, st);
}

}  // namespace pio

extern "C" {

size_t pio_pool_workspace_bytes(int B, int R, int grid) {
  const size_t P = (size_t)grid * grid;
  return pio::align_up((size_t)B * R * 4 * sizeof(int), 256) + pio::align_up((size_t)B * R * sizeof(int), 256) +
         pio::align_up((size_t)B * R * (P > 132 ? P : 132) * sizeof(float), 256) /* per-box weights, or the 132-float factor rows */ +
         pio::align_up((size_t)B * P * sizeof(float), 256) + 1024;
}

int pio_pool_boxes(const float* tokens, long long img_stride, long long row_stride, int B, int grid, int D, const void* boxes,
                   int boxes_dt, int R, int patch_size, int mode, float variance, const float* attn_map, int set_mode,
                   float* out, int* out_bounds, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace pio;
  cudaStream_t st = as_stream(stream);
  PIO_CHECK(D % 32 == 0 && row_stride % 4 == 0 && img_stride % 4 == 0 && (((uintptr_t)tokens) & 15) == 0,
            "pool_boxes: D must be a multiple of 32 and tokens 16-byte aligned");
  PIO_CHECK(boxes_dt == PIO_DT_F32 || boxes_dt == PIO_DT_I32, "pool_boxes: boxes must be float32 or int32");
  PIO_CHECK(mode == PIO_POOL_MEAN || mode == PIO_POOL_GAUSS || mode == PIO_POOL_ATTN, "pool_boxes: bad mode %d", mode);
  PIO_CHECK(mode != PIO_POOL_ATTN || attn_map, "pool_boxes: attention mode needs attn_map");
  PIO_CHECK(!(mode == PIO_POOL_GAUSS && variance == 0.f),
            "pool_boxes: gaussian variance 0 (python-random centre, bbox_utils.py:62-71) is not supported");
  PIO_CHECK(workspace_bytes >= pio_pool_workspace_bytes(B, R, grid), "pool_boxes: workspace too small");
  if (B == 0 || R == 0) return PIO_OK;
  const size_t P = (size_t)grid * grid;
  char* ws = (char*)workspace;
  int* bounds = (int*)ws; ws += align_up((size_t)B * R * 4 * sizeof(int), 256);
  int* skip = (int*)ws;   ws += align_up((size_t)B * R * sizeof(int), 256);
  float* per_box = (float*)ws; ws += align_up((size_t)B * R * (P > 132 ? P : 132) * sizeof(float), 256);
  float* set_map = (float*)ws;
  box_bounds_kernel<<<cdiv(B * R, 128), 128, 0, st>>>(boxes, boxes_dt, B * R, patch_size, grid, bounds, skip);
  PIO_LAUNCHED();
  if (out_bounds) PIO_CUDA(cudaMemcpyAsync(out_bounds, bounds, (size_t)B * R * 4 * sizeof(int), cudaMemcpyDeviceToDevice, st));
  const size_t wsmem = 2 * P * sizeof(float);
  if (set_mode) {
    box_weights_kernel<<<B, 256, wsmem, st>>>(bounds, skip, R, grid, mode, variance, attn_map, nullptr, set_map, 1);
    PIO_LAUNCHED();
    return launch_slab<2>(tokens, img_stride, row_stride, B, grid, D, nullptr, 1, 0.f, set_map, 1.0f, out, st);
  }
  if (mode == PIO_POOL_ATTN) {
    box_weights_kernel<<<B, 256, wsmem, st>>>(bounds, skip, R, grid, mode, variance, attn_map, per_box, nullptr, 0);
    PIO_LAUNCHED();
    return launch_slab<2>(tokens, img_stride, row_stride, B, grid, D, bounds, R, 0.f, per_box, 1.0f, out, st);
  }
  // closed-form weights: warp-per-box kernel (two CTAs per SM when the 16-channel slab fits twice); PIO_POOL_SLAB=1 keeps
  // the 32-channel slab kernel for A/B runs
  static const bool old_slab = [] { const char* e = getenv("PIO_POOL_SLAB"); return e && e[0] == '1'; }();
  const bool box_ok = !old_slab && grid <= 64 && (size_t)grid * grid * 16 * sizeof(float) <= 113 * 1024;
  if (box_ok)  // the factor table lives in the (otherwise unused here) per-box weight region of the workspace
    return launch_box(tokens, img_stride, row_stride, B, grid, D, bounds, R, mode == PIO_POOL_GAUSS ? 1 : 0, variance, per_box, out, st);
  if (mode == PIO_POOL_GAUSS)
    return launch_slab<1>(tokens, img_stride, row_stride, B, grid, D, bounds, R, variance, nullptr, 1.0f, out, st);
  return launch_slab<0>(tokens, img_stride, row_stride, B, grid, D, bounds, R, 0.f, nullptr, 1.0f, out, st);
}

int pio_pool_grid(const float* tokens, long long img_stride, long long row_stride, int B, int grid, int D,
                  const float* weights, int R, float scale, float* out, void* stream) {
  using namespace pio;
  PIO_CHECK(D % 32 == 0 && row_stride % 4 == 0 && img_stride % 4 == 0 && (((uintptr_t)tokens) & 15) == 0,
            "pool_grid: D must be a multiple of 32 and tokens 16-byte aligned");
  if (B == 0 || R == 0) return PIO_OK;
  return launch_slab<2>(tokens, img_stride, row_stride, B, grid, D, nullptr, R, 0.f, weights, scale, out, as_stream(stream));
}

int pio_trace_bins(const double* points_xy, const int* offsets, int T, int grid, const float* attn, float* counts, void* stream) {
  using namespace pio;
  if (T == 0) return PIO_OK;
  trace_bins_kernel<<<T, 256, (size_t)grid * grid * sizeof(int), as_stream(stream)>>>(points_xy, offsets, grid, attn, counts);
  PIO_LAUNCHED();
  return PIO_OK;
}

int pio_cls_head_attention(const void* qkv, int qkv_dt, int B, int N, int D, int num_global, int heads, float scale,
                           float* out_maps, void* stream) {
  using namespace pio;
  PIO_CHECK(qkv && out_maps, "cls_head_attention: null argument");
  PIO_CHECK(heads == 16 && D % (2 * heads) == 0, "cls_head_attention: 16 head groups of an even width are supported (model.py:336)");
  if (B == 0) return PIO_OK;
  cudaStream_t st = as_stream(stream);
  const int P = N - num_global;
  const int blocks = cdiv((long long)B * P * 32, 256);
  if (qkv_dt == PIO_DT_F32)
    cls_head_logits_kernel<float><<<blocks, 256, 0, st>>>((const float*)qkv, B, N, D, num_global, heads, scale, out_maps);
  else
    cls_head_logits_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)qkv, B, N, D, num_global, heads, scale, out_maps);
  PIO_LAUNCHED();
  return softmax_rows(out_maps, out_maps, B * heads, P, 1.0f, st);  // self_attn_maps.softmax(dim=-1), model.py:871
}

int pio_ctx_clean(const float* rows, long long img_stride, long long row_stride, const float* ctx, long long ctx_stride, int B, int P,
                  int D, int mode, float alpha, float eps, int prenorm, float* out, void* stream) {
  using namespace pio;
  PIO_CHECK(rows && ctx && out, "ctx_clean: null argument");
  PIO_CHECK(mode == 0 || mode == 1, "ctx_clean: mode %d (0 orthogonal_projection, 1 contrastive_mask)", mode);
  if (B == 0 || P == 0) return PIO_OK;
  ctx_clean_kernel<<<cdiv((long long)B * P * 32, 256), 256, 0, as_stream(stream)>>>(rows, img_stride, row_stride, ctx, ctx_stride, B, P, D,
                                                                                   mode, alpha, eps, prenorm, out);
  PIO_LAUNCHED();
  return PIO_OK;
}

int pio_region_mean_weights(int grid, float variance, float* weights, void* stream) {
  using namespace pio;
  PIO_CHECK(variance != 0.f, "region_mean_weights: variance 0 (python-random centre, model.py:71-77) is not supported");
  region_mean_weights_kernel<<<1, 256, 0, as_stream(stream)>>>(grid, variance, weights);
  PIO_LAUNCHED();
  return PIO_OK;
}

int pio_cls_attention(const void* qkv, int qkv_dt, int B, int N, int D, int num_global, float* out_attn, void* stream) {
  using namespace pio;
  // logits are staged in out_attn itself (softmax_rows is safe in place: each CTA owns one row)
  return cls_attention(qkv, qkv_dt, B, N, D, num_global, out_attn, out_attn, as_stream(stream));
}
}

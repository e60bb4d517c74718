// decode_fused_sm100.cu -- the whole KV-cached greedy decode at small batch (R <= 64 rows) as ONE persistent, cooperative
// kernel: the north star's "kernels fused per decode step".  Replaces the ~32 dependent launches per position of text.cu
// (decoding_batched, src/decap/decap.py:130-155; greedy_search, src/viecap/search.py:108-191) whose fixed cost -- launch,
// TMEM allocation, barrier init, tensor-map fetch, first TMA round trip, grid completion -- was ~8 us each (8 % of the HBM
// roofline at 32 rows).
//
// One CTA per SM (cooperative launch: the grid-wide phase counters below need co-residency).  A decode step is a sequence of
// PHASES; a phase is done when every CTA that has work in it has added 1 to the phase's counter in global memory (release),
// and whoever needs its results polls that counter (acquire):
//     block 0   :         QKV* | ATTN | PROJ(+res) | FC*(gelu_new) | FC2 (split-K 4, partials left unreduced)
//     blocks 1..:  LN1r | QKV  | ATTN | PROJ(+res) | FC*(gelu_new) | FC2
//     then      :  LNFr | LM-HEAD (per-CTA arg-max partials) | PICK (final arg-max, ids, next embedding -> x)
//   * = the LayerNorm in front of the layer is applied while the CTA loads its activations (x fp32 -> normalised bf16 tile);
//   r = the LayerNorm phase first reduces the previous fc2's split-K partials into x (fixed order: deterministic).
// Dense layers run on the tensor cores with the operands SWAPPED with respect to gemm_sm100.cu: a work unit is 128 OUTPUT
// FEATURES (the UMMA M dimension, rows of the [out, in] weight matrix) x 12 k-blocks of 64, against all R rows (the UMMA N
// dimension, R padded to a multiple of 16): D^T[128, R] = W_tile[128, 768] . X[R, 768]^T accumulated in TMEM.  Every unit of
// every layer is 196 KB of weights; unit u of phase g belongs to CTA (u + 41 g) mod G, so consecutive phases land on
// different SMs and -- because weights never change -- warp 0 of every CTA streams the weight tiles of its FUTURE units
// into a shared-memory ring as fast as slots free up, across phase and step boundaries: the weight stream from HBM never
// waits for a phase.  The activations of a phase (R x 768 bf16, L2 resident) are fetched by the eight compute warps with
// ordinary 16-byte loads straight into the 128B-swizzled UMMA layout -- NOT by TMA: measured (profiles/r02f_*), activation
// TMA loads queue behind the 16 KB weight prefetches in the SM's copy engine (3.7 us per phase); the LSU path is one L2
// round trip.
//   warp 0        weight producer   : cp.async.bulk.tensor of 128 x 64 weight tiles (128B swizzle), runs ahead
//   warp 2        MMA issuer        : tcgen05.mma (UMMA 128 x R_pad x 16), accumulator double-buffered in TMEM
//   warps 4..11   compute warps     : activation loads (+ fused LayerNorm), epilogues (tcgen05.ld: lane = output feature,
//                                     column = row), LayerNorm / reduce rows, KV-cache attention, arg-max, phase arrivals
// Cross-CTA data (x, h, qkv, attention rows, gelu rows, KV cache, partials) is written with ordinary stores and read with
// ld.global.cg (never the non-coherent path: the data changes inside the launch).
// Every wait is bounded (~2 s): on expiry the kernel raises `abort` in global memory and all roles drain, so a logic error
// ends as an error on the host instead of a hung GPU.
#include "tc_ptx.cuh"
#include "decoder.cuh"
#include "gemm_epilogue.cuh"

#include <algorithm>
#include <stdlib.h>
#include <vector>

namespace pio {
using namespace tc;

size_t fused_counter_ints(int L, int steps) { return (size_t)steps * (6 * L + 3) + 64; }

namespace {

constexpr int FT_THREADS = 384;          // 12 warps
constexpr int FT_COMPUTE_WARPS = 8;      // warps 4..11
constexpr int W_TILE_BYTES = 128 * 64 * 2;
constexpr int KB_PER_UNIT = 12;          // 768 / 64
constexpr int MAX_STAGES = 12;
constexpr int NACC = 4;                 // independent accumulators per unit (see the kernel); the epilogue is written for 4

enum PhaseType { P_LN1 = 0, P_QKV, P_ATTN, P_PROJ, P_FC, P_FC2, P_LNF, P_LMHEAD, P_PICK };
constexpr int MAX_PHASES = 6 * 12 + 3;  // phases of one step for up to 12 blocks

struct FusedParams {
  const CUtensorMap* wmaps;   // [4 L + 1]
  const FusedLayer* layers;   // [L]
  const float *lnf_w, *lnf_b, *wte32, *wpe;
  float* x;                   // [R, 768] residual stream
  __nv_bfloat16 *h, *qkv, *att, *f;
  __nv_bfloat16 *kc, *vc;     // [L][R][H][T][hd]
  long long kv_layer;         // elements per layer
  float* part;                // [6 * splits][R_pad][128]
  float* pm_val;              // [G][R_pad]
  int* pm_idx;
  int* phase_cnt;             // [steps * PPS]
  int* abort;
  unsigned long long* timeline;  // debug (PIO_FUSED_TIMELINE): 8 globaltimer stamps per (phase, CTA), or NULL
  int* out_ids;
  int ids_ld;
  int L, H, hd, T, R, R_pad, steps, pos_base, first_gp, G, nsw, gs, stop_gp, PPS, mma_mode, pace_ns, l2_prefetch;  // gs: ring stages per release group
  signed char ptype[MAX_PHASES], player[MAX_PHASES];  // the phases of one step: type and block index
};

// ------------------------------------------------------------------------------------------ small PTX helpers
__device__ __forceinline__ unsigned long long gtime_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int ld_relaxed(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add(int* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
// __threadfence() is membar.gl == fence.sc.gpu; release / acquire ordering is all the phase protocol needs
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ uint4 ld_cg16(const void* p) {
  uint4 v;
  asm volatile("ld.global.cg.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_cg_f4(const float* p) {
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_cg_f(const float* p) {
  float v;
  asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int ld_cg_i(const int* p) {
  int v;
  asm volatile("ld.global.cg.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_cg_u32(const void* p) {
  uint32_t v;
  asm volatile("ld.global.cg.b32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// ------------------------------------------------------------------------------------------ row-wise pieces (one warp per row / unit)
// LayerNorm of x[r] (fp32, 768) -> bf16, eps 1e-5 (GPT-2): to the global row `out` and / or into the CTA's shared-memory
// activation tiles (the B operand of the following UMMAs: 12 k-block tiles of [R_pad rows x 128 bytes], 128B swizzle).
// part != NULL: the previous block's fc2 left its 4 split-K partial tiles unreduced -- this warp first forms
//   x[r] += (p0 + p1 + p2 + p3) + fc2_b   (fixed order: deterministic), stores the row back, then normalises it.
// Lane l owns the float4 at column 4 (l + 32 i), i < 6: output tile i of the [6 x 128]-column layout, offset 4 l.
__device__ __forceinline__ void ln_load(const float* __restrict__ xrow, int lane, float4 (&v)[6]) {
#pragma unroll
  for (int i = 0; i < 6; ++i) v[i] = ld_cg_f4(xrow + 4 * (lane + 32 * i));
}
// normalise the row held in v (lane l: float4 at column 4 (l + 32 i)) and write it out
__device__ __forceinline__ void ln_core(const float4 (&v)[6], const float* __restrict__ w, const float* __restrict__ b,
                                        __nv_bfloat16* __restrict__ out, int lane, int r, uint8_t* sm_tiles, int tile_bytes) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 6; ++i) s += v[i].x + v[i].y + v[i].z + v[i].w;
  const float mean = warp_sum(s) * (1.0f / 768.0f);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const float a = v[i].x - mean, c = v[i].y - mean, d = v[i].z - mean, e = v[i].w - mean;
    q += a * a + c * c + d * d + e * e;
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / 768.0f) + 1e-5f);
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const float4 g = __ldg(reinterpret_cast<const float4*>(w) + lane + 32 * i);
    const float4 bb = __ldg(reinterpret_cast<const float4*>(b) + lane + 32 * i);
    uint2 pk;
    pk.x = pack2((v[i].x - mean) * rstd * g.x + bb.x, (v[i].y - mean) * rstd * g.y + bb.y);
    pk.y = pack2((v[i].z - mean) * rstd * g.z + bb.z, (v[i].w - mean) * rstd * g.w + bb.w);
    if (out != nullptr) reinterpret_cast<uint2*>(out)[lane + 32 * i] = pk;
    if (sm_tiles != nullptr) {
      // columns 4 (lane + 32 i) .. + 3 of row r in the UMMA operand layout: k-block 2 i + lane / 16, 16-byte chunk
      // (lane % 16) / 2 XOR-swizzled with the row (128B swizzle), low or high half of the chunk
      const int kb = 2 * i + (lane >> 4), chunk = (lane & 15) >> 1;
      *reinterpret_cast<uint2*>(sm_tiles + (size_t)kb * tile_bytes + r * 128 + ((chunk ^ (r & 7)) << 4) + (lane & 1) * 8) = pk;
    }
  }
}
// LayerNorm of rows cw, cw + 8, ... of x straight into the CTA's activation tiles; the next row's loads are in flight while a
// row is normalised (a row is one L2 round trip + two warp reductions: serialised, four rows cost 5.8 us -- measured)
__device__ __noinline__ void ln_rows_to_tiles(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                                                 int R, int cw, int lane, uint8_t* sm_tiles, int tile_bytes) {
  float4 nxt[6], cur[6];
  if (cw < R) ln_load(x + (long long)cw * gD, lane, nxt);
  for (int r = cw; r < R; r += FT_COMPUTE_WARPS) {
#pragma unroll
    for (int i = 0; i < 6; ++i) cur[i] = nxt[i];
    if (r + FT_COMPUTE_WARPS < R) ln_load(x + (long long)(r + FT_COMPUTE_WARPS) * gD, lane, nxt);
    ln_core(cur, w, b, nullptr, lane, r, sm_tiles, tile_bytes);
  }
}

// One row, all eight compute warps (the LayerNorm PHASES: LN1r, LNFr -- at most one row per CTA up to 148 rows):
//   x[r] += (p0 + p1 + p2 + p3) + fc2_b  when `part` is given (the previous fc2's split-K partial tiles, summed in split order),
//   then h[r] = LayerNorm(x[r]) in bf16.  Warp w owns columns [96 w, 96 w + 96): 24 lanes x one float4, so the whole row is ONE
//   L2 round trip; the two row statistics meet in shared memory (scratch: 16 floats).  Called by all 256 compute threads.
__device__ __noinline__ void ln_row_cta(float* __restrict__ xrow, const float* __restrict__ w, const float* __restrict__ b,
                                           __nv_bfloat16* __restrict__ out, int cw, int lane, const float* __restrict__ part, int r,
                                           int R_pad, const float* __restrict__ fc2_b, float* scratch) {
  const bool act = lane < 24;
  const int c = 96 * cw + 4 * lane;  // first of this lane's four columns
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (act) {
    v = ld_cg_f4(xrow + c);
    if (part != nullptr) {
      const int tile = c >> 7, nl = c & 127;
      float4 pp[kFusedFc2Splits];
#pragma unroll
      for (int sp = 0; sp < kFusedFc2Splits; ++sp) pp[sp] = ld_cg_f4(part + ((long long)(sp * 6 + tile) * R_pad + r) * 128 + nl);
      float4 a = pp[0];
#pragma unroll
      for (int sp = 1; sp < kFusedFc2Splits; ++sp) { a.x += pp[sp].x; a.y += pp[sp].y; a.z += pp[sp].z; a.w += pp[sp].w; }
      const float4 bb = __ldg(reinterpret_cast<const float4*>(fc2_b + c));
      v.x += a.x + bb.x; v.y += a.y + bb.y; v.z += a.z + bb.z; v.w += a.w + bb.w;
      *reinterpret_cast<float4*>(xrow + c) = v;
    }
  }
  float s = warp_sum(act ? v.x + v.y + v.z + v.w : 0.f);
  if (lane == 0) scratch[cw] = s;
  asm volatile("bar.sync 1, 256;" ::: "memory");
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += scratch[i];  // same order in every thread
  const float mean = tot * (1.0f / 768.0f);
  const float d0 = v.x - mean, d1 = v.y - mean, d2 = v.z - mean, d3 = v.w - mean;
  float q = warp_sum(act ? d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3 : 0.f);
  if (lane == 0) scratch[8 + cw] = q;
  asm volatile("bar.sync 1, 256;" ::: "memory");
  float qt = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) qt += scratch[8 + i];
  const float rstd = rsqrtf(qt * (1.0f / 768.0f) + 1e-5f);
  if (act) {
    const float4 g = __ldg(reinterpret_cast<const float4*>(w + c));
    const float4 bb = __ldg(reinterpret_cast<const float4*>(b + c));
    uint2 pk;
    pk.x = pack2(d0 * rstd * g.x + bb.x, d1 * rstd * g.y + bb.y);
    pk.y = pack2(d2 * rstd * g.z + bb.z, d3 * rstd * g.w + bb.w);
    *reinterpret_cast<uint2*>(out + c) = pk;
  }
  asm volatile("bar.sync 1, 256;" ::: "memory");  // scratch is free for the next row
}

// KV-cache attention of one (row, head), head_dim 192, by the EIGHT compute warps of a CTA (DeCap: T <= 32 positions):
// warp w scores keys w, w + 8, w + 16, w + 24 (all its key / value rows are requested at once: one L2 round trip), keeps a
// local (max, sum, weighted value sum) and the warps meet in shared memory.  Appends this position's key / value to the cache.
// scratch: 8 x (192 + 2) floats.  Must be called by all 256 compute threads (named barrier 1).
__device__ __noinline__ void attn_192_cta(const __nv_bfloat16* __restrict__ qkv_row, __nv_bfloat16* __restrict__ kbase,
                                             __nv_bfloat16* __restrict__ vbase, __nv_bfloat16* __restrict__ out, int h, int H, int t,
                                             int cw, int lane, int ct, float* scratch) {
  constexpr int HDIM = 192, NK = 4;
  const float scale = rsqrtf(192.0f);
  const bool act = lane < HDIM / 8;
  const __nv_bfloat16* row = qkv_row + h * HDIM + lane * 8;
  const uint4 zero = make_uint4(0, 0, 0, 0);
  uint4 qu = zero, knew = zero, vnew = zero, ku[NK], vu[NK];
#pragma unroll
  for (int u = 0; u < NK; ++u) {
    const int j = cw + 8 * u;
    const bool ld = act && j < t;
    ku[u] = ld ? ld_cg16(kbase + (long long)j * HDIM + lane * 8) : zero;
    vu[u] = ld ? ld_cg16(vbase + (long long)j * HDIM + lane * 8) : zero;
  }
  if (act) {
    qu = ld_cg16(row);
    if ((t & 7) == cw) {  // the warp that owns key t takes it from the qkv row and appends it to the cache
      knew = ld_cg16(row + H * HDIM);
      vnew = ld_cg16(row + 2 * H * HDIM);
      *reinterpret_cast<uint4*>(kbase + (long long)t * HDIM + lane * 8) = knew;
      *reinterpret_cast<uint4*>(vbase + (long long)t * HDIM + lane * 8) = vnew;
    }
  }
  float q[8];
  unpack8(qu, q);
  float sc[NK];
#pragma unroll
  for (int u = 0; u < NK; ++u) {
    const int j = cw + 8 * u;
    if (j == t) { ku[u] = knew; vu[u] = vnew; }
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
    for (int u = 0; u < NK; ++u) sc[u] += __shfl_xor_sync(0xffffffffu, sc[u], o);
  }
  float mw = -INFINITY;
#pragma unroll
  for (int u = 0; u < NK; ++u) {
    sc[u] = (cw + 8 * u <= t) ? sc[u] * scale : -INFINITY;
    mw = fmaxf(mw, sc[u]);
  }
  float lw = 0.f, acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int u = 0; u < NK; ++u) {
    const float pj = (cw + 8 * u <= t) ? __expf(sc[u] - mw) : 0.f;  // mw = -inf only when this warp has no key at all
    lw += pj;
    float vf[8];
    unpack8(vu[u], vf);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = fmaf(pj, vf[e], acc[e]);
  }
  float* sm = scratch + cw * (HDIM + 2);
  if (act) {
#pragma unroll
    for (int e = 0; e < 8; ++e) sm[lane * 8 + e] = acc[e];
  }
  if (lane == 0) { sm[HDIM] = mw; sm[HDIM + 1] = lw; }
  asm volatile("bar.sync 1, 256;" ::: "memory");
  if (ct < HDIM) {
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < 8; ++w) M = fmaxf(M, scratch[w * (HDIM + 2) + HDIM]);
    float den = 0.f, num = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      const float mwv = scratch[w * (HDIM + 2) + HDIM];
      const float f = (mwv == -INFINITY) ? 0.f : __expf(mwv - M);
      den = fmaf(f, scratch[w * (HDIM + 2) + HDIM + 1], den);
      num = fmaf(f, scratch[w * (HDIM + 2) + ct], num);
    }
    out[h * HDIM + ct] = __float2bfloat16(num / den);
  }
  asm volatile("bar.sync 1, 256;" ::: "memory");  // scratch is free for the next unit
}

// KV-cache attention of one (row, head), head_dim 192 (DeCap: 4 heads, T <= 32): the arithmetic of
// decode_attention_bf16_kernel (attention.cu) with coherent loads.  Appends this position's key / value to the cache.
__device__ __noinline__ void attn_192(const __nv_bfloat16* __restrict__ qkv_row, __nv_bfloat16* __restrict__ kbase,
                                         __nv_bfloat16* __restrict__ vbase, __nv_bfloat16* __restrict__ out, int h, int H, int t,
                                         int lane) {
  constexpr int HDIM = 192, KB = 8;
  const float scale = rsqrtf(192.0f);
  const bool act = lane < HDIM / 8;
  const __nv_bfloat16* row = qkv_row + h * HDIM + lane * 8;
  kbase += lane * 8;
  vbase += lane * 8;
  float q[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const uint4 zero = make_uint4(0, 0, 0, 0);
  uint4 knew = zero, vnew = zero;
  if (act) {
    const uint4 qu = ld_cg16(row);
    knew = ld_cg16(row + H * HDIM);
    vnew = ld_cg16(row + 2 * H * HDIM);
    unpack8(qu, q);
    *reinterpret_cast<uint4*>(kbase + (long long)t * HDIM) = knew;
    *reinterpret_cast<uint4*>(vbase + (long long)t * HDIM) = vnew;
  }
  const int tl = t > 0 ? t - 1 : 0;
  float my = -INFINITY;
  for (int j0 = 0; j0 <= t; j0 += KB) {
    uint4 ku[KB];
#pragma unroll
    for (int u = 0; u < KB; ++u) ku[u] = act ? ld_cg16(kbase + (long long)min(j0 + u, tl) * HDIM) : zero;
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
    for (int u = 0; u < KB; ++u) vu[u] = act ? ld_cg16(vbase + (long long)min(j0 + u, tl) * HDIM) : zero;
#pragma unroll
    for (int u = 0; u < KB; ++u) {
      const int j = j0 + u;
      const uint4 vv = (j < t) ? vu[u] : (j == t ? vnew : zero);
      const float pj = __shfl_sync(0xffffffffu, p, j & 31);
      if (j <= t) {
        float vf[8];
        unpack8(vv, vf);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = fmaf(pj, vf[e], acc[e]);
      }
    }
  }
  if (act) {
    uint4 o;
    o.x = pack2(acc[0], acc[1]); o.y = pack2(acc[2], acc[3]); o.z = pack2(acc[4], acc[5]); o.w = pack2(acc[6], acc[7]);
    *reinterpret_cast<uint4*>(out + h * HDIM + lane * 8) = o;
  }
}

// head_dim 64 (GPT-2 small: 12 heads, T <= 128): lane j scores keys j, j+32, j+64, j+96 (the arithmetic of
// decode_attention_long_kernel); qs = 64 floats of this warp's shared memory
__device__ __noinline__ void attn_64(const __nv_bfloat16* __restrict__ qkv_row, __nv_bfloat16* __restrict__ kbase,
                                        __nv_bfloat16* __restrict__ vbase, __nv_bfloat16* __restrict__ out, int h, int H, int t,
                                        int lane, float* qs) {
  constexpr int HDIM = 64, MAXC = 4;
  const __nv_bfloat16* row = qkv_row + h * HDIM;
  {
    const uint32_t qq = ld_cg_u32(row + lane * 2);
    const float2 q2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&qq));
    const uint32_t k2 = ld_cg_u32(row + H * HDIM + lane * 2), v2 = ld_cg_u32(row + 2 * H * HDIM + lane * 2);
    qs[lane * 2] = q2.x;
    qs[lane * 2 + 1] = q2.y;
    *reinterpret_cast<uint32_t*>(kbase + (long long)t * HDIM + lane * 2) = k2;
    *reinterpret_cast<uint32_t*>(vbase + (long long)t * HDIM + lane * 2) = v2;
  }
  __threadfence_block();
  __syncwarp();  // q and the appended row are visible to the whole warp
  float sc[MAXC];
  float mx = -INFINITY;
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    const int j = lane + 32 * c;
    sc[c] = -INFINITY;
    if (j <= t) {
      const __nv_bfloat16* kr = kbase + (long long)j * HDIM;
      float a = 0.f;
#pragma unroll
      for (int d = 0; d < HDIM; d += 8) {
        float kf[8];
        unpack8(ld_cg16(kr + d), kf);
#pragma unroll
        for (int e = 0; e < 8; ++e) a = fmaf(qs[d + e], kf[e], a);
      }
      sc[c] = a * 0.125f;
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
    for (int jj = 0; jj < n; ++jj) {
      const float pj = __shfl_sync(0xffffffffu, sc[c], jj);
      const uint32_t vv = ld_cg_u32(vbase + (long long)(32 * c + jj) * HDIM + lane * 2);
      const float2 v2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&vv));
      a0 = fmaf(pj, v2.x, a0);
      a1 = fmaf(pj, v2.y, a1);
    }
  }
  *reinterpret_cast<uint32_t*>(out + h * HDIM + lane * 2) = pack2(a0 * inv, a1 * inv);
  __syncwarp();  // qs is reused by this warp's next unit
}

// ------------------------------------------------------------------------------------------ schedule (same arithmetic in every role)
struct Sched {
  int G, R, H, cta, PPS;
  const signed char* pt;
  __device__ __forceinline__ int ptype(int p) const { return pt[p]; }
  __device__ __forceinline__ int units(int t) const {
    switch (t) {
      case P_QKV: return 18;
      case P_PROJ: return 6;
      case P_FC: return 24;
      case P_FC2: return 6 * kFusedFc2Splits;
      case P_LMHEAD: return (gV + 127) / 128;
      case P_ATTN: return R * H;
      default: return R;  // LN1, LNF, PICK: one warp per row
    }
  }
  __device__ __forceinline__ bool is_gemm(int t) const { return t == P_QKV || t == P_PROJ || t == P_FC || t == P_FC2 || t == P_LMHEAD; }
  __device__ __forceinline__ int rot(int gp) const { return (int)(((long long)gp * 41) % G); }  // 41: coprime with 148
  // first unit of phase gp owned by this CTA; its further units follow at stride G
  __device__ __forceinline__ int first_unit(int gp) const { return (cta - rot(gp) + G) % G; }
  __device__ __forceinline__ int participants(int t) const { return min(G, units(t)); }
};

struct Waiter {
  int* abort;
  bool dead;
  __device__ __forceinline__ void fail() {
    dead = true;
    atomicExch(abort, 1);
  }
  // bounded mbarrier wait.  NOT inlined, like every larger piece of this kernel: the persistent kernel runs each code path
  // once per phase, i.e. instruction-cache cold -- at 222 KB of SASS (everything inlined) instruction fetch, not data, set the
  // pace of every phase (profiles/r02l, r02n: 48 UMMAs took 5 us to ISSUE)
  __device__ __noinline__ void mbar(uint32_t bar, uint32_t parity) {
    if (dead) return;
    uint32_t done = 0;
    unsigned long long t0 = 0;
    for (unsigned it = 0;; ++it) {
      asm volatile(
          "{\n\t"
          ".reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t"
          "}"
          : "=r"(done)
          : "r"(bar), "r"(parity)
          : "memory");
      if (done) return;
      if ((it & 255) == 255) {
        const unsigned long long now = gtime_ns();
        if (t0 == 0) t0 = now;
        if (now - t0 > 2000000000ull || ld_acquire(abort) != 0) { fail(); return; }
      }
    }
  }
  // fast path inline (one try_wait; a call into the bounded loop costs ~100 ns even when the barrier has completed)
  __device__ __forceinline__ void mbar_fast(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done) mbar(bar, parity);
  }
  // bounded wait for a phase counter (acquire loads: the data loads that follow are ordered after the one that succeeds)
  __device__ __noinline__ void phase(const int* cnt, int target) {
    if (dead) return;
    unsigned long long t0 = 0;
    for (unsigned it = 0;; ++it) {
      if (ld_acquire(cnt) >= target) return;
      if ((it & 63) == 63) {
        const unsigned long long now = gtime_ns();
        if (t0 == 0) t0 = now;
        if (now - t0 > 2000000000ull || ld_acquire(abort) != 0) { fail(); return; }
      }
    }
  }
};

// generic-proxy writes to shared memory that the tensor core (async proxy) will read
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// rows [0, R) x columns [k0 + 64 kb0, k0 + 64 (kb0 + nkb)) of a bf16 matrix -> k-blocks [kb0, kb0 + nkb) of the CTA's activation
// tiles (12 k-blocks of [R_pad x 128 B], 128B swizzle), by all 256 compute threads: 16-byte chunks, every thread's loads issued
// before its first store (one L2 round trip per batch of 6: half a unit at 32 rows)
__device__ __noinline__ void load_acts(const __nv_bfloat16* __restrict__ src, int ld, int k0, int R, uint8_t* sm_tiles, int tile_bytes,
                                       int ct, int kb0, int nkb) {
  const int per_row = nkb * 8;  // chunks of 8 bf16 per row in this k range
  const int total = R * per_row;
  constexpr int NB = 6;
  for (int q0 = ct; q0 < total; q0 += 256 * NB) {
    uint4 v[NB];
#pragma unroll
    for (int i = 0; i < NB; ++i) {
      const int q = q0 + 256 * i;
      if (q < total) {
        const int r = q / per_row, cc = q - r * per_row;
        v[i] = ld_cg16(src + (long long)r * ld + k0 + (kb0 * 8 + cc) * 8);
      }
    }
#pragma unroll
    for (int i = 0; i < NB; ++i) {
      const int q = q0 + 256 * i;
      if (q < total) {
        const int r = q / per_row, cc = q - r * per_row, kb = kb0 + (cc >> 3), c = cc & 7;
        *reinterpret_cast<uint4*>(sm_tiles + (size_t)kb * tile_bytes + r * 128 + ((c ^ (r & 7)) << 4)) = v[i];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ the kernel
__global__ void __launch_bounds__(FT_THREADS, 1) decode_fused_kernel(const FusedParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ __align__(8) uint64_t bars[2 * MAX_STAGES + 6];
  __shared__ uint32_t tmem_slot_var;
  __shared__ int s_dead;

  const uint32_t bar_base = smem_u32(bars);
  auto fullW = [&](int s) { return bar_base + 8u * s; };
  auto emptyW = [&](int s) { return bar_base + 8u * (MAX_STAGES + s); };
  auto tfull = [&](int s) { return bar_base + 8u * (2 * MAX_STAGES + s); };
  auto tempty = [&](int s) { return bar_base + 8u * (2 * MAX_STAGES + 2 + s); };
  // the activation tiles of a GEMM phase are in shared memory: k-blocks 0..5 / 6..11 (the UMMAs of the first half run while
  // the second half is still being fetched)
  auto acts_full = [&](int half) { return bar_base + 8u * (2 * MAX_STAGES + 4 + half); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int R = P.R, R_pad = P.R_pad, L = P.L, G = P.G, PPS = P.PPS;
  const int a_tile_bytes = R_pad * 128;
  // dynamic shared memory: weight ring | 12 activation tiles | lm-head running (max, first index) per TMEM lane quarter and
  // row [4][R_pad] x 2 | attention scratch (8 x 194 floats)
  uint8_t* dyn = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t smemA = smem_base + P.nsw * W_TILE_BYTES;
  uint8_t* acts = dyn + P.nsw * W_TILE_BYTES;
  uint8_t* gen = acts + KB_PER_UNIT * a_tile_bytes;
  float* s_bv = reinterpret_cast<float*>(gen);  // [q * R_pad + r]
  int* s_bi = reinterpret_cast<int*>(gen + 16 * R_pad);
  float* s_att = reinterpret_cast<float*>(gen + 32 * R_pad);
  const int gp_begin = P.first_gp, gp_end = min(P.steps * PPS, P.stop_gp);  // [gp_begin, gp_end)
  Sched S{G, R, P.H, (int)blockIdx.x, PPS, P.ptype};
  Waiter wt{P.abort, false};
  // A unit's 48 UMMAs would form ONE dependent chain on a single accumulator: at N = R_pad <= 64 that chain is latency bound
  // (measured 3.8 us per unit at N = 32, ~150 clk per instruction instead of the 16 clk the data path needs).  Hence NACC
  // independent accumulators -- k-step j of every k-block adds into accumulator j -- summed by the epilogue in a fixed order.
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < 2 * NACC * R_pad) tmem_cols <<= 1;

  if (warp == 0 && lane == 0) {
    s_dead = 0;
    for (int s = 0; s < P.nsw; ++s) { mbar_init(fullW(s), 1); mbar_init(emptyW(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull(s), 1); mbar_init(tempty(s), FT_COMPUTE_WARPS); }
    mbar_init(acts_full(0), 1);
    mbar_init(acts_full(1), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) tmem_alloc(smem_u32(&tmem_slot_var), tmem_cols);
  for (int i = threadIdx.x; i < 4 * R_pad; i += FT_THREADS) {
    s_bv[i] = -INFINITY;
    s_bi[i] = 0x7fffffff;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&tmem_slot_var);
  unsigned long long* tl_tail = P.timeline ? P.timeline + (long long)(P.steps * PPS) * G * 16 : nullptr;  // SM clock vs wall clock
  if (tl_tail && blockIdx.x == 0 && threadIdx.x == 0) { tl_tail[0] = clock64(); tl_tail[1] = gtime_ns(); }

  if (warp == 0) {
    // ------------------------------------------------------------------ weight producer: runs ahead of every phase
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int gp = gp_begin; gp < gp_end && !wt.dead; ++gp) {
        const int p = gp % PPS, pt = S.ptype(p);
        if (!S.is_gemm(pt)) continue;
        const int l = P.player[p], nu = S.units(pt);
        const CUtensorMap* map = P.wmaps + (pt == P_LMHEAD ? 4 * L : 4 * l + (pt == P_QKV ? 0 : pt == P_PROJ ? 1 : pt == P_FC ? 2 : 3));
        for (int u = S.first_unit(gp); u < nu; u += G) {
          const int tile = pt == P_FC2 ? u % 6 : u, kb0 = pt == P_FC2 ? (u / 6) * KB_PER_UNIT : 0;
          // the ring is filled and released in GROUPS of gs stages: one (empty, full) barrier pair per group, so that the MMA
          // warp pays one wait and one commit per group instead of per stage (each costs it 100-200 ns, more than the UMMAs
          // of a k-block at N <= 64)
          for (int kb = 0; kb < KB_PER_UNIT; ++kb) {
            const int grp = stage / P.gs;
            if (stage % P.gs == 0) {
              wt.mbar(emptyW(grp), phase ^ 1);
              if (wt.dead) break;
              mbar_expect_tx(fullW(grp), P.gs * W_TILE_BYTES);
            }
            tma_load_2d(smem_base + stage * W_TILE_BYTES, map, fullW(grp), (kb0 + kb) * 64, tile * 128);
            if (++stage == P.nsw) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 2) {
    // ------------------------------------------------------------------ MMA issuer
    // The whole warp walks the schedule with warp-uniform values (so that descriptors live in uniform registers) and ONE elected
    // lane issues.  Instruction count per UMMA matters here: a single warp issues ~1 dependent instruction per 5 ns, and a
    // UMMA at N <= 64 is short -- with ~60 instructions of address arithmetic per UMMA (runtime modulo, per-lane election
    // loops) issuing the 48 UMMAs of a unit took 5-8 us (profiles/r02l, r02o).  Descriptors advance by constant strides.
    const uint32_t idesc = make_idesc(128, R_pad);
    uint32_t leader;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
    const uint64_t adesc0 = make_smem_desc(smem_base), bdesc0 = make_smem_desc(smemA);
    const uint32_t b_stride = (uint32_t)a_tile_bytes >> 4;  // descriptor start-address units (16 B) per activation k-block
    int sw = 0, it = 0, nphase = 0;
    uint32_t phw = 0;
    for (int gp = gp_begin; gp < gp_end && !wt.dead; ++gp) {
      const int p = gp % PPS, pt = S.ptype(p);
      if (!S.is_gemm(pt)) continue;
      const int nu = S.units(pt), u0 = S.first_unit(gp);
      if (u0 >= nu) continue;
      const uint32_t acts_parity = nphase & 1;  // activation tiles are placed once per phase and kept for all its units
      ++nphase;
      bool acts_seen[2] = {false, false};
      unsigned long long* tl = (P.timeline && lane == 0) ? P.timeline + ((long long)(gp - gp_begin) * G + blockIdx.x) * 16 : nullptr;
      wt.mbar_fast(acts_full(0), acts_parity);
      acts_seen[0] = true;
      if (tl) tl[6] = gtime_ns();
      for (int u = u0; u < nu && !wt.dead; u += G, ++it) {
        const int as = it & 1;
        wt.mbar_fast(tempty(as), ((it >> 1) & 1) ^ 1);  // the epilogue has drained this accumulator
        const uint32_t tmem_d = tmem_base + as * (NACC * R_pad);
        uint64_t bdesc = bdesc0;
        for (int kb0 = 0; kb0 < KB_PER_UNIT; kb0 += P.gs) {  // gs divides 12 and nsw: a group never wraps
          const int half_ = kb0 >= KB_PER_UNIT / 2 ? 1 : 0;  // groups never straddle the halves (gs divides 6 or is 12 -> both)
          if (!acts_seen[half_]) { wt.mbar_fast(acts_full(half_), acts_parity); acts_seen[half_] = true; }
          if (kb0 + P.gs > KB_PER_UNIT / 2 && !acts_seen[1]) { wt.mbar_fast(acts_full(1), acts_parity); acts_seen[1] = true; }
          wt.mbar_fast(fullW(sw / P.gs), phw);
          if (wt.dead) break;
          tc_fence_after();
          uint64_t adesc = adesc0 + (uint64_t)(sw * (W_TILE_BYTES >> 4));
          const uint32_t acc = kb0 != 0;
#pragma unroll 1
          for (int j = 0; j < P.gs; ++j) {
            if (leader) {
              umma_f16(tmem_d, adesc, bdesc, idesc, acc | (uint32_t)j);
              umma_f16(tmem_d + R_pad, adesc + 2, bdesc + 2, idesc, acc | (uint32_t)j);
              umma_f16(tmem_d + 2 * R_pad, adesc + 4, bdesc + 4, idesc, acc | (uint32_t)j);
              umma_f16(tmem_d + 3 * R_pad, adesc + 6, bdesc + 6, idesc, acc | (uint32_t)j);
            }
            adesc += W_TILE_BYTES >> 4;
            bdesc += b_stride;
          }
          if (leader) {
            umma_commit(emptyW(sw / P.gs));  // the group of ring stages just consumed
            if (kb0 + P.gs >= KB_PER_UNIT) umma_commit(tfull(as));
          }
          __syncwarp();
          sw += P.gs;
          if (sw >= P.nsw) { sw = 0; phw ^= 1; }
        }
        if (tl && u == u0) tl[7] = gtime_ns();
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ compute warps
    const int cw = warp - 4, quarter = warp & 3, half = cw >> 2;
    const int ct = threadIdx.x - 128;  // 0..255 among the compute threads
    const int n_chunks = R_pad / 16;
    int it = 0;
    auto stamp = [&](int gp, int k) {
      if (P.timeline && ct == 0) P.timeline[((long long)(gp - gp_begin) * G + blockIdx.x) * 16 + k] = gtime_ns();
    };
    // the CTA's arrival at a phase: its eight compute warps meet, then ONE thread publishes -- the barrier orders the other
    // threads' writes before that thread's gpu-scope release (cumulativity; the pattern of a cooperative-groups grid sync)
    auto cta_arrive = [&](int gp) {
      stamp(gp, 2);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (ct == 0) red_release_add(P.phase_cnt + gp, 1);
      stamp(gp, 3);
    };
    // NB no early exit from this loop: the CTA-wide named barriers need all eight warps.  After a time-out (wt.dead, raised
    // CTA-wide through the abort flag) a warp keeps walking the schedule but skips every wait and all work.
    for (int gp = gp_begin; gp < gp_end; ++gp) {
      const int s = gp / PPS, p = gp % PPS, pt = S.ptype(p), l = P.player[p];
      const int nu = S.units(pt), u0 = S.first_unit(gp);
      if (u0 >= nu) continue;  // no work for this CTA in this phase (CTA-uniform)
      const int pos = P.pos_base + s;
      stamp(gp, 0);
      if (gp > gp_begin) {  // the previous phase -- the producer of this phase's input -- is complete
        // ONE polling thread per CTA: with every warp polling, ~1000 pollers hammered one L2 line and a poll round trip grew to
        // ~0.6 us (it also delayed the arrivals queued behind them)
        if (ct == 0) {
          wt.phase(P.phase_cnt + gp - 1, S.participants(S.ptype((gp - 1) % PPS)));
          s_dead = wt.dead ? 1 : 0;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        wt.dead = wt.dead || (*(volatile int*)&s_dead != 0);
      }
      stamp(gp, 1);
      if (!S.is_gemm(pt)) {
        if (pt == P_ATTN && P.hd == 192 && nu <= 2 * G) {
          // few (row, head) units: all eight warps work on one unit at a time (one L2 round trip instead of a chain of them)
          for (int u = u0; u < nu; u += G) {
            const int r = u / P.H, hh = u % P.H;
            __nv_bfloat16* kb_ = P.kc + (long long)l * P.kv_layer + ((long long)(r * P.H + hh) * P.T) * P.hd;
            __nv_bfloat16* vb_ = P.vc + (long long)l * P.kv_layer + ((long long)(r * P.H + hh) * P.T) * P.hd;
            if (!wt.dead) attn_192_cta(P.qkv + (long long)r * 3 * gD, kb_, vb_, P.att + (long long)r * gD, hh, P.H, pos, cw, lane, ct, s_att);
          }
          cta_arrive(gp);
          continue;
        }
        if (pt == P_LN1 || pt == P_LNF) {
          // the fc2 of the block before left its split-K partials unreduced (always for LN1: blocks 1..; for ln_f unless the
          // kernel starts there); one row at a time, all eight warps on it
          const bool pending = pt == P_LN1 || gp > gp_begin;
          const float* fb = pending ? P.layers[pt == P_LN1 ? l - 1 : L - 1].fc2_b : nullptr;
          for (int u = u0; u < nu; u += G)
            if (!wt.dead)
              ln_row_cta(P.x + (long long)u * gD, pt == P_LN1 ? P.layers[l].ln1_w : P.lnf_w, pt == P_LN1 ? P.layers[l].ln1_b : P.lnf_b,
                         P.h + (long long)u * gD, cw, lane, pending ? P.part : nullptr, u, R_pad, fb, s_att);
          cta_arrive(gp);
          continue;
        }
        for (int u = u0 + cw * G; u < nu && !wt.dead; u += FT_COMPUTE_WARPS * G) {  // this CTA's units, one warp each
          if (pt == P_ATTN) {
            const int r = u / P.H, hh = u % P.H;
            __nv_bfloat16* kb_ = P.kc + (long long)l * P.kv_layer + ((long long)(r * P.H + hh) * P.T) * P.hd;
            __nv_bfloat16* vb_ = P.vc + (long long)l * P.kv_layer + ((long long)(r * P.H + hh) * P.T) * P.hd;
            if (P.hd == 192) attn_192(P.qkv + (long long)r * 3 * gD, kb_, vb_, P.att + (long long)r * gD, hh, P.H, pos, lane);
            else attn_64(P.qkv + (long long)r * 3 * gD, kb_, vb_, P.att + (long long)r * gD, hh, P.H, pos, lane, s_att + cw * 64);
          } else {  // P_PICK: final arg-max over the per-CTA partials, id out, next token's embedding
            float best = -INFINITY;
            int bi = 0x7fffffff;
            for (int c = lane; c < G; c += 32) {
              const float v = ld_cg_f(P.pm_val + (long long)c * R_pad + u);
              const int i = ld_cg_i(P.pm_idx + (long long)c * R_pad + u);
              if (v > best || (v == best && i < bi)) { best = v; bi = i; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              const float ov = __shfl_xor_sync(0xffffffffu, best, o);
              const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
              if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
            }
            const int tok = (bi == 0x7fffffff) ? 0 : bi;  // all-NaN row -> 0, like torch.argmax
            if (lane == 0) P.out_ids[(long long)u * P.ids_ld + s] = tok;
            if (s + 1 < P.steps) {
              const float4* a = reinterpret_cast<const float4*>(P.wte32 + (long long)min(max(tok, 0), gV - 1) * gD);
              const float4* b = reinterpret_cast<const float4*>(P.wpe + (long long)(pos + 1) * gD);
              float4* o = reinterpret_cast<float4*>(P.x + (long long)u * gD);
#pragma unroll
              for (int i = 0; i < 6; ++i) {
                const float4 uu = __ldg(a + lane + 32 * i), vv = __ldg(b + lane + 32 * i);
                o[lane + 32 * i] = make_float4(uu.x + vv.x, uu.y + vv.y, uu.z + vv.z, uu.w + vv.w);
              }
            }
          }
        }
        cta_arrive(gp);
        continue;
      }
      // ---- GEMM phase.  (1) this CTA's activation tiles: every MMA that read the previous contents has completed (the epilogue
      // of the CTA's last unit waited for its accumulator), so the tiles can be overwritten right away
      const FusedLayer& ly = P.layers[pt == P_LMHEAD ? 0 : l];
      if (!wt.dead) {
        if (pt == P_QKV && l == 0) {         // LN1 of block 0 applied on the way in (x is complete: prefix embedding / PICK)
          ln_rows_to_tiles(P.x, ly.ln1_w, ly.ln1_b, R, cw, lane, acts, a_tile_bytes);
        } else if (pt == P_FC) {             // LN2 applied on the way in (x is complete after PROJ)
          ln_rows_to_tiles(P.x, ly.ln2_w, ly.ln2_b, R, cw, lane, acts, a_tile_bytes);
        }
      }
      const bool fused_ln = (pt == P_QKV && l == 0) || pt == P_FC;
      if (fused_ln) {
        fence_proxy_async_smem();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (ct == 0) { mbar_arrive(acts_full(0)); mbar_arrive(acts_full(1)); }
      } else {
        // plain rows (attention output, gelu rows of this CTA's K slice, LayerNorm rows): two halves of six k-blocks
        const __nv_bfloat16* src = pt == P_PROJ ? P.att : (pt == P_FC2 ? P.f : P.h);
        const int ld = pt == P_FC2 ? gFF : gD, k0 = pt == P_FC2 ? (u0 / 6) * (KB_PER_UNIT * 64) : 0;
#pragma unroll 1
        for (int half_ = 0; half_ < 2; ++half_) {
          if (!wt.dead) load_acts(src, ld, k0, R, acts, a_tile_bytes, ct, half_ * (KB_PER_UNIT / 2), KB_PER_UNIT / 2);
          fence_proxy_async_smem();
          asm volatile("bar.sync 1, 256;" ::: "memory");
          if (ct == 0) mbar_arrive(acts_full(half_));
        }
      }
      stamp(gp, 4);
      // (2) epilogue of every owned unit
      for (int u = u0; u < nu; u += G, ++it) {
        const int as = it & 1;
        const int tile = pt == P_FC2 ? u % 6 : u;
        const int nl = quarter * 32 + lane, n = tile * 128 + nl;  // this thread's output feature
        wt.mbar(tfull(as), (it >> 1) & 1);
        wt.dead = __any_sync(0xffffffffu, wt.dead);
        if (u == u0) stamp(gp, 5);
        tc_fence_after();
        const uint32_t tacc = tmem_base + as * (NACC * R_pad) + ((uint32_t)(quarter * 32) << 16);
        float bias = 0.f;
        if (pt == P_QKV) bias = __ldg(ly.attn_b + n);
        else if (pt == P_PROJ) bias = __ldg(ly.proj_b + n);
        else if (pt == P_FC) bias = __ldg(ly.fc_b + n);
        for (int c = half; c < n_chunks && !wt.dead; c += 2) {
          uint32_t rr[16];
          {
            uint32_t ra[2][16];  // two accumulators at a time (register budget); fixed order: ((a0 + a1) + a2) + a3
            tmem_ld16(tacc + c * 16, ra[0]);
            tmem_ld16(tacc + R_pad + c * 16, ra[1]);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) rr[j] = __float_as_uint(__uint_as_float(ra[0][j]) + __uint_as_float(ra[1][j]));
            tmem_ld16(tacc + 2 * R_pad + c * 16, ra[0]);
            tmem_ld16(tacc + 3 * R_pad + c * 16, ra[1]);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) rr[j] = __float_as_uint((__uint_as_float(rr[j]) + __uint_as_float(ra[0][j])) + __uint_as_float(ra[1][j]));
          }
          if (pt == P_QKV) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int r = c * 16 + j;
              if (r < R) P.qkv[(long long)r * (3 * gD) + n] = __float2bfloat16(__uint_as_float(rr[j]) + bias);
            }
          } else if (pt == P_PROJ) {
            float xv[16];  // all residual loads first (one round trip), then the stores
#pragma unroll
            for (int j = 0; j < 16; ++j) xv[j] = (c * 16 + j < R) ? ld_cg_f(P.x + (long long)(c * 16 + j) * gD + n) : 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int r = c * 16 + j;
              if (r < R) P.x[(long long)r * gD + n] = xv[j] + (__uint_as_float(rr[j]) + bias);
            }
          } else if (pt == P_FC) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int r = c * 16 + j;
              if (r < R) P.f[(long long)r * gFF + n] = __float2bfloat16(gelu_new_fast(__uint_as_float(rr[j]) + bias));
            }
          } else if (pt == P_FC2) {
            // the four split-K partial tiles stay unreduced: the next LayerNorm phase (LN1 of the following block or ln_f) sums
            // them in split order, adds bias and residual and owns the x update (ln_row)
            float* pp = P.part + ((long long)u * R_pad + c * 16) * 128 + nl;
#pragma unroll
            for (int j = 0; j < 16; ++j) pp[(long long)j * 128] = __uint_as_float(rr[j]);
          } else {  // P_LMHEAD: arg-max over the 32 vocabulary rows of this warp, per decode row
            const bool valid = n < gV;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float v = valid ? __uint_as_float(rr[j]) : -INFINITY;
              int idx = valid ? n : 0x7fffffff;
              if (!(v == v)) { v = -INFINITY; idx = 0x7fffffff; }  // NaN never wins (torch.argmax semantics handled in PICK)
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, v, o);
                const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
                if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
              }
              const int r = c * 16 + j;
              if (lane == (j & 31)) {  // (quarter, row) slots are owned by this warp: no race
                const float cv = s_bv[quarter * R_pad + r];
                const int ci = s_bi[quarter * R_pad + r];
                if (v > cv || (v == cv && idx < ci)) { s_bv[quarter * R_pad + r] = v; s_bi[quarter * R_pad + r] = idx; }
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty(as));
      }
      if (pt == P_LMHEAD) {
        // fold the four lane quarters and publish this CTA's partial (max, first index) per row; reset for the next step
        asm volatile("bar.sync 1, 256;" ::: "memory");
        for (int r = ct; r < R_pad; r += 256) {
          float v = s_bv[r];
          int i = s_bi[r];
#pragma unroll
          for (int q = 1; q < 4; ++q) {
            const float ov = s_bv[q * R_pad + r];
            const int oi = s_bi[q * R_pad + r];
            if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
          }
          P.pm_val[(long long)blockIdx.x * R_pad + r] = v;
          P.pm_idx[(long long)blockIdx.x * R_pad + r] = i;
#pragma unroll
          for (int q = 0; q < 4; ++q) { s_bv[q * R_pad + r] = -INFINITY; s_bi[q * R_pad + r] = 0x7fffffff; }
        }
      }
      cta_arrive(gp);
      if (pt == P_LMHEAD && P.l2_prefetch && s + 1 < P.steps) {
        // The lm-head is 57 % of a step's weight bytes and the only phase that streams more than the ring holds: its units
        // beyond the first are fetched during the phase, from HBM (13.5 us of the 20 us phase).  Ask the L2 for them now, a whole
        // step ahead -- the blocks' weights in between are prefetched into shared memory anyway, so what the L2 evicts of
        // THEM costs nothing.  One 16 KB tile per thread.
        const int gpn = gp + PPS, un0 = S.first_unit(gpn) + G;  // the ring takes this CTA's first unit of the next step
        const int t_ = ct / KB_PER_UNIT, kb_ = ct - t_ * KB_PER_UNIT, un = un0 + t_ * G;
        if (un < nu && ct < 4 * KB_PER_UNIT)
          asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(P.wmaps + 4 * L), "r"(kb_ * 64), "r"(un * 128)
                       : "memory");
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (tl_tail && blockIdx.x == 0 && threadIdx.x == 0) { tl_tail[2] = clock64(); tl_tail[3] = gtime_ns(); }
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------ host side
std::atomic<int> g_fused_max_rows{32}, g_fused_ctas{0};  // pio_set_decode_fused(); 0 CTAs = one per SM

bool decode_fused_eligible(const PioDecoder* h, int R, bool want_logprob) {
  const char* sw = getenv("PIO_DECODE_FUSED");  // read per call: tests and A/B runs flip it inside one process
  const bool on = !(sw && sw[0] == '0');
  if (!on || want_logprob || h->mode != PIO_BF16 || h->fused_wmaps == nullptr) return false;
  if (!((h->H == 4) || (h->H == 12)) || h->L > 12) return false;
  // Measured (tests/test_gpu_fused_decode.py::test_fused_decode_speed_report, B200): 163 / 184 us per step at 8 / 32 rows against
  // 213 / 224 for the kernel-per-op path, but 268 against 243 at 64 rows -- there the 12 activation tiles (96 KB) leave only half
  // a unit of weight ring and the fused LayerNorm does 8 rows per warp.  Default: up to 32 rows; PIO_DECODE_FUSED_MAX_ROWS raises
  // it (at most kFusedMaxRows) for experiments.
  int max_rows = std::min(kFusedMaxRows, g_fused_max_rows.load());
  if (const char* e = getenv("PIO_DECODE_FUSED_MAX_ROWS")) max_rows = std::min(kFusedMaxRows, atoi(e));
  return R >= 1 && R <= max_rows;
}

int decode_fused_build(PioDecoder* h, cudaStream_t st) {
  if (h->mode != PIO_BF16) return PIO_OK;
  const int L = h->L;
  std::vector<CUtensorMap> maps(4 * L + 1);
  std::vector<FusedLayer> layers(L);
  for (int i = 0; i < L; ++i) {
    const PioDecoder::Blk& b = h->blk[i];
    PIO_TRY(make_map_2d(&maps[4 * i + 0], b.attn_w, 3 * gD, gD, gD, 128, 64));
    PIO_TRY(make_map_2d(&maps[4 * i + 1], b.proj_w, gD, gD, gD, 128, 64));
    PIO_TRY(make_map_2d(&maps[4 * i + 2], b.fc_w, gFF, gD, gD, 128, 64));
    PIO_TRY(make_map_2d(&maps[4 * i + 3], b.fc2_w, gD, gFF, gFF, 128, 64));
    layers[i] = FusedLayer{b.ln1_w, b.ln1_b, b.attn_b, b.proj_b, b.ln2_w, b.ln2_b, b.fc_b, b.fc2_b};
  }
  PIO_TRY(make_map_2d(&maps[4 * L], h->wte, gV, gD, gD, 128, 64));
  void* dm = nullptr;
  void* dl = nullptr;
  PIO_CUDA(cudaMalloc(&dm, maps.size() * sizeof(CUtensorMap)));
  h->owned.push_back(dm);
  PIO_CUDA(cudaMalloc(&dl, layers.size() * sizeof(FusedLayer)));
  h->owned.push_back(dl);
  // synchronous copies: the host vectors die with this frame
  PIO_CUDA(cudaMemcpy(dm, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice));
  PIO_CUDA(cudaMemcpy(dl, layers.data(), layers.size() * sizeof(FusedLayer), cudaMemcpyHostToDevice));
  (void)st;
  h->fused_wmaps = dm;
  h->fused_layers = (FusedLayer*)dl;
  return PIO_OK;
}

int decode_fused(PioDecoder* h, const DecodeWs& w, int R, int T, int steps, int pos_base, bool start_at_pick, int* out_ids,
                 cudaStream_t st) {
  const int L = h->L, PPS = 6 * L + 2;
  const int R_pad = std::max(16, (R + 15) / 16 * 16);
  PIO_CHECK(R_pad <= kFusedMaxRows && steps >= 1 && L <= 12, "decode_fused: %d rows / %d steps / %d blocks outside the built range", R, steps, L);
  PIO_CHECK(T <= (h->H == 4 ? 32 : 128), "decode_fused: a cache of %d positions exceeds what the attention routines walk", T);
  int dev = 0, sms = 0;
  PIO_CUDA(cudaGetDevice(&dev));
  PIO_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  int G = std::min(sms, kFusedCtas);
  if (const int c = g_fused_ctas.load()) G = std::max(6 * kFusedFc2Splits, std::min(G, c));
  if (const char* e = getenv("PIO_DECODE_FUSED_CTAS")) G = std::max(1, std::min(G, atoi(e)));
  PIO_CHECK(G >= 6 * kFusedFc2Splits, "decode_fused: %d CTAs; every fc2 unit needs its own (the activation tiles are per K slice)", G);
  // Shared memory: weight ring (16 KB stages; 12 = a whole unit, so that nothing of a unit is fetched after its phase has
  // started) | the 12 activation tiles of a phase (R_pad x 1536 B, resident for all units of the phase) | arg-max slots |
  // attention scratch; 222 KB of dynamic memory at most (+ ~0.3 KB static inside the 227 KB an SM offers).
  const int a_tile = R_pad * 128;
  const int gen_bytes = 32 * R_pad + 8 * 194 * 4;
  int nsw = std::min(MAX_STAGES, (222 * 1024 - 1024 - gen_bytes - KB_PER_UNIT * a_tile) / W_TILE_BYTES);
  if (const char* e = getenv("PIO_DECODE_FUSED_WSTAGES")) nsw = std::max(2, std::min(nsw, atoi(e)));  // A/B runs
  PIO_CHECK(nsw >= 4, "decode_fused: %d rows leave no room for the weight ring", R);
  // the ring is two release groups (see the producer): gs stages each, gs a divisor of the 12 k-blocks of a unit
  int gs = nsw >= 12 ? 6 : (nsw >= 8 ? 4 : (nsw >= 6 ? 3 : 2));
  if (const char* e = getenv("PIO_DECODE_FUSED_GROUP")) { const int g = atoi(e); if (g == 1 || g == 2 || g == 3 || g == 4 || g == 6) gs = std::min(g, nsw); }
  nsw = std::min(nsw / gs, 12 / gs) * gs;
  const size_t smem = (size_t)nsw * W_TILE_BYTES + (size_t)KB_PER_UNIT * a_tile + gen_bytes + 1024;

  FusedParams P;
  memset(&P, 0, sizeof(P));
  P.wmaps = (const CUtensorMap*)h->fused_wmaps;
  P.layers = h->fused_layers;
  P.lnf_w = h->lnf_w; P.lnf_b = h->lnf_b; P.wte32 = h->wte32; P.wpe = h->wpe;
  P.x = w.x; P.h = (__nv_bfloat16*)w.hb; P.qkv = (__nv_bfloat16*)w.qkv; P.att = (__nv_bfloat16*)w.att; P.f = (__nv_bfloat16*)w.f;
  P.kc = (__nv_bfloat16*)w.kc; P.vc = (__nv_bfloat16*)w.vc; P.kv_layer = (long long)(w.kv_layer / 2);
  P.part = w.part; P.pm_val = w.pm_val; P.pm_idx = w.pm_idx;
  PIO_CHECK(fused_counter_ints(L, steps) * sizeof(int) <= w.counters_bytes, "decode_fused: counter region too small for %d steps", steps);
  P.phase_cnt = w.counters;
  P.abort = w.counters + (size_t)steps * PPS;
  P.out_ids = out_ids; P.ids_ld = steps;
  P.L = L; P.H = h->H; P.hd = gD / h->H; P.T = T; P.R = R; P.R_pad = R_pad; P.steps = steps; P.pos_base = pos_base;
  P.G = G; P.nsw = nsw; P.gs = gs; P.PPS = PPS;
  // the phases of one step (see the header comment)
  int np = 0, lnf_at = 0;
  for (int l = 0; l < L; ++l) {
    if (l > 0) { P.ptype[np] = P_LN1; P.player[np++] = (signed char)l; }
    for (int t : {(int)P_QKV, (int)P_ATTN, (int)P_PROJ, (int)P_FC, (int)P_FC2}) { P.ptype[np] = (signed char)t; P.player[np++] = (signed char)l; }
  }
  lnf_at = np;
  for (int t : {(int)P_LNF, (int)P_LMHEAD, (int)P_PICK}) { P.ptype[np] = (signed char)t; P.player[np++] = (signed char)L; }
  PIO_CHECK(np == PPS, "decode_fused: phase table of %d entries, expected %d", np, PPS);
  P.first_gp = start_at_pick ? lnf_at : 0;
  P.stop_gp = 1 << 30;
  if (const char* e = getenv("PIO_FUSED_MMA_MODE")) P.mma_mode = atoi(e);
  if (const char* e = getenv("PIO_DECODE_FUSED_PACE_NS")) P.pace_ns = atoi(e);
  P.l2_prefetch = 0;  // measured: no effect (lm-head phase 20.2 vs 20.4 us at 32 rows, profiles/r02al_*): the phase is bound by HBM + the ring refill cadence, and the L2 does not keep 58 MB for a whole step
  if (const char* e = getenv("PIO_DECODE_FUSED_L2PF")) P.l2_prefetch = atoi(e);
  if (const char* e = getenv("PIO_FUSED_STOP_PHASE")) P.stop_gp = atoi(e) + 1;  // debug: run global phases [first, stop]

  PIO_CUDA(cudaMemsetAsync(w.counters, 0, fused_counter_ints(L, steps) * sizeof(int), st));
  static SmemAttrOnce once;
  PIO_CUDA(once.ensure(decode_fused_kernel, 222 * 1024));
  int per_sm = 0;
  PIO_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, decode_fused_kernel, FT_THREADS, smem));
  PIO_CHECK(per_sm >= 1, "decode_fused: the kernel does not fit an SM (%zu bytes of shared memory)", smem);
  const char* tl_path = getenv("PIO_FUSED_TIMELINE");  // debug: per-(phase, CTA) time stamps -> binary file
  const size_t tl_n = tl_path ? (size_t)(steps * PPS) * G * 16 + 4 : 0;
  if (tl_path) {
    PIO_CUDA(cudaMalloc((void**)&P.timeline, tl_n * 8));
    PIO_CUDA(cudaMemsetAsync(P.timeline, 0, tl_n * 8, st));
  }
  void* args[] = {(void*)&P};
  PIO_CUDA(cudaLaunchCooperativeKernel((const void*)decode_fused_kernel, dim3(G), dim3(FT_THREADS), args, smem, st));
  PIO_LAUNCHED();
  if (tl_path) {
    std::vector<unsigned long long> host(tl_n);
    PIO_CUDA(cudaStreamSynchronize(st));
    PIO_CUDA(cudaMemcpy(host.data(), P.timeline, tl_n * 8, cudaMemcpyDeviceToHost));
    PIO_CUDA(cudaFree(P.timeline));
    if (FILE* f = fopen(tl_path, "wb")) {
      int hdr[8 + MAX_PHASES] = {steps, PPS, G, P.first_gp, L, R, 16, MAX_PHASES};
      for (int i = 0; i < MAX_PHASES; ++i) hdr[8 + i] = i < PPS ? P.ptype[i] * 16 + P.player[i] : -1;
      fwrite(hdr, sizeof(int), 8 + MAX_PHASES, f);
      fwrite(host.data(), 8, tl_n, f);
      fclose(f);
    }
  }
  if (getenv("PIO_FUSED_CHECK")) {  // debug: surface a drained (timed-out) kernel right away
    int flag = 0;
    PIO_CUDA(cudaStreamSynchronize(st));
    PIO_CUDA(cudaMemcpy(&flag, P.abort, sizeof(int), cudaMemcpyDeviceToHost));
    PIO_CHECK(flag == 0, "decode_fused: a wait inside the kernel timed out (abort flag raised)");
  }
  return PIO_OK;
}

}  // namespace pio

extern "C" int pio_set_decode_fused(int max_rows, int ctas) {
  if (max_rows >= 0) pio::g_fused_max_rows.store(max_rows);
  if (ctas >= 0) pio::g_fused_ctas.store(ctas);
  return PIO_OK;
}

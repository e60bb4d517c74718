// attention_sm100.cu -- ViT multi-head self-attention (head_dim 64) on the tcgen05 tensor cores.
//
// One CTA = 128 queries of one (image, head); it walks the keys in tiles of 128 with an online softmax:
//   warp 0      TMA producer : Q tile once, then per key tile K [128 keys x 64] and V^T [64 x 128 keys] into a
//               2-stage shared-memory ring (128-byte swizzle, mbarrier completion)
//   warp 1      MMA issuer   : S_j = Q K_j^T   (UMMA 128x128x16 x4, fp32 in TMEM)
//                              PV_j = P_j V_j  (UMMA 128x64x16 x8, A = P_j from shared memory, fresh accumulator)
//   warps 2..5  softmax      : thread == query row.  ONE pass over S_j (tcgen05.ld in 16-column chunks, the next chunk in
//               flight while this one is processed): P_j = exp2(S_j c - ref) -> bf16 -> shared memory in the UMMA K-major
//               swizzled layout, row sum and raw row max on the way.  Tensor-memory reads run at 64 B/clk per SM, so the
//               128 x 128 fp32 S tile costs 1024 clk to read ONCE -- the same as its 16384 ex2 on the SFU -- and every
//               further read (a separate max pass, a per-tile read of PV) is paid in full.  Hence:
//                 * O accumulates in tensor memory across key tiles (MMA accumulate), it is read once at the end;
//                 * ref is a lazily updated reference (FlashAttention-4 style): exact row max of tile 0, then raised
//                   only when the running max has moved more than 2^8 above it.  Entries of P may exceed 1 (<= 2^8, or
//                   <= 2^64 inside the tile that discovers a new max), harmless in bf16 / fp32.  Raising ref rescales
//                   the O rows in tensor memory (tcgen05.ld / st) and l -- rare, warp-uniform, exact.
// Footprint is trimmed to 112.3 KB of shared memory, 256 TMEM columns and <= 168 registers so that TWO CTAs are
// resident per SM: one CTA's softmax (MUFU/FMA bound) overlaps the other's tensor-core and TMA work.
// V is consumed exactly as the qkv GEMM stored it ([key, dim], dims contiguous): for P V the B operand is MN-major
// (instruction-descriptor bit 16; shared-memory descriptor with 8-key groups 1024 B apart), so no transposed copy exists.
// One in POLY exponentials is evaluated on the FMA pipe (Cody-Waite split + cubic, |rel err| < 7.5e-5, far below the
// bf16 rounding of P) instead of the SFU, which is the busiest pipe of this kernel (ncu: XU 60 % at POLY = 0).
#include "tc_ptx.cuh"
#include <stdlib.h>

namespace pio {
using namespace tc;
namespace {

#ifndef PIO_ATTN_WAIT_PV
#define PIO_ATTN_WAIT_PV 1
#endif
constexpr int HD = 64;        // head dim
constexpr int BQ = 128;       // queries per CTA
constexpr int KV_STAGES = 2;
constexpr int Q_BYTES = BQ * HD * 2;            // 16 KB
constexpr int NUM_BARS = 1 + 2 * KV_STAGES + 4;
constexpr int ATT_THREADS = 192;
// Two tile shapes: 128 keys per step with two CTAs per SM (long sequences: 518 px, N = 1374), and 64 keys per step with three
// CTAs per SM (short sequences: 224 px, N = 261, where 128-key tiles would pad 261 keys to 384).  Measured: 0.655 vs 0.684 ms
// at N = 1374 and 0.091 vs 0.075 ms at N = 261 (B = 64).
template <int BKV> struct AttCfg {
  static constexpr int K_BYTES = BKV * HD * 2;           // 16 / 8 KB
  static constexpr int V_BYTES = HD * BKV * 2;           // 16 / 8 KB, [keys x 64 dims] as stored
  static constexpr int KV_BYTES = K_BYTES + V_BYTES;
  static constexpr int P_BYTES = BQ * BKV * 2;           // 32 / 16 KB ([128 q x 64 keys] sub-tiles)
  static constexpr int SMEM = Q_BYTES + KV_STAGES * KV_BYTES + P_BYTES + NUM_BARS * 8 + 16;  // no static smem: base stays 1024-aligned
  static constexpr uint32_t TM_S0 = 0, TM_PV0 = BKV, TM_COLS = 2 * BKV;  // S at cols [0, BKV), O (sum of P V) at [BKV, BKV + 64)
  static constexpr int CTAS_PER_SM = BKV == 128 ? 2 : 3;
};

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Exact row max of one 128-key tile (only the first tile needs it).  MASK for a partial tile.
template <bool MASK, int BKV>
__device__ __forceinline__ float row_max(uint32_t s_addr, int kbase, int N) {
  float mx0 = -INFINITY, mx1 = -INFINITY;
  uint32_t ra[16], rb[16];
  tmem_ld16(s_addr, ra);
  auto take = [&](const uint32_t (&r)[16], int c) {
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      float v0 = __uint_as_float(r[i]), v1 = __uint_as_float(r[i + 1]);
      if (MASK) {
        if (kbase + c + i >= N) v0 = -INFINITY;
        if (kbase + c + i + 1 >= N) v1 = -INFINITY;
      }
      mx0 = fmaxf(mx0, v0);
      mx1 = fmaxf(mx1, v1);
    }
  };
#pragma unroll
  for (int c = 0; c < BKV; c += 32) {
    tmem_ld_wait();
    tmem_ld16(s_addr + c + 16, rb);
    take(ra, c);
    tmem_ld_wait();
    if (c + 32 < BKV) tmem_ld16(s_addr + c + 32, ra);
    take(rb, c + 16);
  }
  return fmaxf(mx0, mx1);
}

// One pass over S (fp32, TMEM): P = exp2(S * c - ref) -> bf16 -> swizzled K-major shared memory.
// Returns the row sum; tmax receives the raw row max of the tile (for the next tile's reference).
// 2^x on the FMA pipe: n = round(x) through the 1.5 * 2^23 trick, cubic minimax for 2^f on [-0.5, 0.5], exponent add.
__device__ __forceinline__ float exp2_poly(float x) {
  x = fmaxf(x, -126.0f);
  const float t = x + 12582912.0f;
  const float f = x - (t - 12582912.0f);
  float p = fmaf(0.05517132f, f, 0.24261054f);
  p = fmaf(p, f, 0.69326097f);
  p = fmaf(p, f, 0.99992812f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

// wait_h0 / wait_full (0 = no wait): mbarriers (parity wait_par) that say the previous tile's P V has finished reading the first
// 64-key sub-tile / the whole P buffer -- waited on right before the first store into each sub-tile, i.e. as late as possible.
// the same polynomial on a PAIR of exponents, two lanes per instruction: 10 instructions per pair instead of 2 x 8
__device__ __forceinline__ uint64_t exp2_poly2(float x0, float x1) {
  const uint64_t x = pack2(fmaxf(x0, -126.0f), fmaxf(x1, -126.0f));
  const uint64_t t = add2(x, pack2(12582912.0f, 12582912.0f));
  const uint64_t u = add2(t, pack2(-12582912.0f, -12582912.0f));
  const uint64_t f = fma2(u, pack2(-1.0f, -1.0f), x);
  uint64_t p = fma2(pack2(0.05517132f, 0.05517132f), f, pack2(0.24261054f, 0.24261054f));
  p = fma2(p, f, pack2(0.69326097f, 0.69326097f));
  p = fma2(p, f, pack2(0.99992812f, 0.99992812f));
  float t0, t1, p0, p1;
  unpack2(t, t0, t1);
  unpack2(p, p0, p1);
  return pack2(__int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23)), __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23)));
}

// POLY: 0 = every exponential on the SFU; 1..9 = the second element of every POLY-th pair on the FMA pipe (one in 2 POLY);
// 10 + n (packed pass only) = BOTH elements of every n-th pair, evaluated two lanes per instruction (one in n)
template <bool MASK, int POLY, int BKV, bool PACKED = false>
__device__ __forceinline__ float softmax_pass(uint32_t s_addr, uint8_t* prow, int row, int kbase, int N, float scale_log2e,
                                              float ref, float& tmax, uint32_t wait_h0 = 0, uint32_t wait_full = 0,
                                              uint32_t wait_par = 0) {
  float s0 = 0.f, s1 = 0.f, mx0 = -INFINITY, mx1 = -INFINITY;
  uint32_t ra[16], rb[16];
  tmem_ld16(s_addr, ra);
  // 16 keys = 2 chunks of 16 bytes: sub-tile (c / 64), chunk index ((c % 64) / 8 + q) ^ (row % 8)
  const uint64_t sc2 = pack2(scale_log2e, scale_log2e), nref2 = pack2(-ref, -ref);
  uint64_t s01 = pack2(0.f, 0.f);
  auto take = [&](const uint32_t (&r)[16], int c) {
    uint32_t pk[8];
    if constexpr (PACKED) {  // two elements per FFMA2 / FADD2: the 64-key kernel (three CTAs per SM) is short of issue slots
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        const float v0 = __uint_as_float(r[i]), v1 = __uint_as_float(r[i + 1]);
        float x0, x1;
        unpack2(fma2(pack2(v0, v1), sc2, nref2), x0, x1);
        float p0, p1;
        constexpr int PN = (POLY >= 10 && POLY < 100) ? POLY - 10 : 1;
        if ((POLY >= 100 && (((POLY - 100) >> (i >> 1)) & 1)) || (POLY >= 10 && POLY < 100 && ((i >> 1) % PN) == PN - 1)) {  // 100 + mask: pair k of each 16-element chunk if bit k
          unpack2(exp2_poly2(x0, x1), p0, p1);
        } else {
          p0 = ex2_approx(x0);
          p1 = (POLY > 0 && POLY < 10 && ((i >> 1) % (POLY > 0 ? POLY : 1)) == POLY - 1) ? exp2_poly(x1) : ex2_approx(x1);
        }
        float m0 = v0, m1 = v1;
        if (MASK) {
          if (kbase + c + i >= N) { p0 = 0.f; m0 = -INFINITY; }
          if (kbase + c + i + 1 >= N) { p1 = 0.f; m1 = -INFINITY; }
        }
        mx0 = fmaxf(mx0, m0);
        mx1 = fmaxf(mx1, m1);
        s01 = add2(s01, pack2(p0, p1));
        __nv_bfloat162 q2 = __floats2bfloat162_rn(p0, p1);
        pk[i >> 1] = *reinterpret_cast<uint32_t*>(&q2);
      }
    }
#pragma unroll
    for (int i = 0; i < (PACKED ? 0 : 16); i += 2) {
      float v0 = __uint_as_float(r[i]), v1 = __uint_as_float(r[i + 1]);
      const float x0 = fmaf(v0, scale_log2e, -ref), x1 = fmaf(v1, scale_log2e, -ref);
      float p0 = ex2_approx(x0);
      float p1 = (POLY > 0 && ((i >> 1) % POLY) == POLY - 1) ? exp2_poly(x1) : ex2_approx(x1);  // one in 2 * POLY off the SFU
      if (MASK) {
        if (kbase + c + i >= N) { p0 = 0.f; v0 = -INFINITY; }
        if (kbase + c + i + 1 >= N) { p1 = 0.f; v1 = -INFINITY; }
      }
      mx0 = fmaxf(mx0, v0);
      mx1 = fmaxf(mx1, v1);
      s0 += p0;
      s1 += p1;
      __nv_bfloat162 q2 = __floats2bfloat162_rn(p0, p1);
      pk[i >> 1] = *reinterpret_cast<uint32_t*>(&q2);
    }
    uint8_t* sub = prow + (c >> 6) * (BQ * 128);
    if (c == 0 && wait_h0 != 0) mbar_wait(BKV == 64 ? wait_full : wait_h0, wait_par);
    if (c == 64 && wait_full != 0) mbar_wait(wait_full, wait_par);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int chunk = (((c & 63) >> 3) + q) ^ (row & 7);
      *reinterpret_cast<uint4*>(sub + chunk * 16) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
    }
  };
#pragma unroll
  for (int c = 0; c < BKV; c += 32) {
    tmem_ld_wait();
    tmem_ld16(s_addr + c + 16, rb);
    take(ra, c);
    tmem_ld_wait();
    if (c + 32 < BKV) tmem_ld16(s_addr + c + 32, ra);
    take(rb, c + 16);
  }
  tmax = fmaxf(mx0, mx1);
  if constexpr (PACKED) unpack2(s01, s0, s1);
  return s0 + s1;
}

template <int POLY, int BKV, bool PACKED = false>
__global__ void __launch_bounds__(ATT_THREADS, AttCfg<BKV>::CTAS_PER_SM)
vit_attention_tc_kernel(const __grid_constant__ CUtensorMap map_qk, const __grid_constant__ CUtensorMap map_kv,
                        __nv_bfloat16* __restrict__ out, int N, int H, float scale_log2e) {
  using cfg = AttCfg<BKV>;
  constexpr int K_BYTES = cfg::K_BYTES, KV_BYTES = cfg::KV_BYTES, P_BYTES = cfg::P_BYTES;
  constexpr uint32_t TM_S0 = cfg::TM_S0, TM_PV0 = cfg::TM_PV0, TM_COLS = cfg::TM_COLS;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  if ((smem_base & 1023u) != 0) __trap();  // the 128B-swizzle atoms need 1024-byte alignment
  const uint32_t sQ = smem_base;
  const uint32_t sKV = sQ + Q_BYTES;
  const uint32_t sP = sKV + KV_STAGES * KV_BYTES;
  uint8_t* sP_gen = smem_raw + Q_BYTES + KV_STAGES * KV_BYTES;
  const uint32_t bar0 = sP + P_BYTES;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + Q_BYTES + KV_STAGES * KV_BYTES + P_BYTES + NUM_BARS * 8);
  const uint32_t q_full = bar0;
  auto kv_full = [&](int s) { return bar0 + 8u * (1 + s); };
  auto kv_empty = [&](int s) { return bar0 + 8u * (1 + KV_STAGES + s); };
  const uint32_t s_ready = bar0 + 8u * (1 + 2 * KV_STAGES);
  const uint32_t p_ready = s_ready + 8u;
  const uint32_t pv_step = s_ready + 16u;  // one phase per key tile (never > 1 phase ahead)
  const uint32_t pv_half = s_ready + 24u;  // one phase per key tile: the first 64-key sub-tile of P has been consumed

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * BQ, h = blockIdx.y, b = blockIdx.z;
  const int nt = (N + BKV - 1) / BKV;
  const int row_base = b * N;  // rows of the [B*N, 3*H*64] qkv matrix

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_qk) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_kv) : "memory");
    mbar_init(q_full, 1);
    for (int s = 0; s < KV_STAGES; ++s) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 1); }
    mbar_init(s_ready, 1);
    mbar_init(p_ready, 4);
    mbar_init(pv_step, 1);
    mbar_init(pv_half, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot_ptr)), TM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_expect_tx(q_full, Q_BYTES);
      tma_load_2d(sQ, &map_qk, q_full, h * HD, row_base + q0);
      for (int j = 0; j < nt; ++j) {
        const int s = j % KV_STAGES;
        const uint32_t ph = (j / KV_STAGES) & 1;
        mbar_wait(kv_empty(s), ph ^ 1);
        const uint32_t dst = sKV + s * KV_BYTES;
        mbar_expect_tx(kv_full(s), KV_BYTES);
        tma_load_2d(dst, &map_kv, kv_full(s), H * HD + h * HD, row_base + j * BKV);           // K_j  [128 keys x 64]
        tma_load_2d(dst + K_BYTES, &map_kv, kv_full(s), 2 * H * HD + h * HD, row_base + j * BKV);  // V_j  [128 keys x 64], as stored
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc_s = make_idesc(BQ, BKV);  // 128 x 128
    // 128 x 64; with V consumed as stored (keys x dims, dims contiguous) the B operand is MN-major: idesc bit 16
    constexpr uint32_t idesc_o = make_idesc(BQ, HD) | (1u << 16);
    auto issue_qk = [&](int j) {
      const int s = j % KV_STAGES;
      mbar_wait(kv_full(s), (j / KV_STAGES) & 1);
      tc_fence_after();
      if (lane == 0) {
        const uint64_t adesc = make_smem_desc(sQ), bdesc = make_smem_desc(sKV + s * KV_BYTES);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_f16(tmem_base + TM_S0, adesc + 2 * k, bdesc + 2 * k, idesc_s, k != 0);
        umma_commit(s_ready);
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    issue_qk(0);
    for (int j = 0; j < nt; ++j) {
      const int s = j % KV_STAGES;
      mbar_wait(p_ready, j & 1);   // P_j is in shared memory, S_j has been read, O carries the units P_j is in
      tc_fence_after();
      if (j + 1 < nt) issue_qk(j + 1);  // S first: the softmax warps need it next
      if (lane == 0) {
        const uint32_t vbase = sKV + s * KV_BYTES + K_BYTES;
#pragma unroll
        for (int k = 0; k < BKV / 16; ++k) {
          const uint64_t adesc = make_smem_desc(sP + (k >> 2) * (BQ * 128)) + 2 * (k & 3);
          // MN-major V: rows are keys (128 B = 64 dims each), 8-key groups 1024 B apart (SBO); K = 16 keys = 2048 B per step.
          const uint64_t bdesc = make_smem_desc(vbase + k * 2048);
          umma_f16(tmem_base + TM_PV0, adesc, bdesc, idesc_o, (j | k) != 0);  // O accumulates across key tiles
          if (BKV == 128 && k == 3) umma_commit(pv_half);
        }
        umma_commit(pv_step);
        umma_commit(kv_empty(s));
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ softmax / output (warps 2..5)
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;           // query row inside the tile == TMEM lane
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const uint32_t s_addr = tmem_base + lane_addr + TM_S0;
    const uint32_t pv_addr = tmem_base + lane_addr + TM_PV0;
    uint8_t* prow = sP_gen + row * 128;
    float ref = 0.f, l = 0.f, seen = -INFINITY, tmax = -INFINITY;
    // Raise the reference to new_ref (>= ref; equal for rows that keep theirs): O rows in tensor memory and l move to the
    // new units.  j = the tile about to be (re)done: O then holds P V of tiles 0..j-1.  Warp-uniform.
    auto rescale = [&](int j, float new_ref) {
      const float f = ex2_approx(ref - new_ref);
      if (j > 0) {
        mbar_wait(pv_step, (j - 1) & 1);  // P_{j-1} V_{j-1} has landed
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < HD; c += 16) {
          uint32_t r[16];
          tmem_ld16(pv_addr + c, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * f);
          tmem_st16(pv_addr + c, r);
        }
        tmem_st_wait();
      }
      l *= f;
      ref = new_ref;
    };

    for (int j = 0; j < nt; ++j) {
      mbar_wait(s_ready, j & 1);
      tc_fence_after();
      const bool mask = (j + 1) * BKV > N;
      const int kbase = j * BKV;
      if (j == 0) {
        ref = (mask ? row_max<true, BKV>(s_addr, kbase, N) : row_max<false, BKV>(s_addr, kbase, N)) * scale_log2e;  // finite: >= 1 valid key
      } else if (__any_sync(0xffffffffu, fmaf(seen, scale_log2e, -ref) > 8.f)) {
        rescale(j, fmaxf(ref, seen * scale_log2e));
      }
      // P_{j-1} V_{j-1} -- issued AFTER Q K_j^T, so S_j being ready says nothing about it -- must have finished reading a 64-key
      // sub-tile of the (single) P buffer before this tile's probabilities overwrite it.  Round 1 relied on the tensor pipe
      // staying ahead of the P stores; a second CTA's MMAs on the same pipe can delay it.  Round 2: the MMA warp commits after each
      // half of P V, the softmax warps wait right before their first store into each sub-tile.
      const uint32_t wh = (j > 0 && PIO_ATTN_WAIT_PV) ? pv_half : 0u, wf = (j > 0 && PIO_ATTN_WAIT_PV) ? pv_step : 0u;
      const uint32_t wpar = (uint32_t)(j - 1) & 1u;
      float ls;
      for (;;) {
        ls = mask ? softmax_pass<true, POLY, BKV, PACKED>(s_addr, prow, row, kbase, N, scale_log2e, ref, tmax, wh, wf, wpar)
                  : softmax_pass<false, POLY, BKV, PACKED>(s_addr, prow, row, kbase, N, scale_log2e, ref, tmax, wh, wf, wpar);
        // exponent headroom: a row whose tile max sits more than 2^64 above its reference redoes the tile
        if (!__any_sync(0xffffffffu, fmaf(tmax, scale_log2e, -ref) > 64.f)) break;
        rescale(j, fmaxf(ref, tmax * scale_log2e));
      }
      seen = fmaxf(seen, tmax);
      l += ls;
      fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor-core (async) proxy
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready);
    }
    // O / l -> bf16 -> global (the only full read of the output accumulator)
    mbar_wait(pv_step, (nt - 1) & 1);
    tc_fence_after();
    const float inv = 1.0f / l;
    const int q = q0 + row;
    __nv_bfloat16* dst = out + ((long long)(row_base + q)) * (H * HD) + h * HD;
    uint32_t ra[16], rb[16];
    auto emit = [&](const uint32_t (&r)[16], int c) {
      if (q >= N) return;
      uint32_t pk[8];
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        __nv_bfloat162 t = __floats2bfloat162_rn(__uint_as_float(r[i]) * inv, __uint_as_float(r[i + 1]) * inv);
        pk[i >> 1] = *reinterpret_cast<uint32_t*>(&t);
      }
      *reinterpret_cast<uint4*>(dst + c) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      *reinterpret_cast<uint4*>(dst + c + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    };
    tmem_ld16(pv_addr, ra);
#pragma unroll
    for (int c = 0; c < HD; c += 32) {
      tmem_ld_wait();
      tmem_ld16(pv_addr + c + 16, rb);
      emit(ra, c);
      tmem_ld_wait();
      if (c + 32 < HD) tmem_ld16(pv_addr + c + 32, ra);
      emit(rb, c + 16);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TM_COLS);
  }
}


// ------------------------------------------------------------------------------------------------------------------
// Round-2 kernel for long sequences: 64-key tiles with S and P DOUBLE-BUFFERED, separate K and V rings.
//
// In the kernel above one CTA's loop is a chain: softmax(j) -> P ready -> QK(j+1) issued -> S ready -> softmax(j+1); the
// softmax warps idle while S = Q K^T of the next tile is produced (ncu r01d: their top stall is the wait for S), and K(j+1) can
// only be fetched once PV(j-1) has released the shared K/V stage -- about one TMA latency before it is needed.  It also
// re-uses its single P buffer as soon as S(j+1) is ready, although PV(j) -- issued AFTER QK(j+1) -- may still be reading it:
// harmless only as long as PV(j) runs ahead of the first P stores of softmax(j+1), which a second CTA's MMAs on the same tensor
// pipe can break (the run-to-run differences of profiles/r02ab).  Here:
//   * S lives in two 64-column TMEM buffers: QK(j+2) is issued right after PV(j), two tiles ahead of the softmax that reads it,
//     so the softmax warps go from tile to tile without waiting for the tensor pipe;
//   * P lives in two shared-memory buffers; softmax(j) waits for PV(j-2) (mbarrier pv_done) before it overwrites one -- no race;
//   * K and V have their own rings and barriers: a K stage is released by QK (not by the later PV), so K(j+3) is fetched
//     three tiles ahead; one producer lane per ring;
//   * 96 KB of shared memory and 256 TMEM columns (S0 | S1 | O | unused): two CTAs per SM as before.
constexpr int DB_BKV = 64, DB_KST = 3;
constexpr int DB_KB = DB_BKV * HD * 2;                     // 8 KB: one K (or V) tile
constexpr int DB_PB = BQ * DB_BKV * 2;                     // 16 KB: one P buffer
constexpr int DB_NBARS = 1 + 4 * DB_KST + 6;
constexpr int DB_SMEM = Q_BYTES + 2 * DB_KST * DB_KB + 2 * DB_PB + DB_NBARS * 8 + 16;

template <int POLY>
__global__ void __launch_bounds__(ATT_THREADS, 2)
vit_attention_db_kernel(const __grid_constant__ CUtensorMap map_qk, const __grid_constant__ CUtensorMap map_kv,
                        __nv_bfloat16* __restrict__ out, int N, int H, float scale_log2e) {
  constexpr int BKV = DB_BKV, KST = DB_KST;
  constexpr uint32_t TM_S = 0, TM_O = 2 * BKV, TM_COLS = 256;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  if ((smem_base & 1023u) != 0) __trap();
  const uint32_t sQ = smem_base;
  const uint32_t sK = sQ + Q_BYTES;
  const uint32_t sV = sK + KST * DB_KB;
  const uint32_t sP = sV + KST * DB_KB;
  uint8_t* sP_gen = smem_raw + Q_BYTES + 2 * KST * DB_KB;
  const uint32_t bar0 = sP + 2 * DB_PB;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + Q_BYTES + 2 * KST * DB_KB + 2 * DB_PB + DB_NBARS * 8);
  const uint32_t q_full = bar0;
  auto k_full = [&](int s) { return bar0 + 8u * (1 + s); };
  auto k_empty = [&](int s) { return bar0 + 8u * (1 + KST + s); };
  auto v_full = [&](int s) { return bar0 + 8u * (1 + 2 * KST + s); };
  auto v_empty = [&](int s) { return bar0 + 8u * (1 + 3 * KST + s); };
  auto s_ready = [&](int b) { return bar0 + 8u * (1 + 4 * KST + b); };
  auto p_ready = [&](int b) { return bar0 + 8u * (3 + 4 * KST + b); };
  auto pv_done = [&](int b) { return bar0 + 8u * (5 + 4 * KST + b); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * BQ, h = blockIdx.y, b = blockIdx.z;
  const int nt = (N + BKV - 1) / BKV;
  const int row_base = b * N;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_qk) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_kv) : "memory");
    mbar_init(q_full, 1);
    for (int s = 0; s < KST; ++s) { mbar_init(k_full(s), 1); mbar_init(k_empty(s), 1); mbar_init(v_full(s), 1); mbar_init(v_empty(s), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(s_ready(i), 1); mbar_init(p_ready(i), 4); mbar_init(pv_done(i), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot_ptr)), TM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producers: lane 0 the K ring (+ Q), lane 1 the V ring
    if (lane == 0) {
      mbar_expect_tx(q_full, Q_BYTES);
      tma_load_2d(sQ, &map_qk, q_full, h * HD, row_base + q0);
      for (int j = 0; j < nt; ++j) {
        const int s = j % KST;
        mbar_wait(k_empty(s), ((j / KST) & 1) ^ 1);
        mbar_expect_tx(k_full(s), DB_KB);
        tma_load_2d(sK + s * DB_KB, &map_kv, k_full(s), H * HD + h * HD, row_base + j * BKV);
      }
    } else if (lane == 1) {
      for (int j = 0; j < nt; ++j) {
        const int s = j % KST;
        mbar_wait(v_empty(s), ((j / KST) & 1) ^ 1);
        mbar_expect_tx(v_full(s), DB_KB);
        tma_load_2d(sV + s * DB_KB, &map_kv, v_full(s), 2 * H * HD + h * HD, row_base + j * BKV);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc_s = make_idesc(BQ, BKV);
    constexpr uint32_t idesc_o = make_idesc(BQ, HD) | (1u << 16);  // V as stored: MN-major B operand
    auto issue_qk = [&](int j) {  // S[j & 1] = Q K_j^T
      const int s = j % KST;
      mbar_wait(k_full(s), (j / KST) & 1);
      tc_fence_after();
      if (lane == 0) {
        const uint64_t adesc = make_smem_desc(sQ), bdesc = make_smem_desc(sK + s * DB_KB);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_f16(tmem_base + TM_S + (j & 1) * BKV, adesc + 2 * k, bdesc + 2 * k, idesc_s, k != 0);
        umma_commit(s_ready(j & 1));
        umma_commit(k_empty(s));
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    issue_qk(0);
    if (nt > 1) issue_qk(1);
    for (int j = 0; j < nt; ++j) {
      const int s = j % KST, pb = j & 1;
      mbar_wait(p_ready(pb), (j >> 1) & 1);   // P_j is in shared memory and S_j has been read
      mbar_wait(v_full(s), (j / KST) & 1);
      tc_fence_after();
      if (lane == 0) {
#pragma unroll
        for (int k = 0; k < BKV / 16; ++k) {
          const uint64_t adesc = make_smem_desc(sP + pb * DB_PB) + 2 * k;
          const uint64_t bdesc = make_smem_desc(sV + s * DB_KB + k * 2048);  // 16 keys = 2048 B per K step
          umma_f16(tmem_base + TM_O, adesc, bdesc, idesc_o, (j | k) != 0);
        }
        umma_commit(pv_done(pb));
        umma_commit(v_empty(s));
      }
      __syncwarp();
      if (j + 2 < nt) issue_qk(j + 2);  // into the S buffer softmax(j) has just drained
    }
  } else {
    // ------------------------------------------------------------------ softmax / output (warps 2..5)
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const uint32_t pv_addr = tmem_base + lane_addr + TM_O;
    float ref = 0.f, l = 0.f, seen = -INFINITY, tmax = -INFINITY;
    auto wait_pv = [&](int j) { mbar_wait(pv_done(j & 1), (j >> 1) & 1); };  // P_j V_j has landed in O (and P[j & 1] is free)
    auto rescale = [&](int j, float new_ref) {
      const float f = ex2_approx(ref - new_ref);
      if (j > 0) {
        wait_pv(j - 1);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < HD; c += 16) {
          uint32_t r[16];
          tmem_ld16(pv_addr + c, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * f);
          tmem_st16(pv_addr + c, r);
        }
        tmem_st_wait();
      }
      l *= f;
      ref = new_ref;
    };
    for (int j = 0; j < nt; ++j) {
      const int pb = j & 1;
      const uint32_t s_addr = tmem_base + lane_addr + TM_S + pb * BKV;
      uint8_t* prow = sP_gen + pb * DB_PB + row * 128;
      mbar_wait(s_ready(pb), (j >> 1) & 1);
      tc_fence_after();
      const bool mask = (j + 1) * BKV > N;
      const int kbase = j * BKV;
      if (j == 0) {
        ref = (mask ? row_max<true, BKV>(s_addr, kbase, N) : row_max<false, BKV>(s_addr, kbase, N)) * scale_log2e;
      } else if (__any_sync(0xffffffffu, fmaf(seen, scale_log2e, -ref) > 8.f)) {
        rescale(j, fmaxf(ref, seen * scale_log2e));
      }
      if (j >= 2) wait_pv(j - 2);  // the tensor core has finished reading this P buffer
      float ls;
      for (;;) {
        ls = mask ? softmax_pass<true, POLY, BKV>(s_addr, prow, row, kbase, N, scale_log2e, ref, tmax)
                  : softmax_pass<false, POLY, BKV>(s_addr, prow, row, kbase, N, scale_log2e, ref, tmax);
        if (!__any_sync(0xffffffffu, fmaf(tmax, scale_log2e, -ref) > 64.f)) break;
        rescale(j, fmaxf(ref, tmax * scale_log2e));
      }
      seen = fmaxf(seen, tmax);
      l += ls;
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready(pb));
    }
    wait_pv(nt - 1);
    tc_fence_after();
    const float inv = 1.0f / l;
    const int q = q0 + row;
    __nv_bfloat16* dst = out + ((long long)(row_base + q)) * (H * HD) + h * HD;
    uint32_t ra[16], rb[16];
    auto emit = [&](const uint32_t (&r)[16], int c) {
      if (q >= N) return;
      uint32_t pk[8];
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        __nv_bfloat162 t = __floats2bfloat162_rn(__uint_as_float(r[i]) * inv, __uint_as_float(r[i + 1]) * inv);
        pk[i >> 1] = *reinterpret_cast<uint32_t*>(&t);
      }
      *reinterpret_cast<uint4*>(dst + c) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      *reinterpret_cast<uint4*>(dst + c + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    };
    tmem_ld16(pv_addr, ra);
#pragma unroll
    for (int c = 0; c < HD; c += 32) {
      tmem_ld_wait();
      tmem_ld16(pv_addr + c + 16, rb);
      emit(ra, c);
      tmem_ld_wait();
      if (c + 32 < HD) tmem_ld16(pv_addr + c + 32, ra);
      emit(rb, c + 16);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TM_COLS);
  }
}
}  // namespace

size_t vit_attention_tc_workspace(int B, int N, int H) {
  (void)B; (void)N; (void)H;
  return 0;  // V is read in place (MN-major operand): no transposed copy
}

// qkv bf16 [B,N,3*H*64] -> out bf16 [B,N,H*64]; vt_ws holds the transposed V copy.
int vit_attention_tc(const void* qkv, void* out, void* vt_ws, int B, int N, int H, cudaStream_t st) {
  (void)vt_ws;
  // PIO_ATTN_POLY: 0 = every exponential on the SFU; n < 10 = one in 2n on the FMA pipe; 10 + n = both elements of every n-th pair
  // by the two-lane polynomial; 100 + mask = the pairs of each 16-element chunk whose bit is set.  Unset: 18 for the 64-key kernel
  // (the LAST pair of every chunk: 0.648 ms against 0.674 for one-in-six scalar polynomials and 0.705 for none; masks 0xC0 0.646,
  // 0x81 0.653, two spread pairs 0.66 -- profiles/r02bh_attention_pair_poly.txt), 3 for the 128-key kernel.  Same-box sweep of the
  // 64-key kernel at B = 64, N = 1374 (profiles/r02be_attention_variants.txt): scalar arithmetic 0.693 / 0.712 / 0.701 ms at
  // POLY = 0 / 2 / 3; with x = S c - ref and the row sums on two elements per instruction (PIO_ATTN_PACKED, the default)
  // 0.706 / 0.681 / 0.674 ms -- the 64-key kernel with three CTAs per SM is short of issue slots, not of SFU throughput
  static const int poly_env = [] { const char* e = getenv("PIO_ATTN_POLY"); return e ? atoi(e) : -1; }();
  CUtensorMap mqk, mkv;
  PIO_TRY(make_map_2d(&mqk, qkv, (long long)B * N, 3 * H * HD, 3 * H * HD, BQ, HD));
  // tile choice by sequence length: fewer padded keys for short sequences (PIO_ATTN_BKV=64|128 overrides)
  static const int force_bkv = [] { const char* e = getenv("PIO_ATTN_BKV"); return e ? atoi(e) : 0; }();
  // double-buffered kernel (round 2; measured SLOWER: 0.763 vs 0.666 ms at N = 1374, 0.284 vs 0.244 ms at N = 261, so off by
  // default): PIO_ATTN_DB = 0 never (default), 1 long sequences only, 2 every sequence length
  static const int db = [] { const char* e = getenv("PIO_ATTN_DB"); return e ? atoi(e) : 0; }();
  if (!force_bkv && ((db == 1 && N > 640) || db == 2)) {
    PIO_TRY(make_map_2d(&mkv, qkv, (long long)B * N, 3 * H * HD, 3 * H * HD, DB_BKV, HD));
    dim3 grid(cdiv(N, BQ), H, B);
    const float sl2 = 0.125f * 1.4426950408889634f;
#define PIO_ATT_DB_LAUNCH(POLY)                                                                                                  \
  do {                                                                                                                           \
    static SmemAttrOnce once;                                                                                                    \
    PIO_CUDA(once.ensure(vit_attention_db_kernel<POLY>, DB_SMEM));                                                                \
    launch_pdl_k(PDL_KIND_ATTN, vit_attention_db_kernel<POLY>, grid, dim3(ATT_THREADS), DB_SMEM, st, mqk, mkv, (__nv_bfloat16*)out, N, H, sl2); \
  } while (0)
    const int dbpoly = poly_env >= 0 ? poly_env : 0;
    if (dbpoly == 0) PIO_ATT_DB_LAUNCH(0); else if (dbpoly == 2) PIO_ATT_DB_LAUNCH(2); else PIO_ATT_DB_LAUNCH(3);
#undef PIO_ATT_DB_LAUNCH
    PIO_LAUNCHED();
    return PIO_OK;
  }
  // Round 1 dispatched by sequence length (N <= 640 -> 64-key tiles).  With the P-buffer race closed the 128-key kernel pays for
  // its waits (0.666 -> 0.717 ms at B = 64, N = 1374) and the 64-key kernel with three CTAs per SM is the faster one at every
  // length measured (0.698 ms there; profiles/r02be_attention_variants.txt): 64-key tiles everywhere, 128 only when forced.
  const int bkv = force_bkv ? force_bkv : 64;
  const int poly = poly_env >= 0 ? poly_env : (bkv == 64 ? 18 : 3);
  PIO_TRY(make_map_2d(&mkv, qkv, (long long)B * N, 3 * H * HD, 3 * H * HD, bkv, HD));
  dim3 grid(cdiv(N, BQ), H, B);
  const float scale_log2e = 0.125f * 1.4426950408889634f;
#define PIO_ATT_LAUNCH3(POLY, BKVV, PK)                                                                                           \
  do {                                                                                                                            \
    static SmemAttrOnce once;                                                                                                     \
    PIO_CUDA(once.ensure(vit_attention_tc_kernel<POLY, BKVV, PK>, AttCfg<BKVV>::SMEM));                                            \
    launch_pdl_k(PDL_KIND_ATTN, vit_attention_tc_kernel<POLY, BKVV, PK>, grid, dim3(ATT_THREADS), AttCfg<BKVV>::SMEM, st, mqk, mkv, (__nv_bfloat16*)out, N, H, \
               scale_log2e);                                                                                                      \
  } while (0)
#define PIO_ATT_LAUNCH(POLY, BKVV) PIO_ATT_LAUNCH3(POLY, BKVV, false)
  // PIO_ATTN_PACKED (64-key kernel): x = S c - ref and the row sums on two elements per instruction (fma / add .f32x2)
  static const int packed_env = [] { const char* e = getenv("PIO_ATTN_PACKED"); return e ? atoi(e) : -1; }();
  const bool packed = packed_env >= 0 ? packed_env != 0 : true;
  if (bkv == 64 && packed) {
    if (poly == 0) PIO_ATT_LAUNCH3(0, 64, true); else if (poly == 2) PIO_ATT_LAUNCH3(2, 64, true); else if (poly == 3) PIO_ATT_LAUNCH3(3, 64, true);
    else if (poly == 13) PIO_ATT_LAUNCH3(13, 64, true); else if (poly == 100 + 0xC0) PIO_ATT_LAUNCH3(100 + 0xC0, 64, true);
    else PIO_ATT_LAUNCH3(18, 64, true);
  }
  else if (bkv == 64) { if (poly == 0) PIO_ATT_LAUNCH(0, 64); else if (poly == 2) PIO_ATT_LAUNCH(2, 64); else PIO_ATT_LAUNCH(3, 64); }
  else           { if (poly == 0) PIO_ATT_LAUNCH(0, 128); else if (poly == 2) PIO_ATT_LAUNCH(2, 128); else PIO_ATT_LAUNCH(3, 128); }
#undef PIO_ATT_LAUNCH
#undef PIO_ATT_LAUNCH3
  PIO_LAUNCHED();
  return PIO_OK;
}

}  // namespace pio

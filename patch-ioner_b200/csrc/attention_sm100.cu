// attention_sm100.cu -- ViT multi-head self-attention (head_dim 64) on the tcgen05 tensor cores.
//
// One CTA = 128 queries of one (image, head); it walks the keys in tiles of 128 with an online softmax:
//   warp 0      TMA producer : Q tile once, then per key tile K [128 keys x 64] and V^T [64 x 128 keys] into a
//               2-stage shared-memory ring (128-byte swizzle, mbarrier completion)
//   warp 1      MMA issuer   : S_j = Q K_j^T   (UMMA 128x128x16 x4, fp32 in TMEM)
//                              PV_j = P_j V_j  (UMMA 128x64x16 x8, A = P_j from shared memory, fresh accumulator)
//   warps 2..5  softmax      : thread == query row.  tcgen05.ld S_j, running max / sum in fp32 with ex2.approx,
//               O = (O + PV_{j-1}) * alpha_j in registers (no TMEM read-modify-write of the output), then
//               P_j -> bf16 -> shared memory in the UMMA K-major swizzled layout.
// Footprint is trimmed to 112.3 KB of shared memory, 256 TMEM columns and <= 168 registers so that TWO CTAs are
// resident per SM: one CTA's softmax (MUFU/FMA bound) overlaps the other's tensor-core and TMA work.
// V is consumed K-major (keys contiguous) from a transposed copy V^T [B, H, 64, Np] written by transpose_v_kernel,
// so both GEMMs use the same, verified, K-major 128B-swizzle descriptors as gemm_sm100.cu.
#include "tc_ptx.cuh"

namespace pio {
using namespace tc;
namespace {

constexpr int HD = 64;        // head dim
constexpr int BQ = 128;       // queries per CTA
constexpr int BKV = 128;      // keys per tile
constexpr int KV_STAGES = 2;
constexpr int Q_BYTES = BQ * HD * 2;            // 16 KB
constexpr int K_BYTES = BKV * HD * 2;           // 16 KB
constexpr int V_BYTES = HD * BKV * 2;           // 16 KB (two [64 d x 64 keys] sub-tiles)
constexpr int KV_BYTES = K_BYTES + V_BYTES;
constexpr int P_BYTES = BQ * BKV * 2;           // 32 KB (two [128 q x 64 keys] sub-tiles)
constexpr int NUM_BARS = 1 + 2 * KV_STAGES + 3;
constexpr int ATT_SMEM = Q_BYTES + KV_STAGES * KV_BYTES + P_BYTES + NUM_BARS * 8 + 16;  // no static smem: base stays 1024-aligned
constexpr int ATT_THREADS = 192;
constexpr uint32_t TM_S0 = 0, TM_PV0 = 128, TM_COLS = 256;  // S at cols 0..127, PV at cols 128..191

// V^T[b,h,d,n] = V[b,n,h,d]; columns [N, Np) are zero.  grid (ceil(Np/64), H, B), block (64, 4)
__global__ void __launch_bounds__(256) transpose_v_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ vt,
                                                          int N, int Np, int H) {
  __shared__ __nv_bfloat16 tile[64][HD + 2];
  pdl_wait();
  pdl_launch_dependents();
  const int n0 = blockIdx.x * 64, h = blockIdx.y, b = blockIdx.z;
  const int C3 = 3 * H * HD;
  const __nv_bfloat16* src = qkv + (long long)b * N * C3 + 2 * H * HD + h * HD;
  for (int r = threadIdx.y; r < 64; r += 4) {
    const int n = n0 + r;
    tile[r][threadIdx.x] = (n < N) ? src[(long long)n * C3 + threadIdx.x] : __float2bfloat16(0.f);
  }
  __syncthreads();
  __nv_bfloat16* dst = vt + ((long long)(b * H + h) * HD) * Np;
  for (int d = threadIdx.y; d < HD; d += 4) {
    const int n = n0 + threadIdx.x;
    if (n < Np) dst[(long long)d * Np + n] = tile[threadIdx.x][d];
  }
}

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// One query row against one 128-key tile, in 32-column tcgen05.ld chunks (TMEM loads are cheap: re-reading S
// beats keeping 128 values live in registers).  MASK only for the last, partial key tile.
template <bool MASK>
__device__ __forceinline__ float row_max(uint32_t s_addr, int kbase, int N) {
  float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll 1
  for (int c = 0; c < BKV; c += 32) {
    uint32_t r[32];
    tmem_ld32(s_addr + c, r);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      float v0 = __uint_as_float(r[i]), v1 = __uint_as_float(r[i + 1]), v2 = __uint_as_float(r[i + 2]), v3 = __uint_as_float(r[i + 3]);
      if (MASK) {
        if (kbase + c + i >= N) v0 = -INFINITY;
        if (kbase + c + i + 1 >= N) v1 = -INFINITY;
        if (kbase + c + i + 2 >= N) v2 = -INFINITY;
        if (kbase + c + i + 3 >= N) v3 = -INFINITY;
      }
      mx0 = fmaxf(mx0, v0); mx1 = fmaxf(mx1, v1); mx2 = fmaxf(mx2, v2); mx3 = fmaxf(mx3, v3);
    }
  }
  return fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
}

// S (fp32, TMEM) -> P = exp2(S * c - m_new) (bf16, swizzled K-major shared memory); returns the row sum.
template <bool MASK>
__device__ __forceinline__ float write_probs(uint32_t s_addr, uint8_t* prow, int row, int kbase, int N, float scale_log2e,
                                             float m_new) {
  float s0 = 0.f, s1 = 0.f;
#pragma unroll 1
  for (int c = 0; c < BKV; c += 32) {
    uint32_t r[32];
    tmem_ld32(s_addr + c, r);
    tmem_ld_wait();
    uint32_t pk[16];
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      float p0 = ex2_approx(fmaf(__uint_as_float(r[i]), scale_log2e, -m_new));
      float p1 = ex2_approx(fmaf(__uint_as_float(r[i + 1]), scale_log2e, -m_new));
      if (MASK) {
        if (kbase + c + i >= N) p0 = 0.f;
        if (kbase + c + i + 1 >= N) p1 = 0.f;
      }
      s0 += p0;
      s1 += p1;
      __nv_bfloat162 q2 = __floats2bfloat162_rn(p0, p1);
      pk[i >> 1] = *reinterpret_cast<uint32_t*>(&q2);
    }
    // 32 keys = 4 chunks of 16 bytes; sub-tile (c / 64), chunk index ((c % 64) / 8 + q) ^ (row % 8)
    uint8_t* sub = prow + (c >> 6) * (BQ * 128);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int chunk = (((c & 63) >> 3) + q) ^ (row & 7);
      *reinterpret_cast<uint4*>(sub + chunk * 16) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
    }
  }
  return s0 + s1;
}

__global__ void __launch_bounds__(ATT_THREADS, 2)
vit_attention_tc_kernel(const __grid_constant__ CUtensorMap map_qk, const __grid_constant__ CUtensorMap map_vt,
                        __nv_bfloat16* __restrict__ out, int N, int H, float scale_log2e) {
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
  const uint32_t pv_done = s_ready + 16u;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * BQ, h = blockIdx.y, b = blockIdx.z;
  const int nt = (N + BKV - 1) / BKV;
  const int row_base = b * N;  // rows of the [B*N, 3*H*64] qkv matrix

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_qk) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_vt) : "memory");
    mbar_init(q_full, 1);
    for (int s = 0; s < KV_STAGES; ++s) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 1); }
    mbar_init(s_ready, 1);
    mbar_init(p_ready, 4);
    mbar_init(pv_done, 1);
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
      const int vt_row = (b * H + h) * HD;
      for (int j = 0; j < nt; ++j) {
        const int s = j % KV_STAGES;
        const uint32_t ph = (j / KV_STAGES) & 1;
        mbar_wait(kv_empty(s), ph ^ 1);
        const uint32_t dst = sKV + s * KV_BYTES;
        mbar_expect_tx(kv_full(s), KV_BYTES);
        tma_load_2d(dst, &map_qk, kv_full(s), H * HD + h * HD, row_base + j * BKV);           // K_j  [128 keys x 64]
        tma_load_2d(dst + K_BYTES, &map_vt, kv_full(s), j * BKV, vt_row);                     // V^T  [64 x keys 0..63]
        tma_load_2d(dst + K_BYTES + V_BYTES / 2, &map_vt, kv_full(s), j * BKV + 64, vt_row);  // V^T  [64 x keys 64..127]
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc_s = make_idesc(BQ, BKV);  // 128 x 128
    constexpr uint32_t idesc_o = make_idesc(BQ, HD);   // 128 x 64
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
      mbar_wait(p_ready, j & 1);   // P_j is in shared memory; S_j and PV_{j-1} have been read
      tc_fence_after();
      if (j + 1 < nt) issue_qk(j + 1);  // S first: the softmax warps need it next
      if (lane == 0) {
        const uint32_t vbase = sKV + s * KV_BYTES + K_BYTES;
#pragma unroll
        for (int k = 0; k < BKV / 16; ++k) {
          const uint64_t adesc = make_smem_desc(sP + (k >> 2) * (BQ * 128)) + 2 * (k & 3);
          const uint64_t bdesc = make_smem_desc(vbase + (k >> 2) * (HD * 128)) + 2 * (k & 3);
          umma_f16(tmem_base + TM_PV0, adesc, bdesc, idesc_o, k != 0);
        }
        umma_commit(pv_done);
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
    float m = -INFINITY, l = 0.f;
    float o[HD];
#pragma unroll
    for (int i = 0; i < HD; ++i) o[i] = 0.f;

    for (int j = 0; j < nt; ++j) {
      mbar_wait(s_ready, j & 1);
      tc_fence_after();
      const bool mask = (j + 1) * BKV > N;
      const int kbase = j * BKV;
      // pass 1: row max
      const float mx = mask ? row_max<true>(s_addr, kbase, N) : row_max<false>(s_addr, kbase, N);
      const float m_new = fmaxf(m, mx * scale_log2e);  // every tile has >= 1 valid key, so m_new is finite
      const float alpha = ex2_approx(m - m_new);       // m = -inf on the first tile -> 0
      // fold in the previous tile's PV (its P buffer and PV accumulator are then free), rescale to the new max
      if (j > 0) {
        mbar_wait(pv_done, (j - 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < HD; c += 32) {
          uint32_t r[32];
          tmem_ld32(pv_addr + c, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[c + i] = (o[c + i] + __uint_as_float(r[i])) * alpha;
        }
      }
      // pass 2: probabilities -> bf16 -> swizzled shared memory
      const float ls = mask ? write_probs<true>(s_addr, prow, row, kbase, N, scale_log2e, m_new)
                            : write_probs<false>(s_addr, prow, row, kbase, N, scale_log2e, m_new);
      fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor-core (async) proxy
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready);
      l = l * alpha + ls;
      m = m_new;
    }
    {
      mbar_wait(pv_done, (nt - 1) & 1);
      tc_fence_after();
      const float inv = 1.0f / l;
#pragma unroll
      for (int c = 0; c < HD; c += 32) {
        uint32_t r[32];
        tmem_ld32(pv_addr + c, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[c + i] = (o[c + i] + __uint_as_float(r[i])) * inv;
      }
    }
    const int q = q0 + row;
    if (q < N) {
      __nv_bfloat16* dst = out + ((long long)(row_base + q)) * (H * HD) + h * HD;
#pragma unroll
      for (int i = 0; i < HD; i += 8) {
        __nv_bfloat162 a = __floats2bfloat162_rn(o[i], o[i + 1]), c2 = __floats2bfloat162_rn(o[i + 2], o[i + 3]);
        __nv_bfloat162 d = __floats2bfloat162_rn(o[i + 4], o[i + 5]), e = __floats2bfloat162_rn(o[i + 6], o[i + 7]);
        uint4 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&a); pk.y = *reinterpret_cast<uint32_t*>(&c2);
        pk.z = *reinterpret_cast<uint32_t*>(&d); pk.w = *reinterpret_cast<uint32_t*>(&e);
        *reinterpret_cast<uint4*>(dst + i) = pk;
      }
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
  const size_t Np = (size_t)(N + 7) / 8 * 8;
  return (size_t)B * H * HD * Np * 2;
}

// qkv bf16 [B,N,3*H*64] -> out bf16 [B,N,H*64]; vt_ws holds the transposed V copy.
int vit_attention_tc(const void* qkv, void* out, void* vt_ws, int B, int N, int H, cudaStream_t st) {
  const int Np = (N + 7) / 8 * 8;
  {
    dim3 grid(cdiv(Np, 64), H, B);
    launch_pdl(transpose_v_kernel, grid, dim3(64, 4), 0, st, (const __nv_bfloat16*)qkv, (__nv_bfloat16*)vt_ws, N, Np, H);
    PIO_LAUNCHED();
  }
  CUtensorMap mqk, mvt;
  PIO_TRY(make_map_2d(&mqk, qkv, (long long)B * N, 3 * H * HD, 3 * H * HD, BQ, HD));
  PIO_TRY(make_map_2d(&mvt, vt_ws, (long long)B * H * HD, Np, Np, HD, 64));
  static bool attr_set = false;
  if (!attr_set) {
    PIO_CUDA(cudaFuncSetAttribute(vit_attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
    attr_set = true;
  }
  dim3 grid(cdiv(N, BQ), H, B);
  const float scale_log2e = 0.125f * 1.4426950408889634f;
  launch_pdl(vit_attention_tc_kernel, grid, dim3(ATT_THREADS), ATT_SMEM, st, mqk, mvt, (__nv_bfloat16*)out, N, H, scale_log2e);
  PIO_LAUNCHED();
  return PIO_OK;
}

}  // namespace pio

// gemm2_sm100.cu -- the 2-CTA (cta_group::2) variant of the tcgen05 GEMM: a CTA pair (cluster 2x1x1, one TPC) works
// on a 256 x 256 output tile.  Each CTA loads ITS 128 rows of A and ITS 128 rows of W per k-block (32 KB per stage
// instead of 48 KB: 1.5x less L2->SM operand traffic, which bounds the 1-CTA kernel at ~930 TFLOP/s, and a 6-deep ring),
// the leader CTA issues one tcgen05.mma.cta_group::2 (UMMA 256 x 256 x 16) that reads both CTAs' shared memory and
// writes both CTAs' tensor memory, and each CTA runs the epilogue for its own 128 rows.
//
// Synchronisation (all mbarriers live at the same shared-memory offsets in both CTAs):
//   full[s]    leader only; both CTAs' TMA loads complete_tx on it (barrier address with the peer bit cleared)
//   empty[s]   both CTAs; tcgen05.commit.cta_group::2 ... multicast::cluster (mask 0b11) from the leader
//   tfull[a]   both CTAs; commit multicast after the last k-block
//   tempty[a]  leader only; 8 local + 8 remote (mapa) epilogue-warp arrivals
#include "gemm_epilogue.cuh"

namespace pio {
using namespace tc;
namespace {

constexpr int BM2 = 128;            // rows per CTA (256 per pair)
constexpr int BN2 = 256;            // columns per pair tile
constexpr int BNH = 128;            // W rows each CTA loads
constexpr int BK2 = 64;
constexpr int STAGES2 = 5;
constexpr int A2_BYTES = BM2 * BK2 * 2;   // 16 KB
constexpr int B2_BYTES = BNH * BK2 * 2;   // 16 KB
constexpr int STAGE2_BYTES = A2_BYTES + B2_BYTES;
constexpr int NUM_EPI_WARPS2 = 8;
constexpr int OUT_STAGE_BYTES = 4096;     // one 32-row x 128-byte TMA-store box per epilogue warp
constexpr int SMEM2_BYTES = STAGES2 * STAGE2_BYTES + NUM_EPI_WARPS2 * OUT_STAGE_BYTES + 1024;
constexpr int NUM_THREADS2 = 64 + NUM_EPI_WARPS2 * 32;
constexpr uint32_t TMEM_COLS2 = 512;      // 2 accumulators x 256 columns
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address -> CTA 0

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {  // arrives on `bar` in BOTH CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t local_bar, uint32_t target_cta) {
  asm volatile(
      "{\n\t"
      ".reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t"
      "}" ::"r"(local_bar), "r"(target_cta)
      : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS2, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                const __grid_constant__ CUtensorMap map_c, int store_mode, void* C, int M, int N, int K, int ldc, int c_dt, Epilogue epi,
                int prefetch_w) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ __align__(8) uint64_t bars[2 * STAGES2 + 4];
  __shared__ uint32_t tmem_slot_var;
  __shared__ __align__(16) float s_scale[BN2], s_bias[BN2], s_gamma[BN2];
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES2 + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES2 + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES2 + 2 + s); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int m_blocks = (M + 2 * BM2 - 1) / (2 * BM2), n_blocks = (N + BN2 - 1) / BN2;
  const int num_tiles = m_blocks * n_blocks;
  const int k_blocks = (K + BK2 - 1) / BK2;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_c) : "memory");
    for (int s = 0; s < STAGES2; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 2 * NUM_EPI_WARPS2); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc_2sm(smem_u32(&tmem_slot_var), TMEM_COLS2);
  tc_fence_before();
  cluster_sync_all();  // barriers and tensor memory of BOTH CTAs are ready
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&tmem_slot_var);
  pdl_launch_dependents();
  // pdl_wait() is per role (see gemm_sm100.cu): weight tiles are fetched ahead of it, A tiles and C accesses after it

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs, own halves)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int pre = 0;
      if (prefetch_w && pair < num_tiles) {  // PioLinear.w_static, see gemm_sm100.cu
        const int n0 = (pair % n_blocks) * BN2 + (int)rank * BNH;
        pre = min(STAGES2, k_blocks);
        for (int i = 0; i < pre; ++i) {
          const uint32_t lbar = full_bar(i) & PEER_MASK;
          if (leader) mbar_expect_tx(full_bar(i), 2 * STAGE2_BYTES);
          tma_load_2d_2sm(smem_base + i * STAGE2_BYTES + A2_BYTES, &map_w, lbar, i * BK2, n0);
        }
      }
      pdl_wait();
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int m0 = (tile / n_blocks) * (2 * BM2) + (int)rank * BM2;
        const int n0 = (tile % n_blocks) * BN2 + (int)rank * BNH;
        for (int kb = 0; kb < k_blocks; ++kb) {
          const bool w_in_flight = (tile == pair) && (kb < pre);
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = smem_base + stage * STAGE2_BYTES, sb = sa + A2_BYTES;
          const uint32_t lbar = full_bar(stage) & PEER_MASK;  // the LEADER's full barrier
          if (!w_in_flight) {
            if (leader) mbar_expect_tx(full_bar(stage), 2 * STAGE2_BYTES);
            tma_load_2d_2sm(sb, &map_w, lbar, kb * BK2, n0);
          }
          tma_load_2d_2sm(sa, &map_a, lbar, kb * BK2, m0);
          if (++stage == STAGES2) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (leader) {
      constexpr uint32_t idesc = make_idesc(2 * BM2, BN2);  // 256 x 256
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(tempty_bar(as), aphase ^ 1);  // both CTAs' epilogues have drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN2;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          if (lane == 0) {
            const uint32_t sa = smem_base + stage * STAGE2_BYTES, sb = sa + A2_BYTES;
            const uint64_t adesc = make_smem_desc(sa), bdesc = make_smem_desc(sb);
#pragma unroll
            for (int k = 0; k < BK2 / 16; ++k) umma_f16_2sm(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
            umma_commit_2sm(empty_bar(stage));
            if (kb == k_blocks - 1) umma_commit_2sm(tfull_bar(as));
          }
          __syncwarp();
          if (++stage == STAGES2) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..9, both CTAs, own rows)
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    const int et = threadIdx.x - 64;
    const TmaOut to{&map_c, smem_base + STAGES2 * STAGE2_BYTES + (warp - 2) * OUT_STAGE_BYTES, store_mode};
    pdl_wait();
    int it = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int m0 = (tile / n_blocks) * (2 * BM2) + (int)rank * BM2, n0 = (tile % n_blocks) * BN2;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      for (int c = et; c < BN2; c += NUM_EPI_WARPS2 * 32) {
        const int n = n0 + c;
        const bool ok = n < N;
        s_scale[c] = epi.alpha * ((epi.colscale && ok) ? __ldg(epi.colscale + n) : 1.0f);
        s_bias[c] = (epi.bias && ok) ? __ldg(epi.bias + n) : 0.0f;
        s_gamma[c] = (epi.gamma && ok) ? __ldg(epi.gamma + n) : 1.0f;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
      const int m = m0 + quarter * 32 + lane;
      epilogue_tile<BN2, true>(tmem_base + as * BN2, quarter, lane, half, m, M, n0, N, (tile % n_blocks) * 2 + half, C, ldc, c_dt, epi,
                    s_scale, s_bias, s_gamma, to);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(tempty_bar(as));
        else mbar_arrive_remote(tempty_bar(as), 0);
      }
    }
    stage_drain(lane);
  }

  tc_fence_before();
  cluster_sync_all();  // nobody may still be using the pair's tensor memory / barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, TMEM_COLS2);
  }
}

}  // namespace

// true when the shape fills the machine with 256 x 256 pair tiles
bool linear_tc2_eligible(const PioLinear& p) {
  return (long long)cdiv(p.M, 2 * BM2) * cdiv(p.N, BN2) >= kNumSMs / 2;
}

int linear_tc2(const PioLinear& p, cudaStream_t st) {
  CUtensorMap ma, mw, mc;
  PIO_TRY(make_map_2d(&ma, p.A, p.M, p.K, p.lda, BM2, BK2));
  PIO_TRY(make_map_2d(&mw, p.W, p.N, p.K, p.ldw, BNH, BK2));
  const int store_mode = tma_store_enabled() ? pick_store_mode(p) : STORE_DIRECT;
  if (store_mode != STORE_DIRECT) PIO_TRY(make_map_out(&mc, p.C, p.M, p.N, p.ldc, p.c_dt));
  else mc = ma;
  static SmemAttrOnce once;
  PIO_CUDA(once.ensure(gemm_tc2_kernel, SMEM2_BYTES));
  const int tiles = cdiv(p.M, 2 * BM2) * cdiv(p.N, BN2);
  const int pairs = tiles < kNumSMs / 2 ? tiles : kNumSMs / 2;
  launch_pdl_k(PDL_KIND_GEMM2, gemm_tc2_kernel, dim3(2 * pairs), dim3(NUM_THREADS2), SMEM2_BYTES, st, ma, mw, mc, store_mode, p.C, p.M, p.N, p.K, p.ldc,
             p.c_dt, make_epilogue(p), p.w_static ? 1 : 0);
  PIO_LAUNCHED();
  return PIO_OK;
}

}  // namespace pio

// gemm_sm100.cu -- bf16 x bf16 -> fp32 GEMM on the Blackwell 5th-gen tensor cores (sm_100a).
//
//   C[map(m), n] = epilogue( sum_k A[m,k] * W[n,k] )          A [M,K], W [N,K], both K-contiguous
//
// Structure (one persistent CTA per SM, 192 threads, warp specialised):
//   warp 0      TMA producer : cp.async.bulk.tensor.2d (128-byte swizzle) of a 128 x 64 A tile and a
//               BN x 64 W tile per stage into a STAGES-deep shared-memory ring, completion on mbarriers
//   warp 1      MMA issuer   : one elected lane issues tcgen05.mma.cta_group::1.kind::f16 (UMMA 128 x BN x 16)
//               straight from shared memory into a TMEM accumulator; tcgen05.commit releases the ring slot
//               and, after the last k-block, publishes the accumulator
//   warps 2..9  epilogue     : tcgen05.ld the fp32 accumulator (lane == row; warp w reads TMEM lane quarter w%4 and
//               one half of the columns), apply the epilogue and store.  Per-column vectors (colscale, bias, LayerScale)
//               are staged once per tile in shared memory; residual rows are fetched with 16-byte loads issued ahead of
//               the TMEM wait.  The accumulator is double buffered in TMEM (2 x BN columns) so the epilogue of tile i
//               overlaps the MMAs of tile i+1.
// No CUTLASS: descriptors are built by hand (field layout checked against cute/arch/mma_sm100_desc.hpp).
#include "tc_ptx.cuh"
#include "gemm_epilogue.cuh"
#include <map>
#include <mutex>
#include <utility>

#include <algorithm>
#include <mutex>
#include <stdlib.h>

namespace pio {
using namespace tc;
namespace {

constexpr int BM = 128;
constexpr int BK = 64;                 // 64 bf16 = 128 bytes = one swizzle-128B row
constexpr int UMMA_K = 16;
constexpr int NUM_EPI_WARPS = 8;          // 2 per TMEM lane quarter, each owning half of the tile's columns
constexpr int NUM_THREADS = 64 + NUM_EPI_WARPS * 32;

// Debug build only (-DPIO_GEMM_TRACE, tools/gemm_trace.py): %globaltimer stamps of the roles of every CTA of the LAST launch.
#ifdef PIO_GEMM_TRACE
__device__ unsigned long long g_gemm_trace[160 * 16];
__device__ int g_gemm_trace_nk[2];  // only launches of this (N, K) leave stamps (0, 0: every launch)
__device__ __forceinline__ void trace_stamp(int slot, int N, int K) {
  if ((g_gemm_trace_nk[0] | g_gemm_trace_nk[1]) != 0 && (N != g_gemm_trace_nk[0] || K != g_gemm_trace_nk[1])) return;
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  g_gemm_trace[blockIdx.x * 16 + slot] = t;
}
#define PIO_TRACE(slot) trace_stamp(slot, N, K)
#else
#define PIO_TRACE(slot) ((void)0)
#endif

template <int BN> struct Cfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // the 256-wide tile keeps 4 x 48 KB of operand ring and therefore direct stores; the narrower tiles trade one ring
  // slot for the TMA-store staging boxes (one 4 KB box per epilogue warp)
  static constexpr bool TMA_OUT = BN != 256;
  static constexpr int STAGES = (BN == 256) ? 4 : (BN == 192 ? 4 : (BN == 128 ? 5 : 7));
  static constexpr int TMEM_COLS = BN == 192 ? 512 : 2 * BN;  // power of two: 128 / 256 / 512
  static constexpr int OUT_BYTES = TMA_OUT ? NUM_EPI_WARPS * 4096 : 0;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + OUT_BYTES + 1024 /*manual 1024-byte alignment*/;
};

template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
               const __grid_constant__ CUtensorMap map_c, int store_mode, void* C, int M, int N, int K, int ldc, int c_dt, Epilogue epi,
               int k_splits, int prefetch_w) {
  using cfg = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the 128B-swizzle atoms
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ __align__(8) uint64_t bars[2 * cfg::STAGES + 4];
  __shared__ uint32_t tmem_slot_var;
  __shared__ __align__(16) float s_scale[BN], s_bias[BN], s_gamma[BN];  // per-tile column vectors of the epilogue
  __shared__ int s_ticket;
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (cfg::STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * cfg::STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * cfg::STAGES + 2 + s); };
  const uint32_t tmem_slot = smem_u32(&tmem_slot_var);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) PIO_TRACE(0);  // kernel entry
  const int m_blocks = (M + BM - 1) / BM, n_blocks = (N + BN - 1) / BN;
  // split-K (k_splits > 1, accumulate-only epilogue through the copy engine's reduce-add): work item w = split * tiles + tile
  // covers k-blocks [split * kbs, min(k_blocks, (split + 1) * kbs)); the host guarantees that no split is empty
  const int out_tiles = m_blocks * n_blocks;
  const int num_tiles = out_tiles * k_splits;
  const int k_blocks = (K + BK - 1) / BK;
  const int kbs = (k_blocks + k_splits - 1) / k_splits;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    for (int s = 0; s < cfg::STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), NUM_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&tmem_slot_var);
  if (threadIdx.x == 0) PIO_TRACE(1);  // barriers initialised, tensor memory allocated
  pdl_launch_dependents();  // the next kernel may start its own prologue as soon as this grid's CTAs retire
  // pdl_wait() is per role: with prefetch_w (PioLinear.w_static: W is not written by a kernel still in flight -- weights,
  // bank, wte) the producer fetches WEIGHT tiles first; A and the residual always wait; the epilogue waits before it touches
  // C, the MMA warp never touches global memory.

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int pre = 0;  // ring slots of the first work item whose W tile is already in flight (issued ahead of pdl_wait)
      if (prefetch_w && (int)blockIdx.x < num_tiles) {  // only when the caller vouches that W is static (PioLinear.w_static)
        const int tile = blockIdx.x % out_tiles, split = blockIdx.x / out_tiles;
        const int n0 = (tile % n_blocks) * BN, kb0 = split * kbs, kb1 = min(k_blocks, kb0 + kbs);
        pre = min(cfg::STAGES, kb1 - kb0);
        for (int i = 0; i < pre; ++i) {
          mbar_expect_tx(full_bar(i), cfg::STAGE_BYTES);
          tma_load_2d(smem_base + i * cfg::STAGE_BYTES + cfg::A_BYTES, &map_w, full_bar(i), (kb0 + i) * BK, n0);
        }
      }
      pdl_wait();  // the previous kernel's output (our A operand) is complete and visible
      PIO_TRACE(2);  // producer: the previous grid has completed
      for (int w = blockIdx.x; w < num_tiles; w += gridDim.x) {
        const int tile = w % out_tiles, split = w / out_tiles;
        const int m0 = (tile / n_blocks) * BM, n0 = (tile % n_blocks) * BN;
        const int kb0 = split * kbs, kb1 = min(k_blocks, kb0 + kbs);
        for (int kb = kb0; kb < kb1; ++kb) {
          const bool w_in_flight = (w == (int)blockIdx.x) && (kb - kb0 < pre);
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = smem_base + stage * cfg::STAGE_BYTES, sb = sa + cfg::A_BYTES;
          if (!w_in_flight) {
            mbar_expect_tx(full_bar(stage), cfg::STAGE_BYTES);
            tma_load_2d(sb, &map_w, full_bar(stage), kb * BK, n0);
          }
          tma_load_2d(sa, &map_a, full_bar(stage), kb * BK, m0);
          if (++stage == cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
      PIO_TRACE(3);  // producer: last load issued
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = make_idesc(BM, BN);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int w = blockIdx.x; w < num_tiles; w += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(tempty_bar(as), aphase ^ 1);  // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + as * BN;
      const int kb0 = (w / out_tiles) * kbs, kb1 = min(k_blocks, kb0 + kbs);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        if (lane == 0) {
          if (it == 0 && kb == kb0) PIO_TRACE(4);  // MMA: first stage has landed
          const uint32_t sa = smem_base + stage * cfg::STAGE_BYTES, sb = sa + cfg::A_BYTES;
          const uint64_t adesc = make_smem_desc(sa), bdesc = make_smem_desc(sb);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advance 32 bytes (16 bf16) along K inside the swizzle atom: +2 in the (addr >> 4) field
            umma_f16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb > kb0) || (k != 0));
          }
          umma_commit(empty_bar(stage));                       // ring slot reusable once these MMAs retire
          if (kb == kb1 - 1) { umma_commit(tfull_bar(as)); PIO_TRACE(5); }  // accumulator complete (stamp: last commit ISSUED, last tile wins)
        }
        __syncwarp();
        if (++stage == cfg::STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..9)
    const int quarter = warp & 3;              // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;          // which half of the tile's columns
    const int et = threadIdx.x - 64;           // 0..255 among the epilogue threads
    const TmaOut to{&map_c, smem_base + cfg::STAGES * cfg::STAGE_BYTES + (warp - 2) * 4096, cfg::TMA_OUT ? store_mode : STORE_DIRECT};
    pdl_wait();  // C / the residual may still be read or written by the previous kernel
    int it = 0;
    for (int w = blockIdx.x; w < num_tiles; w += gridDim.x, ++it) {
      const int tile = w % out_tiles;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int m0 = (tile / n_blocks) * BM, n0 = (tile % n_blocks) * BN;
      // stage this tile's column vectors (defaults make every epilogue the same arithmetic)
      asm volatile("bar.sync 1, 256;" ::: "memory");  // previous tile's readers are done
      for (int c = et; c < BN; c += NUM_EPI_WARPS * 32) {
        const int n = n0 + c;
        const bool ok = n < N;
        s_scale[c] = epi.alpha * ((epi.colscale && ok) ? __ldg(epi.colscale + n) : 1.0f);
        s_bias[c] = (epi.bias && ok) ? __ldg(epi.bias + n) : 0.0f;
        s_gamma[c] = (epi.gamma && ok) ? __ldg(epi.gamma + n) : 1.0f;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
      if (et == 0) PIO_TRACE(it == 0 ? 6 : 7);  // epilogue: accumulator of the first / a later tile is ready
      const int m = m0 + quarter * 32 + lane;
      const uint32_t tacc = tmem_base + as * BN;
      if (store_mode == STORE_SPLIT_FIXUP) {
        // Deterministic split-K: every split leaves its raw fp32 partial tile in the workspace; the split that arrives last at
        // the tile's counter sums all partials in split order (so the result does not depend on the arrival order) and runs
        // the epilogue.  Only for GEMMs with too few output tiles to occupy the machine (decode steps at small batch).
        float* part = epi.split_ws + (size_t)w * (BM * BN);
#pragma unroll 1
        for (int c = half * (BN / 2); c < (half + 1) * (BN / 2); c += 32) {
          uint32_t r[32];
          tmem_ld32(tacc + ((uint32_t)(quarter * 32) << 16) + c, r);
          tmem_ld_wait();
          float4* dst = reinterpret_cast<float4*>(part + (size_t)(quarter * 32 + lane) * BN + c);
#pragma unroll
          for (int i = 0; i < 8; ++i)
            dst[i] = make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]), __uint_as_float(r[4 * i + 2]),
                                 __uint_as_float(r[4 * i + 3]));
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(as));
        __threadfence();  // this thread's partial values are visible device-wide before the ticket is taken
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (et == 0) s_ticket = atomicAdd(epi.split_cnt + tile, 1);
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (s_ticket == k_splits - 1) {
          __threadfence();
          for (int idx = et; idx < BM * (BN / 4); idx += NUM_EPI_WARPS * 32) {
            const int row = idx / (BN / 4), c4 = (idx % (BN / 4)) * 4;
            const int mm = m0 + row;
            if (mm >= M) break;  // rows ascend with idx
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int sp = 0; sp < k_splits; ++sp) {  // fixed order; .cg: the partials were written by other SMs
              const float4 t = __ldcg(reinterpret_cast<const float4*>(epi.split_ws + ((size_t)(sp * out_tiles + tile) * BM + row) * BN + c4));
              acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
            }
            const long long orow = epi.out_row(mm);
            const float av[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int n = n0 + c4 + j;
              if (n >= N) continue;
              const float v = epi.apply(av[j], mm, orow, n);
              if (c_dt == PIO_DT_F32) reinterpret_cast<float*>(C)[orow * ldc + n] = v;
              else reinterpret_cast<__nv_bfloat16*>(C)[orow * ldc + n] = __float2bfloat16(v);
            }
          }
          if (et == 0) epi.split_cnt[tile] = 0;  // ready for the next launch
        }
        continue;
      }
      epilogue_tile<BN>(tacc, quarter, lane, half, m, M, n0, N, (tile % n_blocks) * 2 + half, C, ldc, c_dt, epi, s_scale, s_bias, s_gamma, to);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(as));
      if (et == 0) PIO_TRACE(8);  // epilogue: tile processed, stores issued (last tile wins)
    }
    stage_drain(lane);
    if (et == 0) PIO_TRACE(9);    // epilogue: bulk stores have read their staging boxes
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, cfg::TMEM_COLS);
  }
  if (threadIdx.x == 0) PIO_TRACE(10);  // exit
}

// ------------------------------------------------------------------------------------------ host side
// Workspace of the deterministic split-K: one per (device, stream), allocated on first use (work items never exceed the SM count,
// a partial tile is at most 128 x 192 floats) and kept for the life of the process.
struct SplitScratch { float* ws; int* cnt; };
inline bool split_fixup_enabled() {
  static const bool on = [] { const char* e = getenv("PIO_GEMM_SPLITK"); return !(e && e[0] == '0'); }();
  return on;
}
inline int split_min_kblocks() {
  // below 12 k-blocks a split could come out empty (k_blocks / 6 == 0 would divide by zero further down) and never pays
  static const int v = [] { const char* e = getenv("PIO_GEMM_SPLITK_MIN_KB"); return std::max(12, e ? atoi(e) : 24); }();
  return v;
}
std::mutex g_split_mu;
std::map<std::pair<int, cudaStream_t>, SplitScratch> g_split_cache;
// found = false (and no allocation) while `st` is being captured into a CUDA graph and no scratch exists yet: the caller then
// runs the GEMM without the split (cudaMalloc is not allowed during a capture).
inline int split_scratch(cudaStream_t st, SplitScratch* out, bool* found) {
  int dev = 0;
  PIO_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(g_split_mu);
  auto it = g_split_cache.find({dev, st});
  if (it == g_split_cache.end()) {
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    PIO_CUDA(cudaStreamIsCapturing(st, &cap));
    if (cap != cudaStreamCaptureStatusNone) { *found = false; return PIO_OK; }
    SplitScratch sc;
    PIO_CUDA(cudaMalloc(&sc.ws, (size_t)kNumSMs * BM * 192 * sizeof(float)));
    PIO_CUDA(cudaMalloc(&sc.cnt, kNumSMs * sizeof(int)));
    // zeroed ON THE LAUNCH STREAM: torch side streams are non-blocking, nothing would order a legacy-stream memset before
    // the first kernel's tickets.  The kernels leave the counters at zero again.
    PIO_CUDA(cudaMemsetAsync(sc.cnt, 0, kNumSMs * sizeof(int), st));
    it = g_split_cache.emplace(std::make_pair(dev, st), sc).first;
  }
  *out = it->second;
  *found = true;
  return PIO_OK;
}

template <int BN>
int launch(const PioLinear& p, cudaStream_t st) {
  using cfg = Cfg<BN>;
  CUtensorMap ma, mw, mc;
  PIO_TRY(make_map_2d(&ma, p.A, p.M, p.K, p.lda, BM, BK));
  PIO_TRY(make_map_2d(&mw, p.W, p.N, p.K, p.ldw, BN, BK));
  // a bf16 staging box is 64 columns wide; with the 64-wide tile an epilogue warp owns only 32 columns
  const bool box_fits = BN >= 128 || p.c_dt != PIO_DT_BF16;
  const int store_mode = (cfg::TMA_OUT && box_fits && tma_store_enabled()) ? pick_store_mode(p) : STORE_DIRECT;
  if (store_mode != STORE_DIRECT) PIO_TRY(make_map_out(&mc, p.C, p.M, p.N, p.ldc, p.c_dt));
  else mc = ma;
  static SmemAttrOnce once;
  PIO_CUDA(once.ensure(gemm_tc_kernel<BN>, cfg::SMEM_BYTES));
  const int tiles = cdiv(p.M, BM) * cdiv(p.N, BN);
  // split-K: a long-K accumulation (C += A W^T, nothing else in the epilogue) with too few output tiles to fill the machine --
  // the recombination GEMM of the memory projection for a handful of queries -- is cut along K; the partial tiles meet in C
  // through the bulk reduce-add.
  int k_splits = 1;
  int mode = store_mode;
  const int k_blocks = cdiv(p.K, BK);
  Epilogue epi = make_epilogue(p);
  if (store_mode == STORE_TMA_ADD && p.bias == nullptr && p.gamma == nullptr && p.colscale == nullptr && p.act == PIO_ACT_NONE &&
      tiles * 2 <= kNumSMs && k_blocks >= 64) {
    k_splits = std::min(kNumSMs / tiles, k_blocks / 32);
    const int kbs = cdiv(k_blocks, k_splits);
    k_splits = cdiv(k_blocks, kbs);  // no empty split
  } else if (split_fixup_enabled() && BN <= 192 && tiles * 2 <= kNumSMs && k_blocks >= split_min_kblocks() && p.act == PIO_ACT_NONE &&
             p.argmax_val == nullptr && p.exp_ref == nullptr) {
    // Deterministic split-K for long-K GEMMs with a handful of output tiles (decode steps at small batch: each CTA would
    // stream all of K through one SM's L2 port).  Partials meet in a workspace and are summed in split order.
    SplitScratch sc;
    bool have = false;
    PIO_TRY(split_scratch(st, &sc, &have));
    k_splits = have ? std::max(1, std::min(kNumSMs / tiles, k_blocks / 6)) : 1;
    const int kbs = cdiv(k_blocks, k_splits);
    k_splits = cdiv(k_blocks, kbs);
    if (k_splits > 1) {
      mode = STORE_SPLIT_FIXUP;
      epi.split_ws = sc.ws;
      epi.split_cnt = sc.cnt;
    }
  }
  const int work = tiles * k_splits;
  const int grid = work < kNumSMs ? work : kNumSMs;
  launch_pdl_k(PDL_KIND_GEMM, gemm_tc_kernel<BN>, dim3(grid), dim3(NUM_THREADS), cfg::SMEM_BYTES, st, ma, mw, mc, mode, p.C, p.M, p.N, p.K, p.ldc,
             p.c_dt, epi, k_splits, p.w_static ? 1 : 0);
  PIO_LAUNCHED();
  return PIO_OK;
}

}  // namespace

int argmax_slabs_tc(int M, int N) { (void)M; return 2 * cdiv(N, 256); }

// frees the split-K scratch of every (device, stream) this process used; the streams must be idle (pio_release_scratch)
void release_split_scratch() {
  std::lock_guard<std::mutex> lock(g_split_mu);
  int cur = 0;
  cudaGetDevice(&cur);
  for (auto& kv : g_split_cache) {
    cudaSetDevice(kv.first.first);
    cudaFree(kv.second.ws);
    cudaFree(kv.second.cnt);
  }
  g_split_cache.clear();
  cudaSetDevice(cur);
}

int linear_tc(const PioLinear& p, cudaStream_t st) {
  PIO_CHECK(p.a_dt == PIO_DT_BF16, "tcgen05 GEMM needs bf16 operands");
  PIO_CHECK(p.lda % 8 == 0 && p.ldw % 8 == 0, "tcgen05 GEMM: lda/ldw must be multiples of 8 (16-byte TMA strides)");
  PIO_CHECK((((uintptr_t)p.A) & 15) == 0 && (((uintptr_t)p.W) & 15) == 0, "tcgen05 GEMM: operands must be 16-byte aligned");
  PIO_CHECK(p.K > 0, "tcgen05 GEMM: K must be positive");
  PIO_CHECK(p.act >= PIO_ACT_NONE && p.act <= PIO_ACT_GELU_NEW, "tcgen05 GEMM: activation %d is built for the fp32 mode only", p.act);
  if (p.M == 0 || p.N == 0) return PIO_OK;
  PIO_CHECK(p.argmax_val == nullptr || (p.argmax_idx && p.argmax_ld >= argmax_slabs_tc(p.M, p.N)),
            "tcgen05 GEMM: fused arg-max needs val/idx (and optionally sumexp) buffers with ld >= pio_argmax_slabs()");
  PIO_CHECK(p.exp_ref == nullptr || (p.exp_psum && p.exp_pmax && p.c_dt == PIO_DT_BF16 && p.exp_ld >= argmax_slabs_tc(p.M, p.N)),
            "tcgen05 GEMM: fused exp needs bf16 C and psum/pmax buffers with ld >= pio_argmax_slabs()");
  static const bool use_2cta = [] { const char* e = getenv("PIO_GEMM_2CTA"); return !(e && e[0] == '0'); }();
  const bool slabs256 = p.argmax_val != nullptr || p.exp_ref != nullptr;  // slab count is defined for 256-wide tiles
  // Tile shape: fewest (waves x per-tile MMA time), with a small penalty for the narrower, less efficient tiles.
  // per-SM cost of one tile ~ its width (every CTA owns 128 rows); 74 CTA pairs or 148 CTAs work per wave.
  const long long mt = cdiv(p.M, BM);
  // Round 2 (second session): the main loop of every tile shape is bound by operand delivery, not by the MMA (tools/gemm_trace.py:
  // 315 ns per k-block for a 128 x 192 tile against 200 ns of tensor-core time), so a wave costs ~ the bytes a CTA fetches per
  // k-block -- (128 + BN) rows of 128 bytes, 256 for a CTA of the pair kernel -- not the tile width.  Measured on the shapes this
  // flips (M = 4096, N = 768 -> the pair kernel instead of 128 x 192 tiles): memory projection 7.66 -> 6.91 ms, dense decode
  // 25.14 -> 24.86 ms.  PIO_GEMM_COST=old keeps the round-1 model (width x penalty) for A/B runs.
  static const bool old_cost = [] { const char* e = getenv("PIO_GEMM_COST"); return e && !strcmp(e, "old"); }();
  auto cost = [&](int bn, double penalty) {
    return (double)cdiv(mt * cdiv(p.N, bn), kNumSMs) * (old_cost ? (double)bn : (double)(BM + bn)) * penalty;
  };
  // (round 1 sent only GEMMs with a full wave of pair tiles to the pair kernel; with the traffic model the cost decides)
  double c2 = (use_2cta && (linear_tc2_eligible(p) || !old_cost))
                  ? (double)cdiv((long long)cdiv(p.M, 256) * cdiv(p.N, 256), kNumSMs / 2) * (old_cost ? 256 * 0.95 : 266.0) : 1e30;
  // long-K GEMMs with a handful of output tiles are cut along K by the 1-CTA kernel (launch<BN>): the pair kernel has no split-K
  if (!old_cost && cdiv(p.M, BM) * cdiv(p.N, 192) * 2 <= kNumSMs && cdiv(p.K, BK) >= split_min_kblocks()) c2 = 1e30;
  if (!old_cost && !linear_tc2_eligible(p) && p.M <= BM) c2 = 1e30;  // a pair tile is 256 rows: with <= 128 the second CTA idles
  const double c256 = cost(256, old_cost ? 1.0 : 1.15), c192 = slabs256 ? 1e30 : cost(192, old_cost ? 1.03 : 1.0),
               c128 = slabs256 ? 1e30 : cost(128, old_cost ? 1.10 : 1.0), c64 = slabs256 ? 1e30 : cost(64, old_cost ? 1.30 : 1.0);
  const double best = std::min(std::min(std::min(c2, c256), std::min(c192, c128)), c64);
  if (const char* force = getenv("PIO_GEMM_TILE")) {  // A/B testing: 2cta | 256 | 192 | 128 | 64
    if (!strcmp(force, "2cta")) return linear_tc2(p, st);
    if (!strcmp(force, "256") || slabs256) return launch<256>(p, st);
    if (!strcmp(force, "192")) return launch<192>(p, st);
    if (!strcmp(force, "128")) return launch<128>(p, st);
    if (!strcmp(force, "64")) return launch<64>(p, st);
  }
  if (best == c2) return linear_tc2(p, st);
  if (best == c256) return launch<256>(p, st);
  if (best == c192) return launch<192>(p, st);
  if (best == c128) return launch<128>(p, st);
  return launch<64>(p, st);
}

namespace tc {
namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

}  // namespace
// 2-D bf16 tensor [rows, cols] with row stride ld elements; box = box_rows x 64 columns, 128-byte swizzle.
int make_map_2d(CUtensorMap* map, const void* ptr, long long rows, long long cols, long long ld, int box_rows, int box_cols) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return fail(PIO_ECUDA, "cuTensorMapEncodeTiled is not available (no CUDA driver?)");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(PIO_ECUDA, "cuTensorMapEncodeTiled failed with %d (rows %lld cols %lld ld %lld)", (int)r, rows, cols, ld);
  return PIO_OK;
}

int make_map_out(CUtensorMap* map, const void* ptr, long long rows, long long cols, long long ld, int c_dt) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return fail(PIO_ECUDA, "cuTensorMapEncodeTiled is not available (no CUDA driver?)");
  const bool bf = c_dt == PIO_DT_BF16;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * (bf ? 2 : 4)};
  cuuint32_t box[2] = {(cuuint32_t)(bf ? 64 : 32), 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, bf ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims,
                   strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(PIO_ECUDA, "cuTensorMapEncodeTiled (output) failed with %d (rows %lld cols %lld ld %lld)", (int)r, rows, cols, ld);
  return PIO_OK;
}
bool tma_store_enabled() {
  static const bool on = [] { const char* e = getenv("PIO_GEMM_TMA_STORE"); return !(e && e[0] == '0'); }();
  return on;
}

int make_map_f32_3d(CUtensorMap* map, const void* ptr, long long d0, long long d1, long long d2, long long stride1_bytes,
                    long long stride2_bytes, int box0, int box1) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return fail(PIO_ECUDA, "cuTensorMapEncodeTiled is not available (no CUDA driver?)");
  cuuint64_t dims[3] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)d2};
  cuuint64_t strides[2] = {(cuuint64_t)stride1_bytes, (cuuint64_t)stride2_bytes};
  cuuint32_t box[3] = {(cuuint32_t)box0, (cuuint32_t)box1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(PIO_ECUDA, "cuTensorMapEncodeTiled (f32 3d) failed with %d", (int)r);
  return PIO_OK;
}

}  // namespace tc

}  // namespace pio

#ifdef PIO_GEMM_TRACE
extern "C" int pio_debug_gemm_trace_read(unsigned long long* host, int n) {
  return cudaMemcpyFromSymbol(host, pio::g_gemm_trace, sizeof(unsigned long long) * (size_t)std::min(n, 160 * 16)) == cudaSuccess ? 0 : -1;
}
extern "C" int pio_debug_gemm_trace_clear(int n, int k) {
  static unsigned long long zeros[160 * 16];
  const int nk[2] = {n, k};
  if (cudaMemcpyToSymbol(pio::g_gemm_trace_nk, nk, sizeof(nk)) != cudaSuccess) return -1;
  return cudaMemcpyToSymbol(pio::g_gemm_trace, zeros, sizeof(zeros)) == cudaSuccess ? 0 : -1;
}
#endif

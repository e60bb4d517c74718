// gemm_simt.cu -- fp32 FFMA GEMM  C = epi(A W^T), both operands K-contiguous.
// This is the PIO_FP32 ("fp32 parity") arithmetic mode: the reference computes in fp32
// (SURVEY.md 8a), and greedy-token parity needs fp32-faithful lm_head / attention.  The
// throughput mode is the tcgen05 kernel in gemm_sm100.cu.
#include "common.cuh"

namespace pio {

namespace {
constexpr int BM = 128, BN = 128, BK = 16, TM = 8, TN = 8;
constexpr int THREADS = 256;

__device__ __forceinline__ void store_out(void* C, int c_dt, long long idx, float v) {
  if (c_dt == PIO_DT_F32)
    reinterpret_cast<float*>(C)[idx] = v;
  else
    reinterpret_cast<__nv_bfloat16*>(C)[idx] = __float2bfloat16(v);
}

__global__ void __launch_bounds__(THREADS) sgemm_tn_kernel(const float* __restrict__ A, const float* __restrict__ W,
                                                           void* C, int M, int N, int K, int lda, int ldw, int ldc,
                                                           int c_dt, Epilogue epi) {
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int lrow = tid >> 2;        // 0..63
  const int lk = (tid & 3) * 4;     // 0,4,8,12
  const int ty = tid >> 4, tx = tid & 15;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float4 ra[2], rb[2];
  auto gload = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int r = m0 + lrow + 64 * i, k = k0 + lk;
      ra[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < M && k < K) ra[i] = *reinterpret_cast<const float4*>(A + (long long)r * lda + k);
      int c = n0 + lrow + 64 * i;
      rb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < N && k < K) rb[i] = __ldg(reinterpret_cast<const float4*>(W + (long long)c * ldw + k));
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int r = lrow + 64 * i;
      As[buf][lk + 0][r] = ra[i].x; As[buf][lk + 1][r] = ra[i].y; As[buf][lk + 2][r] = ra[i].z; As[buf][lk + 3][r] = ra[i].w;
      Bs[buf][lk + 0][r] = rb[i].x; Bs[buf][lk + 1][r] = rb[i].y; Bs[buf][lk + 2][r] = rb[i].z; Bs[buf][lk + 3][r] = rb[i].w;
    }
  };

  const int nk = (K + BK - 1) / BK;
  gload(0);
  sstore(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) gload((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
    long long orow = epi.out_row(m);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (n >= N) continue;
      store_out(C, c_dt, orow * ldc + n, epi.apply(acc[i][j], m, orow, n));
    }
  }
}
}  // namespace

int linear_simt(const PioLinear& p, cudaStream_t st) {
  PIO_CHECK(p.a_dt == PIO_DT_F32, "fp32 GEMM needs fp32 operands");
  PIO_CHECK(p.argmax_val == nullptr && p.exp_ref == nullptr, "the fused arg-max / exp epilogues exist in PIO_BF16 mode only");
  PIO_CHECK(p.K % 4 == 0 && p.lda % 4 == 0 && p.ldw % 4 == 0, "fp32 GEMM needs K, lda, ldw multiples of 4 (K=%d)", p.K);
  PIO_CHECK((((uintptr_t)p.A) & 15) == 0 && (((uintptr_t)p.W) & 15) == 0, "fp32 GEMM operands must be 16-byte aligned");
  if (p.M == 0 || p.N == 0) return PIO_OK;
  dim3 grid(cdiv(p.N, BN), cdiv(p.M, BM));
  sgemm_tn_kernel<<<grid, THREADS, 0, st>>>((const float*)p.A, (const float*)p.W, p.C, p.M, p.N, p.K, p.lda, p.ldw,
                                             p.ldc, p.c_dt, make_epilogue(p));
  PIO_LAUNCHED();
  return PIO_OK;
}

}  // namespace pio

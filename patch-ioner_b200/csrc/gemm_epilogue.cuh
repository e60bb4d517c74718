// gemm_epilogue.cuh -- the epilogue shared by the 1-CTA and the 2-CTA tcgen05 GEMM kernels:
// TMEM accumulator (lane == row) -> registers -> bias / activation / LayerScale / residual -> global,
// plus the fused arg-max and fused exponential variants.
#pragma once
#include "tc_ptx.cuh"

namespace pio {
namespace tc {

template <bool OUT_BF16>
__device__ __forceinline__ void store_chunk(void* C, long long base, const float (&v)[32], int nvalid, bool vec_ok) {
  if constexpr (!OUT_BF16) {
    float* p = reinterpret_cast<float*>(C) + base;
    if (nvalid == 32 && vec_ok) {
#pragma unroll
      for (int i = 0; i < 8; ++i) reinterpret_cast<float4*>(p)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (i < nvalid) p[i] = v[i];
    }
  } else {
    __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(C) + base;
    if (nvalid == 32 && vec_ok) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        __nv_bfloat162 a = __floats2bfloat162_rn(v[8 * i], v[8 * i + 1]), b = __floats2bfloat162_rn(v[8 * i + 2], v[8 * i + 3]);
        __nv_bfloat162 c = __floats2bfloat162_rn(v[8 * i + 4], v[8 * i + 5]), d = __floats2bfloat162_rn(v[8 * i + 6], v[8 * i + 7]);
        uint4 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&a); pk.y = *reinterpret_cast<uint32_t*>(&b);
        pk.z = *reinterpret_cast<uint32_t*>(&c); pk.w = *reinterpret_cast<uint32_t*>(&d);
        reinterpret_cast<uint4*>(p)[i] = pk;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (i < nvalid) p[i] = __float2bfloat16(v[i]);
    }
  }
}

// One epilogue warp, one tile: columns [c_begin, c_end) of TMEM lane quarter `quarter`.
//   out = residual * rowscale + gamma * act(acc * (alpha * colscale) + bias); vectors come from shared memory.
template <int ACT, bool HAS_RES, bool OUT_BF16>
__device__ __forceinline__ void epilogue_cols(uint32_t tmem_acc, int quarter, int lane, int c_begin, int c_end, int m, int M,
                                              int n0, int N, void* C, int ldc, const Epilogue& epi, const float* s_scale,
                                              const float* s_bias, const float* s_gamma) {
  const long long orow = epi.out_row(m < M ? m : 0);
  const bool row_ok = m < M;
  const bool vec_ok = (ldc % 8 == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
  // column vectors that are absent are not read at all (uniform branches)
  const bool use_scale = (epi.colscale != nullptr) || (epi.alpha != 1.0f);
  const bool use_bias = epi.bias != nullptr, use_gamma = epi.gamma != nullptr;
  float rs = 1.0f;
  if constexpr (HAS_RES) {
    if (epi.res_rowscale && row_ok) rs = __ldg(epi.res_rowscale + m);
  }
  const bool res_vec = HAS_RES && (epi.ldres % 4 == 0) && ((reinterpret_cast<uintptr_t>(epi.residual) & 15) == 0);
  const uint32_t taddr = tmem_acc + ((uint32_t)(quarter * 32) << 16);
  auto process = [&](const uint32_t (&r)[32], int c) {
    const int n = n0 + c;
    const int nvalid = min(32, N - n);
    if (!(row_ok && nvalid > 0)) return;
    float res[32];
    if constexpr (HAS_RES) {
      const float* rp = epi.residual + orow * epi.ldres + n;
      if (nvalid == 32 && res_vec) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 t = *reinterpret_cast<const float4*>(rp + 4 * i);
          res[4 * i] = t.x; res[4 * i + 1] = t.y; res[4 * i + 2] = t.z; res[4 * i + 3] = t.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) res[i] = (i < nvalid) ? rp[i] : 0.f;
      }
    }
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), bi = make_float4(0.f, 0.f, 0.f, 0.f), ga = sc;
      if (use_scale) sc = *reinterpret_cast<const float4*>(s_scale + c + i);
      if (use_bias) bi = *reinterpret_cast<const float4*>(s_bias + c + i);
      if (use_gamma) ga = *reinterpret_cast<const float4*>(s_gamma + c + i);
      const float scv[4] = {sc.x, sc.y, sc.z, sc.w}, biv[4] = {bi.x, bi.y, bi.z, bi.w}, gav[4] = {ga.x, ga.y, ga.z, ga.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float t = fmaf(__uint_as_float(r[i + j]), scv[j], biv[j]);
        if constexpr (ACT == PIO_ACT_GELU_ERF) t = gelu_erf(t);
        if constexpr (ACT == PIO_ACT_GELU_NEW) t = gelu_new(t);
        t *= gav[j];
        if constexpr (HAS_RES) t = fmaf(res[i + j], rs, t);
        v[i + j] = t;
      }
    }
    store_chunk<OUT_BF16>(C, orow * ldc + n, v, nvalid, vec_ok);
  };
  // software pipeline over two register buffers: the tcgen05.ld of the next chunk is in flight while this one
  // is processed (static buffer names -- dynamic indexing would push the arrays to local memory)
  uint32_t ra[32], rb[32];
  tmem_ld32(taddr + c_begin, ra);
#pragma unroll
  for (int c = c_begin; c < c_end; c += 64) {
    tmem_ld_wait();
    if (c + 32 < c_end) tmem_ld32(taddr + c + 32, rb);
    process(ra, c);
    if (c + 32 < c_end) {
      tmem_ld_wait();
      if (c + 64 < c_end) tmem_ld32(taddr + c + 64, ra);
      process(rb, c + 32);
    }
  }
}

__device__ __forceinline__ float ex2_approx_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Fused exponential epilogue (streaming softmax numerator): this warp's columns [c_begin, c_end) of its row.
__device__ __forceinline__ void epilogue_exp(uint32_t tmem_acc, int quarter, int c_begin, int c_end, int m, int M, int n0, int N,
                                             int slab, void* C, int ldc, const Epilogue& epi, const float* s_scale) {
  const bool row_ok = m < M;
  const float ref = row_ok ? __ldg(epi.exp_ref + m) : 0.f;
  const bool vec_ok = (ldc % 8 == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
  float sum0 = 0.f, sum1 = 0.f, mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll 1
  for (int c = c_begin; c < c_end; c += 32) {
    uint32_t r[32];
    tmem_ld32(tmem_acc + ((uint32_t)(quarter * 32) << 16) + c, r);
    tmem_ld_wait();
    const int n = n0 + c;
    const int nvalid = min(32, N - n);
    if (row_ok && nvalid > 0) {
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        const float a = __uint_as_float(r[i]) * s_scale[c + i], b = __uint_as_float(r[i + 1]) * s_scale[c + i + 1];
        const bool va = i < nvalid, vb = i + 1 < nvalid;
        mx0 = fmaxf(mx0, va ? a : -INFINITY);
        mx1 = fmaxf(mx1, vb ? b : -INFINITY);
        v[i] = va ? ex2_approx_ftz(a - ref) : 0.f;
        v[i + 1] = vb ? ex2_approx_ftz(b - ref) : 0.f;
        sum0 += v[i];
        sum1 += v[i + 1];
      }
      store_chunk<true>(C, (long long)m * ldc + n, v, nvalid, vec_ok);
    }
  }
  if (row_ok) {
    const long long o = (long long)m * epi.exp_ld + slab;
    epi.exp_psum[o] = sum0 + sum1;
    epi.exp_pmax[o] = fmaxf(mx0, mx1);
  }
}

// Fused arg-max epilogue: this warp's columns [c_begin, c_end) of its row -> (max, first arg-max, sum exp(v - max)).
__device__ __forceinline__ void epilogue_argmax(uint32_t tmem_acc, int quarter, int c_begin, int c_end, int m, int M, int n0, int N,
                                                int slab, const Epilogue& epi, const float* s_scale, const float* s_bias) {
  float best = -INFINITY, sum = 0.f;
  int bidx = 0x7fffffff;
#pragma unroll 1
  for (int c = c_begin; c < c_end; c += 32) {
    uint32_t r[32];
    tmem_ld32(tmem_acc + ((uint32_t)(quarter * 32) << 16) + c, r);
    tmem_ld_wait();
    const int n = n0 + c;
    float v[32];
    float cmax = -INFINITY;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      v[i] = (n + i < N) ? fmaf(__uint_as_float(r[i]), s_scale[c + i], s_bias[c + i]) : -INFINITY;
      cmax = fmaxf(cmax, v[i]);
    }
    if (cmax == -INFINITY) continue;  // chunk entirely beyond N
    if (cmax > best) {  // strictly greater: earlier chunks keep ties (first index wins)
      int ci = 0;
#pragma unroll
      for (int i = 31; i >= 0; --i)
        if (v[i] == cmax) ci = i;
      sum *= __expf(best - cmax);
      best = cmax;
      bidx = n + ci;
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) sum += __expf(v[i] - best);
  }
  if (m < M) {
    const long long o = (long long)m * epi.argmax_ld + slab;
    epi.argmax_val[o] = best;
    epi.argmax_idx[o] = bidx;
    epi.argmax_sumexp[o] = sum;
  }
}


// One epilogue warp, one tile: dispatch on the (uniform) epilogue kind.  `half` selects this warp's column half.
__device__ __forceinline__ void epilogue_tile(uint32_t tacc, int quarter, int lane, int half, int BN, int m, int M, int n0, int N,
                                              int slab, void* C, int ldc, int c_dt, const Epilogue& epi, const float* s_scale,
                                              const float* s_bias, const float* s_gamma) {
  const int cb = half * (BN / 2), ce = cb + BN / 2;
  const bool bf = c_dt == PIO_DT_BF16;
  const bool hr = epi.residual != nullptr;
  if (epi.argmax_val != nullptr) {
    epilogue_argmax(tacc, quarter, cb, ce, m, M, n0, N, slab, epi, s_scale, s_bias);
  } else if (epi.exp_ref != nullptr) {
    epilogue_exp(tacc, quarter, cb, ce, m, M, n0, N, slab, C, ldc, epi, s_scale);
  } else
#define PIO_EPI(ACTV, HR, BF) epilogue_cols<ACTV, HR, BF>(tacc, quarter, lane, cb, ce, m, M, n0, N, C, ldc, epi, s_scale, s_bias, s_gamma)
#define PIO_EPI_ACT(ACTV)                  \
  do {                         \
    if (hr) { if (bf) PIO_EPI(ACTV, true, true); else PIO_EPI(ACTV, true, false); }   \
    else    { if (bf) PIO_EPI(ACTV, false, true); else PIO_EPI(ACTV, false, false); } \
  } while (0)
  if (epi.act == PIO_ACT_GELU_ERF) PIO_EPI_ACT(PIO_ACT_GELU_ERF);
  else if (epi.act == PIO_ACT_GELU_NEW) PIO_EPI_ACT(PIO_ACT_GELU_NEW);
  else PIO_EPI_ACT(PIO_ACT_NONE);
#undef PIO_EPI_ACT
#undef PIO_EPI
}

}  // namespace tc
}  // namespace pio

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


// ---------------------------------------------------------------------------------------------------------
// TMA-store output path.  An epilogue warp owns one 32-row x 128-byte staging box in shared memory (4 KB, 1024-byte
// aligned, 128-byte swizzle: 16-byte chunk j of row r sits at r * 128 + ((j ^ (r & 7)) << 4), which makes the
// lane-per-row st.shared.v4 conflict free).  One lane then issues cp.async.bulk.tensor (or cp.reduce ... .add for the
// in-place residual  x += t) for the whole box: fully coalesced global writes, rows >= M / columns >= N clipped by the
// tensor map, and no global load/store instructions in the epilogue warps at all.
struct TmaOut {
  const CUtensorMap* map;  // tensor map of C: box = 32 rows x 128 bytes, SWIZZLE_128B
  uint32_t smem;           // this warp's staging box
  int mode;                // 0 = direct stores, 1 = TMA store, 2 = TMA reduce-add (C += value)
};
enum { STORE_DIRECT = 0, STORE_TMA = 1, STORE_TMA_ADD = 2, STORE_SPLIT_FIXUP = 3 };

__device__ __forceinline__ void stage_wait_free(int lane) {
  if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  __syncwarp();
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
// 32 consecutive output columns of this lane's row -> staging box.  bf16: 64 bytes = chunks [4 * part, 4 * part + 4);
// fp32: 128 bytes = the whole row.
template <bool OUT_BF16>
__device__ __forceinline__ void stage_write(uint32_t smem, int lane, int part, const float (&v)[32]) {
  const uint32_t row = smem + lane * 128;
  const uint32_t x = lane & 7;
  if constexpr (OUT_BF16) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      st_shared_v4(row + (((uint32_t)(part * 4 + i) ^ x) << 4), pack_bf16x2(v[8 * i], v[8 * i + 1]), pack_bf16x2(v[8 * i + 2], v[8 * i + 3]),
                   pack_bf16x2(v[8 * i + 4], v[8 * i + 5]), pack_bf16x2(v[8 * i + 6], v[8 * i + 7]));
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      st_shared_v4(row + (((uint32_t)i ^ x) << 4), __float_as_uint(v[4 * i]), __float_as_uint(v[4 * i + 1]), __float_as_uint(v[4 * i + 2]),
                   __float_as_uint(v[4 * i + 3]));
  }
}
// make the generic-proxy writes visible to the async proxy, then one lane issues the bulk tensor store of the box
// whose first element is C[m_base, n]
__device__ __forceinline__ void stage_commit(const TmaOut& to, int lane, int n, int m_base, bool add) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  if (lane == 0) {
    if (add)
      asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(to.map),
                   "r"(to.smem), "r"(n), "r"(m_base)
                   : "memory");
    else
      asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(to.map), "r"(to.smem),
                   "r"(n), "r"(m_base)
                   : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
}
// before the kernel exits: every bulk store this thread issued has completed
__device__ __forceinline__ void stage_drain(int lane) {
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  __syncwarp();
}

// Activations of the bf16 tensor-core path.  The exact erff / tanhf of the fp32 path cost 25-30 instructions per element,
// more than the MMA time of the tile they follow; these use the SFU (rcp / ex2 / tanh.approx) instead.
//   erf: Abramowitz-Stegun 7.1.28 (one rcp), |gelu error| <= 9e-7 (far below the bf16 rounding of the result)
//   tanh.approx.f32: relative error 2^-11, below the bf16 rounding (2^-9) of the activation it feeds
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// Exact-erf GELU with ONE special-function op per element (round 2; the Abramowitz-Stegun 7.1.26 form of round 1 needed two --
// rcp and ex2 -- and made the fc1 epilogue SFU-bound: ncu xu_realtime 85.7 %, profiles/r01e_ncu_gemm_tc2_fc1.txt).
//   A&S 7.1.28:  erf(z) = 1 - 1 / (1 + a1 z + ... + a6 z^6)^16,  |error| <= 3e-7 for z >= 0;
//   gelu(x) = max(x, 0) - 0.5 |x| r  with  r = 1 - erf(|x| / sqrt 2) = (...)^-16   (no cancellation in the negative tail).
// Measured against float64 erf over [-12, 12] in fp32 arithmetic: |gelu error| <= 9e-7 (far below the bf16 rounding of the result);
// for |x| >~ 13 the sixteenth power overflows to +inf, rcp gives 0 and gelu saturates to max(x, 0), as it should.
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float ax = fabsf(x), z = ax * 0.70710678118654752440f;
  float t = fmaf(0.0000430638f, z, 0.0002765672f);
  t = fmaf(t, z, 0.0001520143f);
  t = fmaf(t, z, 0.0092705272f);
  t = fmaf(t, z, 0.0422820123f);
  t = fmaf(t, z, 0.0705230784f);
  t = fmaf(t, z, 1.0f);
  t *= t; t *= t; t *= t; t *= t;
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t));
  return fmaf(-0.5f * ax, r, fmaxf(x, 0.0f));
}
// The same formula on TWO elements per instruction (fma.rn.f32x2 / mul.rn.f32x2 -> FFMA2 / FMUL2): the GELU epilogue issues
// 12 packed + 6 scalar instructions per pair instead of 2 x 17 -- the epilogue warps (two per scheduler) are the critical path of
// fc1, not the tensor pipe (B200: fc1 + GELU 1035 -> 1163 TFLOP/s, profiles/r02ap_gemm_probe_gelu_f32x2.txt).
// pack2 / unpack2 / fma2 / mul2 / add2: tc_ptx.cuh
__device__ __forceinline__ uint64_t gelu_erf_fast2(uint64_t x) {
  // coefficients of 7.1.28 times 2^(1/16): the sixteenth power comes out doubled, its reciprocal is r / 2; with nax = -|x|
  // gelu = max(x, 0) + nax (r / 2): one packed multiply less.  |gelu error| <= 7.1e-7 in fp32 arithmetic over [-14, 14].
  float x0, x1;
  unpack2(x, x0, x1);
  const uint64_t nax = pack2(-fabsf(x0), -fabsf(x1));
  const uint64_t z = mul2(nax, pack2(-0.70710678118654752440f, -0.70710678118654752440f));
  uint64_t t = fma2(pack2(4.497039844864048e-05f, 4.497039844864048e-05f), z, pack2(0.0002888118615373969f, 0.0002888118615373969f));
  t = fma2(t, z, pack2(0.0001587445440236479f, 0.0001587445440236479f));
  t = fma2(t, z, pack2(0.009680968709290028f, 0.009680968709290028f));
  t = fma2(t, z, pack2(0.04415399581193924f, 0.04415399581193924f));
  t = fma2(t, z, pack2(0.07364540547132492f, 0.07364540547132492f));
  t = fma2(t, z, pack2(1.0442737340927124f, 1.0442737340927124f));
  t = mul2(t, t); t = mul2(t, t); t = mul2(t, t); t = mul2(t, t);
  float t0, t1, r0, r1;
  unpack2(t, t0, t1);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(t0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(t1));
  return fma2(nax, pack2(r0, r1), pack2(fmaxf(x0, 0.0f), fmaxf(x1, 0.0f)));
}
__device__ __forceinline__ float gelu_new_fast(float x) {
  const float u = 0.79788456080286535588f * fmaf(0.044715f * x * x, x, x);
  float th;
  asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(u));
  const float hx = 0.5f * x;
  return fmaf(hx, th, hx);
}

// One epilogue warp, one tile: columns [c_begin, c_end) of TMEM lane quarter `quarter`.
//   out = residual * rowscale + gamma * act(acc * (alpha * colscale) + bias); vectors come from shared memory.
// STORE_TMA_ADD is the in-place residual (residual == C, no row scale): the value is reduced into C by the copy engine.
// PLAIN: no column scale, alpha == 1 and no LayerScale vector (qkv, fc1, fc of the decoder, ...): the per-column defaults are
// not materialised and not multiplied in -- acc + bias only (bit-identical: x * 1 + b == x + b), ~20 % fewer epilogue instructions
template <int ACT, bool HAS_RES, bool OUT_BF16, int STORE, int NCOLS, bool PLAIN = false>
__device__ __forceinline__ void epilogue_cols(uint32_t tmem_acc, int quarter, int lane, int c_begin, int m, int M,
                                              int n0, int N, void* C, int ldc, const Epilogue& epi, const float* s_scale,
                                              const float* s_bias, const float* s_gamma, const TmaOut& to) {
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
  const int c_end = c_begin + NCOLS;  // NCOLS is a compile-time constant: the chunk loop below unrolls completely
  auto process = [&](const uint32_t (&r)[32], int c) {
    const int n = n0 + c;
    const int nvalid = min(32, N - n);
    if constexpr (STORE != STORE_DIRECT) {
      if (nvalid <= 0) return;  // uniform: the whole chunk lies beyond N (rows >= M are clipped by the tensor map)
    } else {
      if (!(row_ok && nvalid > 0)) return;
    }
    float res[32];
    if constexpr (HAS_RES && STORE == STORE_DIRECT) {
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
    if constexpr (PLAIN) {
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const float4 bi = *reinterpret_cast<const float4*>(s_bias + c + i);  // staged as zeros when there is no bias
        const float biv[4] = {bi.x, bi.y, bi.z, bi.w};
        if constexpr (ACT != PIO_ACT_GELU_NEW && !(HAS_RES && STORE == STORE_DIRECT)) {
#pragma unroll
          for (int j = 0; j < 4; j += 2) {
            uint64_t t = add2(pack2(__uint_as_float(r[i + j]), __uint_as_float(r[i + j + 1])), pack2(biv[j], biv[j + 1]));
            if constexpr (ACT == PIO_ACT_GELU_ERF) t = gelu_erf_fast2(t);
            unpack2(t, v[i + j], v[i + j + 1]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float t = __uint_as_float(r[i + j]) + biv[j];
            if constexpr (ACT == PIO_ACT_GELU_ERF) t = gelu_erf_fast(t);
            if constexpr (ACT == PIO_ACT_GELU_NEW) t = gelu_new_fast(t);
            if constexpr (HAS_RES && STORE == STORE_DIRECT) t = fmaf(res[i + j], rs, t);
            v[i + j] = t;
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < (PLAIN ? 0 : 32); i += 4) {
      float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), bi = make_float4(0.f, 0.f, 0.f, 0.f), ga = sc;
      if (use_scale) sc = *reinterpret_cast<const float4*>(s_scale + c + i);
      if (use_bias) bi = *reinterpret_cast<const float4*>(s_bias + c + i);
      if (use_gamma) ga = *reinterpret_cast<const float4*>(s_gamma + c + i);
      const float scv[4] = {sc.x, sc.y, sc.z, sc.w}, biv[4] = {bi.x, bi.y, bi.z, bi.w}, gav[4] = {ga.x, ga.y, ga.z, ga.w};
      if constexpr (ACT != PIO_ACT_GELU_NEW && !(HAS_RES && STORE == STORE_DIRECT)) {  // two elements per instruction
#pragma unroll
        for (int j = 0; j < 4; j += 2) {
          uint64_t t = fma2(pack2(__uint_as_float(r[i + j]), __uint_as_float(r[i + j + 1])), pack2(scv[j], scv[j + 1]), pack2(biv[j], biv[j + 1]));
          if constexpr (ACT == PIO_ACT_GELU_ERF) t = gelu_erf_fast2(t);
          t = mul2(t, pack2(gav[j], gav[j + 1]));
          unpack2(t, v[i + j], v[i + j + 1]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float t = fmaf(__uint_as_float(r[i + j]), scv[j], biv[j]);
          if constexpr (ACT == PIO_ACT_GELU_ERF) t = gelu_erf_fast(t);
          if constexpr (ACT == PIO_ACT_GELU_NEW) t = gelu_new_fast(t);
          t *= gav[j];
          if constexpr (HAS_RES && STORE == STORE_DIRECT) t = fmaf(res[i + j], rs, t);
          v[i + j] = t;
        }
      }
    }
    if constexpr (STORE == STORE_DIRECT) {
      store_chunk<OUT_BF16>(C, orow * ldc + n, v, nvalid, vec_ok);
    } else {
      const int part = OUT_BF16 ? (((c - c_begin) >> 5) & 1) : 0;
      if (part == 0) stage_wait_free(lane);  // the previous box has been read out of shared memory
      stage_write<OUT_BF16>(to.smem, lane, part, v);
      if (!OUT_BF16 || part == 1 || c + 32 >= c_end || n + 32 >= N)
        stage_commit(to, lane, n - 32 * part, m - lane, STORE == STORE_TMA_ADD);
    }
  };
  // software pipeline over two register buffers: the tcgen05.ld of the next chunk is in flight while this one
  // is processed (static buffer names -- dynamic indexing would push the arrays to local memory)
  uint32_t ra[32], rb[32];
  tmem_ld32(taddr + c_begin, ra);
#pragma unroll
  for (int j = 0; j < NCOLS; j += 64) {
    const int c = c_begin + j;
    tmem_ld_wait();
    if (j + 32 < NCOLS) tmem_ld32(taddr + c + 32, rb);
    process(ra, c);
    if (j + 32 < NCOLS) {
      tmem_ld_wait();
      if (j + 64 < NCOLS) tmem_ld32(taddr + c + 64, ra);
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
// Round 2, second session: full 32-column chunks take a fast path -- column scales as four-wide shared-memory loads, scale /
// subtract / row sum on two elements per instruction (mul / add.rn.f32x2: the same roundings as the scalar code), no validity
// selects -- and the tensor-memory load of the next chunk is in flight while this one is processed (8.5 -> ~4 instructions per
// element; the similarity GEMM of the memory projection ran 15 % below the plain GEMM of its shape).
__device__ __forceinline__ void epilogue_exp(uint32_t tmem_acc, int quarter, int lane, int c_begin, int c_end, int m, int M, int n0,
                                             int N, int slab, void* C, int ldc, const Epilogue& epi, const float* s_scale,
                                             const TmaOut& to) {
  const bool row_ok = m < M;
  const float ref = row_ok ? __ldg(epi.exp_ref + m) : 0.f;
  const bool vec_ok = (ldc % 8 == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
  float sum0 = 0.f, sum1 = 0.f, mx0 = -INFINITY, mx1 = -INFINITY;
  uint64_t sum2 = pack2(0.f, 0.f);
  const uint64_t nref2 = pack2(-ref, -ref);
  const uint32_t taddr = tmem_acc + ((uint32_t)(quarter * 32) << 16);
  auto process = [&](const uint32_t (&r)[32], int c) {
    const int n = n0 + c;
    const int nvalid = min(32, N - n);
    if (nvalid <= 0) return;  // uniform: the chunk lies beyond N
    if (!(row_ok || to.mode != STORE_DIRECT)) return;
    float v[32];
    if (nvalid == 32) {
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const float4 sc = *reinterpret_cast<const float4*>(s_scale + c + i);
        const uint64_t a01 = mul2(pack2(__uint_as_float(r[i]), __uint_as_float(r[i + 1])), pack2(sc.x, sc.y));
        const uint64_t a23 = mul2(pack2(__uint_as_float(r[i + 2]), __uint_as_float(r[i + 3])), pack2(sc.z, sc.w));
        float a0, a1, a2, a3, x0, x1, x2, x3;
        unpack2(a01, a0, a1); unpack2(a23, a2, a3);
        mx0 = fmaxf(fmaxf(mx0, a0), a2);
        mx1 = fmaxf(fmaxf(mx1, a1), a3);
        unpack2(add2(a01, nref2), x0, x1); unpack2(add2(a23, nref2), x2, x3);
        v[i] = ex2_approx_ftz(x0); v[i + 1] = ex2_approx_ftz(x1); v[i + 2] = ex2_approx_ftz(x2); v[i + 3] = ex2_approx_ftz(x3);
        sum2 = add2(add2(sum2, pack2(v[i], v[i + 1])), pack2(v[i + 2], v[i + 3]));
      }
    } else {
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
    }
    if (to.mode == STORE_DIRECT) {
      store_chunk<true>(C, (long long)m * ldc + n, v, nvalid, vec_ok);
    } else {
      const int part = ((c - c_begin) >> 5) & 1;
      if (part == 0) stage_wait_free(lane);
      stage_write<true>(to.smem, lane, part, v);
      if (part == 1 || c + 32 >= c_end || n + 32 >= N) stage_commit(to, lane, n - 32 * part, m - lane, false);
    }
  };
  uint32_t ra[32], rb[32];
  tmem_ld32(taddr + c_begin, ra);
#pragma unroll 1
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
  if (row_ok) {
    float s2a, s2b;
    unpack2(sum2, s2a, s2b);
    const long long o = (long long)m * epi.exp_ld + slab;
    epi.exp_psum[o] = (sum0 + sum1) + (s2a + s2b);
    epi.exp_pmax[o] = fmaxf(mx0, mx1);
  }
}

// Fused arg-max epilogue: this warp's columns [c_begin, c_end) of its row -> (max, first arg-max, sum exp(v - max)).
template <bool WITH_SUMEXP>  // the sum of exponentials (log-prob of the arg-max) is only needed for compute_scores
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
      if (WITH_SUMEXP) sum *= __expf(best - cmax);
      best = cmax;
      bidx = n + ci;
    }
    if (WITH_SUMEXP) {
#pragma unroll
      for (int i = 0; i < 32; ++i) sum += __expf(v[i] - best);
    }
  }
  if (m < M) {
    const long long o = (long long)m * epi.argmax_ld + slab;
    epi.argmax_val[o] = best;
    epi.argmax_idx[o] = bidx;
    if (WITH_SUMEXP) epi.argmax_sumexp[o] = sum;
  }
}


// One epilogue warp, one tile: dispatch on the (uniform) epilogue kind.  `half` selects this warp's column half.
// ALLOW_PLAIN: also instantiate the bias-only specialisation of epilogue_cols (the pair kernel, where the big GEMMs run; the
// 1-CTA kernel's four tile widths would double their compile time for launches that are latency-bound anyway)
template <int BN, bool ALLOW_PLAIN = false>
__device__ __forceinline__ void epilogue_tile(uint32_t tacc, int quarter, int lane, int half, int m, int M, int n0, int N,
                                              int slab, void* C, int ldc, int c_dt, const Epilogue& epi, const float* s_scale,
                                              const float* s_bias, const float* s_gamma, const TmaOut& to) {
  const bool bf = c_dt == PIO_DT_BF16;
  // this warp's columns: half of the tile (the 192-wide tile is cut 128 + 64 for bf16 output: 64-column staging boxes)
  const int cb = half * (BN / 2), ce = cb + BN / 2;
  const bool hr = epi.residual != nullptr;
  const bool plain = ALLOW_PLAIN && epi.colscale == nullptr && epi.alpha == 1.0f && epi.gamma == nullptr;
  if (epi.argmax_val != nullptr) {
    if (epi.argmax_sumexp != nullptr) epilogue_argmax<true>(tacc, quarter, cb, ce, m, M, n0, N, slab, epi, s_scale, s_bias);
    else epilogue_argmax<false>(tacc, quarter, cb, ce, m, M, n0, N, slab, epi, s_scale, s_bias);
  } else if (epi.exp_ref != nullptr) {
    epilogue_exp(tacc, quarter, lane, cb, ce, m, M, n0, N, slab, C, ldc, epi, s_scale, to);
  } else
#define PIO_EPI_N(ACTV, HR, BF, ST, NC, CB)                                                                                      \
  do {                                                                                                                           \
    if constexpr (ALLOW_PLAIN) {                                                                                                 \
      if (plain) { epilogue_cols<ACTV, HR, BF, ST, NC, true>(tacc, quarter, lane, CB, m, M, n0, N, C, ldc, epi, s_scale, s_bias, s_gamma, to); break; } \
    }                                                                                                                            \
    epilogue_cols<ACTV, HR, BF, ST, NC, false>(tacc, quarter, lane, CB, m, M, n0, N, C, ldc, epi, s_scale, s_bias, s_gamma, to); \
  } while (0)
#define PIO_EPI(ACTV, HR, BF, ST)                                                              \
  do {                                                                                         \
    if constexpr (BN == 192 && BF) {                                                           \
      if (half == 0) PIO_EPI_N(ACTV, HR, BF, ST, 128, 0); else PIO_EPI_N(ACTV, HR, BF, ST, 64, 128); \
    } else {                                                                                   \
      PIO_EPI_N(ACTV, HR, BF, ST, BN / 2, cb);                                                 \
    }                                                                                          \
  } while (0)
#define PIO_EPI_ACT(ACTV)                                                                         \
  do {                                                                                            \
    if (to.mode == STORE_TMA_ADD) PIO_EPI(ACTV, true, false, STORE_TMA_ADD); /* fp32, C += value */ \
    else if (to.mode == STORE_TMA) { if (bf) PIO_EPI(ACTV, false, true, STORE_TMA); else PIO_EPI(ACTV, false, false, STORE_TMA); } \
    else if (hr) { if (bf) PIO_EPI(ACTV, true, true, STORE_DIRECT); else PIO_EPI(ACTV, true, false, STORE_DIRECT); }   \
    else         { if (bf) PIO_EPI(ACTV, false, true, STORE_DIRECT); else PIO_EPI(ACTV, false, false, STORE_DIRECT); } \
  } while (0)
  if (epi.act == PIO_ACT_GELU_ERF) PIO_EPI_ACT(PIO_ACT_GELU_ERF);
  else if (epi.act == PIO_ACT_GELU_NEW) PIO_EPI_ACT(PIO_ACT_GELU_NEW);
  else PIO_EPI_ACT(PIO_ACT_NONE);
#undef PIO_EPI_ACT
#undef PIO_EPI
#undef PIO_EPI_N
}

// Host side: which output path a launch may use (the staging boxes need `tma_capable` kernels: see Cfg / gemm2).
inline int pick_store_mode(const PioLinear& p) {
  const size_t es = p.c_dt == PIO_DT_BF16 ? 2 : 4;
  if (p.argmax_val != nullptr || p.C == nullptr || p.rows_per_group != 0) return STORE_DIRECT;
  if ((reinterpret_cast<uintptr_t>(p.C) & 15) != 0 || ((size_t)p.ldc * es) % 16 != 0) return STORE_DIRECT;
  // measured on B200: the bulk tensor store clips the inner dimension in 16-byte units (a chunk that straddles N is
  // written whole), so N must end on a 16-byte boundary
  if (((size_t)p.N * es) % 16 != 0) return STORE_DIRECT;
  if (p.residual == nullptr) return STORE_TMA;
  if (p.residual == p.C && p.ldres == p.ldc && p.res_rowscale == nullptr && p.c_dt == PIO_DT_F32 && p.exp_ref == nullptr)
    return STORE_TMA_ADD;
  return STORE_DIRECT;
}
// tensor map of C for the staging boxes: 32 rows x 128 bytes, 128-byte swizzle
int make_map_out(CUtensorMap* map, const void* ptr, long long rows, long long cols, long long ld, int c_dt);
bool tma_store_enabled();  // PIO_GEMM_TMA_STORE=0 forces the direct-store epilogue (A/B testing)

}  // namespace tc
}  // namespace pio

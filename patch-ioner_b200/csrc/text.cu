// text.cu -- the DeCap text side: caption-memory projection and the KV-cached greedy prefix decoder.
//   pio_project*        replaces Im2TxtProjector.project   (im2txtprojection.py:353-385)
//   pio_decode_greedy   replaces decoding_batched          (src/decap/decap.py:116-160)
#include "decoder.cuh"
#include <stdlib.h>
#include <vector>

namespace pio {
namespace {

// ------------------------------------------------------------------------------------------
// Online-softmax step over one bank chunk (one CTA per query row).
//   S[r, 0:mc] holds  <q^, k^_j> / T  for the chunk;  running (m, l) are updated,
//   P = exp(S - m_new) is written (fp32 in place or bf16 to P16), columns [mc, mc_pad) are zeroed,
//   alpha[r] = exp(m_old - m_new) rescales the running output in the following GEMM.
__global__ void __launch_bounds__(256) softmax_chunk_kernel(float* __restrict__ S, int lds, __nv_bfloat16* __restrict__ P16,
                                                            int ldp, int mc, int mc_pad, float* __restrict__ m,
                                                            float* __restrict__ l, float* __restrict__ alpha) {
  __shared__ float red[8];
  const int r = blockIdx.x;
  float* row = S + (long long)r * lds;
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < mc; i += 256) mx = fmaxf(mx, row[i]);
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) mx = fmaxf(mx, red[i]);
  __syncthreads();
  const float mold = m[r];
  const float mnew = fmaxf(mold, mx);
  float s = 0.f;
  for (int i = threadIdx.x; i < mc_pad; i += 256) {
    float p = (i < mc) ? expf(row[i] - mnew) : 0.f;
    s += p;
    if (P16) P16[(long long)r * ldp + i] = __float2bfloat16(p);
    else row[i] = p;
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i];
    const float a = expf(mold - mnew);  // mold = -inf on the first chunk -> 0
    alpha[r] = a;
    l[r] = l[r] * a + t;
    m[r] = mnew;
  }
}

// Streaming-softmax bookkeeping after one bank chunk of the fused path (everything in log2 units):
//   l = alpha * l + sum(psum);  m' = max(m, max(pmax));  alpha' = 2^(m - m');  m = m'
__global__ void project_update_kernel(const float* __restrict__ psum, const float* __restrict__ pmax, int ld, int slabs, int R,
                                      float* __restrict__ mref, float* __restrict__ m_used, float* __restrict__ l,
                                      float* __restrict__ alpha) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  pdl_wait();
  pdl_launch_dependents();
  if (warp >= R) return;
  float s = 0.f, mx = -INFINITY;
  for (int k = lane; k < slabs; k += 32) {
    s += psum[(long long)warp * ld + k];
    mx = fmaxf(mx, pmax[(long long)warp * ld + k]);
  }
  s = warp_sum(s);
  mx = warp_max(mx);
  if (lane == 0) {
    const float m_old = mref[warp], m_new = fmaxf(m_old, mx);
    m_used[warp] = m_old;  // the reference this chunk's P, and hence O and l, are expressed in
    l[warp] = alpha[warp] * l[warp] + s;
    alpha[warp] = exp2f(m_old - m_new);
    mref[warp] = m_new;
  }
}
// after the last chunk: natural-log running max for the sharded interface
__global__ void scale_vec_kernel(const float* in, float* out, float f, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i] * f;
}

__global__ void fill_kernel(float* p, float v, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) p[i] = v;
}

__global__ void row_inv_norm_kernel(const float* __restrict__ x, float* __restrict__ inv, long long rows, int dim) {
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const float* r = x + warp * dim;
  float s = 0.f;
  for (int i = lane; i < dim; i += 32) s += r[i] * r[i];
  s = warp_sum(s);
  if (lane == 0) inv[warp] = 1.0f / sqrtf(s);
}

// O[r,:] *= alpha[r]   (the running-output rescale, applied ahead of an accumulate-only, split-K recombination GEMM)
__global__ void scale_rows_kernel(float* __restrict__ O, const float* __restrict__ alpha, int R, int D) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= R) return;
  const float a = alpha[warp];
  for (int i = lane; i < D; i += 32) O[(long long)warp * D + i] *= a;
}
// O[r,:] *= f, l[r] *= f with f = exp(m_local - m_global)
__global__ void rescale_rows_kernel(float* __restrict__ O, float* __restrict__ l, const float* __restrict__ ml,
                                    const float* __restrict__ mg, int R, int D) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= R) return;
  const float f = expf(ml[warp] - mg[warp]);
  for (int i = lane; i < D; i += 32) O[(long long)warp * D + i] *= f;
  if (lane == 0) l[warp] *= f;
}
// O[r,:] /= l[r]; optional L2 normalisation
__global__ void finish_rows_kernel(float* __restrict__ O, const float* __restrict__ l, int R, int D, int normalize) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= R) return;
  float* r = O + (long long)warp * D;
  const float inv = 1.0f / l[warp];
  float s = 0.f;
  for (int i = lane; i < D; i += 32) { float v = r[i] * inv; r[i] = v; s += v * v; }
  if (normalize) {
    const float n = sqrtf(warp_sum(s));
    for (int i = lane; i < D; i += 32) r[i] = r[i] / n;
  }
}

// ------------------------------------------------------------------------------------------
// Row arg-max of the logits (first index wins ties, like torch.argmax) + log softmax at the arg-max.
__global__ void __launch_bounds__(256) argmax_rows_kernel(const float* __restrict__ logits, int ld, int V, int* __restrict__ ids,
                                                          int ids_ld, int t, float* __restrict__ logprob_sum) {
  __shared__ float rv[8];
  __shared__ int ri[8];
  __shared__ float rs[8];
  const float* row = logits + (long long)blockIdx.x * ld;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < V; i += 256) {
    float v = row[i];
    if (v > best) { best = v; bi = i; }   // strided ascending: first index kept on ties
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, best, o);
    int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  if ((threadIdx.x & 31) == 0) { rv[threadIdx.x >> 5] = best; ri[threadIdx.x >> 5] = bi; }
  __syncthreads();
  best = rv[0]; bi = ri[0];
#pragma unroll
  for (int i = 1; i < 8; ++i)
    if (rv[i] > best || (rv[i] == best && ri[i] < bi)) { best = rv[i]; bi = ri[i]; }
  if (logprob_sum) {
    float s = 0.f;
    for (int i = threadIdx.x; i < V; i += 256) s += expf(row[i] - best);
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) rs[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      float tot = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) tot += rs[i];
      logprob_sum[blockIdx.x] += -logf(tot);  // log softmax(logits)[argmax] = -log sum exp(l - max)
    }
  }
  // a row of NaN logits (e.g. the NaN embedding of an empty box) never compares greater: torch.argmax answers 0 there
  if (threadIdx.x == 0) ids[(long long)blockIdx.x * ids_ld + t] = (bi == 0x7fffffff) ? 0 : bi;
}

// Reduce the per-slab (max, first index, sum exp) triples of the fused lm-head epilogue: one warp per row.
__global__ void __launch_bounds__(256) argmax_finish_kernel(const float* __restrict__ val, const int* __restrict__ idx,
                                                            const float* __restrict__ sumexp, int ld, int slabs, int M,
                                                            int* __restrict__ ids, int ids_ld, int t, float* __restrict__ logprob_sum) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  pdl_wait();
  pdl_launch_dependents();
  if (row >= M) return;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int s = lane; s < slabs; s += 32) {
    const float v = val[(long long)row * ld + s];
    const int i = idx[(long long)row * ld + s];
    if (v > best || (v == best && i < bi)) { best = v; bi = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  if (logprob_sum) {
    float s = 0.f;
    for (int k = lane; k < slabs; k += 32) {
      const float v = val[(long long)row * ld + k];
      if (v != -INFINITY) s += sumexp[(long long)row * ld + k] * expf(v - best);  // slabs beyond N carry no mass
    }
    s = warp_sum(s);
    if (lane == 0) logprob_sum[row] += -logf(s);
  }
  if (lane == 0) ids[(long long)row * ids_ld + t] = (bi == 0x7fffffff) ? 0 : bi;  // all-NaN row -> 0, like torch.argmax
}

// x[r,:] = wte[ids[r,t],:] + wpe[pos,:]
__global__ void embed_kernel(const float* __restrict__ wte, const float* __restrict__ wpe, const int* __restrict__ ids,
                             int ids_ld, int t, int pos, float* __restrict__ x, int R, int D) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  pdl_wait();
  pdl_launch_dependents();
  if (warp >= R) return;
  const int tok = min(max(ids[(long long)warp * ids_ld + t], 0), 50257 - 1);  // never index outside the table
  const float4* a = reinterpret_cast<const float4*>(wte + (long long)tok * D);
  const float4* b = reinterpret_cast<const float4*>(wpe + (long long)pos * D);
  float4* o = reinterpret_cast<float4*>(x + (long long)warp * D);
  for (int i = lane; i < D / 4; i += 32) {
    float4 u = __ldg(a + i), v = __ldg(b + i);
    o[i] = make_float4(u.x + v.x, u.y + v.y, u.z + v.z, u.w + v.w);
  }
}

// x[r,:] = prompt[r,p,:] + wpe[p,:]   (inputs_embeds of a prompt position; GPT-2 adds the position embedding)
__global__ void prompt_embed_kernel(const float* __restrict__ prompt, int P, int p, const float* __restrict__ wpe,
                                    float* __restrict__ x, int R, int D) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= R) return;
  const float4* a = reinterpret_cast<const float4*>(prompt + ((long long)warp * P + p) * D);
  const float4* b = reinterpret_cast<const float4*>(wpe + (long long)p * D);
  float4* o = reinterpret_cast<float4*>(x + (long long)warp * D);
  for (int i = lane; i < D / 4; i += 32) {
    float4 u = __ldg(a + i), v = __ldg(b + i);
    o[i] = make_float4(u.x + v.x, u.y + v.y, u.z + v.z, u.w + v.w);
  }
}

// x[r*P + p, :] = prompt[r, p, :] + wpe[p, :] for every prompt position (batched prefill)
__global__ void prompt_embed_all_kernel(const float* __restrict__ prompt, int P, const float* __restrict__ wpe, float* __restrict__ x,
                                        int rows, int D) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const float4* a = reinterpret_cast<const float4*>(prompt + (long long)warp * D);
  const float4* b = reinterpret_cast<const float4*>(wpe + (long long)(warp % P) * D);
  float4* o = reinterpret_cast<float4*>(x + (long long)warp * D);
  for (int i = lane; i < D / 4; i += 32) {
    float4 u = __ldg(a + i), v = __ldg(b + i);
    o[i] = make_float4(u.x + v.x, u.y + v.y, u.z + v.z, u.w + v.w);
  }
}

// x[r*n + i, :] = wte[ids[r,i], :] + wpe[i, :]   (token embeddings of whole sentences)
__global__ void embed_all_kernel(const float* __restrict__ wte, const float* __restrict__ wpe, const int* __restrict__ ids, int n,
                                 float* __restrict__ x, int rows, int D) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const int tok = min(max(ids[warp], 0), 50257 - 1);
  const float4* a = reinterpret_cast<const float4*>(wte + (long long)tok * D);
  const float4* b = reinterpret_cast<const float4*>(wpe + (long long)(warp % n) * D);
  float4* o = reinterpret_cast<float4*>(x + (long long)warp * D);
  for (int i = lane; i < D / 4; i += 32) {
    float4 u = __ldg(a + i), v = __ldg(b + i);
    o[i] = make_float4(u.x + v.x, u.y + v.y, u.z + v.z, u.w + v.w);
  }
}

// nll[row] = logsumexp(logits[row - row0, :V]) - logits[row - row0, ids[row + 1]] for the positions that have a next token
// (row = r*n + i with i + 1 < lens[r]); 0 elsewhere.  One CTA per logits row.
__global__ void __launch_bounds__(256) token_nll_kernel(const float* __restrict__ logits, int ld, int V, const int* __restrict__ ids,
                                                        const int* __restrict__ lens, int n, int row0, float* __restrict__ nll) {
  __shared__ float red[8];
  const int row = row0 + blockIdx.x, r = row / n, i = row % n, tid = threadIdx.x;
  if (i + 1 >= lens[r]) {
    if (tid == 0) nll[row] = 0.f;
    return;
  }
  const float* lr = logits + (long long)blockIdx.x * ld;
  float mx = -INFINITY;
  for (int v = tid; v < V; v += 256) mx = fmaxf(mx, lr[v]);
  mx = warp_max(mx);
  if ((tid & 31) == 0) red[tid >> 5] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) mx = fmaxf(mx, red[w]);
  __syncthreads();
  float s = 0.f;
  for (int v = tid; v < V; v += 256) s += expf(lr[v] - mx);
  s = warp_sum(s);
  if ((tid & 31) == 0) red[tid >> 5] = s;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w];
    const int target = min(max(ids[row + 1], 0), V - 1);
    nll[row] = mx + logf(t) - lr[target];
  }
}

// out[r] = mean of nll[r, 0 .. lens[r]-2]  (HF causal-LM loss of one sentence; NaN for fewer than two tokens)
__global__ void nll_mean_kernel(const float* __restrict__ nll, const int* __restrict__ lens, int n, int R, float* __restrict__ out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const int m = lens[r] - 1;
  float s = 0.f;
  for (int i = 0; i < m; ++i) s += nll[(long long)r * n + i];
  out[r] = m > 0 ? s / (float)m : NAN;
}

__global__ void add_vec_kernel(const float* a, const float* b, float* o, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) o[i] = a[i] + b[i];
}


// Running top-n of a score row: one CTA per query.  Candidates = the running list (n, from earlier chunks) + this chunk's
// mc scores; n rounds of a block-wide arg-max (ties: smaller bank row first, like a stable descending sort of equal values).
constexpr int TOPN_MAX = 32;
__global__ void __launch_bounds__(256) topn_chunk_kernel(const float* __restrict__ S, int lds, int mc, long long row0, int n,
                                                         float* __restrict__ best_v, int* __restrict__ best_i) {
  __shared__ float sv[TOPN_MAX];
  __shared__ int si[TOPN_MAX];
  __shared__ float rv[8];
  __shared__ int ri[8];
  __shared__ int taken[TOPN_MAX];  // candidate ids already selected this call (>= 0: chunk column, < 0: -(slot + 1) of the old list)
  const int r = blockIdx.x, tid = threadIdx.x;
  const float* row = S + (long long)r * lds;
  if (tid < n) { sv[tid] = best_v[(long long)r * n + tid]; si[tid] = best_i[(long long)r * n + tid]; }
  __syncthreads();
  float outv = 0.f;
  int outi = 0;
  for (int k = 0; k < n; ++k) {
    // candidate id space: [0, mc) chunk columns, [mc, mc + n) the old list
    float bv = -INFINITY;
    int bc = 0x7fffffff;
    for (int c = tid; c < mc + n; c += 256) {
      bool used = false;
      for (int t = 0; t < k; ++t) used |= (taken[t] == c);
      if (used) continue;
      const float v = c < mc ? row[c] : sv[c - mc];
      if (v > bv || (v == bv && c < bc)) { bv = v; bc = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oc = __shfl_xor_sync(0xffffffffu, bc, o);
      if (ov > bv || (ov == bv && oc < bc)) { bv = ov; bc = oc; }
    }
    if ((tid & 31) == 0) { rv[tid >> 5] = bv; ri[tid >> 5] = bc; }
    __syncthreads();
    if (tid == 0) {
      float v = rv[0];
      int c = ri[0];
      for (int w = 1; w < 8; ++w)
        if (rv[w] > v || (rv[w] == v && ri[w] < c)) { v = rv[w]; c = ri[w]; }
      taken[k] = c;
      rv[0] = v;
      ri[0] = c;
    }
    __syncthreads();
    if (tid == k) {  // thread k keeps the k-th winner until the old list has been fully read
      const int c = ri[0];
      outv = rv[0];
      outi = (c == 0x7fffffff) ? -1 : (c < mc ? (int)(row0 + c) : si[c - mc]);
    }
    __syncthreads();
  }
  if (tid < n) { best_v[(long long)r * n + tid] = outv; best_i[(long long)r * n + tid] = outi; }
}

int linear(int mode, const void* A, const void* W, void* C, int M, int N, int K, int lda, int ldw, int ldc, int a_dt, int c_dt,
           const float* bias, const float* residual, int act, cudaStream_t st) {
  PioLinear p;
  memset(&p, 0, sizeof(p));
  p.A = A; p.W = W; p.C = C; p.M = M; p.N = N; p.K = K; p.lda = lda; p.ldw = ldw; p.ldc = ldc; p.a_dt = a_dt; p.c_dt = c_dt;
  p.bias = bias; p.residual = residual; p.ldres = ldc; p.alpha = 1.0f; p.act = act;
  p.w_static = 1;  // every caller of this helper passes model weights
  return mode == PIO_FP32 ? linear_simt(p, st) : linear_tc(p, st);
}

}  // namespace
}  // namespace pio

// ==========================================================================================
struct PioBank {
  int mode, act_dt, D;
  long long M, Mp;       // rows, padded row count (leading dimension of the transpose)
  void* bank;            // act [M, D]
  void* bankT;           // act [D, Mp], zero padded
  float* inv_norm;       // [M]
};

namespace pio {
namespace {
// bankT[c, r] = bank[r, c] with leading dimension ldt; also the straight copy in the activation dtype
template <typename T>
__global__ void bank_transpose_kernel(const float* __restrict__ in, T* __restrict__ outT, T* __restrict__ outC, long long rows,
                                      int cols, long long ldt) {
  __shared__ float tile[32][33];
  const long long r0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const long long r = r0 + i;
    const int c = c0 + threadIdx.x;
    float v = (r < rows && c < cols) ? in[r * cols + c] : 0.f;
    tile[i][threadIdx.x] = v;
    if (r < rows && c < cols) outC[r * cols + c] = (T)v;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + i;
    const long long r = r0 + threadIdx.x;
    if (c < cols && r < rows) outT[(long long)c * ldt + r] = (T)tile[threadIdx.x][i];
  }
}
}  // namespace
}  // namespace pio

static int pio_bank_fill_(PioBank* h, const float* bank, cudaStream_t st) {
  using namespace pio;
  dim3 grid((unsigned)((h->M + 31) / 32), (unsigned)(h->D / 32));
  if (h->act_dt == PIO_DT_F32)
    bank_transpose_kernel<float><<<grid, dim3(32, 8), 0, st>>>(bank, (float*)h->bankT, (float*)h->bank, h->M, h->D, h->Mp);
  else
    bank_transpose_kernel<__nv_bfloat16><<<grid, dim3(32, 8), 0, st>>>(bank, (__nv_bfloat16*)h->bankT, (__nv_bfloat16*)h->bank, h->M, h->D, h->Mp);
  PIO_LAUNCHED();
  return PIO_OK;
}

extern "C" {

int pio_bank_create(PioBank** out, const float* bank, long long M, int D, int mode, void* stream) {
  using namespace pio;
  PIO_CHECK(out && bank && M > 0, "bank_create: bad arguments");
  PIO_CHECK(D % 128 == 0, "bank_create: D %d must be a multiple of 128", D);
  PIO_CHECK(M < (1ll << 31) - 1024, "bank_create: too many rows for one shard");
  cudaStream_t st = as_stream(stream);
  PioBank* h = new PioBank();
  memset(h, 0, sizeof(*h));
  h->mode = mode; h->act_dt = mode == PIO_FP32 ? PIO_DT_F32 : PIO_DT_BF16; h->D = D; h->M = M;
  h->Mp = (M + 127) / 128 * 128;
  const size_t e = h->act_dt == PIO_DT_F32 ? 4 : 2;
  auto go = [&]() -> int {
    PIO_CUDA(cudaMalloc(&h->bank, (size_t)M * D * e));
    PIO_CUDA(cudaMalloc(&h->bankT, (size_t)D * h->Mp * e));
    PIO_CUDA(cudaMalloc((void**)&h->inv_norm, (size_t)M * sizeof(float)));
    PIO_CUDA(cudaMemsetAsync(h->bankT, 0, (size_t)D * h->Mp * e, st));
    row_inv_norm_kernel<<<cdiv(M * 32, 256), 256, 0, st>>>(bank, h->inv_norm, M, D);
    PIO_LAUNCHED();
    return PIO_OK;
  };
  int rc = go();
  if (rc != PIO_OK) { pio_bank_destroy(h); return rc; }
  rc = pio_bank_fill_(h, bank, st);  // bank + transposed copy in the activation dtype
  if (rc != PIO_OK) { pio_bank_destroy(h); return rc; }
  *out = h;
  return PIO_OK;
}
}

extern "C" {

void pio_bank_destroy(PioBank* h) {
  if (!h) return;
  cudaFree(h->bank); cudaFree(h->bankT); cudaFree(h->inv_norm);
  delete h;
}
long long pio_bank_rows(const PioBank* h) { return h ? h->M : 0; }

static bool project_exact_requested() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("PIO_PROJECT_EXACT"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}

static int project_chunk_rows(int R) {
  // S chunk (R x Mc fp32, or P bf16 with twice the columns on the fused path).  Measured at R = 4096, M = 591 753 (B200):
  // 16 M elements (P = 67 MB, L2 resident) 8.7 ms, 32 M 7.8 ms, 64 M 7.5 ms, 128 M 8.1 ms -- fewer, longer GEMMs beat L2
  // residency (the P round trip runs at ~3 TB/s, well inside HBM bandwidth).  PIO_PROJECT_CHUNK_M overrides for A/B runs.
  // Few queries (R = 32 ... 256: image-level captions, traces, small batches) are launch-bound, not bandwidth-bound: the chunk
  // may then grow to 131072 rows (3 chunks instead of 19 for the 591 753-row bank).  Both limits are read per call so that
  // tests can force many small chunks (PIO_PROJECT_MAX_CHUNK).
  const char* e = getenv("PIO_PROJECT_CHUNK_M");
  const char* c = getenv("PIO_PROJECT_MAX_CHUNK");
  const long long elems = (long long)(e ? atoi(e) : 64) << 20, cap = c ? atoi(c) : 131072;
  long long mc = elems / (R > 0 ? R : 1);
  mc = mc / 128 * 128;
  if (mc > cap) mc = cap / 128 * 128;
  if (mc < 1024) mc = 1024;
  return (int)mc;
}

size_t pio_project_workspace_bytes(const PioBank* h, int R) {
  using namespace pio;
  const size_t Mc = project_chunk_rows(R), e = h->act_dt == PIO_DT_F32 ? 4 : 2;
  return align_up((size_t)R * h->D * e, 1024) /*qn*/ + align_up((size_t)R * h->D * 4, 1024) /*qn fp32*/ +
         align_up((size_t)R * Mc * 4, 1024) /*S, or P + partials on the fused path*/ + align_up((size_t)R * Mc * 2, 1024) /*P16*/ +
         4 * align_up((size_t)R * 4, 1024) + 4096;
}

int pio_project(PioBank* h, const float* q, int R, float temperature, int normalize, float* out, float* part_m, float* part_l,
                void* workspace, size_t workspace_bytes, void* stream) {
  using namespace pio;
  if (R == 0) return PIO_OK;  // nothing to do (empty tensors have null data pointers)
  PIO_CHECK(h && q && out && workspace, "project: null argument");
  PIO_CHECK(temperature > 0.f, "project: temperature must be positive");
  PIO_CHECK(workspace_bytes >= pio_project_workspace_bytes(h, R), "project: workspace too small");
  PIO_CHECK((((uintptr_t)workspace) & 1023) == 0, "project: workspace must be 1024-byte aligned");
  PIO_CHECK((part_m == nullptr) == (part_l == nullptr), "project: part_m and part_l go together");
  if (R == 0) return PIO_OK;
  cudaStream_t st = as_stream(stream);
  const int D = h->D, Mc = project_chunk_rows(R), adt = h->act_dt;
  const size_t e = adt == PIO_DT_F32 ? 4 : 2;
  char* ws = (char*)workspace;
  void* qn = ws;            ws += align_up((size_t)R * D * e, 1024);
  float* qn32 = (float*)ws; ws += align_up((size_t)R * D * 4, 1024);
  float* S = (float*)ws;    ws += align_up((size_t)R * Mc * 4, 1024);
  __nv_bfloat16* P16 = (__nv_bfloat16*)ws; ws += align_up((size_t)R * Mc * 2, 1024);
  float* m = (float*)ws;    ws += align_up((size_t)R * 4, 1024);
  float* l = (float*)ws;    ws += align_up((size_t)R * 4, 1024);
  float* alpha = (float*)ws; ws += align_up((size_t)R * 4, 1024);
  float* m_used = (float*)ws;

  PIO_CUDA(cudaMemcpyAsync(qn32, q, (size_t)R * D * 4, cudaMemcpyDeviceToDevice, st));
  PIO_TRY(l2norm_rows(qn32, R, D, st));                       // q / |q|  (im2txtprojection.py:368)
  if (adt == PIO_DT_F32) qn = qn32; else PIO_TRY(f32_to_bf16(qn32, (__nv_bfloat16*)qn, (long long)R * D, st));
  fill_kernel<<<cdiv(R, 256), 256, 0, st>>>(m, -INFINITY, R); PIO_LAUNCHED();
  PIO_CUDA(cudaMemsetAsync(l, 0, (size_t)R * 4, st));
  PIO_CUDA(cudaMemsetAsync(out, 0, (size_t)R * D * 4, st));

  if (adt == PIO_DT_BF16 && !project_exact_requested()) {
    // ---- fused path: P = exp2(s - m_ref) straight from the similarity GEMM's epilogue (no fp32 score matrix,
    // no separate softmax pass).  m_ref lags by one chunk and never drops below log2e * (1/T - 70), so with
    // |s| <= 1/T the exponent stays <= 70 / ln 2: no overflow for any input; accurate while max cosine >= -0.4.
    const float log2e = 1.4426950408889634f;
    const bool few_queries = R <= 768;
    // the S region (R x Mc fp32) hosts P (bf16, R x 2Mc columns) on this path: twice the chunk, same bytes
    const int Mf = 2 * Mc;
    __nv_bfloat16* P = (__nv_bfloat16*)S;
    const int slabs_max = argmax_slabs_tc(R, Mf);
    float* psum = (float*)P16;                       // [R, slabs_max]
    float* pmax = psum + (size_t)R * slabs_max;      // [R, slabs_max]   (R*Mc*2 bytes are available: plenty)
    // (round 1 zeroed all of P -- 268 MB at R = 4096 -- per call; only the <= 63 padding columns of the LAST chunk are ever read
    // without having been written: they are zeroed right before that chunk's similarity GEMM)
    fill_kernel<<<cdiv(R, 256), 256, 0, st>>>(m, log2e * (1.0f / temperature - 70.0f), R); PIO_LAUNCHED();
    fill_kernel<<<cdiv(R, 256), 256, 0, st>>>(alpha, 1.0f, R); PIO_LAUNCHED();
    for (long long c0 = 0; c0 < h->M; c0 += Mf) {
      const int mc = (int)std::min<long long>(Mf, h->M - c0);
      const int mc_pad = (mc + 63) / 64 * 64;
      const int slabs = argmax_slabs_tc(R, mc);
      if (mc_pad > mc)  // K padding of the recombination GEMM: stale bits there (NaN x 0) would poison O
        PIO_CUDA(cudaMemset2DAsync(P + mc, (size_t)Mf * 2, 0, (size_t)(mc_pad - mc) * 2, R, st));
      PioLinear p;
      memset(&p, 0, sizeof(p));
      p.A = qn; p.W = (const char*)h->bank + (size_t)c0 * D * e; p.C = P; p.M = R; p.N = mc; p.K = D;
      p.lda = D; p.ldw = D; p.ldc = Mf; p.a_dt = adt; p.c_dt = PIO_DT_BF16; p.colscale = h->inv_norm + c0;
      p.alpha = log2e / temperature;
      p.exp_ref = m; p.exp_psum = psum; p.exp_pmax = pmax; p.exp_ld = slabs_max; p.w_static = 1;
      PIO_TRY(linear_tc(p, st));
      memset(&p, 0, sizeof(p));
      p.A = P; p.W = (const char*)h->bankT + (size_t)c0 * e; p.C = out; p.M = R; p.N = D; p.K = mc_pad;
      p.lda = Mf; p.ldw = (int)h->Mp; p.ldc = D; p.a_dt = adt; p.c_dt = PIO_DT_F32; p.residual = out; p.ldres = D;
      p.alpha = 1.0f; p.w_static = 1;
      if (few_queries) {  // O *= alpha first, then a pure accumulation that the GEMM may split along K (too few output tiles)
        scale_rows_kernel<<<cdiv((long long)R * 32, 256), 256, 0, st>>>(out, alpha, R, D); PIO_LAUNCHED();
      } else {
        p.res_rowscale = alpha;
      }
      PIO_TRY(linear_tc(p, st));
      launch_pdl(project_update_kernel, dim3(cdiv((long long)R * 32, 256)), dim3(256), 0, st, psum, pmax, slabs_max, slabs, R, m, m_used, l,
                 alpha);
      PIO_LAUNCHED();
    }
    // O and l are relative to the reference the LAST chunk used (m_used); the ratio O / l does not depend on it
    if (part_m) {
      scale_vec_kernel<<<cdiv(R, 256), 256, 0, st>>>(m_used, part_m, 1.0f / log2e, R); PIO_LAUNCHED();  // natural-log units
      PIO_CUDA(cudaMemcpyAsync(part_l, l, (size_t)R * 4, cudaMemcpyDeviceToDevice, st));
      return PIO_OK;
    }
    finish_rows_kernel<<<cdiv((long long)R * 32, 256), 256, 0, st>>>(out, l, R, D, normalize);
    PIO_LAUNCHED();
    return PIO_OK;
  }

  for (long long c0 = 0; c0 < h->M; c0 += Mc) {
    const int mc = (int)std::min<long long>(Mc, h->M - c0);
    const int mc_pad = (mc + 63) / 64 * 64;
    PioLinear p;
    memset(&p, 0, sizeof(p));
    // S = (q^ . bank_j) * inv_norm_j / T                       (:367,370,376)
    p.A = qn; p.W = (const char*)h->bank + (size_t)c0 * D * e; p.C = S; p.M = R; p.N = mc; p.K = D;
    p.lda = D; p.ldw = D; p.ldc = Mc; p.a_dt = adt; p.c_dt = PIO_DT_F32; p.colscale = h->inv_norm + c0;
    p.alpha = 1.0f / temperature; p.w_static = 1;
    PIO_TRY(h->mode == PIO_FP32 ? linear_simt(p, st) : linear_tc(p, st));
    softmax_chunk_kernel<<<R, 256, 0, st>>>(S, Mc, adt == PIO_DT_F32 ? nullptr : P16, Mc, mc, mc_pad, m, l, alpha);
    PIO_LAUNCHED();
    // O = alpha * O + P . bank[c0:c0+mc]  (raw rows, :377) -- K-contiguous through the transposed copy
    memset(&p, 0, sizeof(p));
    p.A = adt == PIO_DT_F32 ? (const void*)S : (const void*)P16;
    p.W = (const char*)h->bankT + (size_t)c0 * e; p.C = out; p.M = R; p.N = D; p.K = mc_pad;
    p.lda = Mc; p.ldw = (int)h->Mp; p.ldc = D; p.a_dt = adt; p.c_dt = PIO_DT_F32; p.residual = out; p.ldres = D;
    p.res_rowscale = alpha; p.alpha = 1.0f; p.w_static = 1;
    PIO_TRY(h->mode == PIO_FP32 ? linear_simt(p, st) : linear_tc(p, st));
  }
  if (part_m) {
    PIO_CUDA(cudaMemcpyAsync(part_m, m, (size_t)R * 4, cudaMemcpyDeviceToDevice, st));
    PIO_CUDA(cudaMemcpyAsync(part_l, l, (size_t)R * 4, cudaMemcpyDeviceToDevice, st));
    return PIO_OK;
  }
  finish_rows_kernel<<<cdiv((long long)R * 32, 256), 256, 0, st>>>(out, l, R, D, normalize);
  PIO_LAUNCHED();
  return PIO_OK;
}

int pio_best_sims(PioBank* h, const float* q, int R, int n, float* out_sims, int* out_rows, void* workspace,
                  size_t workspace_bytes, void* stream) {
  using namespace pio;
  if (R == 0) return PIO_OK;
  PIO_CHECK(h && q && out_sims && workspace, "best_sims: null argument");
  PIO_CHECK(n >= 1 && n <= TOPN_MAX, "best_sims: n %d outside [1,%d]", n, TOPN_MAX);
  PIO_CHECK(workspace_bytes >= pio_project_workspace_bytes(h, R), "best_sims: workspace too small");
  PIO_CHECK((((uintptr_t)workspace) & 1023) == 0, "best_sims: workspace must be 1024-byte aligned");
  if (R == 0) return PIO_OK;
  cudaStream_t st = as_stream(stream);
  const int D = h->D, Mc = project_chunk_rows(R), adt = h->act_dt;
  const size_t e = adt == PIO_DT_F32 ? 4 : 2;
  char* ws = (char*)workspace;
  void* qn = ws;            ws += align_up((size_t)R * D * e, 1024);
  float* qn32 = (float*)ws; ws += align_up((size_t)R * D * 4, 1024);
  float* S = (float*)ws;    ws += align_up((size_t)R * Mc * 4, 1024);
  int* rows = (int*)ws;     // the P16 region: R * n ints fit easily (n <= 32 << Mc / 2)
  PIO_CUDA(cudaMemcpyAsync(qn32, q, (size_t)R * D * 4, cudaMemcpyDeviceToDevice, st));
  PIO_TRY(l2norm_rows(qn32, R, D, st));
  if (adt == PIO_DT_F32) qn = qn32; else PIO_TRY(f32_to_bf16(qn32, (__nv_bfloat16*)qn, (long long)R * D, st));
  int* best_i = out_rows ? out_rows : rows;
  fill_kernel<<<cdiv((long long)R * n, 256), 256, 0, st>>>(out_sims, -INFINITY, (long long)R * n); PIO_LAUNCHED();
  PIO_CUDA(cudaMemsetAsync(best_i, 0xff, (size_t)R * n * sizeof(int), st));
  for (long long c0 = 0; c0 < h->M; c0 += Mc) {
    const int mc = (int)std::min<long long>(Mc, h->M - c0);
    PioLinear p;
    memset(&p, 0, sizeof(p));
    p.A = qn; p.W = (const char*)h->bank + (size_t)c0 * D * e; p.C = S; p.M = R; p.N = mc; p.K = D;
    p.lda = D; p.ldw = D; p.ldc = Mc; p.a_dt = adt; p.c_dt = PIO_DT_F32; p.colscale = h->inv_norm + c0; p.alpha = 1.0f;
    p.w_static = 1;
    PIO_TRY(h->mode == PIO_FP32 ? linear_simt(p, st) : linear_tc(p, st));
    topn_chunk_kernel<<<R, 256, 0, st>>>(S, Mc, mc, c0, n, out_sims, best_i);
    PIO_LAUNCHED();
  }
  return PIO_OK;
}

int pio_project_rescale(float* O, float* l, const float* m_local, const float* m_global, int R, int D, void* stream) {
  using namespace pio;
  if (R == 0) return PIO_OK;
  rescale_rows_kernel<<<cdiv((long long)R * 32, 256), 256, 0, as_stream(stream)>>>(O, l, m_local, m_global, R, D);
  PIO_LAUNCHED();
  return PIO_OK;
}
int pio_project_finish(float* O, const float* l, int R, int D, int normalize, void* stream) {
  using namespace pio;
  if (R == 0) return PIO_OK;
  finish_rows_kernel<<<cdiv((long long)R * 32, 256), 256, 0, as_stream(stream)>>>(O, l, R, D, normalize);
  PIO_LAUNCHED();
  return PIO_OK;
}
}

// ==========================================================================================
// struct PioDecoder, the workspace layout and the fused-decode entry points: decoder.cuh
namespace pio {
namespace {
int d_own(PioDecoder* h, void** p, size_t bytes) {
  PIO_CUDA(cudaMalloc(p, bytes));
  h->owned.push_back(*p);
  return PIO_OK;
}
int d_f32(PioDecoder* h, const float** dst, const float* src, size_t n, cudaStream_t st) {
  void* p;
  PIO_TRY(d_own(h, &p, n * 4));
  PIO_CUDA(cudaMemcpyAsync(p, src, n * 4, cudaMemcpyDeviceToDevice, st));
  *dst = (const float*)p;
  return PIO_OK;
}
// Conv1D weight [in, out] -> Linear layout [out, in] in the activation dtype
int d_matT(PioDecoder* h, const void** dst, const float* src, int in, int outn, cudaStream_t st) {
  void* p;
  PIO_TRY(d_own(h, &p, (size_t)in * outn * (h->act_dt == PIO_DT_F32 ? 4 : 2)));
  if (h->act_dt == PIO_DT_F32) PIO_TRY(transpose_f32(src, (float*)p, in, outn, st));
  else PIO_TRY(transpose_to_bf16(src, (__nv_bfloat16*)p, in, outn, st));
  *dst = p;
  return PIO_OK;
}
int d_mat(PioDecoder* h, const void** dst, const float* src, size_t n, cudaStream_t st) {
  void* p;
  if (h->act_dt == PIO_DT_F32) return d_f32(h, (const float**)dst, src, n, st);
  PIO_TRY(d_own(h, &p, n * 2));
  PIO_TRY(f32_to_bf16(src, (__nv_bfloat16*)p, (long long)n, st));
  *dst = p;
  return PIO_OK;
}
}  // namespace
}  // namespace pio

extern "C" {

static int decoder_build(PioDecoder** out, const float* wte, const float* wpe, const float* lnf_w, const float* lnf_b,
                         const float* prefix_w, const float* prefix_b, int prefix_size, const PioGptBlock* blocks, int n_layer,
                         int n_head, int max_len, int mode, void* stream) {
  using namespace pio;
  cudaStream_t st = as_stream(stream);
  PioDecoder* h = new PioDecoder();
  h->mode = mode; h->act_dt = mode == PIO_FP32 ? PIO_DT_F32 : PIO_DT_BF16; h->prefix_size = prefix_size;
  h->L = n_layer; h->H = n_head; h->T = max_len;
  h->blk.resize(n_layer);
  h->prefix_w = nullptr; h->prefix_b0 = nullptr;
  auto go = [&]() -> int {
    PIO_TRY(d_f32(h, &h->wte32, wte, (size_t)gV * gD, st));
    if (h->act_dt == PIO_DT_F32) h->wte = h->wte32; else PIO_TRY(d_mat(h, &h->wte, wte, (size_t)gV * gD, st));
    PIO_TRY(d_f32(h, &h->wpe, wpe, (size_t)1024 * gD, st));
    PIO_TRY(d_f32(h, &h->lnf_w, lnf_w, gD, st)); PIO_TRY(d_f32(h, &h->lnf_b, lnf_b, gD, st));
    if (prefix_w) {
      PIO_TRY(d_mat(h, &h->prefix_w, prefix_w, (size_t)gD * prefix_size, st));
      void* pb;
      PIO_TRY(d_own(h, &pb, gD * 4));
      add_vec_kernel<<<cdiv(gD, 256), 256, 0, st>>>(prefix_b, wpe, (float*)pb, gD);
      PIO_LAUNCHED();
      h->prefix_b0 = (const float*)pb;
    }
    for (int i = 0; i < n_layer; ++i) {
      const PioGptBlock& s = blocks[i];
      PioDecoder::Blk& d = h->blk[i];
      PIO_TRY(d_f32(h, &d.ln1_w, s.ln1_w, gD, st)); PIO_TRY(d_f32(h, &d.ln1_b, s.ln1_b, gD, st));
      PIO_TRY(d_f32(h, &d.attn_b, s.attn_b, 3 * gD, st)); PIO_TRY(d_f32(h, &d.proj_b, s.proj_b, gD, st));
      PIO_TRY(d_f32(h, &d.ln2_w, s.ln2_w, gD, st)); PIO_TRY(d_f32(h, &d.ln2_b, s.ln2_b, gD, st));
      PIO_TRY(d_f32(h, &d.fc_b, s.fc_b, gFF, st)); PIO_TRY(d_f32(h, &d.fc2_b, s.fc2_b, gD, st));
      PIO_TRY(d_matT(h, &d.attn_w, s.attn_w, gD, 3 * gD, st));
      PIO_TRY(d_matT(h, &d.proj_w, s.proj_w, gD, gD, st));
      PIO_TRY(d_matT(h, &d.fc_w, s.fc_w, gD, gFF, st));
      PIO_TRY(d_matT(h, &d.fc2_w, s.fc2_w, gFF, gD, st));
    }
    PIO_TRY(decode_fused_build(h, st));  // tensor maps + layer table of the persistent small-batch decode kernel (bf16 mode)
    return PIO_OK;
  };
  int rc = go();
  if (rc != PIO_OK) { pio_decoder_destroy(h); return rc; }
  *out = h;
  return PIO_OK;
}

int pio_decoder_create(PioDecoder** out, const PioDecoderWeights* w, int mode, void* stream) {
  using namespace pio;
  PIO_CHECK(out && w, "decoder_create: null argument");
  PIO_CHECK(mode == PIO_FP32 || mode == PIO_BF16, "decoder_create: unknown mode %d", mode);
  PIO_CHECK(w->prefix_size > 0 && w->prefix_size % 64 == 0, "decoder_create: prefix_size %d must be a multiple of 64", w->prefix_size);
  return decoder_build(out, w->wte, w->wpe, w->lnf_w, w->lnf_b, w->prefix_w, w->prefix_b, w->prefix_size, w->blk, 4, 4, 32, mode,
                       stream);
}

int pio_decoder_create_gpt2(PioDecoder** out, const PioGpt2Weights* w, int mode, void* stream) {
  using namespace pio;
  PIO_CHECK(out && w && w->blk, "decoder_create_gpt2: null argument");
  PIO_CHECK(mode == PIO_FP32 || mode == PIO_BF16, "decoder_create_gpt2: unknown mode %d", mode);
  PIO_CHECK(w->n_layer >= 1 && w->n_layer <= 48, "decoder_create_gpt2: n_layer %d outside [1,48]", w->n_layer);
  PIO_CHECK(w->n_head == 12, "decoder_create_gpt2: n_head %d (only 12 heads x 64 are built)", w->n_head);
  return decoder_build(out, w->wte, w->wpe, w->lnf_w, w->lnf_b, nullptr, nullptr, 0, w->blk, w->n_layer, w->n_head, 128, mode, stream);
}

void pio_decoder_destroy(PioDecoder* h) {
  if (!h) return;
  for (void* p : h->owned) cudaFree(p);
  delete h;
}

namespace {
using pio::DecodeWs;
using pio::decode_ws;
// ------------------------------------------------------------------------------------------ beam search (viecap/search.py:193-285)
constexpr int kBeamMax = 8;
// One CTA per row of the fp32 logits: the W largest log-probabilities log softmax(logits / temperature) (descending; equal
// values: lower token id first) and their token ids.  Pass 1: every thread keeps the top W of its strided share in registers
// together with an online (max, sum of exponentials); pass 2: W rounds of a block-wide arg-max over the threads' list heads.
__global__ void __launch_bounds__(256) beam_topk_kernel(const float* __restrict__ logits, int ld, int V, int W, float inv_temp,
                                                        float* __restrict__ cand_lp, int* __restrict__ cand_id) {
  __shared__ float s_val[8];
  __shared__ int s_idx[8], s_thr[8];
  __shared__ float s_red[8], s_red2[8];
  const int row = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const float* x = logits + (long long)row * ld;
  float tv[kBeamMax];
  int ti[kBeamMax];
#pragma unroll
  for (int k = 0; k < kBeamMax; ++k) { tv[k] = -INFINITY; ti[k] = 0x7fffffff; }
  float mx = -INFINITY, sum = 0.f;
  for (int i = tid; i < V; i += 256) {
    const float v = x[i] * inv_temp;
    if (v > mx) { sum = sum * __expf(mx - v) + 1.f; mx = v; } else sum += __expf(v - mx);
    if (v > tv[kBeamMax - 1]) {  // strict: an equal value with a higher index never displaces (indices ascend per thread)
      tv[kBeamMax - 1] = v; ti[kBeamMax - 1] = i;
#pragma unroll
      for (int k = kBeamMax - 1; k > 0; --k) {
        if (tv[k] > tv[k - 1]) { const float a = tv[k]; tv[k] = tv[k - 1]; tv[k - 1] = a; const int b = ti[k]; ti[k] = ti[k - 1]; ti[k - 1] = b; }
      }
    }
  }
  // block (max, sum)
  float bm = pio::warp_max(mx);
  if (lane == 0) s_red[wid] = bm;
  __syncthreads();
  bm = s_red[0];
#pragma unroll
  for (int k = 1; k < 8; ++k) bm = fmaxf(bm, s_red[k]);
  float bs = pio::warp_sum(mx == -INFINITY ? 0.f : sum * __expf(mx - bm));
  if (lane == 0) s_red2[wid] = bs;
  __syncthreads();
  bs = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) bs += s_red2[k];
  const float lse = bm + logf(bs);
  int head = 0;  // this thread's next unconsumed list entry
  for (int k = 0; k < W; ++k) {
    float v = -INFINITY;
    int idx = 0x7fffffff;
#pragma unroll
    for (int q = 0; q < kBeamMax; ++q)
      if (q == head) { v = tv[q]; idx = ti[q]; }
    int thr = tid;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float v2 = __shfl_xor_sync(0xffffffffu, v, o);
      const int i2 = __shfl_xor_sync(0xffffffffu, idx, o), t2 = __shfl_xor_sync(0xffffffffu, thr, o);
      if (v2 > v || (v2 == v && i2 < idx)) { v = v2; idx = i2; thr = t2; }
    }
    if (lane == 0) { s_val[wid] = v; s_idx[wid] = idx; s_thr[wid] = thr; }
    __syncthreads();
    v = s_val[0]; idx = s_idx[0]; thr = s_thr[0];
#pragma unroll
    for (int q = 1; q < 8; ++q)
      if (s_val[q] > v || (s_val[q] == v && s_idx[q] < idx)) { v = s_val[q]; idx = s_idx[q]; thr = s_thr[q]; }
    if (tid == thr) ++head;
    if (tid == 0) { cand_lp[(long long)row * W + k] = v - lse; cand_id[(long long)row * W + k] = idx; }
    __syncthreads();
  }
}

struct BeamState {
  float* scores; float* seqlen; int* stopped; int* tok;  // [R * W]
  int* ids[2];   // [R * W, steps] generated tokens, ping-pong
  int* anc[2];   // [R * W, T] cache-row table, ping-pong
  int* active;   // [steps] live beams after each step
};

// One thread per region: step 0 opens the W beams from the region's single prompt row (search.py:234-243); later steps choose the
// W best of the W x W candidate continuations by length-normalised score (:244-262).  Writes the NEXT ids / cache-row tables.
__global__ void beam_select_kernel(BeamState b, int cur, const float* __restrict__ cand_lp, const int* __restrict__ cand_id, int R,
                                   int W, int steps, int T, int P, int s, int eos0, int eos1) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const int* ids_in = b.ids[cur];
  int* ids_out = b.ids[cur ^ 1];
  const int* anc_in = b.anc[cur];
  int* anc_out = b.anc[cur ^ 1];
  int live = 0;
  if (s == 0) {
    for (int k = 0; k < W; ++k) {
      const int j = r * W + k, id = cand_id[(long long)r * W + k];
      b.scores[j] = cand_lp[(long long)r * W + k];
      b.seqlen[j] = 1.f;
      const int stop = (id == eos0 || id == eos1);
      b.stopped[j] = stop;
      live += !stop;
      b.tok[j] = id;
      ids_out[(long long)j * steps] = id;
      for (int p = 0; p < P; ++p) anc_out[(long long)j * T + p] = r;  // the prompt lives in cache row r (prefill of R rows)
      if (P < T) anc_out[(long long)j * T + P] = j;
    }
    atomicAdd(b.active + s, live);
    return;
  }
  float sc[kBeamMax], sl[kBeamMax];
  int st[kBeamMax];
  for (int k = 0; k < W; ++k) { const int j = r * W + k; sc[k] = b.scores[j]; st[k] = b.stopped[j]; sl[k] = b.seqlen[j] + (st[k] ? 0.f : 1.f); }
  // selection: W rounds of arg-max over the candidates not taken yet (flattened index beam * V + token breaks ties)
  unsigned long long taken = 0ull;
  for (int k = 0; k < W; ++k) {
    float best = -INFINITY;
    int bb = -1, bc = -1, bid = 0;
    for (int q = 0; q < W; ++q) {
      const int nc = st[q] ? 1 : W;  // a stopped beam has one continuation: token 0 with log-probability 0 (:250-251)
      for (int c = 0; c < nc; ++c) {
        if (taken >> (q * kBeamMax + c) & 1ull) continue;
        const float lp = st[q] ? 0.f : cand_lp[((long long)r * W + q) * W + c];
        const int id = st[q] ? 0 : cand_id[((long long)r * W + q) * W + c];
        const float avg = (sc[q] + lp) / sl[q];
        if (avg > best || (avg == best && bb >= 0 && (q < bb || (q == bb && id < bid)))) { best = avg; bb = q; bc = c; bid = id; }
      }
    }
    if (bb < 0) { bb = 0; bc = 0; bid = 0; best = -INFINITY; }  // fewer than W finite candidates cannot happen (W live continuations or W stopped beams)
    taken |= 1ull << (bb * kBeamMax + bc);
    const int j = r * W + k, src = r * W + bb;
    b.seqlen[j] = sl[bb];
    b.scores[j] = best * sl[bb];
    const int stop = st[bb] | (bid == eos0 || bid == eos1);
    b.stopped[j] = stop;
    live += !stop;
    b.tok[j] = bid;
    for (int i = 0; i < s; ++i) ids_out[(long long)j * steps + i] = ids_in[(long long)src * steps + i];
    ids_out[(long long)j * steps + s] = bid;
    const int npos = min(T, P + s);
    for (int p = 0; p < npos; ++p) anc_out[(long long)j * T + p] = anc_in[(long long)src * T + p];
    if (P + s < T) anc_out[(long long)j * T + P + s] = j;
  }
  atomicAdd(b.active + s, live);
}

// greedy search with an early exit: done[r] latches once row r has emitted an end-of-sentence token; live[s] counts the rows
// still open after step s (the host reads it every few steps)
__global__ void eos_track_kernel(const int* __restrict__ ids, int ids_ld, int s, int eos0, int eos1, int* __restrict__ done,
                                 int* __restrict__ live, int R) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  int open = 0;
  if (r < R) {
    const int id = ids[(long long)r * ids_ld + s];
    const int d = done[r] | (id == eos0 || id == eos1);
    done[r] = d;
    open = !d;
  }
  open = __reduce_add_sync(0xffffffffu, open);
  if ((threadIdx.x & 31) == 0 && open) atomicAdd(live + s, open);
}
__global__ void fill_cols_kernel(int* __restrict__ ids, int ids_ld, int from, int R, int value) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int w = ids_ld - from;
  if (i >= (long long)R * w) return;
  ids[(i / w) * ids_ld + from + (i % w)] = value;
}

// search.py:278-283: length-normalised scores, beams best first (equal scores keep their beam order)
__global__ void beam_finish_kernel(BeamState b, int cur, int R, int W, int steps, int n_done, int* __restrict__ out_ids,
                                   int* __restrict__ out_len, float* __restrict__ out_score) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  float avg[kBeamMax];
  int ord[kBeamMax];
  for (int k = 0; k < W; ++k) { avg[k] = b.scores[r * W + k] / b.seqlen[r * W + k]; ord[k] = k; }
  for (int i = 1; i < W; ++i) {
    const int o = ord[i];
    int q = i - 1;
    while (q >= 0 && avg[ord[q]] < avg[o]) { ord[q + 1] = ord[q]; --q; }
    ord[q + 1] = o;
  }
  const int* ids = b.ids[cur];
  for (int k = 0; k < W; ++k) {
    const int src = r * W + ord[k], j = r * W + k;
    for (int i = 0; i < steps; ++i) out_ids[(long long)j * steps + i] = i < n_done ? ids[(long long)src * steps + i] : 0;
    out_len[j] = (int)b.seqlen[src];
    out_score[j] = avg[ord[k]];
  }
}

// all transformer blocks for the single new position t (KV cache of T positions per head)
int decode_blocks(PioDecoder* h, const DecodeWs& w, int R, int T, int t, cudaStream_t st, const int* anc = nullptr) {
  using namespace pio;
  const int adt = h->act_dt, mode = h->mode;
  float* x = w.x;
  for (int i = 0; i < h->L; ++i) {
    const PioDecoder::Blk& b = h->blk[i];
    PIO_TRY(layernorm(x, gD, b.ln1_w, b.ln1_b, w.hb, adt, gD, R, gD, 1e-5f, st));
    PIO_TRY(linear(mode, w.hb, b.attn_w, w.qkv, R, 3 * gD, gD, gD, gD, 3 * gD, adt, adt, b.attn_b, nullptr, PIO_ACT_NONE, st));
    PIO_TRY(decode_attention(w.qkv, w.kc + i * w.kv_layer, w.vc + i * w.kv_layer, w.hb, adt, R, h->H, T, t, st, anc));
    PIO_TRY(linear(mode, w.hb, b.proj_w, x, R, gD, gD, gD, gD, gD, adt, PIO_DT_F32, b.proj_b, x, PIO_ACT_NONE, st));
    PIO_TRY(layernorm(x, gD, b.ln2_w, b.ln2_b, w.hb, adt, gD, R, gD, 1e-5f, st));
    PIO_TRY(linear(mode, w.hb, b.fc_w, w.f, R, gFF, gD, gD, gD, gFF, adt, adt, b.fc_b, nullptr, PIO_ACT_GELU_NEW, st));
    PIO_TRY(linear(mode, w.f, b.fc2_w, x, R, gD, gFF, gFF, gFF, gD, adt, PIO_DT_F32, b.fc2_b, x, PIO_ACT_NONE, st));
  }
  return PIO_OK;
}

// ln_f + tied lm-head + greedy arg-max -> out_ids[:, col]
int decode_pick(PioDecoder* h, const DecodeWs& w, int R, int* out_ids, int ids_ld, int col, float* out_logprob_sum, cudaStream_t st) {
  using namespace pio;
  const int adt = h->act_dt, mode = h->mode;
  PIO_TRY(layernorm(w.x, gD, h->lnf_w, h->lnf_b, w.hb, adt, gD, R, gD, 1e-5f, st));
  if (mode == PIO_BF16) {
    // lm_head with the arg-max fused in the epilogue: the R x 50257 logits are never materialised
    const int slabs = argmax_slabs_tc(R, gV);
    float* av = w.logits;
    int* ai = (int*)(w.logits + (size_t)R * slabs);
    float* as = w.logits + 2 * (size_t)R * slabs;
    PioLinear p;
    memset(&p, 0, sizeof(p));
    p.A = w.hb; p.W = h->wte; p.C = nullptr; p.M = R; p.N = gV; p.K = gD; p.lda = gD; p.ldw = gD; p.ldc = gVld;
    p.a_dt = adt; p.c_dt = PIO_DT_F32; p.alpha = 1.0f; p.w_static = 1;
    p.argmax_val = av; p.argmax_idx = ai; p.argmax_sumexp = out_logprob_sum ? as : nullptr; p.argmax_ld = slabs;  // sum exp only for scores
    PIO_TRY(linear_tc(p, st));
    launch_pdl(argmax_finish_kernel, dim3(cdiv((long long)R * 32, 256)), dim3(256), 0, st, av, ai, as, slabs, slabs, R, out_ids, ids_ld, col,
               out_logprob_sum);
    PIO_LAUNCHED();
  } else {
    PIO_TRY(linear(mode, w.hb, h->wte, w.logits, R, gV, gD, gD, gD, gVld, adt, PIO_DT_F32, nullptr, nullptr, PIO_ACT_NONE, st));
    argmax_rows_kernel<<<R, 256, 0, st>>>(w.logits, gVld, gV, out_ids, ids_ld, col, out_logprob_sum);
    PIO_LAUNCHED();
  }
  return PIO_OK;
}
}  // namespace

int pio_decode_debug_layout(const PioDecoder* h, int R, long long* offsets, int n) {
  using namespace pio;
  PIO_CHECK(h && offsets && n >= 8, "decode_debug_layout: need room for 8 offsets");
  const DecodeWs w = decode_ws(h, nullptr, R, h->T, (size_t)R * h->prefix_size);
  const char* b = nullptr;
  offsets[0] = (const char*)w.x - b; offsets[1] = (const char*)w.hb - b; offsets[2] = (const char*)w.qkv - b;
  offsets[3] = (const char*)w.f - b; offsets[4] = (const char*)w.att - b; offsets[5] = (const char*)w.kc - b;
  offsets[6] = (const char*)w.vc - b; offsets[7] = (const char*)w.pm_val - b;
  return PIO_OK;
}

size_t pio_decode_workspace_bytes(const PioDecoder* h, int R, int steps) {
  (void)steps;
  return decode_ws(h, nullptr, R, h->T, (size_t)R * h->prefix_size).total;
}

int pio_decode_greedy(PioDecoder* h, const float* prefix, int R, int steps, int* out_ids, float* out_logprob_sum,
                      void* workspace, size_t workspace_bytes, void* stream) {
  using namespace pio;
  if (R == 0) return PIO_OK;
  PIO_CHECK(h && prefix && out_ids && workspace, "decode_greedy: null argument");
  PIO_CHECK(h->prefix_w, "decode_greedy: this decoder has no prefix projection (use pio_decode_greedy_prompt)");
  PIO_CHECK(steps >= 1 && steps <= h->T, "decode_greedy: steps %d outside [1,%d]", steps, h->T);
  PIO_CHECK(workspace_bytes >= pio_decode_workspace_bytes(h, R, steps), "decode_greedy: workspace too small");
  PIO_CHECK((((uintptr_t)workspace) & 1023) == 0, "decode_greedy: workspace must be 1024-byte aligned");
  cudaStream_t st = as_stream(stream);
  const int adt = h->act_dt, mode = h->mode, T = h->T;
  const DecodeWs w = decode_ws(h, (char*)workspace, R, T, (size_t)R * h->prefix_size);

  // prefix embedding at position 0: clip_project(feats) + wpe[0]   (decap.py:124; GPT-2 adds wpe)
  const void* pA = prefix;
  if (adt != PIO_DT_F32) { PIO_TRY(f32_to_bf16(prefix, (__nv_bfloat16*)w.pfx, (long long)R * h->prefix_size, st)); pA = w.pfx; }
  PIO_TRY(linear(mode, pA, h->prefix_w, w.x, R, gD, h->prefix_size, h->prefix_size, h->prefix_size, gD, adt, PIO_DT_F32,
                 h->prefix_b0, nullptr, PIO_ACT_NONE, st));
  if (out_logprob_sum) PIO_CUDA(cudaMemsetAsync(out_logprob_sum, 0, (size_t)R * 4, st));

  // small batches: the whole loop below as ONE persistent kernel (decode_fused_sm100.cu) instead of ~32 launches per position
  if (decode_fused_eligible(h, R, out_logprob_sum != nullptr)) return decode_fused(h, w, R, T, steps, 0, false, out_ids, st);

  for (int t = 0; t < steps; ++t) {
    PIO_TRY(decode_blocks(h, w, R, T, t, st));
    PIO_TRY(decode_pick(h, w, R, out_ids, steps, t, out_logprob_sum, st));
    if (t + 1 < steps) {
      launch_pdl(embed_kernel, dim3(cdiv((long long)R * 32, 256)), dim3(256), 0, st, h->wte32, h->wpe, (const int*)out_ids, steps, t, t + 1,
                 w.x, R, gD);
      PIO_LAUNCHED();
    }
  }
  return PIO_OK;
}

namespace {
// the prompt is prefilled in one batched pass when its causal attention fits the small-attention kernel
bool prompt_prefill_batched(const PioDecoder* h, int prompt_len) {
  static const bool off = [] { const char* e = getenv("PIO_PROMPT_PREFILL"); return e && e[0] == '0'; }();
  return !off && prompt_len >= 2 && pio::small_attention_smem(prompt_len, pio::gD / h->H) <= 48 * 1024;
}
}  // namespace

namespace {
size_t prompt_staging_bytes(int R, bool batched) { return pio::align_up(batched ? (size_t)R * pio::gD * 4 : 0, 256); }
size_t prompt_tail_elems(const PioDecoder* h, int R, bool batched) {
  const size_t e = h->act_dt == PIO_DT_F32 ? 4 : 2;
  return (prompt_staging_bytes(R, batched) + ((size_t)R + 128) * sizeof(int) + e - 1) / e;
}
}  // namespace

size_t pio_decode_prompt_workspace_bytes(const PioDecoder* h, int R, int prompt_len, int steps) {
  const bool batched = prompt_prefill_batched(h, prompt_len);
  // tail: fp32 [R,768] staging row for the last prompt position (prefill), then the early-exit flags ([R] done + [128] live counters)
  return decode_ws(h, nullptr, R, prompt_len + steps - 1, prompt_tail_elems(h, R, batched), batched ? (size_t)R * prompt_len : 0).total;
}

// Greedy continuation of a prompt of P input embeddings per row (ViECap: soft + hard prompt; viecap/search.py:108-191).
// The prompt is consumed one position at a time through the same single-position blocks as the generation steps.
static int decode_greedy_prompt_impl(PioDecoder* h, const float* prompt, int R, int prompt_len, int steps, int eos0, int eos1,
                                     int* out_ids, float* out_logprob_sum, int* out_steps_run, void* workspace, size_t workspace_bytes,
                                     void* stream);
int pio_decode_greedy_prompt(PioDecoder* h, const float* prompt, int R, int prompt_len, int steps, int* out_ids,
                             float* out_logprob_sum, void* workspace, size_t workspace_bytes, void* stream) {
  return decode_greedy_prompt_impl(h, prompt, R, prompt_len, steps, -1, -1, out_ids, out_logprob_sum, nullptr, workspace, workspace_bytes,
                                   stream);
}
// The same search, stopped once EVERY row has emitted an end-of-sentence token: the reference runs all `steps` positions for a batch
// (search.py:173-176 exits early only at batch 1) and then cuts every row after its first '.' (:184-190), so the tokens it keeps
// are the same.  Columns from *out_steps_run on are filled with eos0.  The stream is synchronised every eighth step (one 4-byte
// read-back).  With out_logprob_sum the search runs to the end (the sum covers every step, decap.py:157-160).
int pio_decode_greedy_prompt_eos(PioDecoder* h, const float* prompt, int R, int prompt_len, int steps, int eos0, int eos1, int* out_ids,
                                 float* out_logprob_sum, int* out_steps_run, void* workspace, size_t workspace_bytes, void* stream) {
  return decode_greedy_prompt_impl(h, prompt, R, prompt_len, steps, eos0, eos1, out_ids, out_logprob_sum, out_steps_run, workspace,
                                   workspace_bytes, stream);
}
static int decode_greedy_prompt_impl(PioDecoder* h, const float* prompt, int R, int prompt_len, int steps, int eos0, int eos1,
                                     int* out_ids, float* out_logprob_sum, int* out_steps_run, void* workspace, size_t workspace_bytes,
                                     void* stream) {
  using namespace pio;
  if (out_steps_run) *out_steps_run = 0;
  if (R == 0) return PIO_OK;
  PIO_CHECK(h && prompt && out_ids && workspace, "decode_greedy_prompt: null argument");
  const int T = prompt_len + steps - 1;
  PIO_CHECK(prompt_len >= 1 && steps >= 1 && T <= h->T, "decode_greedy_prompt: %d prompt + %d new positions exceed the cache of %d",
            prompt_len, steps, h->T);
  PIO_CHECK(workspace_bytes >= pio_decode_prompt_workspace_bytes(h, R, prompt_len, steps), "decode_greedy_prompt: workspace too small");
  PIO_CHECK((((uintptr_t)workspace) & 1023) == 0, "decode_greedy_prompt: workspace must be 1024-byte aligned");
  cudaStream_t st = as_stream(stream);
  const bool batched = prompt_prefill_batched(h, prompt_len);
  const DecodeWs w = decode_ws(h, (char*)workspace, R, T, prompt_tail_elems(h, R, batched), batched ? (size_t)R * prompt_len : 0);
  if (out_logprob_sum) PIO_CUDA(cudaMemsetAsync(out_logprob_sum, 0, (size_t)R * 4, st));
  if (batched) {
    // prefill (search.py:150-153): all prompt positions of all rows in one pass, M = R * prompt_len rows per GEMM
    const int adt = h->act_dt, mode = h->mode, P = prompt_len, rows = R * P;
    const size_t e = adt == PIO_DT_F32 ? 4 : 2;
    prompt_embed_all_kernel<<<cdiv((long long)rows * 32, 256), 256, 0, st>>>(prompt, P, h->wpe, w.x, rows, gD);
    PIO_LAUNCHED();
    for (int i = 0; i < h->L; ++i) {
      const PioDecoder::Blk& b = h->blk[i];
      PIO_TRY(layernorm(w.x, gD, b.ln1_w, b.ln1_b, w.hb, adt, gD, rows, gD, 1e-5f, st));
      PIO_TRY(linear(mode, w.hb, b.attn_w, w.qkv, rows, 3 * gD, gD, gD, gD, 3 * gD, adt, adt, b.attn_b, nullptr, PIO_ACT_NONE, st));
      PIO_TRY(small_attention(w.qkv, 3 * gD, (char*)w.qkv + (size_t)gD * e, 3 * gD, w.hb, gD, adt, R, P, h->H, gD / h->H, true,
                              w.kc + i * w.kv_layer, w.vc + i * w.kv_layer, T, st));
      PIO_TRY(linear(mode, w.hb, b.proj_w, w.x, rows, gD, gD, gD, gD, gD, adt, PIO_DT_F32, b.proj_b, w.x, PIO_ACT_NONE, st));
      PIO_TRY(layernorm(w.x, gD, b.ln2_w, b.ln2_b, w.hb, adt, gD, rows, gD, 1e-5f, st));
      PIO_TRY(linear(mode, w.hb, b.fc_w, w.f, rows, gFF, gD, gD, gD, gFF, adt, adt, b.fc_b, nullptr, PIO_ACT_GELU_NEW, st));
      PIO_TRY(linear(mode, w.f, b.fc2_w, w.x, rows, gD, gFF, gFF, gFF, gD, adt, PIO_DT_F32, b.fc2_b, w.x, PIO_ACT_NONE, st));
    }
    // the residual stream of the LAST prompt position becomes the single-position state (rows 0..R-1 of x)
    float* last = (float*)w.pfx;
    PIO_CUDA(cudaMemcpy2DAsync(last, (size_t)gD * 4, w.x + (size_t)(P - 1) * gD, (size_t)P * gD * 4, (size_t)gD * 4, R,
                               cudaMemcpyDeviceToDevice, st));
    PIO_CUDA(cudaMemcpyAsync(w.x, last, (size_t)R * gD * 4, cudaMemcpyDeviceToDevice, st));
  } else {
    for (int p = 0; p < prompt_len; ++p) {
      prompt_embed_kernel<<<cdiv((long long)R * 32, 256), 256, 0, st>>>(prompt, prompt_len, p, h->wpe, w.x, R, gD);
      PIO_LAUNCHED();
      PIO_TRY(decode_blocks(h, w, R, T, p, st));
    }
  }
  // generation: pick(0) on the last prompt position's residual stream, then embed -> blocks -> pick per new position --
  // one persistent kernel at small batch (decode_fused_sm100.cu)
  if (out_steps_run) *out_steps_run = steps;
  if (decode_fused_eligible(h, R, out_logprob_sum != nullptr)) return decode_fused(h, w, R, T, steps, prompt_len - 1, true, out_ids, st);
  const bool early = eos0 >= 0 && out_logprob_sum == nullptr && steps > 8;
  int* done = (int*)((char*)w.pfx + prompt_staging_bytes(R, batched));  // [R] flags, then [steps <= 128] counters
  int* live = done + R;
  if (early) PIO_CUDA(cudaMemsetAsync(done, 0, (size_t)(R + steps) * sizeof(int), st));
  for (int s = 0; s < steps; ++s) {
    PIO_TRY(decode_pick(h, w, R, out_ids, steps, s, out_logprob_sum, st));
    if (early) {
      eos_track_kernel<<<cdiv(R, 128), 128, 0, st>>>(out_ids, steps, s, eos0, eos1, done, live, R);
      PIO_LAUNCHED();
      if ((s & 7) == 7 && s + 1 < steps) {
        int open = 0;
        PIO_CUDA(cudaMemcpyAsync(&open, live + s, sizeof(int), cudaMemcpyDeviceToHost, st));
        PIO_CUDA(cudaStreamSynchronize(st));
        if (open == 0) {
          const long long n = (long long)R * (steps - s - 1);
          fill_cols_kernel<<<cdiv(n, 256), 256, 0, st>>>(out_ids, steps, s + 1, R, eos0);
          PIO_LAUNCHED();
          if (out_steps_run) *out_steps_run = s + 1;
          return PIO_OK;
        }
      }
    }
    if (s + 1 < steps) {
      launch_pdl(embed_kernel, dim3(cdiv((long long)R * 32, 256)), dim3(256), 0, st, h->wte32, h->wpe, (const int*)out_ids, steps, s,
                 prompt_len + s, w.x, R, gD);
      PIO_LAUNCHED();
      PIO_TRY(decode_blocks(h, w, R, T, prompt_len + s, st));
    }
  }
  return PIO_OK;
}

namespace {
struct BeamLayout { size_t staging, cand_lp, cand_id, scores, seqlen, stopped, tok, ids0, ids1, anc0, anc1, active, total; };
BeamLayout beam_layout(int R, int W, int steps, int T) {
  BeamLayout l;
  size_t o = 0;
  auto take = [&](size_t bytes) { const size_t at = o; o += pio::align_up(bytes, 256); return at; };
  const size_t rows = (size_t)R * W;
  l.staging = take((size_t)R * pio::gD * 4);
  l.cand_lp = take(rows * W * 4); l.cand_id = take(rows * W * 4);
  l.scores = take(rows * 4); l.seqlen = take(rows * 4); l.stopped = take(rows * 4); l.tok = take(rows * 4);
  l.ids0 = take(rows * steps * 4); l.ids1 = take(rows * steps * 4);
  l.anc0 = take(rows * T * 4); l.anc1 = take(rows * T * 4);
  l.active = take((size_t)steps * 4);
  l.total = o;
  return l;
}
}  // namespace

size_t pio_decode_beam_workspace_bytes(const PioDecoder* h, int R, int prompt_len, int steps, int beam_width) {
  const int T = prompt_len + steps - 1;
  const size_t e = h->act_dt == PIO_DT_F32 ? 4 : 2;
  const size_t rows = std::max((size_t)R * beam_width, (size_t)R * prompt_len);
  return pio::decode_ws(h, nullptr, R * beam_width, T, (beam_layout(R, beam_width, steps, T).total + e - 1) / e, rows).total;
}

// beam_search (viecap/search.py:193-285) for R prompts at once: W beams per prompt, every step decodes all R * W beam rows as one
// batch with a KV cache (the reference re-runs the whole sequence of each beam, one region at a time).  Beams are re-ordered
// through a per-(row, position) cache-row table instead of copying the cache; per-row top-W candidates come from beam_topk_kernel.
int pio_decode_beam_prompt(PioDecoder* h, const float* prompt, int R, int prompt_len, int steps, int beam_width, int eos0, int eos1,
                           float temperature, int* out_ids, int* out_len, float* out_score, int* out_steps_run, void* workspace,
                           size_t workspace_bytes, void* stream) {
  using namespace pio;
  if (out_steps_run) *out_steps_run = 0;
  if (R == 0) return PIO_OK;
  PIO_CHECK(h && prompt && out_ids && out_len && out_score && workspace, "decode_beam_prompt: null argument");
  PIO_CHECK(h->H == 12, "decode_beam_prompt: built for the 12-head GPT-2 decoder (pio_decoder_create_gpt2)");
  PIO_CHECK(beam_width >= 1 && beam_width <= kBeamMax, "decode_beam_prompt: beam width %d outside [1,%d]", beam_width, kBeamMax);
  const int T = prompt_len + steps - 1, W = beam_width, rows = R * W;
  PIO_CHECK(prompt_len >= 1 && steps >= 1 && T <= h->T, "decode_beam_prompt: %d prompt + %d new positions exceed the cache of %d",
            prompt_len, steps, h->T);
  PIO_CHECK(workspace_bytes >= pio_decode_beam_workspace_bytes(h, R, prompt_len, steps, W), "decode_beam_prompt: workspace too small");
  PIO_CHECK((((uintptr_t)workspace) & 1023) == 0, "decode_beam_prompt: workspace must be 1024-byte aligned");
  cudaStream_t st = as_stream(stream);
  const int adt = h->act_dt, mode = h->mode, P = prompt_len;
  const size_t e = adt == PIO_DT_F32 ? 4 : 2;
  const BeamLayout lay = beam_layout(R, W, steps, T);
  const DecodeWs w = decode_ws(h, (char*)workspace, rows, T, (lay.total + e - 1) / e, std::max((size_t)rows, (size_t)R * P));
  char* tail = (char*)w.pfx;
  float* cand_lp = (float*)(tail + lay.cand_lp);
  int* cand_id = (int*)(tail + lay.cand_id);
  BeamState bs;
  bs.scores = (float*)(tail + lay.scores); bs.seqlen = (float*)(tail + lay.seqlen); bs.stopped = (int*)(tail + lay.stopped);
  bs.tok = (int*)(tail + lay.tok);
  bs.ids[0] = (int*)(tail + lay.ids0); bs.ids[1] = (int*)(tail + lay.ids1);
  bs.anc[0] = (int*)(tail + lay.anc0); bs.anc[1] = (int*)(tail + lay.anc1);
  bs.active = (int*)(tail + lay.active);
  PIO_CUDA(cudaMemsetAsync(bs.active, 0, (size_t)steps * 4, st));
  const float inv_temp = 1.0f / (temperature > 0.f ? temperature : 1.0f);  // search.py:232

  // prefill: all prompt positions of the R prompts in one pass; their keys / values go to cache rows 0..R-1
  {
    const int prow = R * P;
    prompt_embed_all_kernel<<<cdiv((long long)prow * 32, 256), 256, 0, st>>>(prompt, P, h->wpe, w.x, prow, gD);
    PIO_LAUNCHED();
    if (P >= 2 && small_attention_smem(P, gD / h->H) <= 48 * 1024) {
      for (int i = 0; i < h->L; ++i) {
        const PioDecoder::Blk& b = h->blk[i];
        PIO_TRY(layernorm(w.x, gD, b.ln1_w, b.ln1_b, w.hb, adt, gD, prow, gD, 1e-5f, st));
        PIO_TRY(linear(mode, w.hb, b.attn_w, w.qkv, prow, 3 * gD, gD, gD, gD, 3 * gD, adt, adt, b.attn_b, nullptr, PIO_ACT_NONE, st));
        PIO_TRY(small_attention(w.qkv, 3 * gD, (char*)w.qkv + (size_t)gD * e, 3 * gD, w.hb, gD, adt, R, P, h->H, gD / h->H, true,
                                w.kc + i * w.kv_layer, w.vc + i * w.kv_layer, T, st));
        PIO_TRY(linear(mode, w.hb, b.proj_w, w.x, prow, gD, gD, gD, gD, gD, adt, PIO_DT_F32, b.proj_b, w.x, PIO_ACT_NONE, st));
        PIO_TRY(layernorm(w.x, gD, b.ln2_w, b.ln2_b, w.hb, adt, gD, prow, gD, 1e-5f, st));
        PIO_TRY(linear(mode, w.hb, b.fc_w, w.f, prow, gFF, gD, gD, gD, gFF, adt, adt, b.fc_b, nullptr, PIO_ACT_GELU_NEW, st));
        PIO_TRY(linear(mode, w.f, b.fc2_w, w.x, prow, gD, gFF, gFF, gFF, gD, adt, PIO_DT_F32, b.fc2_b, w.x, PIO_ACT_NONE, st));
      }
      float* last = (float*)(tail + lay.staging);
      PIO_CUDA(cudaMemcpy2DAsync(last, (size_t)gD * 4, w.x + (size_t)(P - 1) * gD, (size_t)P * gD * 4, (size_t)gD * 4, R,
                                 cudaMemcpyDeviceToDevice, st));
      PIO_CUDA(cudaMemcpyAsync(w.x, last, (size_t)R * gD * 4, cudaMemcpyDeviceToDevice, st));
    } else {
      for (int p = 0; p < P; ++p) {  // position by position through the single-position blocks (cache rows 0..R-1 of the R * W-row cache)
        prompt_embed_kernel<<<cdiv((long long)R * 32, 256), 256, 0, st>>>(prompt, P, p, h->wpe, w.x, R, gD);
        PIO_LAUNCHED();
        PIO_TRY(decode_blocks(h, w, R, T, p, st));
      }
    }
  }
  int cur = 0, done = 0;
  for (int s = 0; s < steps; ++s) {
    const int m = s == 0 ? R : rows;
    if (s > 0) {
      embed_kernel<<<cdiv((long long)rows * 32, 256), 256, 0, st>>>(h->wte32, h->wpe, bs.tok, 1, 0, P + s - 1, w.x, rows, gD);
      PIO_LAUNCHED();
      PIO_TRY(decode_blocks(h, w, rows, T, P + s - 1, st, bs.anc[cur]));
    }
    PIO_TRY(layernorm(w.x, gD, h->lnf_w, h->lnf_b, w.hb, adt, gD, m, gD, 1e-5f, st));
    PIO_TRY(linear(mode, w.hb, h->wte, w.logits, m, gV, gD, gD, gD, gVld, adt, PIO_DT_F32, nullptr, nullptr, PIO_ACT_NONE, st));
    beam_topk_kernel<<<m, 256, 0, st>>>(w.logits, gVld, gV, W, inv_temp, cand_lp, cand_id);
    PIO_LAUNCHED();
    beam_select_kernel<<<cdiv(R, 64), 64, 0, st>>>(bs, cur, cand_lp, cand_id, R, W, steps, T, P, s, eos0, eos1);
    PIO_LAUNCHED();
    cur ^= 1;
    done = s + 1;
    if ((s & 3) == 3 || s + 1 == steps) {  // search.py:275-276: stop once every beam has ended (checked every fourth step: one small sync)
      int live = 0;
      PIO_CUDA(cudaMemcpyAsync(&live, bs.active + s, sizeof(int), cudaMemcpyDeviceToHost, st));
      PIO_CUDA(cudaStreamSynchronize(st));
      if (live == 0) break;
    }
  }
  beam_finish_kernel<<<cdiv(R, 64), 64, 0, st>>>(bs, cur, R, W, steps, done, out_ids, out_len, out_score);
  PIO_LAUNCHED();
  if (out_steps_run) *out_steps_run = done;
  return PIO_OK;
}

size_t pio_gpt2_score_workspace_bytes(const PioDecoder* h, int R, int n) {
  return decode_ws(h, nullptr, R, 1, (size_t)R * n * 2, (size_t)R * n).total;
}

// Mean negative log-likelihood of each right-padded token row under the language model (ViECap compute_scores:
// entrypoint.py:164-177 -- GPT2LMHeadModel(input_ids, labels=input_ids).loss per sentence; perplexity = exp of it).
int pio_gpt2_score_tokens(PioDecoder* h, const int* ids, const int* lens, int R, int n, float* out_nll_mean, void* workspace,
                          size_t workspace_bytes, void* stream) {
  using namespace pio;
  if (R == 0) return PIO_OK;
  PIO_CHECK(h && ids && lens && out_nll_mean && workspace, "gpt2_score_tokens: null argument");
  PIO_CHECK(n >= 1 && n <= 128, "gpt2_score_tokens: %d tokens per row outside [1,128]", n);
  PIO_CHECK(workspace_bytes >= pio_gpt2_score_workspace_bytes(h, R, n), "gpt2_score_tokens: workspace too small");
  PIO_CHECK((((uintptr_t)workspace) & 1023) == 0, "gpt2_score_tokens: workspace must be 1024-byte aligned");
  cudaStream_t st = as_stream(stream);
  const int adt = h->act_dt, mode = h->mode, rows = R * n;
  const size_t e = adt == PIO_DT_F32 ? 4 : 2;
  const DecodeWs w = decode_ws(h, (char*)workspace, R, 1, (size_t)rows * 2, (size_t)rows);
  float* nll = (float*)w.pfx;
  embed_all_kernel<<<cdiv((long long)rows * 32, 256), 256, 0, st>>>(h->wte32, h->wpe, ids, n, w.x, rows, gD);
  PIO_LAUNCHED();
  for (int i = 0; i < h->L; ++i) {
    const PioDecoder::Blk& b = h->blk[i];
    PIO_TRY(layernorm(w.x, gD, b.ln1_w, b.ln1_b, w.hb, adt, gD, rows, gD, 1e-5f, st));
    PIO_TRY(linear(mode, w.hb, b.attn_w, w.qkv, rows, 3 * gD, gD, gD, gD, 3 * gD, adt, adt, b.attn_b, nullptr, PIO_ACT_NONE, st));
    PIO_TRY(small_attention(w.qkv, 3 * gD, (char*)w.qkv + (size_t)gD * e, 3 * gD, w.hb, gD, adt, R, n, h->H, gD / h->H, true, nullptr,
                            nullptr, 0, st));
    PIO_TRY(linear(mode, w.hb, b.proj_w, w.x, rows, gD, gD, gD, gD, gD, adt, PIO_DT_F32, b.proj_b, w.x, PIO_ACT_NONE, st));
    PIO_TRY(layernorm(w.x, gD, b.ln2_w, b.ln2_b, w.hb, adt, gD, rows, gD, 1e-5f, st));
    PIO_TRY(linear(mode, w.hb, b.fc_w, w.f, rows, gFF, gD, gD, gD, gFF, adt, adt, b.fc_b, nullptr, PIO_ACT_GELU_NEW, st));
    PIO_TRY(linear(mode, w.f, b.fc2_w, w.x, rows, gD, gFF, gFF, gFF, gD, adt, PIO_DT_F32, b.fc2_b, w.x, PIO_ACT_NONE, st));
  }
  PIO_TRY(layernorm(w.x, gD, h->lnf_w, h->lnf_b, w.hb, adt, gD, rows, gD, 1e-5f, st));
  // lm-head in slices of R rows (the logits buffer of the decode workspace holds R x 50264 floats)
  for (int row0 = 0; row0 < rows; row0 += R) {
    const int m = std::min(R, rows - row0);
    PIO_TRY(linear(mode, (const char*)w.hb + (size_t)row0 * gD * e, h->wte, w.logits, m, gV, gD, gD, gD, gVld, adt, PIO_DT_F32, nullptr,
                   nullptr, PIO_ACT_NONE, st));
    token_nll_kernel<<<m, 256, 0, st>>>(w.logits, gVld, gV, ids, lens, n, row0, nll);
    PIO_LAUNCHED();
  }
  nll_mean_kernel<<<cdiv(R, 128), 128, 0, st>>>(nll, lens, n, R, out_nll_mean);
  PIO_LAUNCHED();
  return PIO_OK;
}

}  // extern "C"

// ==========================================================================================
// ViECap mapping network (viecap/ClipCap.py:122-153): Linear(clip -> project_len x 768) tokens + prefix_len learnt tokens
// -> n_layer pre-LN transformer layers (q / kv projections without bias, ReLU MLP) -> the prefix_len last tokens.
struct PioMapper {
  int mode, act_dt, clip, plen, clen, L, H, hidden, n;  // n = plen + clen tokens per region
  std::vector<void*> owned;
  const void* lin_w; const float *lin_b, *prefix_const;
  struct Layer {
    const float *n1_w, *n1_b, *proj_b, *n2_w, *n2_b, *fc1_b, *fc2_b;
    const void *qkv_w /*[3*768,768]: q | k | v*/, *proj_w, *fc1_w, *fc2_w;
  };
  std::vector<Layer> layers;
};

namespace pio {
namespace {
int m_f32(PioMapper* h, const float** dst, const float* src, size_t n, cudaStream_t st) {
  void* p;
  PIO_CUDA(cudaMalloc(&p, n * 4));
  h->owned.push_back(p);
  PIO_CUDA(cudaMemcpyAsync(p, src, n * 4, cudaMemcpyDeviceToDevice, st));
  *dst = (const float*)p;
  return PIO_OK;
}
// [rows, cols] fp32 -> act dtype at dst (+ row offset), dst allocated by the caller
int m_conv(PioMapper* h, void* dst, size_t elem_off, const float* src, size_t n, cudaStream_t st) {
  if (h->act_dt == PIO_DT_F32) PIO_CUDA(cudaMemcpyAsync((float*)dst + elem_off, src, n * 4, cudaMemcpyDeviceToDevice, st));
  else PIO_TRY(f32_to_bf16(src, (__nv_bfloat16*)dst + elem_off, (long long)n, st));
  return PIO_OK;
}
int m_mat(PioMapper* h, const void** dst, const float* src, size_t n, cudaStream_t st) {
  void* p;
  PIO_CUDA(cudaMalloc(&p, n * (h->act_dt == PIO_DT_F32 ? 4 : 2)));
  h->owned.push_back(p);
  PIO_TRY(m_conv(h, p, 0, src, n, st));
  *dst = p;
  return PIO_OK;
}

// x[r, clen + i, :] = prefix_const[i, :]
__global__ void mapper_prefix_kernel(const float* __restrict__ pc, float* __restrict__ x, int R, int n, int clen, int plen, int D) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x, per = (long long)plen * D / 4;
  if (i >= (long long)R * per) return;
  const long long r = i / per, o = i % per;
  reinterpret_cast<float4*>(x + (r * n + clen) * D)[o] = __ldg(reinterpret_cast<const float4*>(pc) + o);
}
template <typename T>
__global__ void relu_kernel(T* __restrict__ x, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = (float)x[i] > 0.f ? x[i] : (T)0.f;
}
}  // namespace
}  // namespace pio

extern "C" {
void pio_mapper_destroy(PioMapper* h) {
  if (!h) return;
  for (void* p : h->owned) cudaFree(p);
  delete h;
}

int pio_mapper_create(PioMapper** out, const PioMapperWeights* w, int mode, void* stream) {
  using namespace pio;
  PIO_CHECK(out && w && w->layers, "mapper_create: null argument");
  PIO_CHECK(mode == PIO_FP32 || mode == PIO_BF16, "mapper_create: unknown mode %d", mode);
  PIO_CHECK(w->clip_size > 0 && w->clip_size % 64 == 0, "mapper_create: clip_size %d must be a multiple of 64", w->clip_size);
  PIO_CHECK(w->n_head > 0 && gD % w->n_head == 0 && gD / w->n_head <= 128, "mapper_create: %d heads do not divide 768", w->n_head);
  PIO_CHECK(w->project_len >= 1 && w->prefix_len >= 1 && w->project_len + w->prefix_len <= 64, "mapper_create: %d + %d tokens (max 64)",
            w->project_len, w->prefix_len);
  PIO_CHECK(w->hidden > 0 && w->hidden % 64 == 0, "mapper_create: mlp hidden size %d must be a multiple of 64", w->hidden);
  cudaStream_t st = as_stream(stream);
  PioMapper* h = new PioMapper();
  h->mode = mode; h->act_dt = mode == PIO_FP32 ? PIO_DT_F32 : PIO_DT_BF16;
  h->clip = w->clip_size; h->clen = w->project_len; h->plen = w->prefix_len; h->L = w->n_layer; h->H = w->n_head; h->hidden = w->hidden;
  h->n = h->clen + h->plen;
  h->layers.resize(h->L);
  const size_t e = h->act_dt == PIO_DT_F32 ? 4 : 2;
  auto go = [&]() -> int {
    PIO_TRY(m_mat(h, &h->lin_w, w->linear_w, (size_t)h->clen * gD * h->clip, st));
    PIO_TRY(m_f32(h, &h->lin_b, w->linear_b, (size_t)h->clen * gD, st));
    PIO_TRY(m_f32(h, &h->prefix_const, w->prefix_const, (size_t)h->plen * gD, st));
    for (int i = 0; i < h->L; ++i) {
      const PioMapperLayer& s = w->layers[i];
      PioMapper::Layer& d = h->layers[i];
      PIO_TRY(m_f32(h, &d.n1_w, s.norm1_w, gD, st)); PIO_TRY(m_f32(h, &d.n1_b, s.norm1_b, gD, st));
      PIO_TRY(m_f32(h, &d.n2_w, s.norm2_w, gD, st)); PIO_TRY(m_f32(h, &d.n2_b, s.norm2_b, gD, st));
      PIO_TRY(m_f32(h, &d.proj_b, s.proj_b, gD, st));
      PIO_TRY(m_f32(h, &d.fc1_b, s.fc1_b, h->hidden, st)); PIO_TRY(m_f32(h, &d.fc2_b, s.fc2_b, gD, st));
      void* qkv;
      PIO_CUDA(cudaMalloc(&qkv, (size_t)3 * gD * gD * e));
      h->owned.push_back(qkv);
      PIO_TRY(m_conv(h, qkv, 0, s.q_w, (size_t)gD * gD, st));                       // to_queries      [768,768]
      PIO_TRY(m_conv(h, qkv, (size_t)gD * gD, s.kv_w, (size_t)2 * gD * gD, st));     // to_keys_values [1536,768]: keys | values
      d.qkv_w = qkv;
      PIO_TRY(m_mat(h, &d.proj_w, s.proj_w, (size_t)gD * gD, st));
      PIO_TRY(m_mat(h, &d.fc1_w, s.fc1_w, (size_t)h->hidden * gD, st));
      PIO_TRY(m_mat(h, &d.fc2_w, s.fc2_w, (size_t)gD * h->hidden, st));
    }
    return PIO_OK;
  };
  int rc = go();
  if (rc != PIO_OK) { pio_mapper_destroy(h); return rc; }
  *out = h;
  return PIO_OK;
}

size_t pio_mapper_workspace_bytes(const PioMapper* h, int R) {
  using namespace pio;
  const size_t e = h->act_dt == PIO_DT_F32 ? 4 : 2, rows = (size_t)R * h->n;
  return align_up(rows * gD * 4, 1024) /*x*/ + align_up(rows * gD * e, 1024) /*h*/ + align_up(rows * 3 * gD * e, 1024) /*qkv*/ +
         align_up(rows * h->hidden * e, 1024) /*f*/ + align_up((size_t)R * h->clip * e, 1024) + 4096;
}

// feats fp32 [R, clip_size] (already L2-normalised by the caller, entrypoint.py:108) -> out fp32 [R, prefix_len, 768]
int pio_mapper_forward(PioMapper* h, const float* feats, int R, float* out, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace pio;
  if (R == 0) return PIO_OK;
  PIO_CHECK(h && feats && out && workspace, "mapper_forward: null argument");
  PIO_CHECK(workspace_bytes >= pio_mapper_workspace_bytes(h, R), "mapper_forward: workspace too small");
  PIO_CHECK((((uintptr_t)workspace) & 1023) == 0, "mapper_forward: workspace must be 1024-byte aligned");
  cudaStream_t st = as_stream(stream);
  const int adt = h->act_dt, mode = h->mode, n = h->n, rows = R * n, hd = gD / h->H;
  const size_t e = adt == PIO_DT_F32 ? 4 : 2;
  char* ws = (char*)workspace;
  float* x = (float*)ws; ws += align_up((size_t)rows * gD * 4, 1024);
  void* hb = ws;         ws += align_up((size_t)rows * gD * e, 1024);
  char* qkv = ws;        ws += align_up((size_t)rows * 3 * gD * e, 1024);
  void* f = ws;          ws += align_up((size_t)rows * h->hidden * e, 1024);
  void* fin = ws;

  const void* pA = feats;
  if (adt != PIO_DT_F32) { PIO_TRY(f32_to_bf16(feats, (__nv_bfloat16*)fin, (long long)R * h->clip, st)); pA = fin; }
  // tokens 0..clen-1 of every region: one GEMM whose output rows are n*768 floats apart
  PIO_TRY(linear(mode, pA, h->lin_w, x, R, h->clen * gD, h->clip, h->clip, h->clip, n * gD, adt, PIO_DT_F32, h->lin_b, nullptr,
                 PIO_ACT_NONE, st));
  {
    const long long tot = (long long)R * h->plen * gD / 4;
    mapper_prefix_kernel<<<cdiv(tot, 256), 256, 0, st>>>(h->prefix_const, x, R, n, h->clen, h->plen, gD);
    PIO_LAUNCHED();
  }
  for (int i = 0; i < h->L; ++i) {
    const PioMapper::Layer& w = h->layers[i];
    PIO_TRY(layernorm(x, gD, w.n1_w, w.n1_b, hb, adt, gD, rows, gD, 1e-5f, st));
    PIO_TRY(linear(mode, hb, w.qkv_w, qkv, rows, 3 * gD, gD, gD, gD, 3 * gD, adt, adt, nullptr, nullptr, PIO_ACT_NONE, st));
    PIO_TRY(small_attention(qkv, 3 * gD, qkv + (size_t)gD * e, 3 * gD, hb, gD, adt, R, n, h->H, hd, false, nullptr, nullptr, 0, st));
    PIO_TRY(linear(mode, hb, w.proj_w, x, rows, gD, gD, gD, gD, gD, adt, PIO_DT_F32, w.proj_b, x, PIO_ACT_NONE, st));
    PIO_TRY(layernorm(x, gD, w.n2_w, w.n2_b, hb, adt, gD, rows, gD, 1e-5f, st));
    PIO_TRY(linear(mode, hb, w.fc1_w, f, rows, h->hidden, gD, gD, gD, h->hidden, adt, adt, w.fc1_b, nullptr, PIO_ACT_NONE, st));
    {
      const long long tot = (long long)rows * h->hidden;
      if (adt == PIO_DT_F32) relu_kernel<float><<<cdiv(tot, 256), 256, 0, st>>>((float*)f, tot);
      else relu_kernel<__nv_bfloat16><<<cdiv(tot, 256), 256, 0, st>>>((__nv_bfloat16*)f, tot);
      PIO_LAUNCHED();
    }
    PIO_TRY(linear(mode, f, w.fc2_w, x, rows, gD, h->hidden, h->hidden, h->hidden, gD, adt, PIO_DT_F32, w.fc2_b, x, PIO_ACT_NONE, st));
  }
  // keep the last prefix_len tokens of every region (ClipCap.py:151)
  PIO_CUDA(cudaMemcpy2DAsync(out, (size_t)h->plen * gD * 4, x + (size_t)h->clen * gD, (size_t)n * gD * 4, (size_t)h->plen * gD * 4, R,
                             cudaMemcpyDeviceToDevice, st));
  return PIO_OK;
}
}  // extern "C"

extern "C" {
int pio_argmax_slabs(int M, int N) { return pio::argmax_slabs_tc(M, N); }

int pio_argmax_finish(const float* val, const int* idx, const float* sumexp, int ld, int slabs, int M, int* ids, int ids_ld,
                      int t, float* logprob_sum, void* stream) {
  using namespace pio;
  PIO_CHECK(val && idx && sumexp && ids, "argmax_finish: null argument");
  if (M == 0) return PIO_OK;
  argmax_finish_kernel<<<cdiv((long long)M * 32, 256), 256, 0, as_stream(stream)>>>(val, idx, sumexp, ld, slabs, M, ids, ids_ld, t,
                                                                                    logprob_sum);
  PIO_LAUNCHED();
  return PIO_OK;
}
}

namespace pio {
namespace {
// one CTA per query row: sims -> shared memory, then warp 0 does the softmax and k rounds of arg-max
__global__ void __launch_bounds__(256) entity_topk_kernel(const float* __restrict__ q, const float* __restrict__ ent, int E, int D,
                                                          float inv_temp, int k, float* __restrict__ out_prob, int* __restrict__ out_idx) {
  extern __shared__ float sm_e[];
  float* qs = sm_e;       // [D]
  float* sims = qs + D;   // [E]
  const int r = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int d = tid; d < D; d += 256) qs[d] = q[(long long)r * D + d];
  __syncthreads();
  for (int e = warp; e < E; e += 8) {
    const float* er = ent + (long long)e * D;
    float a = 0.f;
    for (int d = lane; d < D; d += 32) a = fmaf(qs[d], __ldg(er + d), a);
    a = warp_sum(a);
    if (lane == 0) sims[e] = a * inv_temp;
  }
  __syncthreads();
  if (warp != 0) return;
  float mx = -INFINITY;
  for (int e = lane; e < E; e += 32) mx = fmaxf(mx, sims[e]);
  mx = warp_max(mx);
  float s = 0.f;
  for (int e = lane; e < E; e += 32) {
    const float p = expf(sims[e] - mx);
    sims[e] = p;
    s += p;
  }
  s = warp_sum(s);
  __syncwarp();
  const float inv = 1.0f / s;
  for (int j = 0; j < k; ++j) {
    float bv = -1.f;
    int bi = 0x7fffffff;
    for (int e = lane; e < E; e += 32) {
      const float p = sims[e];
      if (p > bv) { bv = p; bi = e; }  // ascending e per lane: first index wins
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) {
      out_prob[(long long)r * k + j] = bi < E ? bv * inv : 0.f;
      out_idx[(long long)r * k + j] = bi < E ? bi : 0;
      if (bi < E) sims[bi] = -1.f;
    }
    __syncwarp();
  }
}
}  // namespace
}  // namespace pio

extern "C" int pio_entity_topk(const float* q, const float* ent, int R, int n_ent, int D, float temperature, int k, float* out_prob,
                               int* out_idx, void* stream) {
  using namespace pio;
  if (R == 0) return PIO_OK;
  PIO_CHECK(q && ent && out_prob && out_idx, "entity_topk: null argument");
  PIO_CHECK(n_ent >= 1 && n_ent <= 8192 && D >= 1 && D <= 2048, "entity_topk: %d entities x %d dims outside (<=8192, <=2048)", n_ent, D);
  PIO_CHECK(k >= 1 && k <= 32 && k <= n_ent, "entity_topk: k %d outside [1, min(32, n_ent)]", k);
  PIO_CHECK(temperature > 0.f, "entity_topk: temperature must be positive");
  const size_t smem = (size_t)(D + n_ent) * sizeof(float);
  entity_topk_kernel<<<R, 256, smem, as_stream(stream)>>>(q, ent, n_ent, D, 1.0f / temperature, k, out_prob, out_idx);
  PIO_LAUNCHED();
  return PIO_OK;
}

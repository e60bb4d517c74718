// vit.cu -- DINOv2 ViT-B/14 (4 registers) forward, host-side orchestration of the per-layer kernels.
// Replaces `self.dino(imgs, is_training=True)` (Patch-ioner/src/model.py:783) and the qkv forward hook
// (model.py:589-590 -> dino_extraction.py:8-9,24-34).  Arithmetic spec: SURVEY.md Appendix A.3.
//
// HBM layout (B images, N = 5 + g*g tokens, M = B*N rows):
//   x    fp32 [M, 768]   residual stream (always fp32)
//   h    act  [M, 768]   LayerNorm output / attention output
//   qkv  act  [M, 2304]  [q | k | v], head-major inside each third
//   f    act  [M, 3072]  fc1 output (also hosts the im2col matrix [B*P, 640] before block 0)
// act = fp32 in PIO_FP32 mode, bf16 in PIO_BF16 mode.
#include "common.cuh"
#include <vector>

namespace pio {
constexpr int kD = 768, kHeads = 12, kMlp = 3072, kDepth = 12, kNG = 5, kPatchK = 588, kPatchKp = 640;
}

struct PioVit {
  int mode;
  int act_dt;
  std::vector<void*> owned;
  // fp32 vectors
  const float *cls, *reg, *patch_b, *norm_w, *norm_b;
  struct Blk {
    const float *ln1_w, *ln1_b, *qkv_b, *proj_b, *ls1, *ln2_w, *ln2_b, *fc1_b, *fc2_b, *ls2;
    const void *qkv_w, *proj_w, *fc1_w, *fc2_w;  // act dtype, [out, in]
  } blk[12];
  const void* patch_w;  // act dtype [768, 640] zero padded
};

namespace pio {
namespace {

__global__ void pad_patch_weight_kernel(const float* __restrict__ w, void* out, int out_dt) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kD * kPatchKp) return;
  int k = i % kPatchKp, n = i / kPatchKp;
  float v = k < kPatchK ? w[n * kPatchK + k] : 0.f;
  if (out_dt == PIO_DT_F32) ((float*)out)[i] = v; else ((__nv_bfloat16*)out)[i] = __float2bfloat16(v);
}

// patch rows of x start as the positional embedding (the patch-embed GEMM then accumulates onto them)
__global__ void init_patch_pos_kernel(float* __restrict__ x, const float* __restrict__ pos, int B, int N, int P) {
  const long long total4 = (long long)B * P * (kD / 4);
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < total4; i += stride) {
    int d4 = (int)(i % (kD / 4));
    long long r = i / (kD / 4);
    int p = (int)(r % P), b = (int)(r / P);
    reinterpret_cast<float4*>(x + ((long long)b * N + kNG + p) * kD)[d4] =
        __ldg(reinterpret_cast<const float4*>(pos + (long long)(1 + p) * kD) + d4);
  }
}

__global__ void bf16_to_f32_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = __bfloat162float(in[i]);
}

template <typename T>
int dev_alloc(PioVit* h, T** p, size_t n) {
  void* q = nullptr;
  PIO_CUDA(cudaMalloc(&q, n * sizeof(T)));
  h->owned.push_back(q);
  *p = (T*)q;
  return PIO_OK;
}

int own_f32(PioVit* h, const float** dst, const float* src, size_t n, cudaStream_t st) {
  float* p;
  PIO_TRY(dev_alloc(h, &p, n));
  PIO_CUDA(cudaMemcpyAsync(p, src, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  *dst = p;
  return PIO_OK;
}
int own_mat(PioVit* h, const void** dst, const float* src, size_t n, cudaStream_t st) {
  if (h->act_dt == PIO_DT_F32) return own_f32(h, (const float**)dst, src, n, st);
  __nv_bfloat16* p;
  PIO_TRY(dev_alloc(h, &p, n));
  PIO_TRY(f32_to_bf16(src, p, (long long)n, st));
  *dst = p;
  return PIO_OK;
}

int gemm(int mode, const void* A, const void* W, void* C, int M, int N, int K, int lda, int ldw, int ldc, int a_dt, int c_dt,
         const float* bias, const float* gamma, const float* residual, int act, cudaStream_t st, int rpg = 0, int gs = 0,
         int go = 0) {
  PioLinear p;
  memset(&p, 0, sizeof(p));
  p.A = A; p.W = W; p.C = C; p.M = M; p.N = N; p.K = K; p.lda = lda; p.ldw = ldw; p.ldc = ldc;
  p.a_dt = a_dt; p.c_dt = c_dt; p.bias = bias; p.gamma = gamma; p.residual = residual; p.ldres = ldc;
  p.alpha = 1.0f; p.act = act; p.rows_per_group = rpg; p.group_stride = gs; p.group_offset = go;
  p.w_static = 1;  // model weights
  return mode == PIO_FP32 ? linear_simt(p, st) : linear_tc(p, st);
}

}  // namespace
}  // namespace pio

extern "C" {

int pio_vit_create(PioVit** out, const PioVitWeights* w, int mode, void* stream) {
  using namespace pio;
  PIO_CHECK(out && w, "vit_create: null argument");
  PIO_CHECK(mode == PIO_FP32 || mode == PIO_BF16, "vit_create: unknown mode %d", mode);
  cudaStream_t st = as_stream(stream);
  PioVit* h = new PioVit();
  h->mode = mode;
  h->act_dt = mode == PIO_FP32 ? PIO_DT_F32 : PIO_DT_BF16;
  int rc = PIO_OK;
  auto go = [&]() -> int {
    PIO_TRY(own_f32(h, &h->cls, w->cls_token, kD, st));
    PIO_TRY(own_f32(h, &h->reg, w->register_tokens, 4 * kD, st));
    PIO_TRY(own_f32(h, &h->patch_b, w->patch_b, kD, st));
    PIO_TRY(own_f32(h, &h->norm_w, w->norm_w, kD, st));
    PIO_TRY(own_f32(h, &h->norm_b, w->norm_b, kD, st));
    {
      void* pw = nullptr;
      const size_t n = (size_t)kD * kPatchKp;
      PIO_CUDA(cudaMalloc(&pw, n * (h->act_dt == PIO_DT_F32 ? 4 : 2)));
      h->owned.push_back(pw);
      pad_patch_weight_kernel<<<cdiv(n, 256), 256, 0, st>>>(w->patch_w, pw, h->act_dt);
      PIO_LAUNCHED();
      h->patch_w = pw;
    }
    for (int i = 0; i < kDepth; ++i) {
      const PioVitBlock& s = w->blk[i];
      PioVit::Blk& d = h->blk[i];
      PIO_TRY(own_f32(h, &d.ln1_w, s.ln1_w, kD, st));  PIO_TRY(own_f32(h, &d.ln1_b, s.ln1_b, kD, st));
      PIO_TRY(own_f32(h, &d.qkv_b, s.qkv_b, 3 * kD, st)); PIO_TRY(own_f32(h, &d.proj_b, s.proj_b, kD, st));
      PIO_TRY(own_f32(h, &d.ls1, s.ls1, kD, st));
      PIO_TRY(own_f32(h, &d.ln2_w, s.ln2_w, kD, st));  PIO_TRY(own_f32(h, &d.ln2_b, s.ln2_b, kD, st));
      PIO_TRY(own_f32(h, &d.fc1_b, s.fc1_b, kMlp, st)); PIO_TRY(own_f32(h, &d.fc2_b, s.fc2_b, kD, st));
      PIO_TRY(own_f32(h, &d.ls2, s.ls2, kD, st));
      PIO_TRY(own_mat(h, &d.qkv_w, s.qkv_w, (size_t)3 * kD * kD, st));
      PIO_TRY(own_mat(h, &d.proj_w, s.proj_w, (size_t)kD * kD, st));
      PIO_TRY(own_mat(h, &d.fc1_w, s.fc1_w, (size_t)kMlp * kD, st));
      PIO_TRY(own_mat(h, &d.fc2_w, s.fc2_w, (size_t)kD * kMlp, st));
    }
    return PIO_OK;
  };
  rc = go();
  if (rc != PIO_OK) { pio_vit_destroy(h); return rc; }
  *out = h;
  return PIO_OK;
}

void pio_vit_destroy(PioVit* h) {
  if (!h) return;
  for (void* p : h->owned) cudaFree(p);
  delete h;
}

size_t pio_vit_workspace_bytes(const PioVit* h, int B, int S) {
  using namespace pio;
  const size_t g = S / 14, N = kNG + g * g, M = (size_t)B * N, e = h->act_dt == PIO_DT_F32 ? 4 : 2;
  const size_t vt = h->act_dt == PIO_DT_BF16 ? align_up(vit_attention_tc_workspace(B, (int)N, kHeads), 1024) : 0;
  return align_up(M * kD * 4, 1024) + align_up(M * kD * e, 1024) + align_up(M * 3 * kD * e, 1024) +
         align_up(M * kMlp * e, 1024) + vt + 4096;
}

int pio_vit_forward(PioVit* h, const float* imgs, int B, int S, const float* pos_embed, float* out_tokens, float* out_attn,
                    float* out_qkv, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace pio;
  PdlScopeOff serialised;  // see elementwise.cu: the ViT forward is launched without programmatic overlap (reproducibility)
  PIO_CHECK(h && imgs && pos_embed && out_tokens && workspace, "vit_forward: null argument");
  PIO_CHECK(S % 14 == 0 && S >= 14, "vit_forward: image size %d is not a multiple of the patch size 14", S);
  PIO_CHECK(workspace_bytes >= pio_vit_workspace_bytes(h, B, S), "vit_forward: workspace too small");
  PIO_CHECK((((uintptr_t)workspace) & 1023) == 0, "vit_forward: workspace must be 1024-byte aligned");
  if (B == 0) return PIO_OK;
  cudaStream_t st = as_stream(stream);
  const int g = S / 14, P = g * g, N = kNG + P, M = B * N;
  const int adt = h->act_dt, mode = h->mode;
  const size_t e = adt == PIO_DT_F32 ? 4 : 2;
  char* ws = (char*)workspace;
  float* x = (float*)ws;  ws += align_up((size_t)M * kD * 4, 1024);
  void* hb = ws;          ws += align_up((size_t)M * kD * e, 1024);
  void* qkv = ws;         ws += align_up((size_t)M * 3 * kD * e, 1024);
  void* f = ws;            ws += align_up((size_t)M * kMlp * e, 1024);
  void* vt = ws;           // transposed V copy for the tensor-core attention (bf16 mode only)

  // tokens: [cls + pos0 | registers | pos[1+p]] then accumulate the patch embedding onto the patch rows
  PIO_TRY(init_global_tokens(x, h->cls, h->reg, pos_embed, B, N, kD, st));
  init_patch_pos_kernel<<<kNumSMs * 8, 256, 0, st>>>(x, pos_embed, B, N, P);
  PIO_LAUNCHED();
  PIO_TRY(im2col14(imgs, f, adt, B, S, g, kPatchKp, st));
  PIO_TRY(gemm(mode, f, h->patch_w, x, B * P, kD, kPatchKp, kPatchKp, kPatchKp, kD, adt, PIO_DT_F32, h->patch_b, nullptr, x,
               PIO_ACT_NONE, st, P, N, kNG));

  for (int i = 0; i < kDepth; ++i) {
    const PioVit::Blk& w = h->blk[i];
    PIO_TRY(layernorm(x, kD, w.ln1_w, w.ln1_b, hb, adt, kD, M, kD, 1e-6f, st));
    PIO_TRY(gemm(mode, hb, w.qkv_w, qkv, M, 3 * kD, kD, kD, kD, 3 * kD, adt, adt, w.qkv_b, nullptr, nullptr, PIO_ACT_NONE, st));
    if (i == kDepth - 1) {
      if (out_attn) PIO_TRY(cls_attention(qkv, adt, B, N, kD, kNG, out_attn, out_attn, st));
      if (out_qkv) {
        if (adt == PIO_DT_F32) {
          PIO_CUDA(cudaMemcpyAsync(out_qkv, qkv, (size_t)M * 3 * kD * 4, cudaMemcpyDeviceToDevice, st));
        } else {
          bf16_to_f32_kernel<<<kNumSMs * 8, 256, 0, st>>>((const __nv_bfloat16*)qkv, out_qkv, (long long)M * 3 * kD);
          PIO_LAUNCHED();
        }
      }
    }
    if (adt == PIO_DT_BF16) PIO_TRY(vit_attention_tc(qkv, hb, vt, B, N, kHeads, st));
    else PIO_TRY(vit_attention(qkv, hb, adt, B, N, kHeads, st));
    PIO_TRY(gemm(mode, hb, w.proj_w, x, M, kD, kD, kD, kD, kD, adt, PIO_DT_F32, w.proj_b, w.ls1, x, PIO_ACT_NONE, st));
    PIO_TRY(layernorm(x, kD, w.ln2_w, w.ln2_b, hb, adt, kD, M, kD, 1e-6f, st));
    PIO_TRY(gemm(mode, hb, w.fc1_w, f, M, kMlp, kD, kD, kD, kMlp, adt, adt, w.fc1_b, nullptr, nullptr, PIO_ACT_GELU_ERF, st));
    PIO_TRY(gemm(mode, f, w.fc2_w, x, M, kD, kMlp, kMlp, kMlp, kD, adt, PIO_DT_F32, w.fc2_b, w.ls2, x, PIO_ACT_NONE, st));
  }
  PIO_TRY(layernorm(x, kD, h->norm_w, h->norm_b, out_tokens, PIO_DT_F32, kD, M, kD, 1e-6f, st));
  return PIO_OK;
}

size_t pio_vit_block_workspace_bytes(const PioVit* h, int T) {
  using namespace pio;
  const size_t e = h->act_dt == PIO_DT_F32 ? 4 : 2;
  return align_up((size_t)T * kD * e, 1024) + align_up((size_t)T * 3 * kD * e, 1024) + align_up((size_t)T * kMlp * e, 1024) + 4096;
}

// One transformer block of the backbone on PACKED token sequences (double-DINO, bbox_utils.py:300-403 runs blocks[-1] on
// [cls | registers | the patches of a box] for every box): x fp32 [T,768] in/out, rows grouped in buckets of sequences of
// equal length (bucket b: bucket_nseq[b] sequences of bucket_len[b] rows, contiguous), so that attention runs as a regular
// batch per bucket while LayerNorm and the dense layers run once over all T rows.
int pio_vit_block_rows(PioVit* h, int layer, float* x, int T, const int* bucket_nseq, const int* bucket_len, int nbuckets,
                       void* workspace, size_t workspace_bytes, void* stream) {
  using namespace pio;
  if (T == 0) return PIO_OK;
  PdlScopeOff serialised;
  PIO_CHECK(h && x && bucket_nseq && bucket_len && workspace, "vit_block_rows: null argument");
  PIO_CHECK(layer >= -kDepth && layer < kDepth, "vit_block_rows: layer %d outside [-12, 12)", layer);
  PIO_CHECK(workspace_bytes >= pio_vit_block_workspace_bytes(h, T), "vit_block_rows: workspace too small");
  PIO_CHECK((((uintptr_t)workspace) & 1023) == 0, "vit_block_rows: workspace must be 1024-byte aligned");
  long long rows = 0;
  for (int b = 0; b < nbuckets; ++b) {
    PIO_CHECK(bucket_nseq[b] >= 0 && bucket_len[b] > 0, "vit_block_rows: bad bucket %d", b);
    rows += (long long)bucket_nseq[b] * bucket_len[b];
  }
  PIO_CHECK(rows == T, "vit_block_rows: buckets cover %lld rows, T = %d", rows, T);
  cudaStream_t st = as_stream(stream);
  const PioVit::Blk& w = h->blk[layer < 0 ? layer + kDepth : layer];
  const int adt = h->act_dt, mode = h->mode;
  const size_t e = adt == PIO_DT_F32 ? 4 : 2;
  char* ws = (char*)workspace;
  char* hb = ws;   ws += align_up((size_t)T * kD * e, 1024);
  char* qkv = ws;  ws += align_up((size_t)T * 3 * kD * e, 1024);
  void* f = ws;
  PIO_TRY(layernorm(x, kD, w.ln1_w, w.ln1_b, hb, adt, kD, T, kD, 1e-6f, st));
  PIO_TRY(gemm(mode, hb, w.qkv_w, qkv, T, 3 * kD, kD, kD, kD, 3 * kD, adt, adt, w.qkv_b, nullptr, nullptr, PIO_ACT_NONE, st));
  long long r0 = 0;
  for (int b = 0; b < nbuckets; ++b) {
    if (bucket_nseq[b] == 0) continue;
    const void* q = qkv + (size_t)r0 * 3 * kD * e;
    void* o = hb + (size_t)r0 * kD * e;
    if (adt == PIO_DT_BF16) PIO_TRY(vit_attention_tc(q, o, nullptr, bucket_nseq[b], bucket_len[b], kHeads, st));
    else PIO_TRY(vit_attention(q, o, adt, bucket_nseq[b], bucket_len[b], kHeads, st));
    r0 += (long long)bucket_nseq[b] * bucket_len[b];
  }
  PIO_TRY(gemm(mode, hb, w.proj_w, x, T, kD, kD, kD, kD, kD, adt, PIO_DT_F32, w.proj_b, w.ls1, x, PIO_ACT_NONE, st));
  PIO_TRY(layernorm(x, kD, w.ln2_w, w.ln2_b, hb, adt, kD, T, kD, 1e-6f, st));
  PIO_TRY(gemm(mode, hb, w.fc1_w, f, T, kMlp, kD, kD, kD, kMlp, adt, adt, w.fc1_b, nullptr, nullptr, PIO_ACT_GELU_ERF, st));
  PIO_TRY(gemm(mode, f, w.fc2_w, x, T, kD, kMlp, kMlp, kMlp, kD, adt, PIO_DT_F32, w.fc2_b, w.ls2, x, PIO_ACT_NONE, st));
  return PIO_OK;
}

size_t pio_attention_workspace_bytes(int dt, int B, int N, int H) {
  return dt == PIO_DT_BF16 ? pio::align_up(pio::vit_attention_tc_workspace(B, N, H), 1024) : 0;
}

int pio_vit_attention(const void* qkv, void* out, int dt, int B, int N, int H, void* workspace, size_t workspace_bytes,
                      void* stream) {
  using namespace pio;
  PIO_CHECK(qkv && out, "vit_attention: null argument");
  PIO_CHECK(dt == PIO_DT_F32 || dt == PIO_DT_BF16, "vit_attention: dtype must be fp32 or bf16");
  if (B == 0 || N == 0) return PIO_OK;
  if (dt == PIO_DT_BF16) {
    PIO_CHECK(workspace && workspace_bytes >= pio_attention_workspace_bytes(dt, B, N, H), "vit_attention: workspace too small");
    return vit_attention_tc(qkv, out, workspace, B, N, H, as_stream(stream));
  }
  return vit_attention(qkv, out, dt, B, N, H, as_stream(stream));
}
}

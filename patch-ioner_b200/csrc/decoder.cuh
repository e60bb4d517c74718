// decoder.cuh -- the decoder handle and its workspace layout, shared by text.cu (kernel-per-op decode) and
// decode_fused_sm100.cu (one persistent kernel for the whole decode at small batch).
#pragma once
#include "common.cuh"
#include <vector>

namespace pio {
constexpr int gD = 768, gV = 50257, gVld = 50264, gFF = 3072;

// per-layer pointers the fused kernel reads from global memory (one array entry per block)
struct FusedLayer {
  const float *ln1_w, *ln1_b, *attn_b, *proj_b, *ln2_w, *ln2_b, *fc_b, *fc2_b;
};
}  // namespace pio

// L blocks x H heads; T = positions the KV cache of one call may hold (DeCap: 4 x 4 x 32; GPT-2 small for ViECap: 12 x 12 x 128)
struct PioDecoder {
  int mode, act_dt, prefix_size, L, H, T;
  std::vector<void*> owned;
  const float *wte32, *wpe, *lnf_w, *lnf_b, *prefix_b0;  // prefix_b0 = prefix bias + wpe[0]
  const void *wte, *prefix_w;                             // act dtype
  struct Blk {
    const float *ln1_w, *ln1_b, *attn_b, *proj_b, *ln2_w, *ln2_b, *fc_b, *fc2_b;
    const void *attn_w, *proj_w, *fc_w, *fc2_w;  // act dtype, transposed to [out, in]
  };
  std::vector<Blk> blk;
  // fused decode (bf16 mode): device arrays built at create time -- tensor maps of every weight matrix (4 per block: attn,
  // proj, fc, fc2; then wte), 128-row x 64-column boxes, and the per-layer vector pointers
  void* fused_wmaps = nullptr;              // CUtensorMap[4 L + 1]
  pio::FusedLayer* fused_layers = nullptr;  // [L]
};

namespace pio {
struct DecodeWs {
  float* x; void* hb; void* qkv; void* f; float* logits; char* kc; char* vc; void* pfx;
  // fused-decode extras: attention output rows, split-K partials of fc2, per-CTA arg-max partials, phase / tile counters
  void* att; float* part; float* pm_val; int* pm_idx; int* counters; size_t counters_bytes;
  size_t kv_layer, total;
};
constexpr int kFusedMaxRows = 64;      // rows one fused launch handles (their 12 activation tiles stay in shared memory)
constexpr int kFusedCtas = kNumSMs;    // upper bound of its grid
constexpr int kFusedFc2Splits = 4;     // K = 3072 cut into 4 x 12 k-blocks: every GEMM unit is 128 rows x 12 k-blocks
size_t fused_counter_ints(int L, int steps);

// carve the decode workspace for R rows and a KV cache of T positions (same layout for sizing and for use)
// `rows` >= R: rows of the token buffers (R for single-position steps, R * prompt_len for a batched prompt prefill)
inline DecodeWs decode_ws(const PioDecoder* h, char* base, int R, int T, size_t tail_elems, size_t rows = 0, int fused_steps = 0) {
  const size_t e = h->act_dt == PIO_DT_F32 ? 4 : 2;
  DecodeWs w;
  char* ws = base;
  if (rows < (size_t)R) rows = R;
  w.kv_layer = (size_t)R * T * gD * e;
  w.x = (float*)ws;      ws += align_up(rows * gD * 4, 1024);
  w.hb = ws;             ws += align_up(rows * gD * e, 1024);
  w.qkv = ws;            ws += align_up(rows * 3 * gD * e, 1024);
  w.f = ws;              ws += align_up(rows * gFF * e, 1024);
  w.logits = (float*)ws; ws += align_up((size_t)R * gVld * 4, 1024);
  w.kc = ws;             ws += align_up(h->L * w.kv_layer, 1024);
  w.vc = ws;             ws += align_up(h->L * w.kv_layer, 1024);
  w.pfx = ws;            ws += align_up(tail_elems * e, 1024);
  const size_t rp = (size_t)((R + 15) / 16 * 16);
  w.att = ws;               ws += align_up(rp * gD * 2, 1024);
  w.part = (float*)ws;      ws += align_up((size_t)6 * kFusedFc2Splits * rp * 128 * 4, 1024);
  w.pm_val = (float*)ws;    ws += align_up((size_t)kFusedCtas * rp * 4, 1024);
  w.pm_idx = (int*)ws;      ws += align_up((size_t)kFusedCtas * rp * 4, 1024);
  w.counters_bytes = align_up(fused_counter_ints(h->L, fused_steps > 0 ? fused_steps : h->T) * sizeof(int), 1024);
  w.counters = (int*)ws;    ws += w.counters_bytes;
  w.total = (size_t)(ws - base) + 4096;
  return w;
}

// decode_fused_sm100.cu
bool decode_fused_eligible(const PioDecoder* h, int R, bool want_logprob);
int decode_fused_build(PioDecoder* h, cudaStream_t st);  // tensor maps + layer table (decoder create, bf16 mode)
// runs `steps` picks: first_phase = 0 starts with the blocks of position pos_base (DeCap: x holds the prefix embedding);
// first_phase = -1 starts at ln_f (ViECap: x holds the residual stream of the last prompt position, blocks resume at pos_base + 1)
int decode_fused(PioDecoder* h, const DecodeWs& w, int R, int T, int steps, int pos_base, bool start_at_pick, int* out_ids,
                 cudaStream_t st);
}  // namespace pio

// detok.cu -- batched detokenisation of greedy-decode id rows (HOST code; no kernel: byte work of ~1 MB per 4096 captions).
//   pio_detok_rows   replaces the per-row Python loop of decoding_batched's tail (src/decap/decap.py:162-181) and
//                    SimpleTokenizer.decode (src/clip/simple_tokenizer.py:129-131): concatenate the byte strings of a row's
//                    tokens, stop at the end-of-text id, apply the '</w>' -> ' ' substitution on the row's bytes.
// The id -> bytes table is built once on the Python side (patch-ioner_b200/detok.py) from the CLIP BPE merges file, or as
// "<id> " strings when that third-party asset is absent.  UTF-8 decoding (errors='replace') of each row stays in Python: the
// byte substitution commutes with it because every substituted byte is ASCII.
// Two passes over the rows, both split over a few host threads: (1) row status and raw byte count, (2) after a prefix sum of the
// counts, gather + substitution into the row's slot; a last serial pass compacts the rows the substitution shortened.
#include "common.cuh"

#include <algorithm>
#include <thread>
#include <vector>

namespace {
struct DetokArgs {
  const int* ids; int R, T, ld; const unsigned char* table; const long long* offsets; int vocab, eot_id, strip, eow;
  unsigned char* out; long long* row_offsets; int* row_status;
};

// raw byte count of row r before the '</w>' substitution; status as documented in pio.h
inline long long row_measure(const DetokArgs& a, int r, int* status_out) {
  const int* row = a.ids + (long long)r * a.ld;
  long long n = 0;
  int status = 0;
  for (int t = 0; t < a.T; ++t) {
    const int id = row[t];
    if (id == a.eot_id) { status = 2; break; }
    if (id < 0 || id >= a.vocab) { status = 1; break; }
    n += a.offsets[id + 1] - a.offsets[id];
  }
  if (status == 0 && a.strip && n > 0) --n;  // " ".join(...) has no trailing separator
  if (status == 1) n = 0;                     // the caller renders this row itself
  *status_out = status;
  return n;
}

// writes row r at out + start (n raw bytes), returns the byte count after the substitution; *ascii &= all bytes < 128
inline long long row_write(const DetokArgs& a, int r, long long start, long long n, bool* ascii) {
  const int* row = a.ids + (long long)r * a.ld;
  unsigned char* dst = a.out + start;
  long long w = 0;
  for (int t = 0; t < a.T && w < n; ++t) {
    const long long o = a.offsets[row[t]], len = std::min(a.offsets[row[t] + 1] - o, n - w);
    memcpy(dst + w, a.table + o, (size_t)len);
    w += len;
  }
  unsigned char hi = 0;
  for (long long i = 0; i < n; ++i) hi |= dst[i];
  if (hi & 0x80) *ascii = false;
  if (!a.eow) return n;
  // '</w>' -> ' ' in place (the match may straddle token boundaries, as in the reference's str.replace)
  long long i = 0, o = 0;
  while (i < n) {
    if (i + 4 <= n && dst[i] == '<' && dst[i + 1] == '/' && dst[i + 2] == 'w' && dst[i + 3] == '>') { dst[o++] = ' '; i += 4; }
    else dst[o++] = dst[i++];
  }
  return o;
}

template <typename F>
void parallel_rows(int R, F&& f) {
  const unsigned hw = std::thread::hardware_concurrency();
  const int nt = std::max(1, std::min<int>({8, (int)(hw ? hw : 1), R / 1024}));  // ~0.45 us per row: threads pay from a few thousand rows
  if (nt == 1) { f(0, R, 0); return; }
  std::vector<std::thread> th;
  const int per = (R + nt - 1) / nt;
  for (int i = 0; i < nt; ++i) th.emplace_back([&, i] { f(std::min(R, i * per), std::min(R, (i + 1) * per), i); });
  for (auto& t : th) t.join();
}
}  // namespace

extern "C" int pio_detok_rows(const int* ids, int R, int T, int ld, const unsigned char* table, const long long* offsets, int vocab,
                              int eot_id, int strip_trailing_sep, int replace_eow, unsigned char* out, long long out_cap,
                              long long* row_offsets, int* row_status, int* all_ascii) {
  using namespace pio;
  PIO_CHECK(R >= 0 && T >= 0 && ld >= T, "detok_rows: bad shape");
  if (all_ascii) *all_ascii = 1;
  if (R == 0) { if (row_offsets) row_offsets[0] = 0; return PIO_OK; }
  PIO_CHECK(ids && table && offsets && out && row_offsets && row_status, "detok_rows: null argument");
  const DetokArgs a{ids, R, T, ld, table, offsets, vocab, eot_id, strip_trailing_sep, replace_eow, out, row_offsets, row_status};
  std::vector<long long> raw(R);
  parallel_rows(R, [&](int r0, int r1, int) { for (int r = r0; r < r1; ++r) raw[r] = row_measure(a, r, &row_status[r]); });
  long long tot = 0;
  for (int r = 0; r < R; ++r) { row_offsets[r] = tot; tot += raw[r]; }
  row_offsets[R] = tot;
  if (tot > out_cap) return fail(PIO_EINVAL, "detok_rows: output buffer too small (%lld > %lld bytes)", tot, out_cap);
  bool ascii_t[8] = {true, true, true, true, true, true, true, true};
  parallel_rows(R, [&](int r0, int r1, int ti) {
    for (int r = r0; r < r1; ++r) raw[r] = row_write(a, r, row_offsets[r], raw[r], &ascii_t[ti]);
  });
  if (replace_eow) {  // compact: rows moved down over the bytes the substitution freed
    long long w = 0;
    for (int r = 0; r < R; ++r) {
      if (w != row_offsets[r]) memmove(out + w, out + row_offsets[r], (size_t)raw[r]);
      row_offsets[r] = w;
      w += raw[r];
    }
    row_offsets[R] = w;
  }
  if (all_ascii)
    for (bool b : ascii_t) if (!b) *all_ascii = 0;
  return PIO_OK;
}

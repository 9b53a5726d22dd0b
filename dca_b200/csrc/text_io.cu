// Host-side output writer: the TSV files the reference produces with
// pandas.DataFrame.to_csv(sep='\t', float_format='%.6f') in dca/io.py:120-129 (mean.tsv, latent.tsv,
// dispersion.tsv, dropout.tsv), byte for byte, formatted by several threads with a fixed-point digit writer
// instead of one printf per value.  No device code in this file.
#include "dca_internal.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace dca {
namespace {

// '%.6f' of a finite value: |v| * 1e6 rounded to an integer in double precision; whenever the product is within
// its own rounding error of a half-integer (exact ties such as k/128, or a double too close to call) and for
// large magnitudes the value goes through snprintf, so the text always equals printf's.
// NaN -> "" (pandas na_rep), +-inf -> "inf" / "-inf".
inline void append_fixed6(std::string& out, double v) {
  if (v != v) return;
  if (std::isinf(v)) { out += v < 0 ? "-inf" : "inf"; return; }
  const double a = std::fabs(v);
  const double p = a * 1e6;
  if (a >= 8.0e9 || std::fabs((p - std::floor(p)) - 0.5) <= p * 4.5e-16) {
    char buf[400]; const int n = snprintf(buf, sizeof(buf), "%.6f", v); out.append(buf, (size_t)n); return;
  }
  const unsigned long long r = (unsigned long long)std::nearbyint(p);
  unsigned long long ip = r / 1000000ull; unsigned frac = (unsigned)(r % 1000000ull);
  char buf[40]; int pos = 40;
  for (int i = 0; i < 6; ++i) { buf[--pos] = (char)('0' + frac % 10); frac /= 10; }
  buf[--pos] = '.';
  do { buf[--pos] = (char)('0' + ip % 10); ip /= 10; } while (ip);
  if (std::signbit(v)) buf[--pos] = '-';          // printf keeps the sign of -0.0 and of values that round to zero
  out.append(buf + pos, (size_t)(40 - pos));
}

// csv.QUOTE_MINIMAL with '"' as quote character and '\t' as separator
inline void append_label(std::string& out, const char* s) {
  if (!s) return;
  if (strpbrk(s, "\t\"\n\r")) {
    out += '"';
    for (const char* c = s; *c; ++c) { if (*c == '"') out += '"'; out += *c; }
    out += '"';
  } else {
    out += s;
  }
}

template <typename T>
int write_matrix(FILE* f, const T* m, int64_t rows, int64_t cols, int64_t ld, const char* const* row_names,
                 const char* const* col_names, int transpose, int threads) {
  const int64_t out_rows = transpose ? cols : rows, out_cols = transpose ? rows : cols;
  const char* const* rn = transpose ? col_names : row_names;     // labels of the OUTPUT rows / columns
  const char* const* cn = transpose ? row_names : col_names;
  if (cn) {
    std::string h;
    if (rn) h += '\t';                                             // unnamed index: empty first header cell
    for (int64_t j = 0; j < out_cols; ++j) { if (j) h += '\t'; append_label(h, cn[j]); }
    h += '\n';
    if (fwrite(h.data(), 1, h.size(), f) != h.size()) return -1;
  }
  if (threads < 1) threads = 1;
  // rows per thread and round: ~4 MB of text each
  int64_t chunk = (int64_t)(4000000 / (std::max<int64_t>(out_cols, 1) * 9 + 16)); if (chunk < 1) chunk = 1; if (chunk > 4096) chunk = 4096;
  std::vector<std::string> bufs((size_t)threads);
  std::vector<std::vector<T>> tmp((size_t)threads);
  for (int64_t r0 = 0; r0 < out_rows; r0 += chunk * threads) {
    auto work = [&](int t) {
      std::string& out = bufs[(size_t)t]; out.clear();
      const int64_t a = r0 + (int64_t)t * chunk, b = std::min(out_rows, a + chunk);
      if (a >= b) return;
      const T* src = m; int64_t sld = ld;
      if (transpose) {          // gather output rows [a, b) = input columns [a, b) into a small row-major block
        std::vector<T>& tb = tmp[(size_t)t]; tb.resize((size_t)((b - a) * out_cols));
        for (int64_t j = 0; j < out_cols; ++j) { const T* in = m + j * ld + a; for (int64_t i = 0; i < b - a; ++i) tb[(size_t)(i * out_cols + j)] = in[i]; }
        src = tb.data() - a * out_cols; sld = out_cols;
      }
      for (int64_t i = a; i < b; ++i) {
        if (rn) { append_label(out, rn[i]); if (out_cols) out += '\t'; }
        const T* row = src + i * sld;
        for (int64_t j = 0; j < out_cols; ++j) { if (j) out += '\t'; append_fixed6(out, (double)row[j]); }
        out += '\n';
      }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < threads; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
    for (int t = 0; t < threads; ++t)
      if (!bufs[(size_t)t].empty() && fwrite(bufs[(size_t)t].data(), 1, bufs[(size_t)t].size(), f) != bufs[(size_t)t].size()) return -1;
  }
  return 0;
}

}  // namespace
}  // namespace dca

using namespace dca;

extern "C" int dca_write_text_matrix(const char* path, const void* matrix, int32_t is_float64, int64_t rows, int64_t cols,
                                     int64_t ld, const char* const* row_names, const char* const* col_names,
                                     int32_t transpose, int32_t threads) {
  if (!path || !matrix || rows < 0 || cols < 0 || ld < cols) { set_error("dca_write_text_matrix: bad argument"); return DCA_ERR_BAD_ARG; }
  FILE* f = fopen(path, "wb");
  if (!f) { set_error("dca_write_text_matrix: cannot open %s", path); return DCA_ERR_BAD_ARG; }
  if (threads <= 0) { threads = (int)std::thread::hardware_concurrency(); if (threads > 16) threads = 16; if (threads < 1) threads = 1; }
  const int st = is_float64 ? write_matrix<double>(f, (const double*)matrix, rows, cols, ld, row_names, col_names, transpose, threads)
                            : write_matrix<float>(f, (const float*)matrix, rows, cols, ld, row_names, col_names, transpose, threads);
  const int cl = fclose(f);
  if (st != 0 || cl != 0) { set_error("dca_write_text_matrix: write to %s failed", path); return DCA_ERR_CUDA; }
  return DCA_OK;
}

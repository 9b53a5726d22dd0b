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

// ------------------------------------------------------------------------------------------------------------------
// Host-side packer of a raw count matrix into the streaming format of dca_stream_begin_packed (bits per entry +
// row-sorted CSR overflow list): the multi-threaded counterpart of dca_b200/io.py:pack_counts.  No device code.
namespace dca {
namespace {

template <typename T>
inline bool count_value(T v, double& out) { out = (double)v; return out >= 0.0 && out == std::floor(out); }

template <typename F>
void parallel_rows(int64_t rows, int threads, F&& fn) {
  if (threads < 1) threads = 1;
  if (threads > rows) threads = (int)std::max<int64_t>(rows, 1);
  std::vector<std::thread> th;
  const int64_t per = (rows + threads - 1) / threads;
  for (int t = 1; t < threads; ++t) th.emplace_back([&, t] { fn(std::min(rows, t * per), std::min(rows, (t + 1) * per)); });
  fn(0, std::min(rows, per));
  for (auto& x : th) x.join();
}

// pass 1: per row, the number of entries >= 15, >= 255, >= 65535 (the escape values of the three widths); returns
// false when an entry is negative or not an integer
template <typename T>
bool escape_counts(const T* m, int64_t rows, int64_t cols, int64_t ld, int64_t* per_row /* [3][rows] */, int threads) {
  std::vector<int> bad((size_t)std::max(threads, 1) + 1, 0);
  parallel_rows(rows, threads, [&](int64_t a, int64_t b) {
    for (int64_t r = a; r < b; ++r) {
      const T* row = m + r * ld; int64_t c4 = 0, c8 = 0, c16 = 0; bool ok = true;
      for (int64_t j = 0; j < cols; ++j) {
        double v; ok &= count_value(row[j], v);
        c4 += v >= 15.0; c8 += v >= 255.0; c16 += v >= 65535.0;
      }
      per_row[r] = c4; per_row[rows + r] = c8; per_row[2 * rows + r] = c16;
      if (!ok) bad[0] = 1;
    }
  });
  return bad[0] == 0;
}

template <typename T>
void pack_rows(const T* m, int64_t rows, int64_t cols, int64_t ld, int bits, unsigned char* packed, const int64_t* indptr,
               unsigned char* entries, int threads) {
  const double esc = (double)((1u << bits) - 1u);
  const int64_t row_bytes = cols * bits / 8;
  parallel_rows(rows, threads, [&](int64_t a, int64_t b) {
    for (int64_t r = a; r < b; ++r) {
      const T* row = m + r * ld; unsigned char* out = packed + r * row_bytes;
      int64_t k = indptr[r];
      auto emit = [&](int64_t j, double v) {
        int32_t g = (int32_t)j; float c = (float)v;
        memcpy(entries + 8 * k, &g, 4); memcpy(entries + 8 * k + 4, &c, 4); ++k;
      };
      if (bits == 4) {
        for (int64_t j = 0; j < cols; j += 2) {
          double v0 = (double)row[j], v1 = (double)row[j + 1];
          if (v0 >= esc) { emit(j, v0); v0 = esc; }
          if (v1 >= esc) { emit(j + 1, v1); v1 = esc; }
          out[j >> 1] = (unsigned char)((unsigned)v0 | ((unsigned)v1 << 4));
        }
      } else if (bits == 8) {
        for (int64_t j = 0; j < cols; ++j) { double v = (double)row[j]; if (v >= esc) { emit(j, v); v = esc; } out[j] = (unsigned char)v; }
      } else {
        uint16_t* o16 = reinterpret_cast<uint16_t*>(out);
        for (int64_t j = 0; j < cols; ++j) { double v = (double)row[j]; if (v >= esc) { emit(j, v); v = esc; } o16[j] = (uint16_t)v; }
      }
    }
  });
}

// sparse format, pass 1: non-zero entries and entries >= 15 per row
template <typename T>
bool sparse_counts(const T* m, int64_t rows, int64_t cols, int64_t ld, int64_t* nnz, int64_t* esc, int threads) {
  int bad = 0;
  parallel_rows(rows, threads, [&](int64_t a, int64_t b) {
    for (int64_t r = a; r < b; ++r) {
      const T* row = m + r * ld; int64_t c = 0, e = 0; bool ok = true;
      for (int64_t j = 0; j < cols; ++j) { double v; ok &= count_value(row[j], v); c += v != 0.0; e += v >= 15.0; }
      nnz[r] = c; esc[r] = e;
      if (!ok) bad = 1;
    }
  });
  return bad == 0;
}

// sparse format, pass 2: bitmap (cols/8 bytes per row), 4-bit codes (row r from byte nib_indptr[r]), overflow entries
template <typename T>
void sparse_rows(const T* m, int64_t rows, int64_t cols, int64_t ld, unsigned char* bitmap, const int64_t* nib_indptr,
                 unsigned char* nibbles, const int64_t* ovf_indptr, unsigned char* entries, int threads) {
  const int64_t bm_bytes = cols / 8;
  parallel_rows(rows, threads, [&](int64_t a, int64_t b) {
    for (int64_t r = a; r < b; ++r) {
      const T* row = m + r * ld; unsigned char* bm = bitmap + r * bm_bytes; unsigned char* nb = nibbles + nib_indptr[r];
      int64_t k = 0, e = ovf_indptr[r];
      for (int64_t j0 = 0; j0 < cols; j0 += 8) {
        unsigned bits = 0;
        for (int t = 0; t < 8; ++t) {
          const double v = (double)row[j0 + t];
          if (v == 0.0) continue;
          bits |= 1u << t;
          unsigned code = v >= 15.0 ? 15u : (unsigned)v;
          if (v >= 15.0) { int32_t g = (int32_t)(j0 + t); float c = (float)v; memcpy(entries + 8 * e, &g, 4); memcpy(entries + 8 * e + 4, &c, 4); ++e; }
          if (k & 1) nb[k >> 1] = (unsigned char)(nb[k >> 1] | (code << 4)); else nb[k >> 1] = (unsigned char)code;
          ++k;
        }
        bm[j0 >> 3] = (unsigned char)bits;
      }
    }
  });
}

}  // namespace
}  // namespace dca

// Sparse host format (dca_stream_begin_sparse), multi-threaded.  dca_sparse_counts: nnz[r] = non-zero entries of row r,
// esc[r] = entries >= 15; the caller turns them into nib_indptr (bytes: cumsum((nnz + 1) / 2)) and ovf_indptr (cumsum(esc))
// and calls dca_pack_sparse, which fills bitmap [rows x cols/8], nibbles and the overflow entries.
extern "C" int dca_sparse_counts(const void* counts, int32_t dtype, int64_t rows, int64_t cols, int64_t ld, int64_t* nnz,
                                 int64_t* esc, int32_t threads) {
  if (!counts || !nnz || !esc || rows < 0 || cols < 0 || ld < cols) { set_error("dca_sparse_counts: bad argument"); return DCA_ERR_BAD_ARG; }
  if (threads <= 0) { threads = (int)std::thread::hardware_concurrency(); if (threads > 32) threads = 32; if (threads < 1) threads = 1; }
  bool ok = false;
  switch (dtype) {
    case 0: ok = sparse_counts((const float*)counts, rows, cols, ld, nnz, esc, threads); break;
    case 1: ok = sparse_counts((const double*)counts, rows, cols, ld, nnz, esc, threads); break;
    case 2: ok = sparse_counts((const uint16_t*)counts, rows, cols, ld, nnz, esc, threads); break;
    case 3: ok = sparse_counts((const int32_t*)counts, rows, cols, ld, nnz, esc, threads); break;
    case 4: ok = sparse_counts((const int64_t*)counts, rows, cols, ld, nnz, esc, threads); break;
    default: set_error("dca_sparse_counts: unknown dtype %d", dtype); return DCA_ERR_BAD_ARG;
  }
  if (!ok) { set_error("dca_sparse_counts: counts must be non-negative integers"); return DCA_ERR_BAD_ARG; }
  return DCA_OK;
}

extern "C" int dca_pack_sparse(const void* counts, int32_t dtype, int64_t rows, int64_t cols, int64_t ld, void* bitmap,
                               const int64_t* nib_indptr, void* nibbles, const int64_t* ovf_indptr, void* entries, int32_t threads) {
  if (!counts || !bitmap || !nib_indptr || !nibbles || !ovf_indptr || rows < 0 || cols < 0 || ld < cols || (ovf_indptr[rows] > 0 && !entries)) {
    set_error("dca_pack_sparse: bad argument"); return DCA_ERR_BAD_ARG;
  }
  if (cols % 8 != 0) { set_error("dca_pack_sparse: the number of genes must be a multiple of 8"); return DCA_ERR_BAD_ARG; }
  if (threads <= 0) { threads = (int)std::thread::hardware_concurrency(); if (threads > 32) threads = 32; if (threads < 1) threads = 1; }
  unsigned char* bm = (unsigned char*)bitmap; unsigned char* nb = (unsigned char*)nibbles; unsigned char* e = (unsigned char*)entries;
  switch (dtype) {
    case 0: sparse_rows((const float*)counts, rows, cols, ld, bm, nib_indptr, nb, ovf_indptr, e, threads); break;
    case 1: sparse_rows((const double*)counts, rows, cols, ld, bm, nib_indptr, nb, ovf_indptr, e, threads); break;
    case 2: sparse_rows((const uint16_t*)counts, rows, cols, ld, bm, nib_indptr, nb, ovf_indptr, e, threads); break;
    case 3: sparse_rows((const int32_t*)counts, rows, cols, ld, bm, nib_indptr, nb, ovf_indptr, e, threads); break;
    case 4: sparse_rows((const int64_t*)counts, rows, cols, ld, bm, nib_indptr, nb, ovf_indptr, e, threads); break;
    default: set_error("dca_pack_sparse: unknown dtype %d", dtype); return DCA_ERR_BAD_ARG;
  }
  return DCA_OK;
}

// dtype: 0 float32, 1 float64, 2 uint16, 3 int32, 4 int64.  per_row: int64 [3][rows] (escapes at 4 / 8 / 16 bits).
extern "C" int dca_count_escapes(const void* counts, int32_t dtype, int64_t rows, int64_t cols, int64_t ld, int64_t* per_row,
                                 int32_t threads) {
  if (!counts || !per_row || rows < 0 || cols < 0 || ld < cols) { set_error("dca_count_escapes: bad argument"); return DCA_ERR_BAD_ARG; }
  if (threads <= 0) { threads = (int)std::thread::hardware_concurrency(); if (threads > 32) threads = 32; if (threads < 1) threads = 1; }
  bool ok = false;
  switch (dtype) {
    case 0: ok = escape_counts((const float*)counts, rows, cols, ld, per_row, threads); break;
    case 1: ok = escape_counts((const double*)counts, rows, cols, ld, per_row, threads); break;
    case 2: ok = escape_counts((const uint16_t*)counts, rows, cols, ld, per_row, threads); break;
    case 3: ok = escape_counts((const int32_t*)counts, rows, cols, ld, per_row, threads); break;
    case 4: ok = escape_counts((const int64_t*)counts, rows, cols, ld, per_row, threads); break;
    default: set_error("dca_count_escapes: unknown dtype %d", dtype); return DCA_ERR_BAD_ARG;
  }
  if (!ok) { set_error("dca_count_escapes: counts must be non-negative integers"); return DCA_ERR_BAD_ARG; }
  return DCA_OK;
}

// packed: rows x (cols*bits/8) bytes; indptr: int64[rows+1] (exclusive prefix sum of the per-row escape counts of this
// width, as returned by dca_count_escapes); entries: 8 bytes each, indptr[rows] of them.
extern "C" int dca_pack_counts(const void* counts, int32_t dtype, int64_t rows, int64_t cols, int64_t ld, int32_t bits,
                               void* packed, const int64_t* indptr, void* entries, int32_t threads) {
  if (!counts || !packed || !indptr || rows < 0 || cols < 0 || ld < cols || (indptr[rows] > 0 && !entries)) { set_error("dca_pack_counts: bad argument"); return DCA_ERR_BAD_ARG; }
  if (bits != 4 && bits != 8 && bits != 16) { set_error("dca_pack_counts: bits must be 4, 8 or 16"); return DCA_ERR_BAD_ARG; }
  if (cols % 8 != 0) { set_error("dca_pack_counts: the number of genes must be a multiple of 8"); return DCA_ERR_BAD_ARG; }
  if (threads <= 0) { threads = (int)std::thread::hardware_concurrency(); if (threads > 32) threads = 32; if (threads < 1) threads = 1; }
  unsigned char* p = (unsigned char*)packed; unsigned char* e = (unsigned char*)entries;
  switch (dtype) {
    case 0: pack_rows((const float*)counts, rows, cols, ld, bits, p, indptr, e, threads); break;
    case 1: pack_rows((const double*)counts, rows, cols, ld, bits, p, indptr, e, threads); break;
    case 2: pack_rows((const uint16_t*)counts, rows, cols, ld, bits, p, indptr, e, threads); break;
    case 3: pack_rows((const int32_t*)counts, rows, cols, ld, bits, p, indptr, e, threads); break;
    case 4: pack_rows((const int64_t*)counts, rows, cols, ld, bits, p, indptr, e, threads); break;
    default: set_error("dca_pack_counts: unknown dtype %d", dtype); return DCA_ERR_BAD_ARG;
  }
  return DCA_OK;
}

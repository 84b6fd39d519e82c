// CPU restatement of the jxl-rs VarDCT per-group hot path. See oracle.h for
// scope, usage restrictions and parity status. Every function cites the
// reference code it restates (paths relative to /root/reference).
#include "oracle.h"

#ifdef __AVX2__
#include <immintrin.h>
#endif
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstring>
#include <functional>
#include <mutex>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "../jxl_rs_b200/csrc/host/frame.h"  // front-end frame state (host parser), not the hot path

namespace {

thread_local std::string g_last_error;

// ---------------------------------------------------------------------------
// a1. Bit reader — jxl/src/bit_reader.rs:15-219 (optimistic reads: bits past
// the end are zeros; over-read is detected once at the end, :109)
// ---------------------------------------------------------------------------
struct Br {
  const uint8_t* data;
  size_t size, pos = 0;
  uint64_t buf = 0;
  unsigned bits = 0;
  size_t total = 0;
  Br(const uint8_t* d, size_t n) : data(d), size(n) {}
  inline void refill() {
    while (bits <= 56) {
      uint64_t b = pos < size ? data[pos] : 0;
      pos++;
      buf |= b << bits;
      bits += 8;
    }
  }
  inline uint64_t peek(unsigned n) {
    if (bits < n) refill();
    return buf & ((uint64_t(1) << n) - 1);
  }
  inline void consume(unsigned n) {
    buf >>= n;
    bits = bits >= n ? bits - n : 0;
    total += n;
  }
  inline uint64_t read(unsigned n) {
    uint64_t v = peek(n);
    consume(n);
    return v;
  }
  bool overrun() const { return total > size * 8; }
};

inline uint32_t ceil_log2_u(uint64_t x) {
  uint32_t n = 0;
  while ((uint64_t(1) << n) < x) n++;
  return n;
}
inline size_t shrc(size_t v, unsigned s) { return (v + (size_t(1) << s) - 1) >> s; }  // util ShiftRightCeil

// ---------------------------------------------------------------------------
// a4-a7. Symbol reader over the flat tables of a JxgPassDesc —
// entropy_coding/decode.rs:271-332, ans.rs:356-393, huffman.rs:446-457,
// hybrid_uint.rs:87-102
// ---------------------------------------------------------------------------
struct SymReader {
  const JxgPassDesc& p;
  uint32_t state = 0x130000;
  std::vector<uint32_t> window;
  uint32_t num_to_copy = 0, copy_pos = 0, num_decoded = 0;
  bool err_lz77 = false;
  SymReader(const JxgPassDesc& pd, Br& br) : p(pd) {
    if (!p.use_prefix) state = uint32_t(br.read(32));  // ans.rs:431
  }
  static inline uint32_t hybrid(uint32_t cfg, uint32_t token, Br& br) {
    uint32_t split_exponent = cfg & 0xff, msb = (cfg >> 8) & 0xff, lsb = (cfg >> 16) & 0xff;
    uint32_t split_token = 1u << split_exponent;
    if (token < split_token) return token;
    uint32_t bits_in_token = lsb + msb;
    uint32_t nbits = (split_exponent - bits_in_token + ((token - split_token) >> bits_in_token)) & 31;
    uint32_t low = token & ((1u << lsb) - 1);
    uint32_t token_nolow = token >> lsb;
    uint32_t bits = uint32_t(br.read(nbits));
    uint32_t hi = (token_nolow & ((1u << msb) - 1)) | (1u << msb);
    return (((hi << nbits) | bits) << lsb) | low;
  }
  inline uint32_t token(Br& br, uint32_t cluster) {
    if (p.use_prefix) {
      const uint32_t* t = p.huff_entries + p.huff_offset[cluster];
      size_t pos = size_t(br.peek(8));
      uint32_t n_bits = t[pos] & 0xff;
      if (n_bits > 8) {
        br.consume(8);
        n_bits -= 8;
        pos += t[pos] >> 16;
        pos += size_t(br.peek(n_bits));
      }
      uint32_t e = t[pos];
      br.consume(e & 0xff);
      return e >> 16;
    }
    const uint32_t log_bucket = 12 - p.log_alpha_size;
    uint32_t idx = state & 0xfff;
    uint32_t i = idx >> log_bucket;
    uint32_t pos = idx & ((1u << log_bucket) - 1);
    uint64_t b = p.ans_buckets[(size_t(cluster) << p.log_alpha_size) + i];
    uint32_t alias_symbol = uint32_t(b & 0xff), alias_cutoff = uint32_t((b >> 8) & 0xff);
    uint32_t dist = uint32_t((b >> 16) & 0xffff), alias_offset = uint32_t((b >> 32) & 0xffff);
    uint32_t alias_dist_xor = uint32_t((b >> 48) & 0xffff);
    uint32_t map_to_alias = pos >= alias_cutoff;
    uint32_t offset = alias_offset * map_to_alias + pos;
    dist ^= alias_dist_xor * map_to_alias;
    uint32_t symbol = map_to_alias ? alias_symbol : i;
    uint32_t next = (state >> 12) * dist + offset;
    if (next < (1u << 16)) {
      next = (next << 16) | uint32_t(br.peek(16));
      br.consume(16);
    }
    state = next;
    return symbol;
  }
  inline uint32_t read_clustered(Br& br, uint32_t cluster) {
    if (!p.lz77_enabled) return hybrid(p.uint_configs[cluster], token(br, cluster), br);
    constexpr uint32_t kMask = (1u << 20) - 1;
    auto push = [&](uint32_t v) {
      size_t off = num_decoded & kMask;
      if (off < window.size()) window[off] = v;
      else window.push_back(v);
      num_decoded++;
    };
    if (num_to_copy > 0) {
      uint32_t sym = window[copy_pos & kMask];
      copy_pos++;
      num_to_copy--;
      push(sym);
      return sym;
    }
    uint32_t tok = token(br, cluster);
    if (tok < p.lz77_min_symbol) {
      uint32_t sym = hybrid(p.uint_configs[cluster], tok, br);
      push(sym);
      return sym;
    }
    if (num_decoded == 0) {
      err_lz77 = true;
      return 0;
    }
    uint32_t n = hybrid(p.lz77_length_uint, tok - p.lz77_min_symbol, br);
    if (n > 0xffffffffu - p.lz77_min_length) {
      err_lz77 = true;
      return 0;
    }
    n += p.lz77_min_length;
    uint32_t dc = p.lz_dist_cluster;
    uint32_t distance_sym = hybrid(p.uint_configs[dc], token(br, dc), br);
    // HF streams: image_width None => dist_multiplier 0 (group.rs:345-349, decode.rs:204)
    uint32_t distance = std::min(std::min<uint32_t>((1u << 20) - 1, distance_sym) + 1, num_decoded);
    copy_pos = num_decoded - distance;
    num_to_copy = n;
    uint32_t sym = window[copy_pos & kMask];
    copy_pos++;
    num_to_copy--;
    push(sym);
    return sym;
  }
  inline uint32_t read_unsigned(Br& br, size_t ctx) { return read_clustered(br, p.context_map[ctx]); }
};
inline int32_t unpack_signed_u(uint32_t u) { return int32_t((u >> 1) ^ (((~u) & 1) - 1)); }

// ---------------------------------------------------------------------------
// transform geometry — jxl_transforms/src/transform_map.rs:87-116
// ---------------------------------------------------------------------------
const uint8_t kCovX[27] = {1, 1, 1, 1, 2, 4, 1, 2, 1, 4, 2, 4, 1, 1, 1, 1, 1, 1, 8, 4, 8, 16, 8, 16, 32, 16, 32};
const uint8_t kCovY[27] = {1, 1, 1, 1, 2, 4, 2, 1, 4, 1, 4, 2, 1, 1, 1, 1, 1, 1, 8, 8, 4, 16, 16, 8, 32, 32, 16};
const uint8_t kShape[27] = {0, 1, 1, 1, 2, 3, 4, 4, 5, 5, 6, 6, 1, 1, 1, 1, 1, 1, 7, 8, 8, 9, 10, 10, 11, 12, 12};
const uint8_t kOrderT[13] = {0, 1, 4, 5, 7, 9, 11, 18, 20, 21, 23, 24, 26};  // coeff_order.rs:23-37
const uint8_t kQuantTable[27] = {0, 1, 2, 3, 4, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 10, 10, 11, 12, 12, 13, 14, 14, 15, 16, 16};

// a8. natural coefficient order — frame/coeff_order.rs:66-120
std::vector<uint32_t> natural_order(int shape) {
  int t = kOrderT[shape];
  size_t cx = kCovX[t], cy = kCovY[t], xsize = cx * 8;
  size_t xs = cx / cy, xsm = xs - 1, xss = ceil_log2_u(xs);
  std::vector<uint32_t> out(cx * cy * 64);
  size_t cur = cx * cy;
  for (size_t i = 0; i < xsize; i++)
    for (size_t j = 0; j <= i; j++) {
      size_t x = j, y = i - j;
      if (i & 1) std::swap(x, y);
      if (y & xsm) continue;
      y >>= xss;
      size_t val = (x < cx && y < cy) ? y * cx + x : cur++;
      out[val] = uint32_t(y * xsize + x);
    }
  for (size_t ir = 1; ir < xsize; ir++) {
    size_t i = xsize - ir - 1;
    for (size_t j = 0; j <= i; j++) {
      size_t x = xsize - 1 - (i - j), y = xsize - 1 - j;
      if (i & 1) std::swap(x, y);
      if (y & xsm) continue;
      y >>= xss;
      out[cur++] = uint32_t(y * xsize + x);
    }
  }
  return out;
}
const std::vector<uint32_t>& natural_order_cached(int shape) {
  static std::vector<uint32_t> orders[13];
  static std::once_flag once;
  std::call_once(once, [] {
    for (int s = 0; s < 13; s++) orders[s] = natural_order(s);
  });
  return orders[shape];
}

// a3. context tables — frame/block_context_map.rs:20-31
const uint16_t kFreqCtx[64] = {0xBAD, 0,  1,  2,  3,  4,  5,  6,  7,  8,  9,  10, 11, 12, 13, 14, 15, 15, 16, 16, 17, 17,
                               18,    18, 19, 19, 20, 20, 21, 21, 22, 22, 23, 23, 23, 23, 24, 24, 24, 24, 25, 25, 25, 25,
                               26,    26, 26, 26, 27, 27, 27, 27, 28, 28, 28, 28, 29, 29, 29, 29, 30, 30, 30, 30};
const uint16_t kNzCtx[64] = {0xBAD, 0,   31,  62,  62,  93,  93,  93,  93,  123, 123, 123, 123, 152, 152, 152,
                             152,   152, 152, 152, 152, 180, 180, 180, 180, 180, 180, 180, 180, 180, 180, 180,
                             180,   206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206,
                             206,   206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206};

// ---------------------------------------------------------------------------
// a11. inverse DCTs — recursion of jxl_transforms/gen_idct.py:46-127 and
// idct_large.rs:251-310; layouts per tests.rs:123-136 (slow_idct2d)
// ---------------------------------------------------------------------------
struct WcTable {
  float w[9][128];  // [log2 n][i] = 1 / (2 cos((i + 0.5) pi / n))
  WcTable() {
    for (int l = 1; l <= 8; l++) {
      int n = 1 << l;
      for (int i = 0; i < n / 2; i++) w[l][i] = float(1.0 / (2.0 * std::cos((i + 0.5) * M_PI / n)));
    }
  }
};
const WcTable kWc;

void idct1d(float* v, int n, int log_n, float* scratch) {
  if (n == 1) return;
  if (n == 2) {
    float a = v[0], b = v[1];
    v[0] = a + b;
    v[1] = a - b;
    return;
  }
  int half = n / 2;
  float *first = scratch, *second = scratch + half;
  for (int i = 0; i < half; i++) {
    first[i] = v[2 * i];
    second[i] = v[2 * i + 1];
  }
  idct1d(first, half, log_n - 1, scratch + n);
  for (int i = half - 1; i >= 1; i--) second[i] += second[i - 1];
  second[0] *= float(M_SQRT2);
  idct1d(second, half, log_n - 1, scratch + n);
  const float* wc = kWc.w[log_n];
  for (int i = 0; i < half; i++) {
    v[i] = std::fmaf(second[i], wc[i], first[i]);
    v[n - 1 - i] = std::fmaf(-second[i], wc[i], first[i]);
  }
}

int ilog2(int n) {
  int l = 0;
  while ((1 << l) < n) l++;
  return l;
}

#ifdef __AVX2__
// The same recursion with eight independent 1-D transforms per step (one per lane): lane i runs exactly the scalar
// sequence of idct1d on its own data, so the results are bit-identical. Used by idct2d when the fast CPU forms are on
// (jxo_set_fast_cpu): eight rows at a time in the horizontal pass, eight columns at a time in the vertical one.
static std::atomic<int> g_fast_cpu{0};
void idct1d_x8(__m256* v, int n, int log_n, __m256* scratch) {
  if (n == 1) return;
  if (n == 2) {
    const __m256 a = v[0], b = v[1];
    v[0] = _mm256_add_ps(a, b);
    v[1] = _mm256_sub_ps(a, b);
    return;
  }
  const int half = n / 2;
  __m256 *first = scratch, *second = scratch + half;
  for (int i = 0; i < half; i++) {
    first[i] = v[2 * i];
    second[i] = v[2 * i + 1];
  }
  idct1d_x8(first, half, log_n - 1, scratch + n);
  for (int i = half - 1; i >= 1; i--) second[i] = _mm256_add_ps(second[i], second[i - 1]);
  second[0] = _mm256_mul_ps(second[0], _mm256_set1_ps(float(M_SQRT2)));
  idct1d_x8(second, half, log_n - 1, scratch + n);
  const float* wc = kWc.w[log_n];
  for (int i = 0; i < half; i++) {
    const __m256 w = _mm256_set1_ps(wc[i]);
    v[i] = _mm256_fmadd_ps(second[i], w, first[i]);
    v[n - 1 - i] = _mm256_fnmadd_ps(second[i], w, first[i]);  // fma(-second, w, first)
  }
}

// rows and cols multiples of 8, at most 256.
void idct2d_x8(int rows, int cols, float* block) {
  static thread_local std::vector<float> tmp;
  alignas(32) static thread_local __m256 line[256], scratch[2 * 256 + 16];
  tmp.resize(size_t(rows) * cols);
  const bool wide = rows < cols;
  alignas(32) float lanes[8];
  // horizontal pass: rows vf .. vf + 7 in the eight lanes
  for (int vf = 0; vf < rows; vf += 8) {
    for (int hf = 0; hf < cols; hf++) {
      if (wide) {
        for (int i = 0; i < 8; i++) lanes[i] = block[(vf + i) * cols + hf];
        line[size_t(hf)] = _mm256_load_ps(lanes);
      } else {
        line[size_t(hf)] = _mm256_loadu_ps(block + hf * rows + vf);
      }
    }
    idct1d_x8(line, cols, ilog2(cols), scratch);
    for (int x = 0; x < cols; x++) {
      _mm256_store_ps(lanes, line[size_t(x)]);
      for (int i = 0; i < 8; i++) tmp[size_t(vf + i) * cols + x] = lanes[i];
    }
  }
  // vertical pass: columns x .. x + 7 in the eight lanes
  for (int x = 0; x < cols; x += 8) {
    for (int vf = 0; vf < rows; vf++) line[size_t(vf)] = _mm256_loadu_ps(tmp.data() + size_t(vf) * cols + x);
    idct1d_x8(line, rows, ilog2(rows), scratch);
    for (int y = 0; y < rows; y++) _mm256_storeu_ps(block + size_t(y) * cols + x, line[size_t(y)]);
  }
}
#else
static std::atomic<int> g_fast_cpu{0};
#endif

// 2-D IDCT, in place. Coefficient layout: rows < cols -> [vfreq][hfreq] with
// stride cols; otherwise [hfreq][vfreq] with stride rows. Output: rows x cols.
void idct2d(int rows, int cols, float* block) {
#ifdef __AVX2__
  if (g_fast_cpu.load(std::memory_order_relaxed) != 0 && rows % 8 == 0 && cols % 8 == 0) {
    idct2d_x8(rows, cols, block);
    return;
  }
#endif
  std::vector<float> tmp(size_t(rows) * cols), line(std::max(rows, cols)), scratch(2 * std::max(rows, cols) + 16);
  const bool wide = rows < cols;
  // horizontal pass (size cols) for every vertical frequency
  for (int vf = 0; vf < rows; vf++) {
    for (int hf = 0; hf < cols; hf++) line[hf] = wide ? block[vf * cols + hf] : block[hf * rows + vf];
    idct1d(line.data(), cols, ilog2(cols), scratch.data());
    for (int x = 0; x < cols; x++) tmp[size_t(vf) * cols + x] = line[x];
  }
  // vertical pass (size rows)
  for (int x = 0; x < cols; x++) {
    for (int vf = 0; vf < rows; vf++) line[vf] = tmp[size_t(vf) * cols + x];
    idct1d(line.data(), rows, ilog2(rows), scratch.data());
    for (int y = 0; y < rows; y++) block[size_t(y) * cols + x] = line[y];
  }
}

// a10. "reinterpreting" forward DCT of the LF samples —
// jxl_transforms/gen_reinterpreting_dct.py:47-136; output scale constants
// carry 6 decimals exactly as the generated reference code does.
void rdct1d_rec(float* v, int n, int log_n, float* scratch) {
  if (n == 1) return;
  if (n == 2) {
    float a = v[0], b = v[1];
    v[0] = a + b;
    v[1] = a - b;
    return;
  }
  int half = n / 2;
  float *first = scratch, *second = scratch + half;
  for (int i = 0; i < half; i++) {
    first[i] = v[i] + v[n - 1 - i];
    second[i] = v[i] - v[n - 1 - i];
  }
  rdct1d_rec(first, half, log_n - 1, scratch + n);
  const float* wc = kWc.w[log_n];
  for (int i = 0; i < half; i++) second[i] *= wc[i];
  rdct1d_rec(second, half, log_n - 1, scratch + n);
  second[0] = std::fmaf(second[0], float(M_SQRT2), second[1]);
  for (int i = 1; i + 1 < half; i++) second[i] = second[i] + second[i + 1];
  for (int i = 0; i < half; i++) {
    v[2 * i] = first[i];
    v[2 * i + 1] = second[i];
  }
}
float rdct_scale(int i, int n) {
  double s = std::cos(i / (16.0 * n) * M_PI) * std::cos(i / (8.0 * n) * M_PI) * std::cos(i / (4.0 * n) * M_PI) * n;
  double inv = 1.0 / s;
  return float(std::round(inv * 1e6) / 1e6);  // "%f" in the generator
}
void rdct1d(float* v, int n, float* scratch) {
  if (n == 1) return;
  rdct1d_rec(v, n, ilog2(n), scratch);
  for (int i = 0; i < n; i++) v[i] *= rdct_scale(i, n);
}
// rows x cols LF samples -> top-left of the coefficient block (tests.rs:154-180):
// rows < cols: out[vf * stride + hf]; else out[hf * stride + vf].
void reinterpreting_dct2d(int rows, int cols, const float* in, float* out, int out_stride) {
  std::vector<float> tmp(size_t(rows) * cols), line(std::max(rows, cols)), scratch(2 * std::max(rows, cols) + 16);
  for (int y = 0; y < rows; y++) {
    for (int x = 0; x < cols; x++) line[x] = in[y * cols + x];
    rdct1d(line.data(), cols, scratch.data());
    for (int x = 0; x < cols; x++) tmp[size_t(y) * cols + x] = line[x];
  }
  const bool wide = rows < cols;
  for (int hf = 0; hf < cols; hf++) {
    for (int y = 0; y < rows; y++) line[y] = tmp[size_t(y) * cols + hf];
    rdct1d(line.data(), rows, scratch.data());
    for (int vf = 0; vf < rows; vf++) {
      if (wide) out[vf * out_stride + hf] = line[vf];
      else out[hf * out_stride + vf] = line[vf];
    }
  }
}

const float kAfvBasis[256] = {
#include "afv_basis.inc"
};

// transform.rs:14-32
void idct2_top_block(int s, const float* in, float* out) {
  int n = s / 2;
  for (int y = 0; y < n; y++)
    for (int x = 0; x < n; x++) {
      float c00 = in[y * 8 + x], c01 = in[y * 8 + n + x], c10 = in[(y + n) * 8 + x], c11 = in[(y + n) * 8 + n + x];
      out[y * 2 * 8 + x * 2] = c00 + c01 + c10 + c11;
      out[y * 2 * 8 + x * 2 + 1] = c00 + c01 - c10 - c11;
      out[(y * 2 + 1) * 8 + x * 2] = c00 - c01 + c10 - c11;
      out[(y * 2 + 1) * 8 + x * 2 + 1] = c00 - c01 - c10 + c11;
    }
}

// transform.rs:306-374
void afv_transform(int kind, const float* co, float* px) {
  int afv_x = kind & 1, afv_y = kind / 2;
  float b00 = co[0], b01 = co[1], b10 = co[8];
  float dcs[3] = {(b00 + b10 + b01) * 4.0f, b00 + b10 - b01, b00 - b10};
  float coeff[16], block[32];
  for (int iy = 0; iy < 4; iy++)
    for (int ix = 0; ix < 4; ix++) coeff[iy * 4 + ix] = (ix == 0 && iy == 0) ? dcs[0] : co[iy * 2 * 8 + ix * 2];
  for (int i = 0; i < 16; i++) {  // avfidct4x4, transform.rs:295-304
    float p = 0.0f;
    for (int j = 0; j < 16; j++) p += coeff[j] * kAfvBasis[j * 16 + i];
    block[i] = p;
  }
  for (int iy = 0; iy < 4; iy++) {
    int by = afv_y ? 3 - iy : iy;
    for (int ix = 0; ix < 4; ix++) {
      int bx = afv_x ? 3 - ix : ix;
      px[(iy + afv_y * 4) * 8 + afv_x * 4 + ix] = block[by * 4 + bx];
    }
  }
  for (int iy = 0; iy < 4; iy++)
    for (int ix = 0; ix < 4; ix++) block[iy * 4 + ix] = (ix == 0 && iy == 0) ? dcs[1] : co[iy * 2 * 8 + ix * 2 + 1];
  idct2d(4, 4, block);
  for (int iy = 0; iy < 4; iy++)
    for (int ix = 0; ix < 4; ix++) px[(iy + afv_y * 4) * 8 + (1 - afv_x) * 4 + ix] = block[iy * 4 + ix];
  for (int iy = 0; iy < 4; iy++)
    for (int ix = 0; ix < 8; ix++) block[iy * 8 + ix] = (ix == 0 && iy == 0) ? dcs[2] : co[(1 + iy * 2) * 8 + ix];
  idct2d(4, 8, block);
  for (int iy = 0; iy < 4; iy++)
    for (int ix = 0; ix < 8; ix++) px[(iy + (1 - afv_y) * 4) * 8 + ix] = block[iy * 8 + ix];
}

// transform.rs:377-664 — buf holds the dequantised coefficients on entry and
// rows x cols pixels (stride = cols) on return.
void transform_to_pixels(int t, const float* lf, float* buf) {
  int cx = kCovX[t], cy = kCovY[t];
  switch (t) {
    case 0:  // DCT
      buf[0] = lf[0];
      idct2d(8, 8, buf);
      return;
    case 1: {  // IDENTITY ("Hornuss"), transform.rs:530-571
      buf[0] = lf[0];
      float co[64];
      memcpy(co, buf, sizeof(co));
      float b00 = co[0], b01 = co[1], b10 = co[8], b11 = co[9];
      float dcs[4] = {b00 + b01 + b10 + b11, b00 + b01 - b10 - b11, b00 - b01 + b10 - b11, b00 - b01 - b10 + b11};
      for (int y = 0; y < 2; y++)
        for (int x = 0; x < 2; x++) {
          float block_dc = dcs[y * 2 + x];
          float residual_sum = 0.0f;
          for (int iy = 0; iy < 4; iy++)
            for (int ix = 0; ix < 4; ix++) {
              if (ix == 0 && iy == 0) continue;
              residual_sum += co[(y + iy * 2) * 8 + x + ix * 2];
            }
          float center = block_dc - residual_sum * (1.0f / 16.0f);
          buf[(4 * y + 1) * 8 + 4 * x + 1] = center;
          for (int iy = 0; iy < 4; iy++)
            for (int ix = 0; ix < 4; ix++) {
              if (ix == 1 && iy == 1) continue;
              buf[(y * 4 + iy) * 8 + x * 4 + ix] = co[(y + iy * 2) * 8 + x + ix * 2] + center;
            }
          buf[y * 4 * 8 + x * 4] = co[(y + 2) * 8 + x + 2] + center;
        }
      return;
    }
    case 2: {  // DCT2X2, transform.rs:572-578
      buf[0] = lf[0];
      float tmp[64];
      memcpy(tmp, buf, sizeof(tmp));
      idct2_top_block(2, tmp, buf);
      idct2_top_block(4, buf, tmp);
      idct2_top_block(8, tmp, buf);
      return;
    }
    case 3: {  // DCT4X4, transform.rs:579-612
      buf[0] = lf[0];
      float co[64];
      memcpy(co, buf, sizeof(co));
      float b00 = co[0], b01 = co[1], b10 = co[8], b11 = co[9];
      float dcs[4] = {b00 + b01 + b10 + b11, b00 + b01 - b10 - b11, b00 - b01 + b10 - b11, b00 - b01 - b10 + b11};
      for (int y = 0; y < 2; y++)
        for (int x = 0; x < 2; x++) {
          float block[16];
          for (int iy = 0; iy < 4; iy++)
            for (int ix = 0; ix < 4; ix++) block[iy * 4 + ix] = co[(y + iy * 2) * 8 + x + ix * 2];
          block[0] = dcs[y * 2 + x];
          idct2d(4, 4, block);
          for (int iy = 0; iy < 4; iy++)
            for (int ix = 0; ix < 4; ix++) buf[(y * 4 + iy) * 8 + x * 4 + ix] = block[iy * 4 + ix];
        }
      return;
    }
    case 12:    // DCT4X8, transform.rs:638-661
    case 13: {  // DCT8X4, transform.rs:613-637
      buf[0] = lf[0];
      float co[64];
      memcpy(co, buf, sizeof(co));
      float dcs[2] = {co[0] + co[8], co[0] - co[8]};
      for (int h = 0; h < 2; h++) {
        float block[32];
        for (int iy = 0; iy < 4; iy++)
          for (int ix = 0; ix < 8; ix++) block[iy * 8 + ix] = (ix == 0 && iy == 0) ? dcs[h] : co[(h + iy * 2) * 8 + ix];
        if (t == 13) {
          idct2d(8, 4, block);
          for (int iy = 0; iy < 8; iy++)
            for (int ix = 0; ix < 4; ix++) buf[iy * 8 + h * 4 + ix] = block[iy * 4 + ix];
        } else {
          idct2d(4, 8, block);
          for (int iy = 0; iy < 4; iy++)
            for (int ix = 0; ix < 8; ix++) buf[(h * 4 + iy) * 8 + ix] = block[iy * 8 + ix];
        }
      }
      return;
    }
    case 14: case 15: case 16: case 17: {  // AFV0..3
      buf[0] = lf[0];
      float co[64];
      memcpy(co, buf, sizeof(co));
      afv_transform(t - 14, co, buf);
      return;
    }
    default: {  // plain DCTs of (8*cy) x (8*cx)
      int rows = 8 * cy, cols = 8 * cx;
      reinterpreting_dct2d(cy, cx, lf, buf, 8 * std::max(cx, cy));
      idct2d(rows, cols, buf);
      return;
    }
  }
}

// ---------------------------------------------------------------------------
// frame-level helpers
// ---------------------------------------------------------------------------
struct Geometry {
  uint32_t width, height, xb, yb, xg, yg, num_groups;
  explicit Geometry(const JxgFrameDesc& d) {
    width = d.width;
    height = d.height;
    xb = (width + 7) / 8;
    yb = (height + 7) / 8;
    xg = (width + 255) / 256;
    yg = (height + 255) / 256;
    num_groups = xg * yg;
  }
};

void parallel_for(int n, int num_threads, const std::function<void(int)>& f) {
  if (num_threads <= 1 || n <= 1) {
    for (int i = 0; i < n; i++) f(i);
    return;
  }
  std::atomic<int> next{0};
  std::vector<std::thread> th;
  for (int t = 0; t < std::min(num_threads, n); t++)
    th.emplace_back([&] {
      for (;;) {
        int i = next.fetch_add(1);
        if (i >= n) break;
        f(i);
      }
    });
  for (auto& t : th) t.join();
}

const float* dequant_table(const JxgFrameDesc& d, int idx) {
  if (d.dequant_tables[idx]) return d.dequant_tables[idx];
  return jxg::library_dequant_table(idx).data();  // host-side table construction (not on the hot path)
}

// a9. group.rs:85-177: adjust_quant_bias, dequantisation with the per-block scale and the table weight, chroma from
// luma (X and B get x_cc / b_cc times the dequantised Y coefficient through mul_add).
inline void dequant_block(size_t num_coeffs, const int32_t* qx, const int32_t* qy, const int32_t* qb, const float* mat, float sx,
                          float sy, float sb, float x_cc, float b_cc, const float* bias, float* ox, float* oy, float* ob) {
  auto adj = [&](int c, int32_t q) {
    float qf = float(q);
    return std::abs(q) < 2 ? qf * bias[c] : qf - bias[3] / qf;
  };
  for (size_t k = 0; k < num_coeffs; k++) {
    float dy = adj(1, qy[k]) * (mat[num_coeffs + k] * sy);
    float dxc = adj(0, qx[k]) * (mat[k] * sx);
    float dbc = adj(2, qb[k]) * (mat[2 * num_coeffs + k] * sb);
    oy[k] = dy;
    ox[k] = std::fmaf(x_cc, dy, dxc);
    ob[k] = std::fmaf(b_cc, dy, dbc);
  }
}

// a2, a3, a8, a9: one HF group — jxl/src/frame/group.rs:383-632
int decode_group(const JxgFrameDesc& d, const Geometry& geo, uint32_t g, const uint8_t* hf, const uint64_t* sec_off,
                 const uint32_t* sec_len, int32_t* coeffs /* [3][65536] */, float* planes[3], size_t plane_stride) {
  const uint32_t gx = g % geo.xg, gy = g / geo.xg;
  const uint32_t bx0 = gx * 32, by0 = gy * 32;
  const uint32_t gw = std::min(32u, geo.xb - bx0), gh = std::min(32u, geo.yb - by0);
  const float x_dm = std::pow(1.0f / 1.25f, float(d.x_qm_scale) - 2.0f);
  const float b_dm = std::pow(1.0f / 1.25f, float(d.b_qm_scale) - 2.0f);
  const float inv_global_scale = 65536.0f / float(d.global_scale);
  const size_t num_ac_contexts = size_t(d.num_block_contexts) * (37 + 458);
  const uint32_t cxb = (geo.xb + 7) / 8;

  struct Pass {
    Br br;
    std::unique_ptr<SymReader> reader;
    size_t histogram_index;
    uint32_t nz[3][32 * 32];
  };
  std::vector<std::unique_ptr<Pass>> passes;
  for (uint32_t p = 0; p < d.num_passes; p++) {
    size_t s = size_t(p) * geo.num_groups + g;
    auto ps = std::make_unique<Pass>(Pass{Br(hf + sec_off[s], sec_len[s]), nullptr, 0, {{0}}});
    ps->histogram_index = size_t(ps->br.read(ceil_log2_u(d.num_histograms)));  // group.rs:333-341
    if (ps->histogram_index >= d.num_histograms) return JXG_ERR_INVALID_HISTOGRAM_INDEX;
    ps->reader = std::make_unique<SymReader>(d.passes[p], ps->br);
    passes.push_back(std::move(ps));
  }
  memset(coeffs, 0, sizeof(int32_t) * 3 * 65536);
  size_t coeffs_offset = 0;
  std::vector<float> tbuf[3];
  for (auto& t : tbuf) t.resize(65536);
  float lfbuf[32 * 32];

  for (uint32_t by = 0; by < gh; by++) {
    for (uint32_t bx = 0; bx < gw; bx++) {
      const size_t bidx = size_t(by0 + by) * geo.xb + bx0 + bx;
      uint8_t raw_t = d.transform_map[bidx];
      if (raw_t < 128) continue;
      int t = raw_t & 127;
      if (t >= 27) return JXG_ERR_INVALID_TRANSFORM;
      const uint32_t cx = kCovX[t], cy = kCovY[t];
      const int shape = kShape[t];
      const uint32_t raw_quant = uint32_t(d.raw_quant_map[bidx]);
      const uint32_t quant_lf = d.quant_lf[bidx];
      const size_t num_blocks = size_t(cx) * cy, num_coeffs = num_blocks * 64;
      unsigned log_num_blocks = 0;
      while ((size_t(1) << log_num_blocks) < num_blocks) log_num_blocks++;
      for (uint32_t p = 0; p < d.num_passes; p++) {
        Pass& ps = *passes[p];
        const JxgPassDesc& pd = d.passes[p];
        const size_t context_offset = ps.histogram_index * num_ac_contexts;
        for (int c : {1, 0, 2}) {
          uint32_t* nz = ps.nz[c];
          // predict_num_nonzeros, group.rs:70-83
          size_t predicted;
          if (bx == 0) predicted = by == 0 ? 32 : nz[(by - 1) * 32];
          else if (by == 0) predicted = nz[bx - 1];
          else predicted = (nz[(by - 1) * 32 + bx] + nz[by * 32 + bx - 1] + 1) / 2;
          // block_context, block_context_map.rs:127-139
          size_t qf_idx = 0;
          for (uint32_t i = 0; i < d.num_qf_thresholds; i++) qf_idx += raw_quant > d.qf_thresholds[i];
          size_t idx = c < 2 ? size_t(c ^ 1) : 2;
          idx = idx * 13 + size_t(shape);
          idx = idx * (d.num_qf_thresholds + 1) + qf_idx;
          idx = idx * d.num_lf_contexts + quant_lf;
          size_t block_context = d.block_ctx_map[idx];
          // nonzero_context, :141-150
          size_t nzc = predicted < 8 ? predicted : predicted < 64 ? 4 + predicted / 2 : 36;
          size_t nonzero_context = nzc * d.num_block_contexts + block_context + context_offset;
          size_t nonzeros = ps.reader->read_unsigned(ps.br, nonzero_context);
          if (nonzeros + num_blocks > num_coeffs) return JXG_ERR_INVALID_NUM_NONZEROS;
          for (uint32_t iy = 0; iy < cy; iy++)
            for (uint32_t ix = 0; ix < cx; ix++) nz[(by + iy) * 32 + bx + ix] = uint32_t(shrc(nonzeros, log_num_blocks));
          size_t histo_offset = size_t(d.num_block_contexts) * 37 + 458 * block_context + context_offset;
          size_t prev = nonzeros > num_coeffs / 16 ? 0 : 1;
          const uint32_t* order = pd.coeff_order ? pd.coeff_order + pd.coeff_order_offset[shape * 3 + c]
                                                 : natural_order_cached(shape).data();
          int32_t* cur = coeffs + size_t(c) * 65536 + coeffs_offset;
          for (size_t k = num_blocks; k < num_coeffs; k++) {
            if (nonzeros == 0) break;
            // zero_density_context, block_context_map.rs:34-46
            size_t ctx = histo_offset +
                         (kNzCtx[shrc(nonzeros, log_num_blocks) & 63] + kFreqCtx[(k >> log_num_blocks) & 63]) * 2 + prev;
            uint32_t u = ps.reader->read_unsigned(ps.br, ctx);
            int32_t coeff = int32_t(uint32_t(unpack_signed_u(u)) << pd.shift);
            prev = coeff != 0;
            nonzeros -= prev;
            cur[order[k]] += coeff;
          }
          if (nonzeros != 0) return JXG_ERR_RESIDUAL_NONZEROS;
        }
      }
      if (planes[0]) {
        // a9: dequant_block / dequant_lane, group.rs:100-177
        const float x_cc = d.base_correlation_x + float(d.ytox_map[size_t((by0 + by) / 8) * cxb + (bx0 + bx) / 8]) / float(d.color_factor);
        const float b_cc = d.base_correlation_b + float(d.ytob_map[size_t((by0 + by) / 8) * cxb + (bx0 + bx) / 8]) / float(d.color_factor);
        const float sy = inv_global_scale / float(raw_quant), sx = sy * x_dm, sb = sy * b_dm;
        const float* mat = dequant_table(d, kQuantTable[t]);
        const int32_t *qx = coeffs + coeffs_offset, *qy = coeffs + 65536 + coeffs_offset, *qb = coeffs + 2 * 65536 + coeffs_offset;
        dequant_block(num_coeffs, qx, qy, qb, mat, sx, sy, sb, x_cc, b_cc, d.quant_biases, tbuf[0].data(), tbuf[1].data(), tbuf[2].data());
        for (int c : {1, 0, 2}) {
          for (uint32_t y = 0; y < cy; y++)
            for (uint32_t x = 0; x < cx; x++) lfbuf[y * cx + x] = d.lf[c][size_t(by0 + by + y) * geo.xb + bx0 + bx + x];
          transform_to_pixels(t, lfbuf, tbuf[c].data());
          const size_t w = 8 * cx, h = 8 * cy;
          for (size_t y = 0; y < h; y++)
            memcpy(planes[c] + (size_t(by0 + by) * 8 + y) * plane_stride + size_t(bx0 + bx) * 8, &tbuf[c][y * w], w * sizeof(float));
        }
      }
      coeffs_offset += num_coeffs;
    }
  }
  for (auto& ps : passes) {  // check_final_state, decode.rs:400
    if (ps->reader->err_lz77) return JXG_ERR_LZ77;
    if (ps->br.overrun()) return JXG_ERR_OUT_OF_BOUNDS;
    if (!ps->reader->p.use_prefix && ps->reader->state != 0x130000) return JXG_ERR_ANS_CHECKSUM;
  }
  return 0;
}

// AVX2 forms of Gaborish, EPF 0 / 1 / 2 and the sRGB u8 store (8 pixels per step, the per-pixel arithmetic and its order
// unchanged: outputs are bit-identical to the scalar definitions below, tests/test_cpu_paths.py checks that). Off by
// default - the checker stays the plain restatement; bench.py's CPU arm turns them on (jxo_set_fast_cpu) so that the CPU
// baseline is not handicapped by scalar filter loops the reference runs as SIMD (render/stages/epf/*.rs, gaborish.rs).
inline size_t mirror(ptrdiff_t v, size_t s) {  // util/mirror.rs:8
  for (;;) {
    if (v < 0) v = -v - 1;
    else if (size_t(v) >= s) v = ptrdiff_t(s) * 2 - v - 1;
    else return size_t(v);
  }
}

// a12. Gaborish — render/stages/gaborish.rs:19-88, whole-image semantics of
// render/simple_pipeline/run_stage.rs:127-134 (mirror at image edges)
void gaborish(size_t w, size_t h, const float* in, size_t in_stride, float* out, size_t out_stride, float w1, float w2,
              int num_threads) {
  float total = 1.0f + w1 * 4.0f + w2 * 4.0f;
  float k0 = 1.0f / total, k1 = w1 / total, k2 = w2 / total;
  const bool fast = g_fast_cpu.load() != 0;
  parallel_for(int(h), num_threads, [&](int yi) {
    size_t y = size_t(yi);
    const float* t = in + mirror(ptrdiff_t(y) - 1, h) * in_stride;
    const float* c = in + y * in_stride;
    const float* b = in + mirror(ptrdiff_t(y) + 1, h) * in_stride;
    size_t x_vec0 = w, x_vec1 = w;  // [x_vec0, x_vec1): done 8 at a time below
#ifdef __AVX2__
    if (fast && w >= 18) {
      x_vec0 = 1;
      x_vec1 = 1 + (w - 2) / 8 * 8;
      const __m256 vk0 = _mm256_set1_ps(k0), vk1 = _mm256_set1_ps(k1), vk2 = _mm256_set1_ps(k2);
      for (size_t x = x_vec0; x < x_vec1; x += 8) {
        const __m256 cc = _mm256_loadu_ps(c + x), cl = _mm256_loadu_ps(c + x - 1), cr = _mm256_loadu_ps(c + x + 1);
        const __m256 tt = _mm256_loadu_ps(t + x), tl = _mm256_loadu_ps(t + x - 1), tr = _mm256_loadu_ps(t + x + 1);
        const __m256 bb = _mm256_loadu_ps(b + x), bl = _mm256_loadu_ps(b + x - 1), br = _mm256_loadu_ps(b + x + 1);
        __m256 sum = _mm256_mul_ps(cc, vk0);
        sum = _mm256_fmadd_ps(vk1, _mm256_add_ps(_mm256_add_ps(_mm256_add_ps(tt, cl), bb), cr), sum);
        sum = _mm256_fmadd_ps(vk2, _mm256_add_ps(_mm256_add_ps(_mm256_add_ps(tl, tr), bl), br), sum);
        _mm256_storeu_ps(out + y * out_stride + x, sum);
      }
    }
#else
    (void)fast;
#endif
    for (size_t x = 0; x < w; x++) {
      if (x == x_vec0) {
        x = x_vec1 - 1;
        continue;
      }
      size_t xl = mirror(ptrdiff_t(x) - 1, w), xr = mirror(ptrdiff_t(x) + 1, w);
      float sum = c[x] * k0;
      sum = std::fmaf(k1, t[x] + c[xl] + b[x] + c[xr], sum);
      sum = std::fmaf(k2, t[xl] + t[xr] + b[xl] + b[xr], sum);
      out[y * out_stride + x] = sum;
    }
  });
}

constexpr float kInvSigmaNum = -1.1715728752538099024f;  // features/epf.rs:26
constexpr float kMinSigma = -3.90524291751269967465540850526868f;  // jxl/src/lib.rs:28

// a13. features/epf.rs:35-86
std::vector<float> sigma_image(const JxgFrameDesc& d, const Geometry& geo) {
  std::vector<float> s(size_t(geo.xb) * geo.yb);
  float quant_scale = 1.0f / (65536.0f / float(d.global_scale));
  for (size_t i = 0; i < s.size(); i++) {
    float sigma_quant = d.epf_quant_mul / (quant_scale * float(d.raw_quant_map[i]) * kInvSigmaNum);
    float sigma = std::min(sigma_quant * d.epf_sharp_lut[d.epf_map[i]], -1e-4f);
    s[i] = 1.0f / sigma;
  }
  return s;
}

// a14. EPF stages — render/stages/epf/{epf0,epf1,epf2,common}.rs
struct Planes {
  float* p[3];
  size_t stride;
};
void epf(int stage, const JxgFrameDesc& d, const Geometry& geo, const std::vector<float>& sigma, const Planes& in,
         const Planes& out, int num_threads) {
  const size_t w = geo.width, h = geo.height;
  float sigma_scale = stage == 0 ? d.epf_pass0_sigma_scale : stage == 1 ? 1.0f : d.epf_pass2_sigma_scale;
  const float sm = sigma_scale * 1.65f, bsm = sm * d.epf_border_sad_mul;
  // neighbour offsets in the reference's accumulation order
  static const int kOff0[12][2] = {{0, -2}, {-1, -1}, {0, -1}, {1, -1}, {-2, 0}, {-1, 0}, {1, 0}, {2, 0}, {-1, 1}, {0, 1}, {1, 1}, {0, 2}};
  static const int kOff1[4][2] = {{0, -1}, {-1, 0}, {1, 0}, {0, 1}};
  static const int kPlus[5][2] = {{0, 0}, {0, -1}, {-1, 0}, {1, 0}, {0, 1}};
  // SAD term order inside a neighbour (epf0.rs:157-168, epf1.rs:98-101): the
  // reference sums the 5 plus-shaped positions top, left, centre, right, bottom.
  static const int kPlusOrder[5][2] = {{0, -1}, {-1, 0}, {0, 0}, {1, 0}, {0, 1}};
  (void)kPlus;
  // One pixel; `at(c, x, y)` reads the input with whole-image mirroring (SimpleRenderPipeline semantics) or, for
  // pixels whose 7x7 neighbourhood lies inside the image, directly — the arithmetic and its order are the same.
  auto pixel = [&](ptrdiff_t x, ptrdiff_t y, auto&& at) __attribute__((always_inline)) {
    float inv_sigma_px = sigma[size_t(y / 8) * geo.xb + size_t(x / 8)];
    if (inv_sigma_px < kMinSigma) {
      for (int c = 0; c < 3; c++) out.p[c][size_t(y) * out.stride + size_t(x)] = at(c, x, y);
      return;
    }
    bool border = (y % 8 == 0 || y % 8 == 7) || (x % 8 == 0 || x % 8 == 7);
    float inv_s = inv_sigma_px * (border ? bsm : sm);
    if (stage == 2) {  // epf2.rs:53-125
      float cc[3] = {at(0, x, y), at(1, x, y), at(2, x, y)};
      float wacc = 1.0f, acc[3] = {cc[0], cc[1], cc[2]};
      for (auto& o : kOff1) {
        float nb[3] = {at(0, x + o[0], y + o[1]), at(1, x + o[0], y + o[1]), at(2, x + o[0], y + o[1])};
        float sad = std::fmaf(std::fabs(nb[0] - cc[0]), d.epf_channel_scale[0],
                              std::fmaf(std::fabs(nb[1] - cc[1]), d.epf_channel_scale[1],
                                        std::fabs(nb[2] - cc[2]) * d.epf_channel_scale[2]));
        float wt = std::max(std::fmaf(sad, inv_s, 1.0f), 0.0f);
        wacc += wt;
        for (int c = 0; c < 3; c++) acc[c] = std::fmaf(wt, nb[c], acc[c]);
      }
      float inv_w = 1.0f / wacc;
      for (int c = 0; c < 3; c++) out.p[c][size_t(y) * out.stride + size_t(x)] = acc[c] * inv_w;
      return;
    }
    const int n = stage == 0 ? 12 : 4;
    const int(*off)[2] = stage == 0 ? kOff0 : kOff1;
    float sads[12];
    for (int k = 0; k < n; k++) sads[k] = 0.0f;
    for (int c = 0; c < 3; c++) {
      float scale = d.epf_channel_scale[c];
      for (int k = 0; k < n; k++) {
        float s = 0.0f;
        for (auto& pl : kPlusOrder)
          s += std::fabs(at(c, x + pl[0], y + pl[1]) - at(c, x + pl[0] + off[k][0], y + pl[1] + off[k][1]));
        sads[k] = std::fmaf(scale, s, sads[k]);
      }
    }
    float wsum = 1.0f;
    for (int k = 0; k < n; k++) {
      sads[k] = std::max(std::fmaf(sads[k], inv_s, 1.0f), 0.0f);
      wsum += sads[k];
    }
    float inv_w = 1.0f / wsum;
    for (int c = 0; c < 3; c++) {
      float o = at(c, x, y);
      for (int k = n - 1; k >= 0; k--) o = std::fmaf(at(c, x + off[k][0], y + off[k][1]), sads[k], o);
      out.p[c][size_t(y) * out.stride + size_t(x)] = o * inv_w;
    }
  };
  auto at_mirror = [&](int c, ptrdiff_t xx, ptrdiff_t yy) { return in.p[c][mirror(yy, h) * in.stride + mirror(xx, w)]; };
  auto at_direct = [&](int c, ptrdiff_t xx, ptrdiff_t yy) { return in.p[c][size_t(yy) * in.stride + size_t(xx)]; };
#ifdef __AVX2__
  // Eight pixels of one 8x8 block row at a time (x0 a multiple of 8, neighbourhood inside the image): the sigma is one
  // value for the chunk, the border rule a lane mask; every lane runs the scalar sequence above.
  const bool fast = g_fast_cpu.load() != 0;
  const __m256 vabs = _mm256_castsi256_ps(_mm256_set1_epi32(0x7fffffff));
  auto chunk = [&](ptrdiff_t x0, ptrdiff_t y) {
    const float inv_sigma_px = sigma[size_t(y / 8) * geo.xb + size_t(x0 / 8)];
    if (inv_sigma_px < kMinSigma) {
      for (int c = 0; c < 3; c++) _mm256_storeu_ps(out.p[c] + size_t(y) * out.stride + x0, _mm256_loadu_ps(in.p[c] + size_t(y) * in.stride + x0));
      return;
    }
    const bool rowb = (y % 8 == 0 || y % 8 == 7);
    const __m256 mul = rowb ? _mm256_set1_ps(bsm) : _mm256_setr_ps(bsm, sm, sm, sm, sm, sm, sm, bsm);
    const __m256 inv_s = _mm256_mul_ps(_mm256_set1_ps(inv_sigma_px), mul);
    auto ld = [&](int c, ptrdiff_t dx, ptrdiff_t dy) { return _mm256_loadu_ps(in.p[c] + size_t(y + dy) * in.stride + size_t(x0 + dx)); };
    const __m256 one = _mm256_set1_ps(1.0f), zero = _mm256_setzero_ps();
    if (stage == 2) {
      const __m256 cc[3] = {ld(0, 0, 0), ld(1, 0, 0), ld(2, 0, 0)};
      __m256 wacc = one, acc[3] = {cc[0], cc[1], cc[2]};
      for (auto& o : kOff1) {
        const __m256 nb[3] = {ld(0, o[0], o[1]), ld(1, o[0], o[1]), ld(2, o[0], o[1])};
        const __m256 a0 = _mm256_and_ps(_mm256_sub_ps(nb[0], cc[0]), vabs), a1 = _mm256_and_ps(_mm256_sub_ps(nb[1], cc[1]), vabs),
                     a2 = _mm256_and_ps(_mm256_sub_ps(nb[2], cc[2]), vabs);
        const __m256 sad = _mm256_fmadd_ps(a0, _mm256_set1_ps(d.epf_channel_scale[0]),
                                           _mm256_fmadd_ps(a1, _mm256_set1_ps(d.epf_channel_scale[1]),
                                                           _mm256_mul_ps(a2, _mm256_set1_ps(d.epf_channel_scale[2]))));
        const __m256 wt = _mm256_max_ps(_mm256_fmadd_ps(sad, inv_s, one), zero);
        wacc = _mm256_add_ps(wacc, wt);
        for (int c = 0; c < 3; c++) acc[c] = _mm256_fmadd_ps(wt, nb[c], acc[c]);
      }
      const __m256 inv_w = _mm256_div_ps(one, wacc);
      for (int c = 0; c < 3; c++) _mm256_storeu_ps(out.p[c] + size_t(y) * out.stride + x0, _mm256_mul_ps(acc[c], inv_w));
      return;
    }
    const int nn = stage == 0 ? 12 : 4;  // stage 0: the 12 neighbours of epf0.rs:182-195, same sequence as the scalar form
    const int(*offv)[2] = stage == 0 ? kOff0 : kOff1;
    __m256 sads[12];
    for (int k = 0; k < nn; k++) sads[k] = zero;
    for (int c = 0; c < 3; c++) {
      const __m256 scale = _mm256_set1_ps(d.epf_channel_scale[c]);
      for (int k = 0; k < nn; k++) {
        __m256 sacc = zero;
        for (auto& pl : kPlusOrder)
          sacc = _mm256_add_ps(sacc, _mm256_and_ps(_mm256_sub_ps(ld(c, pl[0], pl[1]), ld(c, pl[0] + offv[k][0], pl[1] + offv[k][1])), vabs));
        sads[k] = _mm256_fmadd_ps(scale, sacc, sads[k]);
      }
    }
    __m256 wsum = one;
    for (int k = 0; k < nn; k++) {
      sads[k] = _mm256_max_ps(_mm256_fmadd_ps(sads[k], inv_s, one), zero);
      wsum = _mm256_add_ps(wsum, sads[k]);
    }
    const __m256 inv_w = _mm256_div_ps(one, wsum);
    for (int c = 0; c < 3; c++) {
      __m256 o = ld(c, 0, 0);
      for (int k = nn - 1; k >= 0; k--) o = _mm256_fmadd_ps(ld(c, offv[k][0], offv[k][1]), sads[k], o);
      _mm256_storeu_ps(out.p[c] + size_t(y) * out.stride + x0, _mm256_mul_ps(o, inv_w));
    }
  };
#else
  const bool fast = false;
  auto chunk = [&](ptrdiff_t, ptrdiff_t) {};
#endif
  parallel_for(int(h), num_threads, [&](int yi) {
    const ptrdiff_t y = yi;
    const bool row_inside = y >= 3 && y + 3 < ptrdiff_t(h);
    for (ptrdiff_t x = 0; x < ptrdiff_t(w); x++) {
      if (fast && row_inside && (x & 7) == 0 && x >= 8 && x + 8 + 3 <= ptrdiff_t(w)) {
        chunk(x, y);
        x += 7;
        continue;
      }
      if (row_inside && x >= 3 && x + 3 < ptrdiff_t(w)) pixel(x, y, at_direct);
      else pixel(x, y, at_mirror);
    }
  });
}

// a15. render/stages/xyb.rs:145-241
inline void xyb_to_linear_px(float& x, float& y, float& b, const float* mat, const float* bias_cbrt,
                             const float* scaled_bias, float intensity_scale) {
  float l = y + x - bias_cbrt[0], m = y - x - bias_cbrt[1], s = b - bias_cbrt[2];
  float l2 = l * l, m2 = m * m, s2 = s * s;
  float sl = l * intensity_scale, smm = m * intensity_scale, ss = s * intensity_scale;
  l = std::fmaf(l2, sl, scaled_bias[0]);
  m = std::fmaf(m2, smm, scaled_bias[1]);
  s = std::fmaf(s2, ss, scaled_bias[2]);
  x = std::fmaf(mat[0], l, std::fmaf(mat[1], m, mat[2] * s));
  y = std::fmaf(mat[3], l, std::fmaf(mat[4], m, mat[5] * s));
  b = std::fmaf(mat[6], l, std::fmaf(mat[7], m, mat[8] * s));
}
// a16. color/tf.rs:13-44 + util/rational_poly.rs:20-35
inline float linear_to_srgb(float v) {
  const float P[5] = {-5.135152395e-4f, 5.287254571e-3f, 3.903842876e-1f, 1.474205315f, 7.352629620e-1f};
  const float Q[5] = {1.004519624e-2f, 3.036675394e-1f, 1.340816930f, 9.258482155e-1f, 2.424867759e-2f};
  float a = std::fabs(v), r;
  if (a < 0.0031308f) {
    r = a * 12.92f;
  } else {
    float s = std::sqrt(a);
    float yp = P[4], yq = Q[4];
    for (int i = 3; i >= 0; i--) {
      yp = std::fmaf(yp, s, P[i]);
      yq = std::fmaf(yq, s, Q[i]);
    }
    r = yp / yq;
  }
  return std::copysign(r, v);
}
// util/rational_poly.rs:20-35 (SIMD form: mul_add chains)
inline float rational_poly5(float x, const float* p, const float* q) {
  float yp = p[4], yq = q[4];
  for (int i = 3; i >= 0; i--) {
    yp = std::fmaf(yp, x, p[i]);
    yq = std::fmaf(yq, x, q[i]);
  }
  return yp / yq;
}
// util/fast_math.rs:98-114 (fast_pow2f_simd), :140-149 (fast_log2f_simd), :158-160
inline float fast_pow2f(float x) {
  const float kNum[3] = {1.01749063e1f, 4.88687798e1f, 9.85506591e1f};
  const float kDen[4] = {2.10242958e-1f, -2.22328856e-2f, -1.94414990e1f, 9.85506633e1f};
  const float x_floor = std::floor(x);
  const uint32_t ebits = uint32_t(int32_t(x_floor) + 127) << 23;
  float exp;
  memcpy(&exp, &ebits, 4);
  const float frac = x - x_floor;
  float num = frac + kNum[0];
  num = std::fmaf(num, frac, kNum[1]);
  num = std::fmaf(num, frac, kNum[2]);
  num = num * exp;
  float den = std::fmaf(kDen[0], frac, kDen[1]);
  den = std::fmaf(den, frac, kDen[2]);
  den = std::fmaf(den, frac, kDen[3]);
  return num / den;
}
inline float fast_log2f(float x) {
  const float kP[3] = {-1.8503833400518310e-6f, 1.4287160470083755f, 7.4245873327820566e-1f};
  const float kQ[3] = {9.9032814277590719e-1f, 1.0096718572241148f, 1.7409343003366853e-1f};
  int32_t x_bits;
  memcpy(&x_bits, &x, 4);
  const int32_t exp_bits = int32_t(uint32_t(x_bits) - 0x3f2aaaabu);
  const int32_t exp_shifted = exp_bits >> 23;
  const uint32_t mbits = uint32_t(x_bits) - (uint32_t(exp_shifted) << 23);
  float mantissa;
  memcpy(&mantissa, &mbits, 4);
  const float m1 = mantissa - 1.0f;
  const float yp = std::fmaf(std::fmaf(kP[2], m1, kP[1]), m1, kP[0]);
  const float yq = std::fmaf(std::fmaf(kQ[2], m1, kQ[1]), m1, kQ[0]);
  return yp / yq + float(exp_shifted);
}
inline float fast_powf(float base, float e) { return fast_pow2f(fast_log2f(base) * e); }

// a16 (other encodings). render/stages/from_linear.rs:56-112 on one RGB triple.
inline void from_linear_other(const JxgFrameDesc& d, float* v) {
  switch (d.output_tf) {
    case JXG_TF_GAMMA:  // from_linear.rs:97-109
      for (int c = 0; c < 3; c++) v[c] = std::copysign(fast_powf(std::fabs(v[c]), d.output_gamma), v[c]);
      break;
    case JXG_TF_BT709: {  // color/tf.rs:114-150
      const float P[5] = {-9.625309705734253e-2f, -2.2635456919670105e-1f, 1.935774803161621e1f, 5.897886276245117e1f, 2.3947298049926758e1f};
      const float Q[5] = {1.0f, 1.877663230895996e1f, 5.5292449951171875e1f, 2.6565317153930664e1f, 3.269049823284149e-1f};
      for (int c = 0; c < 3; c++) {
        const float a = std::fabs(v[c]);
        v[c] = std::copysign(0.018f > a ? a * 4.5f : rational_poly5(std::sqrt(a), P, Q), v[c]);
      }
      break;
    }
    case JXG_TF_PQ: {  // color/tf.rs:236-304
      const float P[5] = {1.351392e-2f, -1.095778f, 5.522776e1f, 1.492516e2f, 4.838434e1f};
      const float Q[5] = {1.012416f, 2.016708e1f, 9.26371e1f, 1.120607e2f, 2.590418e1f};
      const float PS[5] = {9.863406e-6f, 3.881234e-1f, 1.352821e2f, 6.889862e4f, -2.864824e5f};
      const float QS[5] = {3.371868e1f, 1.477719e3f, 1.608477e4f, -4.389884e4f, -2.072546e5f};
      const float y_mult = d.intensity_target * (1.0f / 10000.0f);
      for (int c = 0; c < 3; c++) {
        const float a = std::fabs(v[c]);
        const float a14 = std::sqrt(std::sqrt(a * y_mult));
        v[c] = std::copysign(1e-4f > a ? rational_poly5(a14, PS, QS) : rational_poly5(a14, P, Q), v[c]);
      }
      break;
    }
    case JXG_TF_HLG: {  // color/tf.rs:381-395, 458-470, 481-497
      const float system_gamma = 1.2f * std::pow(1.111f, std::log2(d.intensity_target / 1e3f));
      const float e = (1.0f - system_gamma) / system_gamma;
      if (!(std::fabs(e) < 0.1f)) {
        const float mixed = std::fmaf(v[0], d.output_luminances[0], std::fmaf(v[1], d.output_luminances[1], v[2] * d.output_luminances[2]));
        const float mult = fast_powf(mixed, e);
        for (int c = 0; c < 3; c++) v[c] *= mult;
      }
      const double kA = 0.17883277, kB = 1.0 - 4.0 * kA, kC = 0.5599107295;
      for (int c = 0; c < 3; c++) {
        const float a = std::fabs(v[c]);
        float y;
        if (a <= 1.0f / 12.0f) y = std::sqrt(3.0f * a);
        else y = float(kA * 0.693147180559945309417) * fast_log2f(12.0f * a - float(kB)) + float(kC);
        v[c] = std::copysign(y, v[c]);
      }
      break;
    }
    default: break;  // JXG_TF_LINEAR: no stage (frame/render.rs:761)
  }
}
// util/float16.rs:82-141 f16::from_f32: round to nearest even for normal halves, TRUNCATION into the subnormal range
// (2^-24 <= |f| < 2^-14; the reference shifts one bit too far there, so those values also come out halved - kept),
// zero below it and for f32 subnormals, infinity above 65504 + half an ulp.
static uint16_t f32_to_f16_bits(float f) {
  uint32_t bits;
  memcpy(&bits, &f, 4);
  const uint16_t sign = uint16_t((bits >> 31) & 1);
  const int exp = int((bits >> 23) & 0xff);
  const uint32_t mant = bits & 0x007fffffu;
  if (exp == 0) return uint16_t(sign << 15);
  if (exp == 255) return uint16_t((sign << 15) | (0x1f << 10) | (mant ? 0x0200 : 0));
  const int unbiased = exp - 127;
  if (unbiased < -24) return uint16_t(sign << 15);
  if (unbiased < -14) return uint16_t((sign << 15) | uint16_t((mant | 0x00800000u) >> (uint32_t(-14 - unbiased) + 14)));
  if (unbiased > 15) return uint16_t((sign << 15) | (0x1f << 10));
  const uint16_t h_exp = uint16_t(unbiased + 15);
  uint16_t h_mant = uint16_t(mant >> 13);
  const uint32_t round_bit = (mant >> 12) & 1, sticky = mant & 0x0fff;
  if (round_bit && (sticky || (h_mant & 1))) h_mant++;
  if (h_mant > 0x3ff) return h_exp >= 30 ? uint16_t((sign << 15) | (0x1f << 10)) : uint16_t((sign << 15) | ((h_exp + 1) << 10));
  return uint16_t((sign << 15) | (h_exp << 10) | h_mant);
}

const float kDither[32 * 32] = {
#include "dither_table.inc"

};

}  // namespace

// ---------------------------------------------------------------------------
// C API
// ---------------------------------------------------------------------------
extern "C" {

const char* jxo_last_error(void) { return g_last_error.c_str(); }

int jxo_decode_frame(const JxgFrameDesc* desc, const uint8_t* hf_bytes, const uint64_t* sec_off,
                     const uint32_t* sec_len, uint32_t n_sections, void* out, size_t out_row_stride,
                     const JxoTaps* taps, int num_threads, uint32_t* bad_group) {
  const JxgFrameDesc& d = *desc;
  Geometry geo(d);
  if (n_sections != geo.num_groups * d.num_passes) return JXG_ERR_ARGUMENT;
  if (num_threads <= 0) num_threads = int(std::thread::hardware_concurrency());
  static const bool timing = getenv("JXO_TIMING") != nullptr;  // stage split of the CPU port on stderr
  auto tnow = [] { return std::chrono::steady_clock::now(); };
  auto t_begin = tnow();
  auto lap = [&](const char* what) {
    if (!timing) return;
    auto t = tnow();
    fprintf(stderr, "[oracle] %s %.1f ms\n", what, std::chrono::duration<double, std::milli>(t - t_begin).count());
    t_begin = t;
  };
  const size_t pstride = size_t(geo.xb) * 8, prow = size_t(geo.yb) * 8;
  // Fast CPU forms: the two plane sets come from a per-thread pool (a frame-parallel caller decodes many frames per
  // thread; 200 MB of fresh zero pages per 4K frame cost ~15 % of the decode). Every plane element that is read is
  // written first (the IDCTs fill whole padded blocks, the filters read inside the image only), so no clearing is needed.
  const bool pooled = g_fast_cpu.load() != 0;
  static thread_local std::vector<float> pool_a[3], pool_b[3];
  std::vector<float> local_a[3], local_b[3];
  std::vector<float>* plane_a = pooled ? pool_a : local_a;
  std::vector<float>* plane_b = pooled ? pool_b : local_b;
  for (int c = 0; c < 3; c++) {
    if (pooled) {
      if (plane_a[c].size() < pstride * prow) plane_a[c].resize(pstride * prow);
    } else {
      plane_a[c].assign(pstride * prow, 0.0f);
    }
  }
  float* planes[3] = {plane_a[0].data(), plane_a[1].data(), plane_a[2].data()};
  std::atomic<int> err{0};
  std::atomic<uint32_t> errg{0};
  // per-group entropy decode + dequant + IDCT (frame/render.rs:461-479 fan-out)
  std::vector<std::vector<int32_t>> scratch(size_t(std::max(1, num_threads)));
  std::atomic<int> slot{0};
  parallel_for(int(geo.num_groups), num_threads, [&](int g) {
    thread_local std::vector<int32_t> local;
    int32_t* co;
    if (taps && taps->coeffs) co = taps->coeffs + size_t(g) * 3 * 65536;
    else {
      local.resize(3 * 65536);
      co = local.data();
    }
    int r = decode_group(d, geo, uint32_t(g), hf_bytes, sec_off, sec_len, co, planes, pstride);
    if (r != 0) {
      int expected = 0;
      if (err.compare_exchange_strong(expected, r)) errg = uint32_t(g);
    }
  });
  (void)slot;
  if (err.load()) {
    if (bad_group) *bad_group = errg.load();
    return err.load();
  }
  lap("entropy + dequant + IDCT");
  if (taps && taps->xyb_idct)
    for (int c = 0; c < 3; c++) memcpy(taps->xyb_idct + size_t(c) * pstride * prow, planes[c], pstride * prow * sizeof(float));

  // render stages (frame/render.rs:579-620)
  for (int c = 0; c < 3; c++) {
    if (pooled) {
      if (plane_b[c].size() < pstride * prow) plane_b[c].resize(pstride * prow);
    } else {
      plane_b[c].assign(pstride * prow, 0.0f);
    }
  }
  Planes cur{{plane_a[0].data(), plane_a[1].data(), plane_a[2].data()}, pstride};
  Planes nxt{{plane_b[0].data(), plane_b[1].data(), plane_b[2].data()}, pstride};
  if (d.gab) {
    for (int c = 0; c < 3; c++)
      gaborish(geo.width, geo.height, cur.p[c], pstride, nxt.p[c], pstride, d.gab_w1[c], d.gab_w2[c], num_threads);
    std::swap(cur, nxt);
  }
  lap("gaborish");
  if (d.epf_iters > 0) {
    std::vector<float> sigma = sigma_image(d, geo);
    if (d.epf_iters >= 3) {
      epf(0, d, geo, sigma, cur, nxt, num_threads);
      std::swap(cur, nxt);
    }
    epf(1, d, geo, sigma, cur, nxt, num_threads);
    std::swap(cur, nxt);
    if (d.epf_iters >= 2) {
      epf(2, d, geo, sigma, cur, nxt, num_threads);
      std::swap(cur, nxt);
    }
  }
  lap("epf");
  if (taps && taps->xyb_filtered)
    for (int c = 0; c < 3; c++)
      for (size_t y = 0; y < geo.height; y++)
        memcpy(taps->xyb_filtered + (size_t(c) * geo.height + y) * geo.width, cur.p[c] + y * pstride, geo.width * sizeof(float));
  if (!out) return 0;
  if (d.output_format == JXG_FORMAT_XYB_F32_PLANAR) {
    for (int c = 0; c < 3; c++)
      for (size_t y = 0; y < geo.height; y++)
        memcpy(static_cast<uint8_t*>(out) + (size_t(c) * geo.height + y) * out_row_stride, cur.p[c] + y * pstride, geo.width * sizeof(float));
    return 0;
  }
  // XYB -> linear -> (sRGB) -> output (xyb.rs, from_linear.rs, convert.rs:574-598, save)
  float bias_cbrt[3], scaled_bias[3];
  float intensity_scale = 255.0f / d.intensity_target;
  for (int i = 0; i < 3; i++) {
    bias_cbrt[i] = std::cbrt(d.opsin_biases[i]);
    scaled_bias[i] = d.opsin_biases[i] * intensity_scale;
  }
  // Orientation (render/save.rs + headers/image_metadata.rs:85-96): the coded image is produced into a tight staging
  // image first and every pixel moved to display_pixel(x, y) afterwards.
  const uint32_t orientation = d.orientation == 0 ? 1u : d.orientation;
  const bool fmt16 = d.output_format == JXG_FORMAT_RGB_U16 || d.output_format == JXG_FORMAT_RGB_F16;
  const size_t obpp = d.output_format == JXG_FORMAT_RGB_F32 ? 12 : (d.output_format == JXG_FORMAT_RGBA_U8 ? 4 : (fmt16 ? 6 : 3));
  std::vector<uint8_t> staging;
  uint8_t* obase = static_cast<uint8_t*>(out);
  size_t ostride = out_row_stride;
  if (orientation != 1) {
    staging.resize(geo.width * geo.height * obpp);
    obase = staging.data();
    ostride = geo.width * obpp;
  }
  const bool fast_store = g_fast_cpu.load() != 0 && d.output_tf == JXG_TF_SRGB &&
                          (d.output_format == JXG_FORMAT_RGB_U8 || d.output_format == JXG_FORMAT_RGBA_U8);
  parallel_for(int(geo.height), num_threads, [&](int yi) {
    size_t y = size_t(yi);
    uint8_t* row = obase + y * ostride;
    size_t x_begin = 0;
#ifdef __AVX2__
    if (fast_store) {  // eight pixels per step through the same per-pixel sequence (xyb.rs:197-241, tf.rs:13-44, convert.rs:574-598)
      const int nc = d.output_format == JXG_FORMAT_RGBA_U8 ? 4 : 3;
      const __m256 vabs = _mm256_castsi256_ps(_mm256_set1_epi32(0x7fffffff)), vsign = _mm256_castsi256_ps(_mm256_set1_epi32(int(0x80000000u)));
      const float* m = d.opsin_inverse_matrix;
      const __m256 is = _mm256_set1_ps(intensity_scale);
      static const float P[5] = {-5.135152395e-4f, 5.287254571e-3f, 3.903842876e-1f, 1.474205315f, 7.352629620e-1f};
      static const float Q[5] = {1.004519624e-2f, 3.036675394e-1f, 1.340816930f, 9.258482155e-1f, 2.424867759e-2f};
      for (; x_begin + 8 <= geo.width; x_begin += 8) {
        const size_t x = x_begin;
        const __m256 vx = _mm256_loadu_ps(cur.p[0] + y * pstride + x), vy = _mm256_loadu_ps(cur.p[1] + y * pstride + x),
                     vb = _mm256_loadu_ps(cur.p[2] + y * pstride + x);
        __m256 l = _mm256_sub_ps(_mm256_add_ps(vy, vx), _mm256_set1_ps(bias_cbrt[0]));
        __m256 mm = _mm256_sub_ps(_mm256_sub_ps(vy, vx), _mm256_set1_ps(bias_cbrt[1]));
        __m256 sb = _mm256_sub_ps(vb, _mm256_set1_ps(bias_cbrt[2]));
        const __m256 l2 = _mm256_mul_ps(l, l), m2 = _mm256_mul_ps(mm, mm), s2 = _mm256_mul_ps(sb, sb);
        l = _mm256_fmadd_ps(l2, _mm256_mul_ps(l, is), _mm256_set1_ps(scaled_bias[0]));
        mm = _mm256_fmadd_ps(m2, _mm256_mul_ps(mm, is), _mm256_set1_ps(scaled_bias[1]));
        sb = _mm256_fmadd_ps(s2, _mm256_mul_ps(sb, is), _mm256_set1_ps(scaled_bias[2]));
        __m256 rgb[3];
        for (int c = 0; c < 3; c++)
          rgb[c] = _mm256_fmadd_ps(_mm256_set1_ps(m[3 * c]), l,
                                   _mm256_fmadd_ps(_mm256_set1_ps(m[3 * c + 1]), mm, _mm256_mul_ps(_mm256_set1_ps(m[3 * c + 2]), sb)));
        alignas(32) float px[3][8];
        for (int c = 0; c < 3; c++) {
          const __m256 a = _mm256_and_ps(rgb[c], vabs);
          const __m256 lin = _mm256_mul_ps(a, _mm256_set1_ps(12.92f));
          const __m256 sq = _mm256_sqrt_ps(a);
          __m256 yp = _mm256_set1_ps(P[4]), yq = _mm256_set1_ps(Q[4]);
          for (int i = 3; i >= 0; i--) {
            yp = _mm256_fmadd_ps(yp, sq, _mm256_set1_ps(P[i]));
            yq = _mm256_fmadd_ps(yq, sq, _mm256_set1_ps(Q[i]));
          }
          __m256 r = _mm256_blendv_ps(_mm256_div_ps(yp, yq), lin, _mm256_cmp_ps(a, _mm256_set1_ps(0.0031308f), _CMP_LT_OQ));
          r = _mm256_or_ps(r, _mm256_and_ps(rgb[c], vsign));  // copysign: r is non-negative
          alignas(32) float dth[8];
          for (int i = 0; i < 8; i++) dth[i] = kDither[((y + 13 * size_t(c)) % 32) * 32 + (x + size_t(i) + 23 * size_t(c)) % 32];
          __m256 sc = _mm256_add_ps(_mm256_mul_ps(r, _mm256_set1_ps(255.0f)), _mm256_load_ps(dth));
          sc = _mm256_min_ps(_mm256_max_ps(sc, _mm256_setzero_ps()), _mm256_set1_ps(255.0f));
          _mm256_store_ps(px[c], _mm256_round_ps(sc, _MM_FROUND_TO_NEAREST_INT | _MM_FROUND_NO_EXC));
        }
        for (int i = 0; i < 8; i++) {
          uint8_t* o = row + (x + size_t(i)) * size_t(nc);
          o[0] = uint8_t(px[0][i]);
          o[1] = uint8_t(px[1][i]);
          o[2] = uint8_t(px[2][i]);
          if (nc == 4) o[3] = 255;
        }
      }
    }
#else
    (void)fast_store;
#endif
    for (size_t x = x_begin; x < geo.width; x++) {
      float v[3] = {cur.p[0][y * pstride + x], cur.p[1][y * pstride + x], cur.p[2][y * pstride + x]};
      xyb_to_linear_px(v[0], v[1], v[2], d.opsin_inverse_matrix, bias_cbrt, scaled_bias, intensity_scale);
      if (d.output_tf == JXG_TF_SRGB)
        for (float& f : v) f = linear_to_srgb(f);
      else
        from_linear_other(d, v);
      if (d.output_format == JXG_FORMAT_RGB_F32) {
        memcpy(row + x * 12, v, 12);
      } else if (d.output_format == JXG_FORMAT_RGB_U16) {  // convert.rs:739-762, bit depth 16
        uint16_t px[3];
        for (int c = 0; c < 3; c++) px[c] = uint16_t(std::nearbyint(std::min(std::max(v[c], 0.0f), 1.0f) * 65535.0f));
        memcpy(row + x * 6, px, 6);
      } else if (d.output_format == JXG_FORMAT_RGB_F16) {  // convert.rs:831-857; clamp ranges frame/render.rs:746-750
        uint16_t px[3];
        for (int c = 0; c < 3; c++) {
          float f = v[c];
          if (d.output_tf == JXG_TF_PQ) f = std::min(std::max(f, 0.0f), 1.0f);
          else if (d.output_tf == JXG_TF_HLG) f = std::min(std::max(f, -0.074f), 1.1f);
          px[c] = f32_to_f16_bits(f);
        }
        memcpy(row + x * 6, px, 6);
      } else {
        int nc = d.output_format == JXG_FORMAT_RGBA_U8 ? 4 : 3;
        for (int c = 0; c < 3; c++) {
          float dv = kDither[((y + 13 * size_t(c)) % 32) * 32 + (x + 23 * size_t(c)) % 32];
          float s = v[c] * 255.0f + dv;
          s = std::min(std::max(s, 0.0f), 255.0f);
          row[x * nc + c] = uint8_t(std::nearbyint(s));  // round-to-nearest-even (AVX store path, avx.rs:609)
        }
        if (nc == 4) row[x * 4 + 3] = 255;
      }
    }
  });
  lap("colour + store");
  if (orientation != 1)
    jxo_orient_image(staging.data(), geo.width, geo.height, obpp, orientation, static_cast<uint8_t*>(out), out_row_stride);
  return 0;
}

int jxo_file_info(const uint8_t* data, size_t size, JxgImageInfo* info) {
  try {
    auto fs = jxg::parse_vardct_file(data, size);
    info->coded_width = fs->header.xsize();
    info->coded_height = fs->header.ysize();
    info->orientation = fs->file.orientation;
    info->width = fs->file.orientation >= 5 ? info->coded_height : info->coded_width;
    info->height = fs->file.orientation >= 5 ? info->coded_width : info->coded_height;
    info->num_groups = fs->header.num_groups();
    info->num_passes = fs->header.passes.num_passes;
    info->encoding = 0;
    info->hf_bytes = 0;
    for (auto l : fs->hf_len) info->hf_bytes += l;
    return 0;
  } catch (jxg::Error& e) {
    g_last_error = e.what();
    return e.code;
  }
}

int jxo_decode_file(const uint8_t* data, size_t size, uint32_t output_format, void* out, size_t out_row_stride,
                    const JxoTaps* taps, int num_threads) {
  try {
    auto fs = jxg::parse_vardct_file(data, size);
    JxgFrameDesc d;
    fs->fill_desc(&d, output_format);
    uint32_t bad = 0;
    int r = jxo_decode_frame(&d, fs->codestream.data(), fs->hf_off.data(), fs->hf_len.data(), uint32_t(fs->hf_off.size()),
                             out, out_row_stride, taps, num_threads, &bad);
    if (r) g_last_error = "hot-path error " + std::to_string(r) + " in group " + std::to_string(bad);
    return r;
  } catch (jxg::Error& e) {
    g_last_error = e.what();
    return e.code;
  }
}

void jxo_orient_image(const uint8_t* src, size_t w, size_t h, size_t bpp, uint32_t orientation, uint8_t* dst, size_t dst_stride) {
  for (size_t y = 0; y < h; y++)
    for (size_t x = 0; x < w; x++) {
      size_t dx = x, dy = y;
      switch (orientation) {
        case 2: dx = w - 1 - x; break;                  // FlipHorizontal
        case 3: dx = w - 1 - x; dy = h - 1 - y; break;  // Rotate180
        case 4: dy = h - 1 - y; break;                  // FlipVertical
        case 5: dx = y; dy = x; break;                  // Transpose
        case 6: dx = h - 1 - y; dy = x; break;          // Rotate90Cw
        case 7: dx = h - 1 - y; dy = w - 1 - x; break;  // AntiTranspose
        case 8: dx = y; dy = w - 1 - x; break;          // Rotate90Ccw
        default: break;
      }
      memcpy(dst + dy * dst_stride + dx * bpp, src + (y * w + x) * bpp, bpp);
    }
}

void jxo_idct2d(int rows, int cols, float* block) { idct2d(rows, cols, block); }
void jxo_reinterpreting_dct2d(int rows, int cols, const float* in, float* out, int out_stride) {
  reinterpreting_dct2d(rows, cols, in, out, out_stride);
}
void jxo_transform_to_pixels(int transform, const float* lf, float* buf) { transform_to_pixels(transform, lf, buf); }
void jxo_gaborish(int w, int h, const float* in, float* out, float w1, float w2) {
  gaborish(size_t(w), size_t(h), in, size_t(w), out, size_t(w), w1, w2, 1);
}
void jxo_xyb_to_linear(int n, float* x, float* y, float* b, const float* opsin_matrix, const float* opsin_biases,
                       float intensity_target) {
  float bias_cbrt[3], scaled_bias[3], is = 255.0f / intensity_target;
  for (int i = 0; i < 3; i++) {
    bias_cbrt[i] = std::cbrt(opsin_biases[i]);
    scaled_bias[i] = opsin_biases[i] * is;
  }
  for (int i = 0; i < n; i++) xyb_to_linear_px(x[i], y[i], b[i], opsin_matrix, bias_cbrt, scaled_bias, is);
}
// Float stages on caller-supplied data, for the independent float64 restatements of tests/test_kat_float_stages.py.
void jxo_dequant_block(uint32_t num_coeffs, const int32_t* qx, const int32_t* qy, const int32_t* qb, const float* mat, float sx,
                       float sy, float sb, float x_cc, float b_cc, const float* bias, float* out3) {
  dequant_block(num_coeffs, qx, qy, qb, mat, sx, sy, sb, x_cc, b_cc, bias, out3, out3 + num_coeffs, out3 + 2 * size_t(num_coeffs));
}
// inv-sigma image (features/epf.rs:35-86) of an xb x yb block grid.
void jxo_sigma_image(uint32_t xb, uint32_t yb, uint32_t global_scale, const int32_t* raw_quant, const uint8_t* sharpness,
                     float quant_mul, const float* sharp_lut, float* out) {
  JxgFrameDesc d;
  memset(&d, 0, sizeof(d));
  d.global_scale = global_scale;
  d.raw_quant_map = raw_quant;
  d.epf_map = sharpness;
  d.epf_quant_mul = quant_mul;
  memcpy(d.epf_sharp_lut, sharp_lut, sizeof(d.epf_sharp_lut));
  d.width = xb * 8;
  d.height = yb * 8;
  Geometry geo(d);
  std::vector<float> s = sigma_image(d, geo);
  memcpy(out, s.data(), s.size() * 4);
}
// One EPF stage (0, 1, 2) over a w x h image of 3 planes (tight rows) with whole-image mirroring; inv_sigma: one value per
// 8x8 block (ceil(w/8) x ceil(h/8)).
void jxo_epf_stage(int stage, uint32_t w, uint32_t h, const float* in3, float* out3, const float* inv_sigma, const float* channel_scale,
                   float pass0_sigma_scale, float pass2_sigma_scale, float border_sad_mul, int threads) {
  JxgFrameDesc d;
  memset(&d, 0, sizeof(d));
  memcpy(d.epf_channel_scale, channel_scale, 12);
  d.epf_pass0_sigma_scale = pass0_sigma_scale;
  d.epf_pass2_sigma_scale = pass2_sigma_scale;
  d.epf_border_sad_mul = border_sad_mul;
  d.width = w;
  d.height = h;
  Geometry geo(d);
  std::vector<float> sigma(inv_sigma, inv_sigma + size_t(geo.xb) * geo.yb);
  const size_t n = size_t(w) * h;
  Planes in{{const_cast<float*>(in3), const_cast<float*>(in3) + n, const_cast<float*>(in3) + 2 * n}, w};
  Planes out{{out3, out3 + n, out3 + 2 * n}, w};
  epf(stage, d, geo, sigma, in, out, threads);
}
// from_linear stage of the oracle on n RGB triples (interleaved), for the known-answer tests of the other encodings.
void jxo_from_linear(uint32_t tf, float gamma, float intensity_target, const float* luminances, int n, float* rgb) {
  JxgFrameDesc d;
  memset(&d, 0, sizeof(d));
  d.output_tf = tf;
  d.output_gamma = gamma;
  d.intensity_target = intensity_target;
  memcpy(d.output_luminances, luminances, 12);
  for (int i = 0; i < n; i++) {
    if (tf == JXG_TF_SRGB)
      for (int c = 0; c < 3; c++) rgb[3 * i + c] = linear_to_srgb(rgb[3 * i + c]);
    else
      from_linear_other(d, rgb + 3 * i);
  }
}
// 1: run Gaborish / EPF 1-2 / the sRGB u8 store through their AVX2 forms (same results, see g_fast_cpu). Process-wide.
void jxo_set_fast_cpu(int on) { g_fast_cpu.store(on); }
// The oracle's f32 -> f16 conversion on n values (known-answer test of the F16 store).
void jxo_f32_to_f16(int n, const float* v, uint16_t* out) {
  for (int i = 0; i < n; i++) out[i] = f32_to_f16_bits(v[i]);
}
void jxo_linear_to_srgb(int n, float* v) {
  for (int i = 0; i < n; i++) v[i] = linear_to_srgb(v[i]);
}

}  // extern "C"

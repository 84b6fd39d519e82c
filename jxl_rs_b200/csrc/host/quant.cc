// Dequantisation matrices (library defaults + custom encodings), transform
// geometry tables and natural coefficient orders.
//
// Reference: jxl/src/frame/quant_weights.rs:29-1180, jxl_transforms/src/
// transform_map.rs:12-116, jxl/src/frame/coeff_order.rs:23-120.
#include <cmath>
#include <mutex>

#include "frame.h"
#include "quant.h"

namespace jxg {

#include "quant_params.inc"

// transform_map.rs:87-116
const uint8_t kCoveredBlocksX[27] = {1, 1, 1, 1, 2, 4, 1, 2, 1, 4, 2, 4, 1, 1, 1, 1, 1, 1, 8, 4, 8, 16, 8, 16, 32, 16, 32};
const uint8_t kCoveredBlocksY[27] = {1, 1, 1, 1, 2, 4, 2, 1, 4, 1, 4, 2, 1, 1, 1, 1, 1, 1, 8, 8, 4, 16, 16, 8, 32, 32, 16};
const uint8_t kBlockShapeId[27] = {0, 1, 1, 1, 2, 3, 4, 4, 5, 5, 6, 6, 1, 1, 1, 1, 1, 1, 7, 8, 8, 9, 10, 10, 11, 12, 12};
// coeff_order.rs:23-37 (TRANSFORM_TYPE_LUT as HfTransformType ids)
const uint8_t kOrderTransform[kNumOrders] = {0, 1, 4, 5, 7, 9, 11, 18, 20, 21, 23, 24, 26};
// quant_weights.rs:1146-1150
const uint8_t kQuantTableRows[kNumQuantTables] = {1, 1, 1, 1, 2, 4, 1, 1, 2, 1, 1, 8, 4, 16, 8, 32, 16};
const uint8_t kQuantTableCols[kNumQuantTables] = {1, 1, 1, 1, 2, 4, 2, 4, 4, 1, 1, 8, 8, 16, 16, 32, 32};

int quant_table_for_transform(int t) {  // quant_weights.rs:311-336
  static const uint8_t lut[27] = {0, 1, 2, 3, 4, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 10, 10, 11, 12, 12, 13, 14, 14, 15, 16, 16};
  return lut[t];
}

const std::vector<uint32_t>& natural_coeff_order_cached(int order_idx) {
  // computed once per process (the 256x256 order alone is 65 536 entries; frames with custom orders start from all 39)
  static std::vector<uint32_t> cache[kNumOrders];
  static std::once_flag once[kNumOrders];
  std::call_once(once[order_idx], [order_idx] { cache[order_idx] = natural_coeff_order(order_idx); });
  return cache[order_idx];
}

std::vector<uint32_t> natural_coeff_order(int order_idx) {  // coeff_order.rs:66-120
  int t = kOrderTransform[order_idx];
  size_t cx = kCoveredBlocksX[t], cy = kCoveredBlocksY[t];
  size_t xsize = cx * 8;
  size_t xs = cx / cy, xsm = xs - 1, xss = ceil_log2(xs);
  std::vector<uint32_t> out(cx * cy * 64, 0);
  size_t cur = cx * cy;
  for (size_t i = 0; i < xsize; i++) {
    for (size_t j = 0; j <= i; j++) {
      size_t x = j, y = i - j;
      if (i % 2) std::swap(x, y);
      if (y & xsm) continue;
      y >>= xss;
      size_t val;
      if (x < cx && y < cy) val = y * cx + x;
      else val = cur++;
      out[val] = uint32_t(y * xsize + x);
    }
  }
  for (size_t ir = 1; ir < xsize; ir++) {
    size_t ip = xsize - ir, i = ip - 1;
    for (size_t j = 0; j <= i; j++) {
      size_t x = xsize - 1 - (i - j), y = xsize - 1 - j;
      if (i % 2) std::swap(x, y);
      if (y & xsm) continue;
      y >>= xss;
      out[cur++] = uint32_t(y * xsize + x);
    }
  }
  return out;
}

// ---------------------------------------------------------------------------

static constexpr float kAlmostZero = 1e-8f;

static float mult(float v) { return v > 0.0f ? 1.0f + v : 1.0f / (1.0f - v); }

static float interpolate_vec(float scaled_pos, const float* array) {  // quant_weights.rs:1165
  float idxf = std::floor(scaled_pos);
  float frac = scaled_pos - idxf;
  size_t idx = size_t(idxf);
  float a = array[idx], b = array[idx + 1];
  return std::pow(b / a, frac) * a;
}
static float interpolate(float pos, float max, const float* array, size_t len) {  // :1174
  float scaled_pos = pos * float(len - 1) / max;
  size_t idx = size_t(scaled_pos);
  float a = array[idx], b = array[idx + 1];
  return a * std::pow(b / a, scaled_pos - float(idx));
}

static void get_quant_weights(size_t rows, size_t cols, const DctParams& p, float* out) {  // :1119-1163
  for (int c = 0; c < 3; c++) {
    float bands[17] = {0};
    bands[0] = p.params[c][0];
    if (bands[0] < kAlmostZero) fail("invalid distance band");
    for (size_t i = 1; i < p.num_bands; i++) {
      bands[i] = bands[i - 1] * mult(p.params[c][i]);
      if (bands[i] < kAlmostZero) fail("invalid distance band");
    }
    float scale = float(p.num_bands - 1) / (float(M_SQRT2) + 1e-6f);
    float rcpcol = scale / float(cols - 1);
    float rcprow = scale / float(rows - 1);
    for (size_t y = 0; y < rows; y++) {
      float dy = float(y) * rcprow;
      float dy2 = dy * dy;
      for (size_t x = 0; x < cols; x++) {
        float dx = float(x) * rcpcol;
        float dist = std::sqrt(dx * dx + dy2);
        out[c * cols * rows + y * cols + x] = p.num_bands == 1 ? bands[0] : interpolate_vec(dist, bands);
      }
    }
  }
}

template <size_t N>
static DctParams params_from(const float (&v)[3][N]) {
  DctParams p;
  p.num_bands = N;
  for (int c = 0; c < 3; c++)
    for (size_t i = 0; i < N; i++) p.params[c][i] = v[c][i];
  return p;
}

QuantEncoding library_encoding(int idx) {  // quant_weights.rs:347-880
  QuantEncoding e;
  auto dct = [&](auto& arr) {
    e.mode = QuantEncoding::kDct;
    e.dct = params_from(arr);
  };
  switch (idx) {
    case 0: dct(k_dct_0); break;
    case 1:
      e.mode = QuantEncoding::kIdentity;
      for (int c = 0; c < 3; c++)
        for (int i = 0; i < 3; i++) e.weights[c][i] = k_id_0[c][i];
      break;
    case 2:
      e.mode = QuantEncoding::kDct2;
      for (int c = 0; c < 3; c++)
        for (int i = 0; i < 6; i++) e.weights[c][i] = k_dct2x2_0[c][i];
      break;
    case 3:
      e.mode = QuantEncoding::kDct4;
      e.dct = params_from(k_dct4x4_0);
      for (int c = 0; c < 3; c++)
        for (int i = 0; i < 2; i++) e.weights[c][i] = k_dct4x4_1[c][i];
      break;
    case 4: dct(k_dct16x16_0); break;
    case 5: dct(k_dct32x32_0); break;
    case 6: dct(k_dct8x16_0); break;
    case 7: dct(k_dct8x32_0); break;
    case 8: dct(k_dct16x32_0); break;
    case 9:
      e.mode = QuantEncoding::kDct4x8;
      e.dct = params_from(k_dct4x8_0);
      for (int c = 0; c < 3; c++) e.weights[c][0] = 1.0f;
      break;
    case 10:
      e.mode = QuantEncoding::kAfv;
      e.dct = params_from(k_dct4x8_0);
      e.dct4x4 = params_from(k_dct4x4_0);
      for (int c = 0; c < 3; c++)
        for (int i = 0; i < 9; i++) e.weights[c][i] = k_afv0_0[c][i];
      break;
    case 11: dct(k_dct64x64_0); break;
    case 12: dct(k_dct32x64_0); break;
    case 13: dct(k_dct128x128_0); break;
    case 14: dct(k_dct64x128_0); break;
    case 15: dct(k_dct256x256_0); break;
    case 16: dct(k_dct128x256_0); break;
  }
  return e;
}

std::vector<float> compute_dequant_table(const QuantEncoding& e, int idx) {  // quant_weights.rs:894-1079
  size_t wrows = 8 * kQuantTableRows[idx], wcols = 8 * kQuantTableCols[idx];
  size_t num = wrows * wcols;
  std::vector<float> w(3 * num, 0.0f);
  switch (e.mode) {
    case QuantEncoding::kLibrary: fail("library encoding has no parameters");
    case QuantEncoding::kIdentity:
      for (int c = 0; c < 3; c++) {
        for (int i = 0; i < 64; i++) w[64 * c + i] = e.weights[c][0];
        w[64 * c + 1] = e.weights[c][1];
        w[64 * c + 8] = e.weights[c][1];
        w[64 * c + 9] = e.weights[c][2];
      }
      break;
    case QuantEncoding::kDct2:
      for (int c = 0; c < 3; c++) {
        float* s = &w[c * 64];
        const float* xw = e.weights[c];
        s[0] = float(0xBAD);
        s[1] = xw[0];
        s[8] = xw[0];
        s[9] = xw[1];
        for (int y = 0; y < 2; y++)
          for (int x = 0; x < 2; x++) {
            s[y * 8 + x + 2] = xw[2];
            s[(y + 2) * 8 + x] = xw[2];
          }
        for (int y = 0; y < 2; y++)
          for (int x = 0; x < 2; x++) s[(y + 2) * 8 + x + 2] = xw[3];
        for (int y = 0; y < 4; y++)
          for (int x = 0; x < 4; x++) {
            s[y * 8 + x + 4] = xw[4];
            s[(y + 4) * 8 + x] = xw[4];
          }
        for (int y = 0; y < 4; y++)
          for (int x = 0; x < 4; x++) s[(y + 4) * 8 + x + 4] = xw[5];
      }
      break;
    case QuantEncoding::kDct4: {
      float w44[3 * 16];
      get_quant_weights(4, 4, e.dct, w44);
      for (int c = 0; c < 3; c++) {
        for (int y = 0; y < 8; y++)
          for (int x = 0; x < 8; x++) w[c * num + y * 8 + x] = w44[c * 16 + (y / 2) * 4 + (x / 2)];
        w[c * num + 1] /= e.weights[c][0];
        w[c * num + 8] /= e.weights[c][0];
        w[c * num + 9] /= e.weights[c][1];
      }
      break;
    }
    case QuantEncoding::kDct4x8: {
      float w48[3 * 32];
      get_quant_weights(4, 8, e.dct, w48);
      for (int c = 0; c < 3; c++) {
        for (int y = 0; y < 8; y++)
          for (int x = 0; x < 8; x++) w[c * num + y * 8 + x] = w48[c * 32 + (y / 2) * 8 + x];
        w[c * num + 8] /= e.weights[c][0];
      }
      break;
    }
    case QuantEncoding::kDct: get_quant_weights(wrows, wcols, e.dct, w.data()); break;
    case QuantEncoding::kRaw:
      if (e.qtable.size() != 3 * num) fail("invalid raw quant table");
      for (size_t i = 0; i < 3 * num; i++) w[i] = 1.0f / (e.qtable_den * float(e.qtable[i]));
      break;
    case QuantEncoding::kAfv: {
      static const float kFreqs[16] = {float(0xBAD), float(0xBAD), 0.8517778890324296f, 5.37778436506804f,
                                       float(0xBAD), float(0xBAD), 4.734747904497923f,  5.449245381693219f,
                                       1.6598270267479331f, 4.0f, 7.275749096817861f, 10.423227632456525f,
                                       2.662932286148962f, 7.630657783650829f, 8.962388608184032f, 12.97166202570235f};
      float w48[3 * 32], w44[3 * 16];
      get_quant_weights(4, 8, e.dct, w48);
      get_quant_weights(4, 4, e.dct4x4, w44);
      const float lo = 0.8517778890324296f;
      const float hi = 12.97166202570235f - lo + 1e-6f;
      for (int c = 0; c < 3; c++) {
        float bands[4];
        bands[0] = e.weights[c][5];
        if (bands[0] < kAlmostZero) fail("invalid distance band");
        for (int i = 1; i < 4; i++) {
          bands[i] = bands[i - 1] * mult(e.weights[c][i + 5]);
          if (bands[i] < kAlmostZero) fail("invalid distance band");
        }
        float* s = &w[c * 64];
        s[0] = 1.0f;
        auto set = [&](int x, int y, float v) { s[y * 8 + x] = v; };
        set(0, 1, e.weights[c][0]);
        set(1, 0, e.weights[c][1]);
        set(0, 2, e.weights[c][2]);
        set(2, 0, e.weights[c][3]);
        set(2, 2, e.weights[c][4]);
        for (int y = 0; y < 4; y++)
          for (int x = 0; x < 4; x++) {
            if (x < 2 && y < 2) continue;
            set(2 * x, 2 * y, interpolate(kFreqs[y * 4 + x] - lo, hi, bands, 4));
          }
        for (int y = 0; y < 4; y++)
          for (int x = 0; x < 8; x++) {
            if (x == 0 && y == 0) continue;
            w[c * num + (2 * y + 1) * 8 + x] = w48[c * 32 + y * 8 + x];
          }
        for (int y = 0; y < 4; y++)
          for (int x = 0; x < 4; x++) {
            if (x == 0 && y == 0) continue;
            w[c * num + (2 * y) * 8 + 2 * x + 1] = w44[c * 16 + y * 4 + x];
          }
      }
      break;
    }
  }
  for (float& v : w) {
    if (!(v >= kAlmostZero && v <= 1.0f / kAlmostZero)) fail("invalid quantisation table weight");
    v = 1.0f / v;
  }
  return w;
}

const std::vector<float>& library_dequant_table(int idx) {
  static std::vector<float> tables[kNumQuantTables];
  static std::once_flag flags[kNumQuantTables];
  std::call_once(flags[idx], [idx] { tables[idx] = compute_dequant_table(library_encoding(idx), idx); });
  return tables[idx];
}

static DctParams read_dct_params(BitReader& br) {  // quant_weights.rs:54-69
  DctParams p;
  p.num_bands = size_t(br.read(4)) + 1;
  for (int c = 0; c < 3; c++) {
    for (size_t i = 0; i < p.num_bands; i++) p.params[c][i] = f16_bits_to_float(uint16_t(br.read(16)));
    if (p.params[c][0] < kAlmostZero) fail("HF quant factor too small");
    p.params[c][0] *= 64.0f;
  }
  return p;
}

QuantEncoding read_quant_encoding(int idx, BitReader& br, const FrameHeader& fh, const ModularTree* global_tree) {
  // quant_weights.rs:111-247
  QuantEncoding e;
  size_t required = size_t(kQuantTableRows[idx]) * kQuantTableCols[idx];
  uint32_t mode = uint32_t(br.read(3));
  auto need1 = [&] {
    if (required != 1) fail("invalid quant encoding for this table size");
  };
  auto rdw = [&](int n, bool check, float scale) {
    for (int c = 0; c < 3; c++)
      for (int i = 0; i < n; i++) {
        float v = f16_bits_to_float(uint16_t(br.read(16)));
        if (check && std::fabs(v) < kAlmostZero) fail("HF quant factor too small");
        e.weights[c][i] = v * scale;
      }
  };
  switch (mode) {
    case 0: e.mode = QuantEncoding::kLibrary; break;
    case 1: need1(); e.mode = QuantEncoding::kIdentity; rdw(3, true, 64.0f); break;
    case 2: need1(); e.mode = QuantEncoding::kDct2; rdw(6, true, 64.0f); break;
    case 3: need1(); e.mode = QuantEncoding::kDct4; rdw(2, true, 1.0f); e.dct = read_dct_params(br); break;
    case 4:
      need1();
      e.mode = QuantEncoding::kDct4x8;
      for (int c = 0; c < 3; c++) {
        float v = f16_bits_to_float(uint16_t(br.read(16)));
        if (std::fabs(v) < kAlmostZero) fail("HF quant factor too small");
        e.weights[c][0] = v;
      }
      e.dct = read_dct_params(br);
      break;
    case 5:
      need1();
      e.mode = QuantEncoding::kAfv;
      for (int c = 0; c < 3; c++) {
        for (int i = 0; i < 9; i++) e.weights[c][i] = f16_bits_to_float(uint16_t(br.read(16)));
        for (int i = 0; i < 6; i++) e.weights[c][i] *= 64.0f;
      }
      e.dct = read_dct_params(br);
      e.dct4x4 = read_dct_params(br);
      break;
    case 6: e.mode = QuantEncoding::kDct; e.dct = read_dct_params(br); break;
    case 7: {
      e.mode = QuantEncoding::kRaw;
      e.qtable_den = f16_bits_to_float(uint16_t(br.read(16)));
      if (e.qtable_den < kAlmostZero) fail("invalid raw quant table");
      // modular/mod.rs:1083-1121 (decode_quant_table)
      uint32_t sx = 8u * kQuantTableRows[idx], sy = 8u * kQuantTableCols[idx];
      std::vector<ModularChannel> ch;
      for (int c = 0; c < 3; c++) ch.emplace_back(sx, sy);
      size_t stream_id = 1 + size_t(fh.num_lf_groups()) * 3 + size_t(idx);
      decode_modular_subbitstream(ch, stream_id, global_tree, br);
      for (auto& c : ch)
        for (int32_t v : c.data) {
          if (v <= 0) fail("invalid raw quant table");
          e.qtable.push_back(v);
        }
      break;
    }
  }
  return e;
}

}  // namespace jxg

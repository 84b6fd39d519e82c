// Test hooks (extern "C") into the host front-end primitives and the oracle, used
// by tests/test_kat_*.py to replay the reference's own known-answer tests.
// Test infrastructure only — lives in liboracle.so, never in libjxgpu.so.
#include <cstring>
#include <vector>

#include "../jxl_rs_b200/csrc/host/frame.h"
#include "../jxl_rs_b200/csrc/host/quant.h"

using namespace jxg;

extern "C" {

// ans.rs:463-485: parse one ANS histogram; out_dist[i] = probability of symbol i. Returns 0, or -1 on a parse error.
int jxo_t_ans_histogram(const uint8_t* bytes, size_t len, uint32_t log_alpha, uint16_t* out_dist, int32_t* single) {
  try {
    BitReader br(bytes, len);
    std::vector<AnsBucket> b;
    int32_t s = EntropyCode::decode_ans_histogram_for_test(br, log_alpha, b);
    for (size_t i = 0; i < b.size(); i++) out_dist[i] = b[i].dist;
    if (single) *single = s;
    return 0;
  } catch (Error&) {
    return -1;
  }
}

// huffman.rs:516-527: build prefix codes from `hist`, then read `n` symbols from `data` with cluster 0.
int jxo_t_prefix_read(const uint8_t* hist, size_t hist_len, const uint8_t* data, size_t data_len, uint32_t n, uint32_t* out) {
  try {
    BitReader hbr(hist, hist_len);
    EntropyCode code = EntropyCode::decode_prefix_codes_for_test(1, hbr);
    BitReader br(data, data_len);
    SymbolReader r(code, br, 0);
    for (uint32_t i = 0; i < n; i++) out[i] = r.read_clustered(br, 0);
    return 0;
  } catch (Error&) {
    return -1;
  }
}

// Full histogram-set + symbol stream decode (decode.rs:487 + :271): returns values, checks the final state.
int jxo_t_decode_stream(const uint8_t* bytes, size_t len, uint32_t num_contexts, const uint32_t* ctxs, uint32_t n, uint32_t* out) {
  try {
    BitReader br(bytes, len);
    EntropyCode code = EntropyCode::decode(num_contexts, br, true);
    SymbolReader r(code, br, 0);
    for (uint32_t i = 0; i < n; i++) out[i] = r.read_unsigned(br, ctxs[i]);
    r.check_final_state(br);
    return 0;
  } catch (Error&) {
    return -1;
  }
}

int jxo_t_hybrid_decode_config(const uint8_t* bytes, size_t len, uint32_t skip_bits, uint32_t log_alpha, uint32_t token, uint32_t* value) {
  try {
    BitReader br(bytes, len);
    br.skip_bits(skip_bits);
    HybridUint u = HybridUint::decode(log_alpha, br);
    *value = u.read(token, br);
    return 0;
  } catch (Error&) {
    return -1;
  }
}

uint64_t jxo_t_bitreader(const uint8_t* bytes, size_t len, const uint32_t* nbits, uint32_t n, uint64_t* out) {
  BitReader br(bytes, len);
  for (uint32_t i = 0; i < n; i++) out[i] = br.read(nbits[i]);
  return br.total_bits_read();
}

uint32_t jxo_t_natural_order(int order_idx, uint32_t* out, size_t cap) {
  std::vector<uint32_t> o = natural_coeff_order(order_idx);
  if (o.size() <= cap) memcpy(out, o.data(), o.size() * 4);
  return uint32_t(o.size());
}

uint32_t jxo_t_dequant_table(int transform, int c, float* out, size_t cap) {
  int qt = quant_table_for_transform(transform);
  const std::vector<float>& t = library_dequant_table(qt);
  size_t n = t.size() / 3;
  if (n <= cap) memcpy(out, t.data() + size_t(c) * n, n * 4);
  return uint32_t(n);
}

int jxo_t_lehmer(const uint32_t* code, uint32_t n, uint32_t skip, uint32_t size, uint32_t* out) {
  try {
    std::vector<uint32_t> p = apply_lehmer(std::vector<uint32_t>(code, code + n), skip, size);
    memcpy(out, p.data(), p.size() * 4);
    return 0;
  } catch (Error&) {
    return -1;
  }
}

// predict.rs:564-593 (predict_and_update_errors golden)
void jxo_t_wp_golden(int64_t* preds, int32_t* props) {
  struct Rnd {
    int64_t out = 1;
    int64_t next() {
      out = out * 48271 % 0x7fffffff;
      return out;
    }
  } rng;
  WeightedHeader h;
  h.p1c = uint32_t(rng.next() % 32);
  h.p2c = uint32_t(rng.next() % 32);
  h.p3ca = uint32_t(rng.next() % 32);
  h.p3cb = uint32_t(rng.next() % 32);
  h.p3cc = uint32_t(rng.next() % 32);
  h.p3cd = uint32_t(rng.next() % 32);
  h.p3ce = uint32_t(rng.next() % 32);
  for (auto& w : h.w) w = uint32_t(rng.next() % 16);
  const size_t xs = 8, ys = 8;
  WpState st(h, xs);
  for (int i = 0; i < 4; i++) {
    size_t x = size_t(rng.next()) % xs, y = size_t(rng.next()) % ys;
    int32_t top = int32_t(rng.next()) % 256, left = int32_t(rng.next()) % 256, topright = int32_t(rng.next()) % 256,
            topleft = int32_t(rng.next()) % 256, toptop = int32_t(rng.next()) % 256;
    st.predict(x, y, top, left, topright, topleft, toptop, preds[i], props[i]);
    st.update(int32_t(rng.next() % 256), x, y);
  }
}

// Modular sub-bitstream decode of `nch` channels of w x h (test for the host modular decoder via synth streams).
int jxo_t_parse_ok(const uint8_t* data, size_t size) {
  try {
    auto fs = parse_vardct_file(data, size);
    return 0;
  } catch (Error& e) {
    return e.code;
  }
}

// Differential tests of the host front-end's specialised Modular walks: 1 routes every channel through the generic
// all-properties loop.
void jxo_t_force_generic_walk(int on) { set_force_generic_walk(on != 0); }

// frame/adaptive_lf_smoothing.rs on caller planes (3 x ys x xs f32, in place), for the known-answer test.
void jxo_t_adaptive_lf_smoothing(uint32_t xs, uint32_t ys, uint32_t global_scale, uint32_t quant_lf,
                                 const float* lf_quant, float* planes, int threads) {
  FrameState fs;
  fs.xb = xs;
  fs.yb = ys;
  fs.global_scale = global_scale;
  fs.quant_lf = quant_lf;
  for (int c = 0; c < 3; c++) {
    fs.lf_quant[c] = lf_quant[c];
    fs.lf[c].assign(planes + size_t(c) * xs * ys, planes + size_t(c + 1) * xs * ys);
  }
  adaptive_lf_smoothing(fs, threads);
  for (int c = 0; c < 3; c++) std::copy(fs.lf[c].begin(), fs.lf[c].end(), planes + size_t(c) * xs * ys);
}

// FNV-1a digest over everything the front-end hands to the hot path (planes, maps, section table), for tests that
// compare parses (e.g. serial vs multi-threaded LF groups). Returns 0 on a parse error.
uint64_t jxo_t_parse_digest(const uint8_t* data, size_t size, int threads) {
  try {
    auto fs = parse_vardct_file(data, size, threads);
    uint64_t h = 1469598103934665603ull;
    auto mix = [&h](const void* p, size_t n) {
      const uint8_t* b = static_cast<const uint8_t*>(p);
      for (size_t i = 0; i < n; i++) h = (h ^ b[i]) * 1099511628211ull;
    };
    for (int c = 0; c < 3; c++) mix(fs->lf[c].data(), fs->lf[c].size() * 4);
    mix(fs->transform_map.data(), fs->transform_map.size());
    mix(fs->raw_quant_map.data(), fs->raw_quant_map.size() * 4);
    mix(fs->epf_map.data(), fs->epf_map.size());
    mix(fs->quant_lf_map.data(), fs->quant_lf_map.size());
    mix(fs->ytox_map.data(), fs->ytox_map.size());
    mix(fs->ytob_map.data(), fs->ytob_map.size());
    mix(fs->hf_off.data(), fs->hf_off.size() * 8);
    mix(fs->hf_len.data(), fs->hf_len.size() * 4);
    recycle_frame_state(fs.release());
    return h ? h : 1;
  } catch (Error&) {
    return 0;
  }
}

// 0: serial parses decode their LF groups one at a time instead of in lockstep pairs (differential tests).
void jxo_t_pair_lf_groups(int on) { set_pair_lf_groups(on != 0); }

// Output colour derivation of the host front-end (headers.cc resolve_output_colour <- render/stages/xyb.rs:65-140) for
// a non-ICC colour encoding given by its header fields; the opsin inverse matrix is the default one. Returns the error
// code on refusal. out: 9 matrix entries, 3 luminances, tf, gamma.
int jxo_t_output_colour(uint32_t color_space, uint32_t white_point, const int32_t* white_xy, uint32_t primaries,
                        const int32_t* prim_xy, int have_gamma, uint32_t gamma, uint32_t tf, float* out) {
  try {
    FileHeader fh;
    ColorEncoding& c = fh.color_encoding;
    c.all_default = false;
    c.color_space = ColorSpace(color_space);
    c.white_point = white_point;
    c.primaries = primaries;
    if (white_xy) memcpy(c.white_xy, white_xy, sizeof(c.white_xy));
    if (prim_xy) memcpy(c.primaries_xy, prim_xy, sizeof(c.primaries_xy));
    c.have_gamma = have_gamma != 0;
    c.gamma = gamma;
    c.tf = TransferFunction(tf);
    const OutputColour oc = resolve_output_colour(fh);
    memcpy(out, oc.matrix, 36);
    memcpy(out + 9, oc.luminances, 12);
    out[12] = float(oc.tf);
    out[13] = oc.gamma;
    return 0;
  } catch (Error& e) {
    return e.code;
  }
}

// modular/mod.rs:837-929 dequant_lf through the front-end's dequant_lf_rect on caller data: q = 3 planes (Y, X, B order of
// the coded channels) of w x h quantised LF integers; out: X, Y, B f32 planes; qlf: context buckets.
void jxo_t_dequant_lf(uint32_t w, uint32_t h, const int32_t* q, uint32_t global_scale, uint32_t quant_lf, const float* lf_quant,
                      uint32_t extra_precision, float base_x, float base_b, int32_t ytox_lf, int32_t ytob_lf, uint32_t color_factor,
                      const int32_t* thr, const uint32_t* nthr, float* out, uint8_t* qlf) {
  FrameState fs;
  fs.xb = w;
  fs.yb = h;
  fs.global_scale = global_scale;
  fs.quant_lf = quant_lf;
  for (int c = 0; c < 3; c++) fs.lf_quant[c] = lf_quant[c];
  fs.base_correlation_x = base_x;
  fs.base_correlation_b = base_b;
  fs.ytox_lf = ytox_lf;
  fs.ytob_lf = ytob_lf;
  fs.color_factor = color_factor;
  fs.num_lf_contexts = 1;
  for (int c = 0; c < 3; c++) {
    fs.lf_thresholds[c].assign(thr, thr + nthr[c]);
    thr += nthr[c];
    fs.num_lf_contexts *= nthr[c] + 1;
    fs.lf[c].assign(size_t(w) * h, 0.0f);
  }
  fs.quant_lf_map.assign(size_t(w) * h, 0);
  const size_t n = size_t(w) * h;
  dequant_lf_rect(fs, q, q + n, q + 2 * n, w, w, h, 1.0f / float(1u << extra_precision), 0);
  for (int c = 0; c < 3; c++) memcpy(out + size_t(c) * n, fs.lf[c].data(), n * 4);
  memcpy(qlf, fs.quant_lf_map.data(), n);
}

// The front-end's varblock placement of one LF-group rect (frame.cc place_varblocks): maps come back w x h with 27 =
// uncovered; returns 0 or the front-end's error code for an invalid block list.
int jxo_t_place_varblocks(uint32_t w, uint32_t h, uint32_t count, const int32_t* raw_transforms, const int32_t* raw_quants,
                          uint8_t* transform_map, int32_t* raw_quant_map) {
  for (size_t i = 0; i < size_t(w) * h; i++) {
    transform_map[i] = 27;
    raw_quant_map[i] = 0;
  }
  try {
    jxg::place_varblocks(w, h, w, count, raw_transforms, raw_quants, transform_map, raw_quant_map);
    return 0;
  } catch (jxg::Error& e) {
    return e.code ? e.code : -1;
  }
}
}

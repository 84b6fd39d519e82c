// See headers.h.
#include "headers.h"

#include <algorithm>
#include <cmath>

#include "entropy.h"

namespace jxg {

namespace {

// encodings.rs U32Coder: one distribution = (nbits, offset); nbits==0 && value only.
struct U32D {
  uint32_t bits, off;
};
inline uint32_t u2s(BitReader& br, U32D a, U32D b, U32D c, U32D d) {
  const U32D ds[4] = {a, b, c, d};
  const U32D& s = ds[br.read(2)];
  return uint32_t(br.read(s.bits)) + s.off;
}
constexpr U32D V(uint32_t v) { return U32D{0, v}; }
constexpr U32D B(uint32_t n, uint32_t off = 0) { return U32D{n, off}; }

// jxl_macros/src/lib.rs:671-679: default enum coder
inline uint32_t read_enum(BitReader& br) { return u2s(br, V(0), V(1), B(4, 2), B(6, 18)); }

std::string read_string(BitReader& br) {  // encodings.rs:141-171
  uint32_t len = u2s(br, V(0), B(4), B(5, 16), B(10, 48));
  std::string s;
  for (uint32_t i = 0; i < len; i++) s.push_back(char(br.read(8)));
  return s;
}

BitDepth read_bit_depth(BitReader& br) {  // bit_depth.rs:13-27
  BitDepth d;
  d.floating_point = br.read_bool();
  if (d.floating_point) {
    d.bits_per_sample = u2s(br, V(32), V(16), V(24), B(6, 1));
    d.exponent_bits = uint32_t(br.read(4)) + 1;
  } else {
    d.bits_per_sample = u2s(br, V(8), V(10), V(12), B(6, 1));
  }
  return d;
}

void read_size(BitReader& br, uint32_t& xs, uint32_t& ys) {  // size.rs:31-47
  bool small = br.read_bool();
  if (small) ys = (uint32_t(br.read(5)) + 1) * 8;
  else ys = 1 + u2s(br, B(9), B(13), B(18), B(30));
  uint32_t ratio = uint32_t(br.read(3));
  if (ratio == 0) {
    if (small) xs = (uint32_t(br.read(5)) + 1) * 8;
    else xs = 1 + u2s(br, B(9), B(13), B(18), B(30));
  } else {
    static const uint32_t num[8] = {0, 1, 12, 4, 3, 16, 5, 2}, den[8] = {1, 1, 10, 3, 2, 9, 4, 1};
    xs = uint32_t(uint64_t(ys) * num[ratio] / den[ratio]);
  }
}

void read_preview(BitReader& br) {  // size.rs:49-66
  bool div8 = br.read_bool();
  if (div8) u2s(br, V(16), V(32), B(5, 1), B(9, 33));
  else u2s(br, B(6), B(8, 64), B(10, 320), B(12, 1344));
  uint32_t ratio = uint32_t(br.read(3));
  if (ratio == 0) {
    if (div8) u2s(br, V(16), V(32), B(5, 1), B(9, 33));
    else u2s(br, B(6), B(8, 64), B(10, 320), B(12, 1344));
  }
}

void read_custom_xy(BitReader& br) {
  for (int i = 0; i < 2; i++) u2s(br, B(19), B(19, 524288), B(20, 1048576), B(21, 2097152));
}

ColorEncoding read_color_encoding(BitReader& br) {  // color_encoding.rs:166-196
  ColorEncoding c;
  c.all_default = br.read_bool();
  if (c.all_default) return c;
  c.want_icc = br.read_bool();
  c.color_space = ColorSpace(read_enum(br));
  bool not_xyb = c.color_space != ColorSpace::XYB;
  if (!c.want_icc && not_xyb) {
    c.white_point = read_enum(br);
    if (c.white_point == 2) read_custom_xy(br);
  }
  if (!c.want_icc && not_xyb && c.color_space != ColorSpace::Gray) {
    c.primaries = read_enum(br);
    if (c.primaries == 2)
      for (int i = 0; i < 3; i++) read_custom_xy(br);
  }
  if (!c.want_icc) {
    if (not_xyb) c.have_gamma = br.read_bool();
    if (c.have_gamma) c.gamma = uint32_t(br.read(24));
    if (!c.have_gamma && not_xyb) c.tf = TransferFunction(read_enum(br));
    c.rendering_intent = read_enum(br);
  }
  return c;
}

ExtraChannelInfo read_extra_channel(BitReader& br) {  // extra_channels.rs:36-56
  ExtraChannelInfo e;
  if (br.read_bool()) return e;
  e.type = read_enum(br);
  e.bit_depth = read_bit_depth(br);
  e.dim_shift = u2s(br, V(0), V(3), V(4), B(3, 1));
  read_string(br);
  if (e.type == 0) e.alpha_associated = br.read_bool();
  if (e.type == 2)
    for (int i = 0; i < 4; i++) read_f16(br);
  if (e.type == 5) u2s(br, V(1), B(2), B(4, 3), B(8, 19));
  if (e.dim_shift > 3) fail("dim_shift too large");
  return e;
}

void skip_icc(BitReader& br) {  // icc/mod.rs:100-190 (decoded and discarded)
  uint64_t len = read_u64(br);
  if (len > (1ull << 24)) fail("ICC too large");
  EntropyCode code = EntropyCode::decode(41, br, true);
  SymbolReader reader(code, br, 0);
  uint8_t b1 = 0, b2 = 0;
  auto is_alpha = [](uint8_t b) { return (b >= 'a' && b <= 'z') || (b >= 'A' && b <= 'Z'); };
  auto is_num = [](uint8_t b) { return (b >= '0' && b <= '9') || b == '.' || b == ','; };
  for (uint64_t i = 0; i < len; i++) {
    uint32_t ctx = 0;
    if (i > 128) {
      uint32_t p1 = is_alpha(b1) ? 0 : is_num(b1) ? 1 : b1 <= 1 ? 2 + b1 : b1 <= 15 ? 4 : (b1 >= 241 && b1 <= 254) ? 5 : b1 == 255 ? 6 : 7;
      uint32_t p2 = is_alpha(b2) ? 0 : is_num(b2) ? 1 : b2 <= 15 ? 2 : b2 >= 241 ? 3 : 4;
      ctx = 1 + p1 + 8 * p2;
    }
    uint32_t sym = reader.read_unsigned(br, ctx);
    if (sym >= 256) fail("invalid ICC symbol");
    b2 = b1;
    b1 = uint8_t(sym);
  }
  reader.check_final_state(br);
}

}  // namespace

float f16_bits_to_float(uint16_t h) {
  uint32_t sign = h >> 15, exp = (h >> 10) & 31, man = h & 1023;
  float v;
  if (exp == 0) v = std::ldexp(float(man), -24);
  else if (exp == 31) v = man ? NAN : INFINITY;
  else v = std::ldexp(float(man | 1024), int(exp) - 25);
  return sign ? -v : v;
}

float read_f16(BitReader& br) {
  float v = f16_bits_to_float(uint16_t(br.read(16)));
  if (!std::isfinite(v)) fail("NaN or Inf in header float");
  return v;
}

uint64_t read_u64(BitReader& br) {  // encodings.rs:111-139
  switch (br.read(2)) {
    case 0: return 0;
    case 1: return 1 + br.read(4);
    case 2: return 17 + br.read(8);
    default: {
      uint64_t result = br.read(12);
      unsigned shift = 12;
      while (br.read(1) == 1) {
        if (shift >= 60) return result | (br.read(4) << shift);
        result |= br.read(8) << shift;
        shift += 8;
      }
      return result;
    }
  }
}

void read_extensions(BitReader& br) {  // encodings.rs:377-407
  uint64_t selector = read_u64(br);
  uint64_t total = 0;
  for (int i = 0; i < 64; i++)
    if (selector & (1ull << i)) total += read_u64(br);
  if (total > br.size_bits()) fail("extension size overflow", kErrOutOfBounds);
  br.skip_bits(size_t(total));
}

std::vector<uint8_t> extract_codestream(const uint8_t* data, size_t size) {
  std::vector<uint8_t> out;
  extract_codestream(data, size, out);
  return out;
}

void extract_codestream(const uint8_t* data, size_t size, std::vector<uint8_t>& out) {
  out.clear();
  if (size >= 2 && data[0] == 0xff && data[1] == 0x0a) {
    out.assign(data, data + size);
    return;
  }
  static const uint8_t kSig[12] = {0, 0, 0, 0xc, 'J', 'X', 'L', ' ', 0xd, 0xa, 0x87, 0xa};
  if (size < 12 || memcmp(data, kSig, 12) != 0) fail("not a JPEG XL file");
  size_t pos = 0;
  while (pos + 8 <= size) {
    uint64_t box_size = (uint64_t(data[pos]) << 24) | (uint64_t(data[pos + 1]) << 16) | (uint64_t(data[pos + 2]) << 8) | data[pos + 3];
    const uint8_t* ty = data + pos + 4;
    size_t header = 8;
    if (box_size == 1) {
      if (pos + 16 > size) fail("truncated box header");
      box_size = 0;
      for (int i = 0; i < 8; i++) box_size = (box_size << 8) | data[pos + 8 + i];
      header = 16;
    }
    size_t end = box_size == 0 ? size : pos + size_t(box_size);
    if (end > size || end < pos + header) fail("bad box size");
    if (!memcmp(ty, "jxlc", 4)) out.insert(out.end(), data + pos + header, data + end);
    else if (!memcmp(ty, "jxlp", 4)) {
      if (end < pos + header + 4) fail("bad jxlp box");
      // in-order jxlp only (box_parser.rs handles out-of-order; test files exercising it are rejected upstream of here)
      out.insert(out.end(), data + pos + header + 4, data + end);
    }
    pos = end;
  }
  if (out.empty()) fail("no codestream box");
}

FileHeader read_file_header(BitReader& br) {
  FileHeader h;
  if (br.read(8) != 0xff || br.read(8) != 0x0a) fail("invalid signature");
  read_size(br, h.xsize, h.ysize);
  // ImageMetadata (image_metadata.rs:197-236)
  bool all_default = br.read_bool();
  bool extra_fields = false;
  if (!all_default) {
    extra_fields = br.read_bool();
    if (extra_fields) {
      h.orientation = uint32_t(br.read(3)) + 1;
      if (br.read_bool()) {  // intrinsic size
        uint32_t a, b;
        read_size(br, a, b);
      }
      h.have_preview = br.read_bool();
      if (h.have_preview) read_preview(br);
      h.have_animation = br.read_bool();
      if (h.have_animation) {
        u2s(br, V(100), V(1000), B(10, 1), B(30, 1));
        u2s(br, V(1), V(1001), B(8, 1), B(10, 1));
        u2s(br, V(0), B(3), B(16), B(32));
        h.have_timecodes = br.read_bool();
      }
    }
    h.bit_depth = read_bit_depth(br);
    h.modular_16bit_sufficient = br.read_bool();
    uint32_t num_ec = u2s(br, V(0), V(1), B(4, 2), B(12, 1));
    for (uint32_t i = 0; i < num_ec; i++) h.extra_channels.push_back(read_extra_channel(br));
    h.xyb_encoded = br.read_bool();
    h.color_encoding = read_color_encoding(br);
    if (extra_fields) {  // ToneMapping
      if (!br.read_bool()) {
        h.intensity_target = read_f16(br);
        read_f16(br);
        br.read_bool();
        read_f16(br);
        if (h.intensity_target <= 0) fail("invalid intensity target");
      }
    }
    read_extensions(br);
  }
  // CustomTransformData (transform_data.rs:322-345)
  if (!br.read_bool()) {
    if (h.xyb_encoded) {
      if (!br.read_bool()) {
        for (float& v : h.opsin.inverse_matrix) v = read_f16(br);
        for (float& v : h.opsin.opsin_biases) v = read_f16(br);
        for (float& v : h.opsin.quant_biases) v = read_f16(br);
      }
    }
    h.custom_upsampling_mask = uint32_t(br.read(3));
    if (h.custom_upsampling_mask & 1)
      for (int i = 0; i < 15; i++) read_f16(br);
    if (h.custom_upsampling_mask & 2)
      for (int i = 0; i < 55; i++) read_f16(br);
    if (h.custom_upsampling_mask & 4)
      for (int i = 0; i < 210; i++) read_f16(br);
  }
  if (h.color_encoding.want_icc) skip_icc(br);
  br.check();
  return h;
}

static BlendingInfo read_blending(BitReader& br, uint32_t num_ec, bool full_frame) {  // frame_header.rs:111-139
  BlendingInfo b;
  b.mode = u2s(br, V(0), V(1), V(2), B(2, 3));
  if (b.mode > 4) fail("invalid blend mode");
  bool uses_alpha = num_ec > 0 && (b.mode == 2 || b.mode == 3);
  if (uses_alpha) b.alpha_channel = u2s(br, V(0), V(1), V(2), B(3, 3));
  if (uses_alpha || b.mode == 4) b.clamp = br.read_bool();
  if (!(full_frame && b.mode == 0)) b.source = u2s(br, V(0), V(1), V(2), V(3));
  return b;
}

FrameHeader read_frame_header(BitReader& br, const FileHeader& fh) {
  br.jump_to_byte_boundary();  // #[aligned], frame_header.rs:264
  FrameHeader f;
  uint32_t num_ec = uint32_t(fh.extra_channels.size());
  f.num_extra_channels = num_ec;
  f.ec_upsampling.assign(num_ec, 1);
  f.ec_blending.assign(num_ec, BlendingInfo{});
  bool all_default = br.read_bool();
  if (!all_default) {
    f.frame_type = uint32_t(br.read(2));
    f.encoding = uint32_t(br.read(1));
    f.flags = read_u64(br);
    if (!fh.xyb_encoded) f.do_ycbcr = br.read_bool();
    bool use_lf = f.has_lf_frame();
    if (f.do_ycbcr && !use_lf)
      for (auto& u : f.jpeg_upsampling) u = uint32_t(br.read(2));
    if (!use_lf) {
      f.upsampling = u2s(br, V(1), V(2), V(4), V(8));
      for (auto& u : f.ec_upsampling) u = u2s(br, V(1), V(2), V(4), V(8));
    }
    if (f.encoding == 1) f.group_size_shift = uint32_t(br.read(2));
    if (f.encoding == 0 && fh.xyb_encoded) {
      f.x_qm_scale = uint32_t(br.read(3));
      f.b_qm_scale = uint32_t(br.read(3));
    }
    if (f.frame_type != 2) {  // Passes (frame_header.rs:46-75)
      Passes& p = f.passes;
      p.num_passes = u2s(br, V(1), V(2), V(3), B(3, 4));
      if (p.num_passes != 1) {
        p.num_ds = u2s(br, V(0), V(1), V(2), B(1, 3));
        p.shift.resize(p.num_passes - 1);
        for (auto& s : p.shift) s = uint32_t(br.read(2));
        p.downsample.resize(p.num_ds);
        for (auto& s : p.downsample) s = u2s(br, V(1), V(2), V(4), V(8));
        p.last_pass.resize(p.num_ds);
        for (auto& s : p.last_pass) s = u2s(br, V(0), V(1), V(2), B(3));
      }
    }
    if (f.frame_type == 1) f.lf_level = u2s(br, V(1), V(2), V(3), V(4));
    if (f.frame_type != 1) f.have_crop = br.read_bool();
    if (f.have_crop) {
      if (f.frame_type != 2) {
        f.x0 = unpack_signed(u2s(br, B(8), B(11, 256), B(14, 2304), B(30, 18688)));
        f.y0 = unpack_signed(u2s(br, B(8), B(11, 256), B(14, 2304), B(30, 18688)));
      }
      f.frame_width = u2s(br, B(8), B(11, 256), B(14, 2304), B(30, 18688));
      f.frame_height = u2s(br, B(8), B(11, 256), B(14, 2304), B(30, 18688));
    }
    bool covers = f.x0 <= 0 && f.y0 <= 0 && int64_t(f.frame_width) + f.x0 >= int64_t(fh.xsize) &&
                  int64_t(f.frame_height) + f.y0 >= int64_t(fh.ysize);
    bool full_frame = !f.have_crop || covers;
    bool normal = f.frame_type == 0 || f.frame_type == 3;
    if (normal) {
      f.blending = read_blending(br, num_ec, full_frame);
      for (auto& b : f.ec_blending) b = read_blending(br, num_ec, full_frame);
      if (fh.have_animation) f.duration = u2s(br, V(0), V(1), B(8), B(32));
      if (fh.have_timecodes) br.read(32);
      f.is_last = br.read_bool();
    } else {
      f.is_last = false;
    }
    if (f.frame_type != 1 && !f.is_last) f.save_as_reference = uint32_t(br.read(2));
    bool can_be_referenced = !f.is_last && f.frame_type != 1 && (f.duration == 0 || f.save_as_reference != 0);
    bool sbct_def_false = can_be_referenced && f.blending.mode == 0 && full_frame && normal;
    f.save_before_ct = f.frame_type == 1;
    if (f.frame_type == 2 || sbct_def_false) f.save_before_ct = br.read_bool();
    f.name = read_string(br);
    // RestorationFilter (frame_header.rs:146-234)
    RestorationFilter& r = f.rf;
    if (!br.read_bool()) {
      r.gab = br.read_bool();
      if (r.gab && br.read_bool()) {
        for (int c = 0; c < 3; c++) {
          r.gab_w1[c] = read_f16(br);
          r.gab_w2[c] = read_f16(br);
        }
      }
      r.epf_iters = uint32_t(br.read(2));
      bool vardct = f.encoding == 0;
      if (r.epf_iters > 0 && vardct && br.read_bool())
        for (float& v : r.epf_sharp_lut) v = read_f16(br);
      if (r.epf_iters > 0 && br.read_bool()) {
        for (float& v : r.epf_channel_scale) v = read_f16(br);
        read_f16(br);  // epf_pass1_zeroflush
        read_f16(br);  // epf_pass2_zeroflush
      }
      if (r.epf_iters > 0 && br.read_bool()) {
        if (vardct) r.epf_quant_mul = read_f16(br);
        r.epf_pass0_sigma_scale = read_f16(br);
        r.epf_pass2_sigma_scale = read_f16(br);
        r.epf_border_sad_mul = read_f16(br);
      }
      if (r.epf_iters > 0 && !vardct) r.epf_sigma_for_modular = read_f16(br);
      read_extensions(br);
    }
    read_extensions(br);
  }
  f.width = f.frame_width ? f.frame_width : fh.xsize;
  f.height = f.frame_height ? f.frame_height : fh.ysize;
  // postprocess (frame_header.rs:667-677)
  if (f.upsampling > 1)
    for (uint32_t i = 0; i < num_ec; i++) f.ec_upsampling[i] <<= fh.extra_channels[i].dim_shift;
  if (f.encoding != 0 || !fh.xyb_encoded) f.x_qm_scale = 2;
  br.check();
  return f;
}

Toc read_toc(BitReader& br, uint32_t num_entries) {
  Toc toc;
  bool permuted = br.read_bool();
  std::vector<uint32_t> perm;
  if (permuted) {  // encodings.rs:173-194
    EntropyCode code = EntropyCode::decode(8, br, true);
    SymbolReader reader(code, br, 0);
    perm = decode_permutation(num_entries, 0, code, br, reader);
    reader.check_final_state(br);
  }
  br.jump_to_byte_boundary();
  toc.sizes.resize(num_entries);
  for (auto& s : toc.sizes) s = u2s(br, B(10), B(14, 1024), B(22, 17408), B(30, 4211712));
  br.jump_to_byte_boundary();
  br.check();
  // frame/decode.rs:263-285: logical section i is bitstream entry perm[i].
  std::vector<uint64_t> bs_off(num_entries);
  uint64_t off = 0;
  for (uint32_t i = 0; i < num_entries; i++) {
    bs_off[i] = off;
    off += toc.sizes[i];
  }
  toc.offsets.assign(num_entries, 0);
  toc.lengths.assign(num_entries, 0);
  for (uint32_t i = 0; i < num_entries; i++) {
    uint32_t src = permuted ? perm[i] : i;
    toc.offsets[i] = bs_off[src];
    toc.lengths[i] = toc.sizes[src];
  }
  return toc;
}

}  // namespace jxg

// See entropy.h. Table construction follows the reference so that alias
// tables and prefix LUTs are entry-for-entry identical (ANS alias mapping is
// normative: encoder and decoder must agree on slot -> symbol).
#include "entropy.h"

#include <algorithm>
#include <array>

namespace jxg {

namespace {

constexpr uint32_t kSumProbs = 1u << kAnsLogSumProbs;
constexpr uint32_t kRleMarkerSym = kAnsLogSumProbs + 1;
constexpr uint32_t kHuffmanMaxBits = 15;
constexpr uint32_t kHuffTableBits = 8;

// decode.rs:20
uint32_t decode_varint16(BitReader& br) {
  if (br.read(1)) {
    uint32_t nbits = uint32_t(br.read(4));
    if (nbits == 0) return 1;
    return (1u << nbits) + uint32_t(br.read(nbits));
  }
  return 0;
}

// ans.rs:316 (read_u8)
uint32_t ans_read_u8(BitReader& br) {
  if (br.read(1)) {
    uint32_t n = uint32_t(br.read(3));
    return ((1u << n) + uint32_t(br.read(n))) & 0xff;
  }
  return 0;
}

// ans.rs:325: fixed prefix code for the log-counts of a histogram.
uint32_t ans_read_prefix(BitReader& br) {
  // (symbol, nbits) indexed by the next 7 bits. Built from the code lengths
  // {10:3, 7:3, 6:3, 8:3, 9:3, 3:4, 5:4, 4:4, 1:4, 2:4, 0:5, 11:6, 12:7, 13:7}.
  static const uint8_t kTable[128][2] = {
      {10, 3}, {12, 7}, {7, 3}, {3, 4}, {6, 3}, {8, 3}, {9, 3}, {5, 4}, {10, 3}, {4, 4}, {7, 3}, {1, 4}, {6, 3},
      {8, 3},  {9, 3},  {2, 4}, {10, 3}, {0, 5}, {7, 3}, {3, 4}, {6, 3}, {8, 3}, {9, 3}, {5, 4}, {10, 3}, {4, 4},
      {7, 3},  {1, 4},  {6, 3}, {8, 3}, {9, 3}, {2, 4}, {10, 3}, {11, 6}, {7, 3}, {3, 4}, {6, 3}, {8, 3}, {9, 3},
      {5, 4},  {10, 3}, {4, 4}, {7, 3}, {1, 4}, {6, 3}, {8, 3}, {9, 3}, {2, 4}, {10, 3}, {0, 5}, {7, 3}, {3, 4},
      {6, 3},  {8, 3},  {9, 3}, {5, 4}, {10, 3}, {4, 4}, {7, 3}, {1, 4}, {6, 3}, {8, 3}, {9, 3}, {2, 4}, {10, 3},
      {13, 7}, {7, 3},  {3, 4}, {6, 3}, {8, 3}, {9, 3}, {5, 4}, {10, 3}, {4, 4}, {7, 3}, {1, 4}, {6, 3}, {8, 3},
      {9, 3},  {2, 4},  {10, 3}, {0, 5}, {7, 3}, {3, 4}, {6, 3}, {8, 3}, {9, 3}, {5, 4}, {10, 3}, {4, 4}, {7, 3},
      {1, 4},  {6, 3},  {8, 3}, {9, 3}, {2, 4}, {10, 3}, {11, 6}, {7, 3}, {3, 4}, {6, 3}, {8, 3}, {9, 3}, {5, 4},
      {10, 3}, {4, 4},  {7, 3}, {1, 4}, {6, 3}, {8, 3}, {9, 3}, {2, 4}, {10, 3}, {0, 5}, {7, 3}, {3, 4}, {6, 3},
      {8, 3},  {9, 3},  {5, 4}, {10, 3}, {4, 4}, {7, 3}, {1, 4}, {6, 3}, {8, 3}, {9, 3}, {2, 4},
  };
  uint32_t idx = uint32_t(br.peek(7));
  br.consume(kTable[idx][1]);
  return kTable[idx][0];
}

// ans.rs:98-195 (general histogram with RLE and omitted largest count)
size_t decode_dist_complex(BitReader& br, std::vector<uint16_t>& dist) {
  size_t table_size = dist.size();
  unsigned len = 0;
  while (len < 3 && br.read(1)) len++;
  int shift = int(br.read(len)) + (1 << len) - 1;
  if (shift > 13) fail("invalid ANS histogram (shift)");
  size_t alphabet_size = ans_read_u8(br) + 3;
  if (alphabet_size > table_size) fail("invalid ANS histogram (alphabet size)");

  struct Range {
    size_t begin, end;
  };
  std::vector<Range> repeat_ranges;
  bool have_omit = false;
  uint16_t omit_log = 0;
  size_t omit_pos = 0;
  size_t idx = 0;
  while (idx < alphabet_size) {
    dist[idx] = uint16_t(ans_read_prefix(br));
    if (dist[idx] == kRleMarkerSym) {
      size_t repeat_count = ans_read_u8(br) + 4;
      if (idx + repeat_count > alphabet_size) fail("invalid ANS histogram (rle)");
      repeat_ranges.push_back({idx, idx + repeat_count});
      idx += repeat_count;
      continue;
    }
    if (!have_omit) {
      have_omit = true;
      omit_log = dist[idx];
      omit_pos = idx;
    } else if (dist[idx] > omit_log) {
      omit_log = dist[idx];
      omit_pos = idx;
    }
    idx++;
  }
  if (!have_omit) fail("invalid ANS histogram (no omit)");
  if (omit_pos + 1 < table_size && dist[omit_pos + 1] == kRleMarkerSym) fail("invalid ANS histogram (omit+rle)");

  size_t range_idx = 0;
  uint32_t acc = 0;
  uint16_t prev_dist = 0;
  for (size_t i = 0; i < table_size; i++) {
    uint16_t& code = dist[i];
    if (range_idx < repeat_ranges.size() && repeat_ranges[range_idx].begin <= i) {
      if (repeat_ranges[range_idx].end == i) {
        range_idx++;
      } else {
        code = prev_dist;
        acc += code;
        if (acc >= kSumProbs) fail("invalid ANS histogram (sum)");
        continue;
      }
    }
    if (code == 0) {
      prev_dist = 0;
      continue;
    }
    if (i == omit_pos) {
      prev_dist = 0;
      continue;
    }
    if (code > 1) {
      int zeros = int(code) - 1;
      int bitcount = std::clamp(shift - ((int(kAnsLogSumProbs) - zeros) >> 1), 0, zeros);
      code = uint16_t((1u << zeros) + (uint32_t(br.read(unsigned(bitcount))) << (zeros - bitcount)));
    }
    prev_dist = code;
    acc += code;
    if (acc >= kSumProbs) fail("invalid ANS histogram (sum)");
  }
  dist[omit_pos] = uint16_t(kSumProbs - acc);
  return alphabet_size;
}

// ans.rs:197-266. The pairing order (two LIFO stacks filled in index order)
// decides the slot->symbol mapping and must match the encoder.
void build_alias_map(size_t alphabet_size, uint32_t log_bucket_size, const std::vector<uint16_t>& dist,
                     AnsBucket* out) {
  struct Working {
    uint16_t dist, alias_symbol, alias_offset, alias_cutoff;
  };
  const uint16_t bucket_size = uint16_t(1u << log_bucket_size);
  size_t n = dist.size();
  std::vector<Working> b(n);
  for (size_t i = 0; i < n; i++) b[i] = {dist[i], uint16_t(i < alphabet_size ? i : 0), 0, dist[i]};
  std::vector<size_t> underfull, overfull;
  for (size_t i = 0; i < n; i++) {
    if (dist[i] < bucket_size) underfull.push_back(i);
    else if (dist[i] > bucket_size) overfull.push_back(i);
  }
  while (!overfull.empty() && !underfull.empty()) {
    size_t o = overfull.back();
    overfull.pop_back();
    size_t u = underfull.back();
    underfull.pop_back();
    uint16_t by = uint16_t(bucket_size - b[u].alias_cutoff);
    b[o].alias_cutoff = uint16_t(b[o].alias_cutoff - by);
    b[u].alias_symbol = uint16_t(o);
    b[u].alias_offset = b[o].alias_cutoff;
    if (b[o].alias_cutoff < bucket_size) underfull.push_back(o);
    else if (b[o].alias_cutoff > bucket_size) overfull.push_back(o);
  }
  if (!overfull.empty() || !underfull.empty()) fail("ANS alias table construction failed");
  for (size_t i = 0; i < n; i++) {
    if (b[i].alias_cutoff == bucket_size) {
      out[i] = {uint8_t(i), 0, b[i].dist, 0, 0};
    } else {
      out[i] = {uint8_t(b[i].alias_symbol), uint8_t(b[i].alias_cutoff), b[i].dist,
                uint16_t(b[i].alias_offset - b[i].alias_cutoff), uint16_t(b[i].dist ^ b[b[i].alias_symbol].dist)};
    }
  }
}

// ans.rs:269-314
int32_t decode_ans_histogram(BitReader& br, uint32_t log_alpha_size, AnsBucket* out) {
  size_t table_size = size_t(1) << log_alpha_size;
  uint32_t log_bucket_size = kAnsLogSumProbs - log_alpha_size;
  uint16_t bucket_size = uint16_t(1u << log_bucket_size);
  std::vector<uint16_t> dist(table_size, 0);
  size_t alphabet_size;
  if (br.read(1)) {
    if (br.read(1)) {  // two symbols
      size_t v0 = ans_read_u8(br), v1 = ans_read_u8(br);
      if (v0 == v1) fail("invalid ANS histogram (two equal symbols)");
      alphabet_size = std::max(v0, v1) + 1;
      if (alphabet_size > table_size) fail("invalid ANS histogram (alphabet size)");
      uint16_t prob = uint16_t(br.read(kAnsLogSumProbs));
      dist[v0] = prob;
      dist[v1] = uint16_t(kSumProbs - prob);
    } else {  // single symbol
      size_t v = ans_read_u8(br);
      alphabet_size = v + 1;
      if (alphabet_size > table_size) fail("invalid ANS histogram (alphabet size)");
      dist[v] = uint16_t(kSumProbs);
    }
  } else if (br.read(1)) {  // flat
    alphabet_size = ans_read_u8(br) + 1;
    if (alphabet_size > table_size) fail("invalid ANS histogram (alphabet size)");
    size_t base = kSumProbs / alphabet_size, rem = kSumProbs % alphabet_size;
    for (size_t i = 0; i < alphabet_size; i++) dist[i] = uint16_t(base + (i < rem ? 1 : 0));
  } else {
    alphabet_size = decode_dist_complex(br, dist);
  }
  int32_t single = -1;
  for (size_t i = 0; i < table_size; i++)
    if (dist[i] == kSumProbs) {
      single = int32_t(i);
      break;
    }
  if (single >= 0) {
    for (size_t i = 0; i < table_size; i++)
      out[i] = {uint8_t(single), 0, dist[i], uint16_t(bucket_size * i), uint16_t(dist[i] ^ kSumProbs)};
  } else {
    build_alias_map(alphabet_size, log_bucket_size, dist, out);
  }
  return single;
}

// ---- prefix codes (huffman.rs; the Brotli canonical-code table builder) ----

struct HE {
  uint8_t bits;
  uint16_t value;
};

uint32_t get_next_key(uint32_t key, uint32_t len) {
  uint32_t step = 1u << (len - 1);
  while (key & step) step >>= 1;
  return (key & (step - 1)) + step;
}
uint32_t next_table_bit_size(const uint16_t* count, uint32_t len, uint32_t root_bits) {
  int left = 1 << (len - root_bits);
  while (len < kHuffmanMaxBits) {
    if (left <= int(count[len])) break;
    left -= count[len];
    len++;
    left <<= 1;
  }
  return len - root_bits;
}

// huffman.rs:276-400
std::vector<HE> build_huffman(uint32_t root_bits, const std::vector<uint8_t>& code_lengths) {
  if (code_lengths.size() > (1u << kHuffmanMaxBits)) fail("invalid prefix code");
  uint16_t counts[kHuffmanMaxBits + 1] = {0};
  for (uint8_t v : code_lengths) counts[v]++;
  std::vector<uint16_t> sorted(code_lengths.size(), 0);
  uint32_t offset[kHuffmanMaxBits + 1] = {0};
  uint32_t max_length = 1;
  {
    uint32_t sum = 0;
    for (uint32_t len = 1; len <= kHuffmanMaxBits; len++) {
      offset[len] = sum;
      if (counts[len]) {
        sum += counts[len];
        max_length = len;
      }
    }
  }
  for (size_t s = 0; s < code_lengths.size(); s++) {
    uint8_t len = code_lengths[s];
    if (len) sorted[offset[len]++] = uint16_t(s);
  }
  uint32_t table_bits = root_bits;
  size_t table_size = size_t(1) << table_bits;
  size_t table_pos = 0;
  std::vector<HE> table(table_size, HE{0, 0});
  if (offset[kHuffmanMaxBits] == 1) {  // single symbol
    for (auto& e : table) e = {0, sorted[0]};
    return table;
  }
  if (table_bits > max_length) {
    table_bits = max_length;
    table_size = size_t(1) << table_bits;
  }
  uint32_t key = 0;
  size_t symbol = 0;
  uint32_t bits = 1;
  size_t step = 2;
  do {
    while (counts[bits]) {
      HE v{uint8_t(bits), sorted[symbol++]};
      // fill within the (possibly reduced) root table only
      for (size_t i = key; i < table_size; i += step) table[i] = v;
      key = get_next_key(key, bits);
      counts[bits]--;
    }
    step <<= 1;
    bits++;
  } while (bits <= table_bits);
  while (table.size() != table_size) {
    for (size_t i = 0; i < table_size; i++) table[i + table_size] = table[i];
    table_size <<= 1;
  }
  uint32_t mask = uint32_t(table.size() - 1);
  uint32_t low = ~0u;
  step = 2;
  for (uint32_t len = root_bits + 1; len <= max_length; len++) {
    while (counts[len]) {
      if ((key & mask) != low) {
        table_pos += table_size;
        table_bits = next_table_bit_size(counts, len, root_bits);
        table_size = size_t(1) << table_bits;
        low = key & mask;
        table[low].bits = uint8_t(table_bits + root_bits);
        table[low].value = uint16_t(table_pos - low);
        if (table.size() < table_pos + table_size) table.resize(table_pos + table_size, HE{0, 0});
      }
      counts[len]--;
      HE v{uint8_t(len - root_bits), sorted[symbol++]};
      size_t pos = table_pos + (key >> root_bits);
      // replicate inside the current 2nd-level table
      for (size_t i = pos; i < table_pos + table_size; i += step) table[i] = v;
      key = get_next_key(key, len);
    }
    step <<= 1;
  }
  return table;
}

// huffman.rs:65-188
std::vector<HE> decode_simple_table(size_t al_size, BitReader& br) {
  uint32_t max_bits = ceil_log2(al_size);
  size_t num_symbols = size_t(br.read(2)) + 1;
  uint16_t symbols[4] = {0, 0, 0, 0};
  for (size_t i = 0; i < num_symbols; i++) {
    size_t sym = size_t(br.read(max_bits));
    if (sym >= al_size) fail("invalid prefix code (symbol)");
    symbols[i] = uint16_t(sym);
  }
  for (size_t i = 0; i + 1 < num_symbols; i++)
    for (size_t j = 0; j < i; j++)
      if (symbols[j] == symbols[i + 1]) fail("invalid prefix code (duplicate)");
  bool special4 = num_symbols == 4 ? br.read(1) != 0 : false;
  const size_t kSize = size_t(1) << kHuffTableBits;
  std::vector<HE> ret(kSize);
  auto fill = [&](std::initializer_list<HE> pattern) {
    size_t n = pattern.size(), i = 0;
    while (i < kSize)
      for (const HE& e : pattern) {
        ret[i++] = e;
        (void)n;
      }
  };
  switch (num_symbols) {
    case 1:
      fill({HE{0, symbols[0]}});
      break;
    case 2:
      std::sort(symbols, symbols + 2);
      fill({HE{1, symbols[0]}, HE{1, symbols[1]}});
      break;
    case 3:
      std::sort(symbols + 1, symbols + 3);
      fill({HE{1, symbols[0]}, HE{2, symbols[1]}, HE{1, symbols[0]}, HE{2, symbols[2]}});
      break;
    default:
      if (!special4) {
        std::sort(symbols, symbols + 4);
        fill({HE{2, symbols[0]}, HE{2, symbols[2]}, HE{2, symbols[1]}, HE{2, symbols[3]}});
      } else {
        std::sort(symbols + 2, symbols + 4);
        fill({HE{1, symbols[0]}, HE{2, symbols[1]}, HE{1, symbols[0]}, HE{3, symbols[2]}, HE{1, symbols[0]},
              HE{2, symbols[1]}, HE{1, symbols[0]}, HE{3, symbols[3]}});
      }
  }
  return ret;
}

// huffman.rs:190-274
std::vector<uint8_t> decode_code_lengths(const uint8_t (&cl_cl)[18], size_t al_size, BitReader& br) {
  std::vector<uint8_t> clv(cl_cl, cl_cl + 18);
  std::vector<HE> table = build_huffman(5, clv);
  size_t symbol = 0;
  uint8_t prev_code_len = 8;
  size_t repeat = 0;
  uint8_t repeat_code_len = 0;
  int64_t space = 1 << 15;
  std::vector<uint8_t> code_lengths(al_size, 0);
  while (symbol < al_size && space > 0) {
    size_t idx = size_t(br.peek(5));
    br.consume(table[idx].bits);
    uint8_t code_len = uint8_t(table[idx].value);
    if (code_len < 16) {
      repeat = 0;
      code_lengths[symbol++] = code_len;
      if (code_len != 0) {
        prev_code_len = code_len;
        space -= 32768 >> code_len;
        if (space < 0) fail("invalid prefix code (space)");
      }
    } else {
      uint32_t extra_bits = code_len - 14;
      uint8_t new_len = code_len == 16 ? prev_code_len : 0;
      if (repeat_code_len != new_len) {
        repeat = 0;
        repeat_code_len = new_len;
      }
      size_t old_repeat = repeat;
      if (repeat > 0) {
        repeat -= 2;
        repeat <<= extra_bits;
      }
      repeat += size_t(br.read(extra_bits)) + 3;
      size_t delta = repeat - old_repeat;
      if (symbol + delta > al_size) fail("invalid prefix code (repeat)");
      for (size_t i = 0; i < delta; i++) code_lengths[symbol + i] = repeat_code_len;
      symbol += delta;
      if (repeat_code_len != 0) {
        space -= int64_t(delta) << (15 - repeat_code_len);
        if (space < 0) fail("invalid prefix code (space)");
      }
    }
  }
  if (space != 0) fail("invalid prefix code (incomplete)");
  return code_lengths;
}

// huffman.rs:404-443
std::vector<HE> decode_huffman_table(size_t al_size, BitReader& br) {
  if (al_size == 1) return std::vector<HE>(size_t(1) << kHuffTableBits, HE{0, 0});
  uint32_t simple_code_or_skip = uint32_t(br.read(2));
  if (simple_code_or_skip == 1) return decode_simple_table(al_size, br);
  uint8_t cl_cl[18] = {0};
  int space = 32;
  static const uint8_t kStaticBits[16] = {2, 2, 2, 3, 2, 2, 2, 4, 2, 2, 2, 3, 2, 2, 2, 4};
  static const uint8_t kStaticVals[16] = {0, 4, 3, 2, 0, 4, 3, 1, 0, 4, 3, 2, 0, 4, 3, 5};
  static const uint8_t kOrder[18] = {1, 2, 3, 4, 0, 5, 17, 6, 16, 7, 8, 9, 10, 11, 12, 13, 14, 15};
  int num_codes = 0;
  for (uint32_t i = simple_code_or_skip; i < 18; i++) {
    if (space <= 0) break;
    size_t idx = size_t(br.peek(4));
    br.consume(kStaticBits[idx]);
    uint8_t v = kStaticVals[idx];
    cl_cl[kOrder[i]] = v;
    if (v != 0) {
      space -= 32 >> v;
      num_codes++;
    }
  }
  if (num_codes != 1 && space != 0) fail("invalid prefix code (code-length code)");
  std::vector<uint8_t> code_lengths = decode_code_lengths(cl_cl, al_size, br);
  return build_huffman(kHuffTableBits, code_lengths);
}

}  // namespace

HybridUint HybridUint::decode(uint32_t log_alpha_size, BitReader& br) {
  HybridUint u;
  u.split_exponent = uint32_t(br.read(ceil_log2(log_alpha_size + 1)));
  if (u.split_exponent != log_alpha_size) {
    u.msb = uint32_t(br.read(ceil_log2(u.split_exponent + 1)));
    if (u.msb > u.split_exponent) fail("invalid hybrid uint config");
    u.lsb = uint32_t(br.read(ceil_log2(u.split_exponent - u.msb + 1)));
  }
  if (u.lsb + u.msb > u.split_exponent) fail("invalid hybrid uint config");
  return u;
}

std::vector<uint8_t> decode_context_map(size_t num_contexts, BitReader& br) {
  std::vector<uint8_t> map(num_contexts, 0);
  bool is_simple = br.read(1);
  if (is_simple) {
    uint32_t bits = uint32_t(br.read(2));
    if (bits)
      for (auto& m : map) m = uint8_t(br.read(bits));
    // NOTE: the reference does not verify "no holes" for simple maps either.
    return map;
  }
  bool use_mtf = br.read(1);
  EntropyCode code = EntropyCode::decode(1, br, num_contexts > 2);
  SymbolReader reader(code, br, 0);
  for (auto& m : map) {
    uint32_t v = reader.read_unsigned(br, 0);
    if (v > 255) fail("invalid context map entry");
    m = uint8_t(v);
  }
  reader.check_final_state(br);
  if (use_mtf) {  // context_map.rs:18-27
    uint8_t mtf[256];
    for (int i = 0; i < 256; i++) mtf[i] = uint8_t(i);
    for (auto& m : map) {
      uint8_t index = m;
      uint8_t value = mtf[index];
      m = value;
      if (index) {
        for (int i = index; i > 0; i--) mtf[i] = mtf[i - 1];
        mtf[0] = value;
      }
    }
  }
  // context_map.rs:31-41: every cluster id below the maximum must be used.
  bool seen[256] = {false};
  uint32_t mx = 0;
  for (uint8_t m : map) {
    seen[m] = true;
    mx = std::max<uint32_t>(mx, m);
  }
  for (uint32_t i = 0; i <= mx; i++)
    if (!seen[i]) fail("context map has holes");
  return map;
}

EntropyCode EntropyCode::decode(size_t num_contexts, BitReader& br, bool allow_lz77) {
  EntropyCode c;
  // Lz77Params (decode.rs:35-45)
  c.lz77_enabled = br.read(1);
  if (c.lz77_enabled) {
    switch (br.read(2)) {
      case 0: c.lz77_min_symbol = 224; break;
      case 1: c.lz77_min_symbol = 512; break;
      case 2: c.lz77_min_symbol = 4096; break;
      default: c.lz77_min_symbol = uint32_t(br.read(15)) + 8;
    }
    switch (br.read(2)) {
      case 0: c.lz77_min_length = 3; break;
      case 1: c.lz77_min_length = 4; break;
      case 2: c.lz77_min_length = uint32_t(br.read(2)) + 5; break;
      default: c.lz77_min_length = uint32_t(br.read(8)) + 9;
    }
  }
  if (!allow_lz77 && c.lz77_enabled) fail("LZ77 not allowed here");
  if (c.lz77_enabled) {
    num_contexts += 1;
    c.lz77_length_uint = HybridUint::decode(8, br);
  }
  if (num_contexts > 1) c.context_map = decode_context_map(num_contexts, br);
  else c.context_map.assign(1, 0);
  if (c.lz77_enabled) c.lz_dist_cluster = c.context_map.back();
  c.use_prefix = br.read(1);
  c.log_alpha_size = c.use_prefix ? kHuffmanMaxBits : uint32_t(br.read(2)) + 5;
  c.num_clusters = uint32_t(*std::max_element(c.context_map.begin(), c.context_map.end())) + 1;
  c.uint_configs.resize(c.num_clusters);
  for (auto& u : c.uint_configs) u = HybridUint::decode(c.log_alpha_size, br);
  c.single_symbol.assign(c.num_clusters, -1);
  if (c.use_prefix) {
    std::vector<size_t> al(c.num_clusters);
    size_t mx = 0;
    for (auto& a : al) {
      a = decode_varint16(br) + 1;
      mx = std::max(mx, a);
    }
    if (mx >= (1u << kHuffmanMaxBits)) fail("prefix alphabet too large");
    c.huff_offset.resize(c.num_clusters);
    for (uint32_t i = 0; i < c.num_clusters; i++) {
      std::vector<HE> t = decode_huffman_table(al[i], br);
      c.huff_offset[i] = uint32_t(c.huff_entries.size());
      for (const HE& e : t) c.huff_entries.push_back(uint32_t(e.bits) | (uint32_t(e.value) << 16));
      if (t[0].bits == 0) c.single_symbol[i] = t[0].value;
    }
  } else {
    c.ans_buckets.resize(size_t(c.num_clusters) << c.log_alpha_size);
    for (uint32_t i = 0; i < c.num_clusters; i++)
      c.single_symbol[i] =
          decode_ans_histogram(br, c.log_alpha_size, &c.ans_buckets[size_t(i) << c.log_alpha_size]);
  }
  br.check();
  return c;
}

bool EntropyCode::is_rle() const {
  return single_symbol[lz_dist_cluster] == 1 && uint_configs[lz_dist_cluster].split_exponent == 0;
}

// ---------------------------------------------------------------------------

SymbolReader::SymbolReader(const EntropyCode& code, BitReader& br, size_t dist_multiplier)
    : code_(code), dist_multiplier_(uint32_t(dist_multiplier)) {
  if (!code.use_prefix) state_ = uint32_t(br.read(32));  // ans.rs:431
}

uint32_t SymbolReader::read_clustered_lz77(BitReader& br, uint32_t cluster) {
  constexpr uint32_t kWindowMask = (1u << 20) - 1;
  auto push = [&](uint32_t v) {
    size_t off = num_decoded_ & kWindowMask;
    if (off < window_.size()) window_[off] = v;
    else window_.push_back(v);
    num_decoded_++;
  };
  if (num_to_copy_ > 0) {
    uint32_t sym = window_[copy_pos_ & kWindowMask];
    copy_pos_++;
    num_to_copy_--;
    push(sym);
    return sym;
  }
  uint32_t token = read_token(br, cluster);
  if (token < code_.lz77_min_symbol) {
    uint32_t sym = code_.uint_configs[cluster].read(token, br);
    push(sym);
    return sym;
  }
  if (num_decoded_ == 0) {
    err_lz77_repeat_ = true;
    return 0;
  }
  uint32_t num_to_copy = code_.lz77_length_uint.read(token - code_.lz77_min_symbol, br);
  if (num_to_copy > 0xffffffffu - code_.lz77_min_length) {
    err_overflow_ = true;
    return 0;
  }
  num_to_copy += code_.lz77_min_length;
  uint32_t dc = code_.lz_dist_cluster;
  uint32_t distance_sym = code_.uint_configs[dc].read(read_token(br, dc), br);
  // decode.rs:103-118 with SPECIAL_DISTANCES (decode.rs:87-101)
  static const int8_t kSpecial[120][2] = {
      {0, 1},  {1, 0},  {1, 1},  {-1, 1}, {0, 2},  {2, 0},  {1, 2},  {-1, 2}, {2, 1},  {-2, 1}, {2, 2},  {-2, 2},
      {0, 3},  {3, 0},  {1, 3},  {-1, 3}, {3, 1},  {-3, 1}, {2, 3},  {-2, 3}, {3, 2},  {-3, 2}, {0, 4},  {4, 0},
      {1, 4},  {-1, 4}, {4, 1},  {-4, 1}, {3, 3},  {-3, 3}, {2, 4},  {-2, 4}, {4, 2},  {-4, 2}, {0, 5},  {3, 4},
      {-3, 4}, {4, 3},  {-4, 3}, {5, 0},  {1, 5},  {-1, 5}, {5, 1},  {-5, 1}, {2, 5},  {-2, 5}, {5, 2},  {-5, 2},
      {4, 4},  {-4, 4}, {3, 5},  {-3, 5}, {5, 3},  {-5, 3}, {0, 6},  {6, 0},  {1, 6},  {-1, 6}, {6, 1},  {-6, 1},
      {2, 6},  {-2, 6}, {6, 2},  {-6, 2}, {4, 5},  {-4, 5}, {5, 4},  {-5, 4}, {3, 6},  {-3, 6}, {6, 3},  {-6, 3},
      {0, 7},  {7, 0},  {1, 7},  {-1, 7}, {5, 5},  {-5, 5}, {7, 1},  {-7, 1}, {4, 6},  {-4, 6}, {6, 4},  {-6, 4},
      {2, 7},  {-2, 7}, {7, 2},  {-7, 2}, {3, 7},  {-3, 7}, {7, 3},  {-7, 3}, {5, 6},  {-5, 6}, {6, 5},  {-6, 5},
      {8, 0},  {4, 7},  {-4, 7}, {7, 4},  {-7, 4}, {8, 1},  {8, 2},  {6, 6},  {-6, 6}, {8, 3},  {5, 7},  {-5, 7},
      {7, 5},  {-7, 5}, {8, 4},  {6, 7},  {-6, 7}, {7, 6},  {-7, 6}, {8, 5},  {7, 7},  {-7, 7}, {8, 6},  {8, 7},
  };
  uint32_t distance_sub_1;
  if (dist_multiplier_ == 0) {
    distance_sub_1 = distance_sym;
  } else if (distance_sym >= 120) {
    distance_sub_1 = distance_sym - 120;
  } else {
    int64_t d = int64_t(dist_multiplier_) * kSpecial[distance_sym][1] + (kSpecial[distance_sym][0] - 1);
    // u32 arithmetic in the reference: multiplier*dist wraps, then checked add.
    uint32_t prod = uint32_t(uint64_t(dist_multiplier_) * uint64_t(uint8_t(kSpecial[distance_sym][1])));
    int64_t sum = int64_t(prod) + (kSpecial[distance_sym][0] - 1);
    (void)d;
    distance_sub_1 = (sum < 0 || sum > 0xffffffffLL) ? 0 : uint32_t(sum);
  }
  uint32_t distance = std::min(std::min<uint32_t>((1u << 20) - 1, distance_sub_1) + 1, num_decoded_);
  copy_pos_ = num_decoded_ - distance;
  num_to_copy_ = num_to_copy;
  // pull one
  uint32_t sym = window_[copy_pos_ & kWindowMask];
  copy_pos_++;
  num_to_copy_--;
  push(sym);
  return sym;
}

void SymbolReader::check_final_state(BitReader& br) const {
  if (err_lz77_repeat_) fail("unexpected LZ77 repeat");
  if (err_overflow_) fail("LZ77 arithmetic overflow");
  br.check();
  if (!code_.use_prefix && state_ != kAnsChecksum) fail("ANS checksum mismatch");
}

// ---------------------------------------------------------------------------

static inline size_t perm_context(uint32_t x) { return std::min<uint32_t>(ceil_log2(uint64_t(x) + 1), 7); }

std::vector<uint32_t> decode_permutation(uint32_t size, uint32_t skip, const EntropyCode& code, BitReader& br,
                                         SymbolReader& reader) {
  (void)code;
  uint32_t end = reader.read_unsigned(br, perm_context(size));
  if (end > size - skip) fail("invalid permutation size");
  std::vector<uint32_t> lehmer(end);
  uint32_t prev = 0;
  for (uint32_t idx = skip; idx < skip + end; idx++) {
    uint32_t v = reader.read_unsigned(br, perm_context(prev));
    br.check();
    if (v >= size - idx) fail("invalid Lehmer code");
    lehmer[idx - skip] = v;
    prev = v;
  }
  return apply_lehmer(lehmer, skip, size);
}

std::vector<uint32_t> apply_lehmer(const std::vector<uint32_t>& lehmer, uint32_t skip, uint32_t size) {
  std::vector<uint32_t> perm(size);
  for (uint32_t i = 0; i < size; i++) perm[i] = i;
  // Plain O(n*k) Lehmer decode of the tail [skip, size): element i takes the
  // lehmer[i]-th remaining value (permutation.rs:103-160 does the same with a
  // Fenwick tree).
  std::vector<uint32_t> remaining(perm.begin() + skip, perm.end());
  uint32_t end = uint32_t(lehmer.size());
  for (uint32_t i = 0; i < end; i++) {
    uint32_t k = lehmer[i];
    if (k >= remaining.size()) fail("invalid Lehmer code");
    perm[skip + i] = remaining[k];
    remaining.erase(remaining.begin() + k);
  }
  for (size_t i = 0; i < remaining.size(); i++) perm[skip + end + i] = remaining[i];
  return perm;
}

int32_t EntropyCode::decode_ans_histogram_for_test(BitReader& br, uint32_t log_alpha_size, std::vector<AnsBucket>& out) {
  out.resize(size_t(1) << log_alpha_size);
  int32_t s = decode_ans_histogram(br, log_alpha_size, out.data());
  br.check();
  return s;
}

EntropyCode EntropyCode::decode_prefix_codes_for_test(size_t num_clusters, BitReader& br) {
  EntropyCode c;
  c.use_prefix = true;
  c.log_alpha_size = kHuffmanMaxBits;
  c.num_clusters = uint32_t(num_clusters);
  c.context_map.assign(num_clusters, 0);
  for (size_t i = 0; i < num_clusters; i++) c.context_map[i] = uint8_t(i);
  c.uint_configs.assign(num_clusters, HybridUint{15, 0, 0});
  c.single_symbol.assign(num_clusters, -1);
  std::vector<size_t> al(num_clusters);
  for (auto& a : al) a = decode_varint16(br) + 1;
  c.huff_offset.resize(num_clusters);
  for (uint32_t i = 0; i < num_clusters; i++) {
    std::vector<HE> t = decode_huffman_table(al[i], br);
    c.huff_offset[i] = uint32_t(c.huff_entries.size());
    for (const HE& e : t) c.huff_entries.push_back(uint32_t(e.bits) | (uint32_t(e.value) << 16));
  }
  return c;
}

}  // namespace jxg

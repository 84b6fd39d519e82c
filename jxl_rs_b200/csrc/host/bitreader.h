// Host-side LSB-first bit reader for the JPEG XL front-end (headers, TOC,
// entropy tables, LF / HF-metadata modular streams).
//
// Behaviour follows the reference bit reader (jxl/src/bit_reader.rs:15-219):
// little-endian bytes, bits consumed LSB first, at most 56 bits per call,
// reads past the end return zero bits and are detected afterwards through
// `overrun()` (the reference's `check_for_error`, bit_reader.rs:109).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>

namespace jxg {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};

// Error codes shared with include/jxg.h (JXG_ERR_*).
enum : int {
  kErrBitstream = -1,
  kErrUnsupported = -2,
  kErrOutOfBounds = -3,
};

[[noreturn]] inline void fail(const std::string& what, int code = kErrBitstream) { throw Error(code, what); }

class BitReader {
 public:
  BitReader() = default;
  BitReader(const uint8_t* data, size_t size) : data_(data), size_(size) {}

  // Peek `n` (<= 56) bits without consuming.
  inline uint64_t peek(unsigned n) {
    if (bits_ < n) refill();
    return buf_ & ((uint64_t(1) << n) - 1);
  }
  inline void consume(unsigned n) {
    buf_ >>= n;
    bits_ = bits_ >= n ? bits_ - n : 0;
    total_ += n;
  }
  inline uint64_t read(unsigned n) {
    uint64_t v = peek(n);
    consume(n);
    return v;
  }
  inline bool read_bool() { return read(1) != 0; }

  size_t total_bits_read() const { return total_; }
  size_t size_bits() const { return size_ * 8; }
  bool overrun() const { return total_ > size_ * 8; }
  void check() const {
    if (overrun()) fail("bitstream over-read", kErrOutOfBounds);
  }

  void skip_bits(size_t n) {
    while (n > 0) {
      unsigned step = n > 56 ? 56 : unsigned(n);
      read(step);
      n -= step;
    }
  }
  // headers/encodings.rs:194, bit_reader.rs:195 — padding bits must be zero.
  void jump_to_byte_boundary() {
    unsigned n = unsigned((8 - (total_ & 7)) & 7);
    if (read(n) != 0) fail("non-zero padding");
  }
  // Byte position (only valid on a byte boundary).
  size_t byte_pos() const { return total_ / 8; }

  // Register-resident copy of the reader for the per-pixel Modular loops (the members of a BitReader reached
  // through a reference are re-loaded and re-stored around every int32 sample store, which puts store-to-load
  // forwarding on the symbol chain). Same semantics as the member functions above; `commit` writes it back.
  struct Local {
    const uint8_t* data;
    size_t size, pos;
    uint64_t buf;
    unsigned bits;
    // Guarantees >= 56 valid bits (zero bits past the end, like refill()).
    inline void fill() {
      if (pos + 8 <= size) {
        uint64_t w;
        memcpy(&w, data + pos, 8);
        buf |= w << bits;
        const unsigned nbytes = (63 - bits) >> 3;
        pos += nbytes;
        bits += nbytes * 8;
        return;
      }
      while (bits <= 56) {
        uint64_t byte = pos < size ? data[pos] : 0;
        pos++;
        buf |= byte << bits;
        bits += 8;
      }
    }
    // fill() without the end-of-data test; only valid while room() holds for the symbols still to be read.
    inline void fill_unchecked() {
      uint64_t w;
      memcpy(&w, data + pos, 8);
      buf |= w << bits;
      const unsigned nbytes = (63 - bits) >> 3;
      pos += nbytes;
      bits += nbytes * 8;
    }
    // True if `nsym` more symbols of at most 48 bits each can be read with fill_unchecked(): the reader never
    // holds more than 8 bytes beyond the consumed position, and every load is 8 bytes wide.
    inline bool room(size_t nsym) const { return pos + nsym * 6 + 16 <= size; }
    inline uint64_t peek(unsigned n) const { return buf & ((uint64_t(1) << n) - 1); }  // caller keeps bits >= n
    inline void consume(unsigned n) {
      buf >>= n;
      bits -= n;
    }
  };
  // Bits consumed so far == position of the next unread bit (see commit() below for why this holds).
  size_t bit_pos() const { return pos_ * 8 - bits_; }
  const uint8_t* data() const { return data_; }
  size_t size_bytes() const { return size_; }
  // Repositions the reader at absolute bit `bitpos` of its data (may lie past the end: the over-read is then
  // reported by check(), like after reading there).
  void seek_bits(size_t bitpos) {
    pos_ = bitpos / 8;
    buf_ = 0;
    bits_ = 0;
    total_ = pos_ * 8;
    read(unsigned(bitpos % 8));
  }

  // The count of consumed bits is not carried by Local: every byte brought into the window advances pos and bits
  // together (zero bytes past the end included), so the bits consumed through a Local are the change of
  // 8 * pos - bits.
  Local local() const { return Local{data_, size_, pos_, buf_, bits_}; }
  void commit(const Local& l) {
    total_ += (l.pos * 8 - l.bits) - (pos_ * 8 - bits_);
    pos_ = l.pos;
    buf_ = l.buf;
    bits_ = l.bits;
  }

 private:
  inline void refill() {
    if (pos_ + 8 <= size_) {  // whole-word refill (little-endian host); bits above bits_ are re-read identically later
      uint64_t w;
      memcpy(&w, data_ + pos_, 8);
      buf_ |= w << bits_;
      const unsigned nbytes = (63 - bits_) >> 3;
      pos_ += nbytes;
      bits_ += nbytes * 8;
      return;
    }
    while (bits_ <= 56) {
      uint64_t byte = pos_ < size_ ? data_[pos_] : 0;
      pos_++;
      buf_ |= byte << bits_;
      bits_ += 8;
    }
  }
  const uint8_t* data_ = nullptr;
  size_t size_ = 0;
  size_t pos_ = 0;
  uint64_t buf_ = 0;
  unsigned bits_ = 0;
  size_t total_ = 0;
};

inline uint32_t floor_log2(uint64_t x) {  // 0 for x <= 1
  return x > 1 ? 63u - uint32_t(__builtin_clzll(x)) : 0u;
}
inline uint32_t ceil_log2(uint64_t x) {  // util CeilLog2: smallest n with 2^n >= x
  return x > 1 ? floor_log2(x - 1) + 1 : 0u;
}
// entropy_coding/decode.rs:31
inline int32_t unpack_signed(uint32_t u) { return int32_t((u >> 1) ^ (((~u) & 1) - 1)); }

}  // namespace jxg

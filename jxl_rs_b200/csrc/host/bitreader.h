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

inline uint32_t ceil_log2(uint64_t x) {  // util CeilLog2: smallest n with 2^n >= x
  uint32_t n = 0;
  while ((uint64_t(1) << n) < x) n++;
  return n;
}
inline uint32_t floor_log2(uint64_t x) {
  uint32_t n = 0;
  while (x >>= 1) n++;
  return n;
}
// entropy_coding/decode.rs:31
inline int32_t unpack_signed(uint32_t u) { return int32_t((u >> 1) ^ (((~u) & 1) - 1)); }

}  // namespace jxg

// RbspReader — bit reader over an un-escaped RBSP plus emulation-prevention removal.
// Host-side restatement of the reference's src/hevc/rbsp_reader.rs (same method names and the
// same error behaviour: reads past the end fail with "unexpected EOF" instead of panicking).
#pragma once
#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "heic_b200.h"

namespace heic {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};
[[noreturn]] inline void bail(int code, const std::string& msg) { throw Error(code, msg); }
inline void ensure(bool ok, int code, const char* msg) {
  if (!ok) throw Error(code, msg);
}

class RbspReader {
 public:
  RbspReader(const uint8_t* data, size_t len) : data_(data), len_(len) {}

  // rbsp_reader.rs:11-39.  The 0x03 of a 00 00 03 triple is dropped only when the following byte is
  // <= 0x03 or the triple ends the payload (pinned by the reference's tests at rbsp_reader.rs:235-249).
  // epb (optional) receives the position, in `data`, of every removed byte — the reference never
  // needs this because it never converts entry points; we do.
  static std::vector<uint8_t> remove_emulation_prevention(const uint8_t* data, size_t len,
                                                          std::vector<uint32_t>* epb = nullptr) {
    std::vector<uint8_t> out;
    out.reserve(len);
    size_t i = 0;
    while (i < len) {
      uint8_t b = data[i];
      if (b == 0 && i + 2 < len && data[i + 1] == 0 && data[i + 2] == 3 &&
          (i + 3 >= len || data[i + 3] <= 3)) {
        out.push_back(0);
        out.push_back(0);
        if (epb) epb->push_back(static_cast<uint32_t>(i + 2));
        i += 3;
      } else {
        out.push_back(b);
        i += 1;
      }
    }
    return out;
  }

  bool is_byte_aligned() const { return bit_pos_ == 0; }
  size_t byte_position() const { return byte_pos_; }
  unsigned bit_position() const { return bit_pos_; }
  size_t bits_left() const { return (len_ - byte_pos_) * 8 - bit_pos_; }

  uint32_t read_bit() {
    if (byte_pos_ >= len_) bail(HEIC_E_BITSTREAM, "unexpected EOF");
    uint32_t bit = (data_[byte_pos_] >> (7 - bit_pos_)) & 1u;
    if (++bit_pos_ == 8) {
      bit_pos_ = 0;
      ++byte_pos_;
    }
    return bit;
  }
  bool read_flag() { return read_bit() == 1; }
  uint64_t read_bits(unsigned n) {
    uint64_t v = 0;
    for (unsigned i = 0; i < n; ++i) v = (v << 1) | read_bit();
    return v;
  }
  uint8_t read_u8(unsigned n) {
    if (n > 8) bail(HEIC_E_INVALID_ARG, "cannot read more than 8 bits into u8");
    return static_cast<uint8_t>(read_bits(n));
  }
  uint32_t read_u32(unsigned n) {
    if (n > 32) bail(HEIC_E_INVALID_ARG, "cannot read more than 32 bits into u32");
    return static_cast<uint32_t>(read_bits(n));
  }
  // 9.2 ue(v) / se(v); rbsp_reader.rs:87-118, vectors at rbsp_reader.rs:144-184.
  uint32_t read_ue() {
    unsigned leading = 0;
    while (!read_flag()) {
      if (++leading > 32) bail(HEIC_E_BITSTREAM, "ue(v) prefix longer than 32 bits");
    }
    if (leading == 0) return 0;
    uint64_t suffix = read_bits(leading);
    return static_cast<uint32_t>(((uint64_t{1} << leading) - 1) + suffix);
  }
  int32_t read_se() {
    uint32_t k = read_ue();
    if (k == 0) return 0;
    return (k & 1) ? static_cast<int32_t>((k + 1) / 2) : -static_cast<int32_t>(k / 2);
  }
  // 7.3.1.6 byte_alignment(): one 1 bit then zeros to the byte boundary (rbsp_reader.rs:53-63).
  void byte_alignment() {
    if (read_bit() != 1) bail(HEIC_E_BITSTREAM, "byte_alignment: missing alignment_bit_equal_to_one");
    while (!is_byte_aligned())
      if (read_bit() != 0) bail(HEIC_E_BITSTREAM, "byte_alignment: non-zero alignment bit");
  }

 private:
  const uint8_t* data_;
  size_t len_;
  size_t byte_pos_ = 0;
  unsigned bit_pos_ = 0;
};

}  // namespace heic

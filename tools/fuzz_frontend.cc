// Dev tool: mutation fuzzing of the host front-end under AddressSanitizer / UBSan.
//   H=jxl_rs_b200/csrc/host
//   g++ -std=c++17 -O1 -g -fsanitize=address,undefined -fno-sanitize-recover=undefined -march=x86-64-v3 \
//       -ffp-contract=off -Ioracle -o /tmp/fuzz_frontend tools/fuzz_frontend.cc oracle/modular_oracle.cc oracle/oracle.cc \
//       $H/entropy.cc $H/headers.cc $H/modular.cc $H/quant.cc $H/frame.cc $H/modular_frame.cc -pthread
//   /tmp/fuzz_frontend 1000 tests/golden/jxl/*.jxl
// Every input is mutated `iters` times (bit flips, random bytes, 0xff bytes, truncation; a third of the mutations land
// in the first 2 KB, where the headers and entropy tables live), copied into an exact-size heap block so that
// over-reads of the input trip ASan, and run through parse_vardct_file or — for Modular frames — parse_modular_file
// plus the CPU decode of every group stream. The front-end must either succeed or throw jxg::Error.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iterator>
#include <random>
#include <vector>

#include "../jxl_rs_b200/csrc/host/frame.h"

extern "C" int jxo_decode_modular_file(const uint8_t* data, size_t size, uint8_t* out, size_t out_row_stride, int32_t* planes);
extern "C" int jxo_modular_info(const uint8_t* data, size_t size, uint32_t* width, uint32_t* height, uint32_t* channels,
                                uint32_t* groups);

int main(int argc, char** argv) {
  if (argc < 3) {
    fprintf(stderr, "usage: %s iters file.jxl...\n", argv[0]);
    return 2;
  }
  const int iters = atoi(argv[1]);
  std::mt19937_64 rng(12345);
  int ok = 0, rejected = 0;
  for (int a = 2; a < argc; a++) {
    std::ifstream f(argv[a], std::ios::binary);
    std::vector<uint8_t> orig((std::istreambuf_iterator<char>(f)), {});
    if (orig.size() < 8) continue;
    for (int it = 0; it < iters; it++) {
      std::vector<uint8_t> d = orig;
      const int nmut = 1 + int(rng() % 4);
      for (int m = 0; m < nmut; m++) {
        const size_t pos = (rng() % 3 == 0) ? rng() % std::min<size_t>(d.size(), 2000) : rng() % d.size();
        switch (rng() % 3) {
          case 0: d[pos] ^= uint8_t(1u << (rng() % 8)); break;
          case 1: d[pos] = uint8_t(rng()); break;
          default:
            if (rng() % 8 == 0) d.resize(std::max<size_t>(4, pos));
            else d[pos] = 0xff;
        }
      }
      uint8_t* p = new uint8_t[d.size()];
      memcpy(p, d.data(), d.size());
      bool good = false;
      try {
        auto fs = jxg::parse_vardct_file(p, d.size(), 1 + int(rng() % 3));
        jxg::recycle_frame_state(fs.release());
        good = true;
      } catch (jxg::Error&) {
        uint32_t w = 0, h = 0, c = 0, g = 0;
        if (jxo_modular_info(p, d.size(), &w, &h, &c, &g) == 0 && size_t(w) * h <= (size_t(1) << 24)) {
          std::vector<uint8_t> out(size_t(w) * h * 3 + 16);
          good = jxo_decode_modular_file(p, d.size(), out.data(), size_t(w) * 3, nullptr) == 0;
        }
      } catch (std::bad_alloc&) {
      } catch (std::length_error&) {
      }
      (good ? ok : rejected)++;
      delete[] p;
    }
  }
  printf("accepted %d, rejected %d\n", ok, rejected);
  return 0;
}

// Dequant-matrix encodings (quant_weights.rs:73-100).
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

#include "bitreader.h"
#include "headers.h"
#include "modular.h"

namespace jxg {

struct DctParams {  // DctQuantWeightParams, quant_weights.rs:33-69
  float params[3][17] = {{0}};
  size_t num_bands = 0;
};

struct QuantEncoding {
  enum Mode { kLibrary, kIdentity, kDct2, kDct4, kDct4x8, kAfv, kDct, kRaw } mode = kLibrary;
  float weights[3][9] = {{0}};  // xyb_weights / xyb_mul / afv weights
  DctParams dct;                // params / params4x8
  DctParams dct4x4;             // AFV only
  std::vector<int32_t> qtable;
  float qtable_den = 0;
};

QuantEncoding library_encoding(int idx);
std::vector<float> compute_dequant_table(const QuantEncoding& e, int idx);
QuantEncoding read_quant_encoding(int idx, BitReader& br, const FrameHeader& fh, const ModularTree* global_tree);

}  // namespace jxg

// Host-side Modular sub-bitstream decoder: MA tree + predictors + weighted
// predictor + inverse RCT / palette / squeeze. In a VarDCT frame this decodes
// the LF image, the HF metadata (block types, quant field, EPF sharpness, CfL
// maps) and raw quant tables — "the step before the hot path" (SURVEY §8 n1).
//
// Reference: jxl/src/frame/modular/{tree,predict}.rs, decode/{bitstream,channel,
// common}.rs, transforms/{rct,palette,squeeze,apply_local}.rs, headers/modular.rs
#pragma once
#include <cstdint>
#include <memory>
#include <vector>

#include "bitreader.h"
#include "entropy.h"

namespace jxg {

// Sample buffers of destroyed channels are kept in a small per-thread pool and handed to new channels: the LF /
// HF-metadata streams of every frame otherwise allocate and free ~2 MB of i32 planes (mmap + page faults + munmap).
std::vector<int32_t> take_channel_buffer(size_t n);  // n zero samples
void give_channel_buffer(std::vector<int32_t>&& buf);

struct ModularChannel {
  uint32_t w = 0, h = 0;
  int32_t hshift = 0, vshift = 0;  // < 0: meta channel
  std::vector<int32_t> data;
  ModularChannel() = default;
  ModularChannel(uint32_t w_, uint32_t h_, int32_t hs = 0, int32_t vs = 0)
      : w(w_), h(h_), hshift(hs), vshift(vs), data(take_channel_buffer(size_t(w_) * h_)) {}
  ModularChannel(const ModularChannel&) = default;
  ModularChannel(ModularChannel&&) = default;
  ModularChannel& operator=(const ModularChannel&) = default;
  ModularChannel& operator=(ModularChannel&& o) {
    if (this != &o) {
      give_channel_buffer(std::move(data));
      w = o.w;
      h = o.h;
      hshift = o.hshift;
      vshift = o.vshift;
      data = std::move(o.data);
    }
    return *this;
  }
  ~ModularChannel() { give_channel_buffer(std::move(data)); }
  int32_t* row(uint32_t y) { return data.data() + size_t(y) * w; }
  const int32_t* row(uint32_t y) const { return data.data() + size_t(y) * w; }
};

struct WeightedHeader {  // headers/modular.rs:14-60
  uint32_t p1c = 16, p2c = 10, p3ca = 7, p3cb = 7, p3cc = 7, p3cd = 0, p3ce = 0;
  uint32_t w[4] = {0xd, 0xc, 0xc, 0xc};
};

struct SqueezeParams {
  bool horizontal, in_place;
  uint32_t begin_channel, num_channels;
};

struct ModularTransform {
  uint32_t id = 0;  // 0 RCT, 1 palette, 2 squeeze
  uint32_t begin_channel = 0, rct_type = 6;
  uint32_t num_channels = 3, num_colors = 256, num_deltas = 0, predictor_id = 0;
  std::vector<SqueezeParams> squeezes;
};

struct GroupHeader {
  bool use_global_tree = false;
  WeightedHeader wp;
  std::vector<ModularTransform> transforms;
  static GroupHeader read(BitReader& br);
};

struct TreeNode {
  // property < 0: leaf {predictor, offset, multiplier, ctx}; else split on
  // property > val ? left : right   (tree.rs:13-26, walk :360-390)
  int32_t property;
  int32_t val;       // split value | leaf offset
  uint32_t left;     // left child  | predictor
  uint32_t right;    // right child | multiplier
  uint32_t ctx;      // leaf id
};

struct ModularTree {
  std::vector<TreeNode> nodes;
  EntropyCode code;
  uint32_t num_properties = 0;
  bool uses_wp = false;
  static ModularTree read(BitReader& br, size_t size_limit);
};

// Weighted ("self-correcting") predictor state, predict.rs:221-527.
struct WpState {
  WpState(const WeightedHeader& h, size_t xsize);
  // returns (prediction, property 15)
  void predict(size_t x, size_t y, int32_t top, int32_t left, int32_t topright, int32_t topleft, int32_t toptop,
               int64_t& pred_out, int32_t& prop_out);
  void update(int32_t val, size_t x, size_t y);
  int64_t prediction[4] = {0, 0, 0, 0};
  int64_t pred = 0;
  size_t xsize;
  std::vector<uint32_t> pred_errors;  // [pos][4]
  std::vector<int32_t> error;
  WeightedHeader hdr;
};

// decode/bitstream.rs:134: decodes all channels of one sub-bitstream in place
// (reads the GroupHeader, applies local transforms, undoes them afterwards).
void decode_modular_subbitstream(std::vector<ModularChannel>& channels, size_t stream_id,
                                 const ModularTree* global_tree, BitReader& br);

// One sub-bitstream decoded channel by channel, so that two independent sub-bitstreams (the LF image or the HF
// metadata of two LF groups) can advance in lockstep: a big static-leaf ANS channel is one serial dependency chain per
// stream, and a core has issue slots for two of them (decode_substreams_paired). begin + finish is
// decode_modular_subbitstream. A SubStream must stay where it is between begin and finish.
class SymbolReader;
struct SubStream {
  SubStream() = default;
  SubStream(const SubStream&) = delete;
  SubStream& operator=(const SubStream&) = delete;
  ~SubStream();
  std::vector<ModularChannel>* channels = nullptr;
  size_t stream_id = 0;
  BitReader* br = nullptr;
  bool empty = true;
  GroupHeader header;
  ModularTree local;
  const ModularTree* tree = nullptr;
  SymbolReader* reader = nullptr;  // owned
  std::vector<ModularChannel*> ptrs;
  size_t next = 0;  // next channel to decode
};
void substream_begin(SubStream& s, std::vector<ModularChannel>& channels, size_t stream_id,
                     const ModularTree* global_tree, BitReader& br);
void substream_finish(SubStream& s);  // remaining channels, final-state check, inverse local transforms
void decode_substreams_paired(SubStream& a, SubStream& b);  // all remaining channels of both

// Pieces used by the Modular-frame path (global image split over groups).
// allocate = false: only channel shapes are tracked (planes stay empty; Modular-frame front end).
void meta_apply_transforms(std::vector<ModularChannel>& channels, uint32_t& nb_meta, GroupHeader& header,
                           bool allocate = true);
void undo_transforms(std::vector<ModularChannel>& channels, const GroupHeader& header, uint32_t bit_depth);
void decode_modular_channels(std::vector<ModularChannel*>& channels, size_t stream_id, const GroupHeader& header,
                             const ModularTree& tree, BitReader& br);
// Test hook: route every channel through the generic all-properties loop instead of the specialised walks
// (static leaf / direct-table ANS / lazy properties), for differential tests of the fast paths.
void set_force_generic_walk(bool on);

}  // namespace jxg

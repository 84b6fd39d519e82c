/*
 * oracle.h — CPU restatement of the jxl-rs VarDCT per-group hot path.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load it, and only as the checker / CPU baseline.
 * The product path (libjxgpu.so) never links or calls it.
 *
 * Parity status: "parity pinned on integers, unpinned on pixels" — the
 * reference (Rust) cannot be built in this container (no cargo), and the tree
 * holds no decoded-pixel goldens for VarDCT files. The oracle is pinned by
 *   (1) the reference's own known-answer tests restated in tests/ (ANS
 *       histograms, Huffman, natural coefficient orders, dequant tables,
 *       IDCT/reinterpreting-DCT vs f64 definitions with the reference's
 *       tolerances, Gaborish checkerboard, XYB primaries, sRGB TF, weighted
 *       predictor golden), and
 *   (2) self-verification on the reference's real .jxl fixtures: every ANS
 *       stream must end in state 0x130000, every block must consume exactly
 *       its non-zero count, no section may be over-read.
 */
#ifndef JXO_ORACLE_H_
#define JXO_ORACLE_H_
#include <stddef.h>
#include <stdint.h>

#include "../include/jxg.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Optional taps; any pointer may be NULL. */
typedef struct JxoTaps {
  int32_t* coeffs;      /* [num_groups][3][65536] i32, decode order (group.rs:53-55)        */
  float* xyb_idct;      /* [3][yb*8][xb*8] planes after dequant+IDCT                        */
  float* xyb_filtered;  /* [3][height][width] planes after Gaborish+EPF                     */
} JxoTaps;

/* Full hot path for one frame (decode_vardct_group for every group, then the
 * render stages with SimpleRenderPipeline semantics). Output layout follows
 * desc->output_format. Returns 0 or a JXG_ERR_* code; *bad_group is set on
 * entropy errors. */
int jxo_decode_frame(const JxgFrameDesc* desc, const uint8_t* hf_bytes, const uint64_t* sec_off,
                     const uint32_t* sec_len, uint32_t n_sections, void* out, size_t out_row_stride,
                     const JxoTaps* taps, int num_threads, uint32_t* bad_group);

/* Host front-end + oracle: decode a whole file. */
int jxo_decode_file(const uint8_t* data, size_t size, uint32_t output_format, void* out, size_t out_row_stride,
                    const JxoTaps* taps, int num_threads);
int jxo_file_info(const uint8_t* data, size_t size, JxgImageInfo* info);
const char* jxo_last_error(void);

/* Unit-test hooks into the restated primitives. */
void jxo_idct2d(int rows, int cols, float* block);                  /* in place, layouts as transform.rs */
void jxo_reinterpreting_dct2d(int rows, int cols, const float* in, float* out, int out_stride);
void jxo_transform_to_pixels(int transform, const float* lf, float* coeffs_in_pixels_out);
void jxo_gaborish(int w, int h, const float* in, float* out, float w1, float w2);
void jxo_xyb_to_linear(int n, float* x, float* y, float* b, const float* opsin_matrix, const float* opsin_biases,
                       float intensity_target);
void jxo_linear_to_srgb(int n, float* v);
void jxo_dequant_block(uint32_t num_coeffs, const int32_t* qx, const int32_t* qy, const int32_t* qb, const float* mat, float sx,
                       float sy, float sb, float x_cc, float b_cc, const float* bias, float* out3);
void jxo_sigma_image(uint32_t xb, uint32_t yb, uint32_t global_scale, const int32_t* raw_quant, const uint8_t* sharpness,
                     float quant_mul, const float* sharp_lut, float* out);
void jxo_epf_stage(int stage, uint32_t w, uint32_t h, const float* in3, float* out3, const float* inv_sigma, const float* channel_scale,
                   float pass0_sigma_scale, float pass2_sigma_scale, float border_sad_mul, int threads);
void jxo_set_fast_cpu(int on);
void jxo_f32_to_f16(int n, const float* v, uint16_t* out);
void jxo_from_linear(uint32_t tf, float gamma, float intensity_target, const float* luminances, int n, float* rgb);
/* ImageMetadata.orientation (headers/image_metadata.rs:85-96 display_pixel): pixel (x, y) of the tight w x h source goes
 * to display_pixel(x, y) of `dst` (row stride dst_stride; h x w for orientations 5..8). */
void jxo_orient_image(const uint8_t* src, size_t w, size_t h, size_t bpp, uint32_t orientation, uint8_t* dst, size_t dst_stride);

#ifdef __cplusplus
}
#endif
#endif

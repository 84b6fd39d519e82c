/* Minimal C host of libjxgpu.so: decode a list of .jxl files (VarDCT) as one batch to interleaved RGB8 and write them
 * as binary PPM. The same call sequence a Rust host makes through the bindings of INTEGRATION.md section 1, here with
 * the in-tree front-end doing the parsing (jxg_parse_file) instead of the jxl crate.
 *   gcc -std=c99 -O2 -Iinclude examples/decode_files.c -o decode_files -Ljxl_rs_b200 -ljxgpu -Wl,-rpath,$PWD/jxl_rs_b200
 *   ./decode_files a.jxl b.jxl            ->  a.jxl.ppm b.jxl.ppm */
#include <stdio.h>
#include <stdlib.h>

#include "jxg.h"

static unsigned char* read_file(const char* path, size_t* size) {
  FILE* f = fopen(path, "rb");
  if (!f) return NULL;
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  unsigned char* p = (unsigned char*)malloc(n > 0 ? (size_t)n : 1);
  if (p && fread(p, 1, (size_t)n, f) != (size_t)n) {
    free(p);
    p = NULL;
  }
  fclose(f);
  *size = (size_t)n;
  return p;
}

int main(int argc, char** argv) {
  if (argc < 2) {
    fprintf(stderr, "usage: %s file.jxl...\n", argv[0]);
    return 2;
  }
  const int n = argc - 1;
  void *ctx = NULL, *batch = NULL;
  int rc = jxg_init(0, &ctx);
  if (rc != JXG_OK) {
    fprintf(stderr, "jxg_init: %d (%s)\n", rc, jxg_last_error());
    return 1;
  }
  void** parsed = (void**)calloc((size_t)n, sizeof(void*));
  unsigned char** pixels = (unsigned char**)calloc((size_t)n, sizeof(unsigned char*));
  JxgImageInfo* info = (JxgImageInfo*)calloc((size_t)n, sizeof(JxgImageInfo));
  rc = jxg_batch_begin(ctx, (uint32_t)n, &batch);
  for (int i = 0; rc == JXG_OK && i < n; i++) {
    size_t size = 0;
    unsigned char* data = read_file(argv[i + 1], &size);
    if (!data) {
      fprintf(stderr, "cannot read %s\n", argv[i + 1]);
      rc = JXG_ERR_ARGUMENT;
      break;
    }
    rc = jxg_parse_file(data, size, &parsed[i], &info[i]); /* headers, TOC, LF groups, HF tables: host work */
    free(data);                                            /* the parsed state owns what it needs */
    if (rc != JXG_OK) break;
    pixels[i] = (unsigned char*)malloc((size_t)info[i].width * info[i].height * 3);
    rc = jxg_batch_add_parsed(batch, parsed[i], JXG_FORMAT_RGB_U8, pixels[i], (size_t)info[i].width * 3, /*out_is_device=*/0);
  }
  uint32_t bad_frame = 0, bad_group = 0;
  if (rc == JXG_OK) rc = jxg_batch_run(batch, NULL);                   /* H2D, kernels, D2H: asynchronous */
  if (rc == JXG_OK) rc = jxg_batch_wait(batch, &bad_frame, &bad_group); /* first stream error, if any */
  if (rc != JXG_OK) fprintf(stderr, "decode failed: %d (%s), frame %u group %u\n", rc, jxg_last_error(), bad_frame, bad_group);
  for (int i = 0; rc == JXG_OK && i < n; i++) {
    char name[1024];
    snprintf(name, sizeof(name), "%s.ppm", argv[i + 1]);
    FILE* f = fopen(name, "wb");
    if (!f) continue;
    fprintf(f, "P6\n%u %u\n255\n", info[i].width, info[i].height);
    fwrite(pixels[i], 3, (size_t)info[i].width * info[i].height, f);
    fclose(f);
  }
  if (batch) jxg_batch_end(batch);
  for (int i = 0; i < n; i++) {
    if (parsed[i]) jxg_parsed_free(parsed[i]);
    free(pixels[i]);
  }
  free(parsed);
  free(pixels);
  free(info);
  jxg_shutdown(ctx);
  return rc == JXG_OK ? 0 : 1;
}

// Dev tool: parse files with the host front-end and print a summary.
#include <cstdio>
#include <fstream>
#include <iterator>
#include "../jxl_rs_b200/csrc/host/frame.h"
int main(int argc, char** argv) {
  int ok = 0, bad = 0;
  for (int i = 1; i < argc; i++) {
    std::ifstream f(argv[i], std::ios::binary);
    std::vector<uint8_t> d((std::istreambuf_iterator<char>(f)), {});
    try {
      auto fs = jxg::parse_vardct_file(d.data(), d.size());
      size_t hf = 0; for (auto l : fs->hf_len) hf += l;
      printf("OK   %-60s %ux%u groups=%u passes=%zu hist=%u gs=%u qlf=%u prefix=%d clusters=%u hf=%zu epf=%u gab=%d\n", argv[i], fs->header.xsize(), fs->header.ysize(),
             fs->header.num_groups(), fs->passes.size(), fs->num_histograms, fs->global_scale, fs->quant_lf, (int)fs->passes[0].code.use_prefix, fs->passes[0].code.num_clusters, hf, fs->header.rf.epf_iters, (int)fs->header.rf.gab);
      ok++;
    } catch (jxg::Error& e) {
      printf("%s %-60s %s\n", e.code == -2 ? "SKIP" : "FAIL", argv[i], e.what());
      bad++;
    }
  }
  printf("%d ok, %d not\n", ok, bad);
}

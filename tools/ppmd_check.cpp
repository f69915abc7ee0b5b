// tools/ppmd_check.cpp - pins cmix_b200/csrc/ppmd_model.h against per-byte dumps of the unmodified
// reference (oracle/_ref/oracle_dump dump ... level>=1 writes <prefix>.ppmd.f32, the distribution
// PPMD::ByteUpdate leaves after every byte, ppmd.cpp:1328-1338).
//   g++ -O2 -std=c++17 -ffp-contract=off tools/ppmd_check.cpp -o build/ppmd_check && build/ppmd_check <prefix>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../cmix_b200/csrc/ppmd_model.h"

using namespace cmixb200;

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: ppmd_check <dump prefix>\n"); return 2; }
  const std::string prefix = argv[1];
  size_t n_bytes = 0; std::string vocab;
  {
    FILE* f = fopen((prefix + ".meta").c_str(), "r");
    if (!f) { perror("meta"); return 2; }
    char k[64], v[4096];
    while (fscanf(f, "%63s %4095s", k, v) == 2) { if (!strcmp(k, "n_bytes")) n_bytes = strtoull(v, 0, 10); if (!strcmp(k, "vocab")) vocab = v; }
    fclose(f);
  }
  std::vector<unsigned char> stream(n_bytes);
  { FILE* f = fopen((prefix + ".stream").c_str(), "rb"); if (!f || fread(stream.data(), 1, n_bytes, f) != n_bytes) { perror("stream"); return 2; } fclose(f); }
  std::vector<float> want(n_bytes * 256);
  { FILE* f = fopen((prefix + ".ppmd.f32").c_str(), "rb"); if (!f || fread(want.data(), 4, want.size(), f) != want.size()) { perror("ppmd"); return 2; } fclose(f); }

  PpmdModel* m = new PpmdModel();
  m->ctx_cap = 1u << 22; m->pool_cap = 1u << 24; m->text_cap = 1u << 24;
  m->ctx = (PpmdCtx*)calloc(m->ctx_cap, sizeof(PpmdCtx));
  m->pool = (PpmdSt*)calloc(m->pool_cap, sizeof(PpmdSt));
  m->text = (uint8_t*)calloc(m->text_cap, 1);
  ppmd_init(*m);
  size_t bad = 0;
  for (size_t t = 0; t < n_bytes; ++t) {
    ppmd_update_byte(*m, stream[t]);
    ppmd_prepare_byte(*m);
    if (m->error) { printf("arena exhausted at byte %zu\n", t); return 1; }
    float probs[256];
    for (int i = 0; i < 256; ++i) { probs[i] = (float)m->sqp[i]; if (probs[i] < 1) probs[i] = 1; }
    for (int i = 0; i < 256; ++i) if (vocab[i] != '1') probs[i] = 0;
    float sum = probs[0];
    for (int i = 1; i < 256; ++i) sum += probs[i];
    for (int i = 0; i < 256; ++i) probs[i] /= sum;
    if (memcmp(probs, &want[t * 256], 1024) != 0) {
      if (bad < 5) {
        int k = 0; while (k < 256 && !memcmp(&probs[k], &want[t * 256 + k], 4)) ++k;
        printf("byte %zu (0x%02x): first differing symbol %d: got %.9g want %.9g\n", t, stream[t], k, probs[k], want[t * 256 + k]);
      }
      ++bad;
    }
  }
  printf("%zu bytes, %zu mismatching distributions; contexts %u, states %u, text %u\n", n_bytes, bad, m->ctx_top, m->pool_top, m->text_pos);
  return bad ? 1 : 0;
}

// tools/ppmd_host.cpp - host build of cmix_b200/csrc/ppmd_model.h behind a two-function C API, for
// tests/test_ppmd_model.py (pins the model against fixtures made from dumps of the unmodified reference).
//   g++ -O2 -std=c++17 -ffp-contract=off -shared -fPIC tools/ppmd_host.cpp -o build/libppmd_host.so
#include <cstdlib>
#include <cstring>

#include "../cmix_b200/csrc/ppmd_model.h"

using namespace cmixb200;

extern "C" {

// Runs the model over `stream` and writes PPMD::ByteUpdate's distribution after every byte (ppmd.cpp:1328-1338)
// to out[n][256]. Returns 0, or 1 when the arenas (arena_mb in total) are exhausted.
int ppmd_host_run(const unsigned char* stream, size_t n, const unsigned char* vocab, float* out, unsigned arena_mb) {
  PpmdModel* m = new PpmdModel();
  memset(m, 0, sizeof *m);
  const size_t bytes = (size_t)arena_mb << 20;
  m->ctx_cap = (uint32_t)(bytes / 4 / sizeof(PpmdCtx));
  m->pool_cap = (uint32_t)(bytes / 2 / sizeof(PpmdSt));
  m->text_cap = (uint32_t)(bytes / 4);
  m->ctx = (PpmdCtx*)calloc(m->ctx_cap, sizeof(PpmdCtx));
  m->pool = (PpmdSt*)calloc(m->pool_cap, sizeof(PpmdSt));
  m->text = (uint8_t*)calloc(m->text_cap, 1);
  ppmd_init(*m);
  int rc = 0;
  for (size_t t = 0; t < n && !rc; ++t) {
    ppmd_update_byte(*m, stream[t]);
    ppmd_prepare_byte(*m);
    if (m->error) { rc = 1; break; }
    float* probs = out + t * 256;
    for (int i = 0; i < 256; ++i) { probs[i] = (float)m->sqp[i]; if (probs[i] < 1) probs[i] = 1; if (!vocab[i]) probs[i] = 0; }
    float sum = probs[0];
    for (int i = 1; i < 256; ++i) sum += probs[i];
    for (int i = 0; i < 256; ++i) probs[i] /= sum;
  }
  free(m->ctx); free(m->pool); free(m->text);
  delete m;
  return rc;
}

}  // extern "C"

// oracle/port/coder.cpp — TEST INFRASTRUCTURE.
// Restatement of the 32-bit binary arithmetic coder (reference
// src/coder/encoder.cpp:10-39, src/coder/decoder.cpp:5-39). The product links
// the reference's coder unchanged; this copy exists so tests can turn a
// p-stream into archive bytes (and back) without the reference tree.
#include "internal.h"

struct op_encoder { uint32_t x1 = 0, x2 = 0xffffffffu; std::vector<uint8_t> out; };
struct op_decoder { uint32_t x1 = 0, x2 = 0xffffffffu, x = 0; const uint8_t* d; size_t n, pos = 0; };

static inline uint32_t discretize16(float p) { return 1 + 65534 * p; }   // encoder.cpp:10-12
static inline uint32_t split(uint32_t x1, uint32_t x2, uint32_t p) {
  return x1 + ((x2 - x1) >> 16) * p + (((x2 - x1) & 0xffff) * p >> 16);
}

extern "C" {

op_encoder* op_enc_create(void) { return new op_encoder(); }
void op_enc_destroy(op_encoder* e) { delete e; }
void op_enc_encode(op_encoder* e, float pr, int bit) {
  const uint32_t xmid = split(e->x1, e->x2, discretize16(pr));
  if (bit) e->x2 = xmid; else e->x1 = xmid + 1;
  while (((e->x1 ^ e->x2) & 0xff000000u) == 0) {
    e->out.push_back(e->x2 >> 24);
    e->x1 <<= 8;
    e->x2 = (e->x2 << 8) + 255;
  }
}
size_t op_enc_finish(op_encoder* e, uint8_t* out, size_t cap) {
  while (((e->x1 ^ e->x2) & 0xff000000u) == 0) {
    e->out.push_back(e->x2 >> 24);
    e->x1 <<= 8;
    e->x2 = (e->x2 << 8) + 255;
  }
  e->out.push_back(e->x2 >> 24);
  if (cap >= e->out.size()) memcpy(out, e->out.data(), e->out.size());
  return e->out.size();
}

static int next_byte(op_decoder* d) { return d->pos < d->n ? d->d[d->pos++] : 0; }   // decoder.cpp:10-14
op_decoder* op_dec_create(const uint8_t* data, size_t n) {
  op_decoder* d = new op_decoder();
  d->d = data; d->n = n;
  for (int i = 0; i < 4; ++i) d->x = (d->x << 8) + (next_byte(d) & 0xff);
  return d;
}
void op_dec_destroy(op_decoder* d) { delete d; }
int op_dec_decode(op_decoder* d, float pr) {
  const uint32_t xmid = split(d->x1, d->x2, discretize16(pr));
  int bit = 0;
  if (d->x <= xmid) { bit = 1; d->x2 = xmid; } else d->x1 = xmid + 1;
  while (((d->x1 ^ d->x2) & 0xff000000u) == 0) {
    d->x1 <<= 8;
    d->x2 = (d->x2 << 8) + 255;
    d->x = (d->x << 8) + next_byte(d);
  }
  return bit;
}

}  // extern "C"

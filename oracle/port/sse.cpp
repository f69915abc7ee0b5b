// oracle/port/sse.cpp — TEST INFRASTRUCTURE.
// Restatement of the final SSE/APM stage (reference src/mixer/sse.cpp, the
// M_T1 model: two 7-bucket interpolated SSE tables + two 1-weight integer
// mixers), SURVEY §8 row a7. All arithmetic after the float->15-bit conversion
// is integer; the stretch/squash tables are built with double libm exactly as
// sse.cpp:78-135 does.
#include "internal.h"

#include <math.h>

namespace op {

namespace {

const int kLog = 15, kScale = 1 << kLog, kHalf = kScale / 2, kMask = kScale - 1;

struct StSq {
  uint16_t st[kScale], sq[kScale];
  StSq() {
    const double log2e = 1.44269504088896340736;
    const double st_coef = (kHalf - 1) / (log2e * log((double)(kScale - 1)));   // sse.cpp:92
    const double sq_coef = 1.0 / st_coef;
    memset(st, 0, sizeof(st)); memset(sq, 0, sizeof(sq));
    for (unsigned i = 1; i < (unsigned)kScale; ++i) {                          // sq_i, sse.cpp:100-103
      double a = double(int(i) - kHalf) * sq_coef;
      double v = 1.0 / (1.0 + exp(a / log2e));
      unsigned p = v * kScale;
      sq[i] = (uint16_t)p;
    }
    unsigned x = 0;
    st[0] = 0;
    for (unsigned i = 1; i < (unsigned)kScale; ++i) {                          // sse.cpp:120-130
      double pr = double(i) / kScale;
      unsigned s = (log2e * log((1 - pr) / pr)) * st_coef + kHalf;            // st_i, sse.cpp:95-98
      st[i] = (uint16_t)s;
      if (s != st[x]) {
        unsigned y = i - 1;
        sq[st[x]] = (uint16_t)((x + y + 1) / 2);
        x = i;
      }
    }
  }
};
const StSq& stsq() { static StSq t; return t; }

int Extrap(int p1, int C) {                                                    // sse.cpp:141-146
  p1 = (((p1 - kHalf) * C) >> 13) + kHalf;
  if (p1 < 1) p1 = 1;
  if (p1 > kMask) p1 = kMask;
  return p1;
}
int rdiv(int x, int a, int d) { return x >= 0 ? (x + a) >> d : -((-x + a) >> d); }
int Mixup(int w, int s1, int s0) {                                             // sse.cpp:171-175
  int x = s1 + rdiv((w - kHalf) * (s0 - s1), 1 << (kLog - 1), kLog);
  return (x > 0) ? (x < kScale) ? x : kScale - 1 : 1;
}

struct Interp {   // one SSEi<7> use (sse.cpp:17-63)
  int P, sw; uint16_t* c1;
};
int SsePred(uint16_t* bucket, int iP, Interp& X) {
  int q = (6 * iP) >> kLog;
  X.sw = (6 * iP) & kMask;
  X.c1 = bucket + q;
  int f = (((kScale - X.sw) * X.c1[0] + X.sw * X.c1[1]) >> kLog) - 8192;
  if (f <= 0) f = 1;
  if (f >= kScale) f = kMask;
  X.P = f;
  return f;
}
void SseUpdate(int c, int wr0, Interp& X) {
  X.P = X.P * (kScale - wr0) >> kLog;
  if (c == 0) X.P += wr0;
  int dC = X.c1[0] - X.c1[1];
  int sw_dC = (X.sw * dC + kMask) >> kLog;
  X.c1[0] = (uint16_t)(X.P + sw_dC + 8192);
  X.c1[1] = (uint16_t)(X.P - (dC - sw_dC) + 8192);
}

}  // namespace

// sse.cpp:181-208 constants
enum { f0C = 10240, f1C = 7935, f2C = 9592, sm6wrB = 106, sm6mw = 0, sm6C1 = 8092, x1W0 = 7649,
       x1wr = 6202, f3C = 8200, f4C = 7677, sm7wrB = 127, sm7mw = 8192, sm7C1 = 8202, x2W0 = 2561,
       x2wr = 8320 };
const size_t kMix1Vol = 4ull * 256 * 8 * 79, kMix2Vol = 3ull * 2 * 256 * 256,
             kSm6Vol = 3ull * 128 * 256 * 256, kSm7Vol = 3ull * 32 * 256 * 255;

struct Sse {
  std::vector<uint16_t> s6, s7;     // 7 u16 per bucket set
  std::vector<int> x1, x2;
  unsigned j = 1, pc = 0, ffl = 0;
  Interp su6, su7;
  size_t sm6x = 0, mix1 = 0, sm7x = 0, mix2 = 0;
  int mix1_s0 = 0, mix1_s1 = 0, mix1_p = 0, mix2_s0 = 0, mix2_s1 = 0, mix2_p = 0;
  uint8_t mx1mask[256], sm7mask[256];
};

Sse* sse_create() {
  Sse* s = new Sse();
  auto init = [](std::vector<uint16_t>& t, size_t vol, int Wi) {   // SSEi::Init, sse.cpp:25-31
    t.resize(vol * 7);
    int SCw = (kScale - Wi) / 6, INC = Wi / 2 + 8192;
    uint16_t row[7];
    int p1 = INC;
    for (int i = 0; i < 7; ++i, p1 += SCw) row[i] = (uint16_t)p1;
    for (size_t v = 0; v < vol; ++v) memcpy(&t[v * 7], row, sizeof(row));
  };
  init(s->s6, kSm6Vol, sm6mw);
  init(s->s7, kSm7Vol, sm7mw);
  s->x1.assign(kMix1Vol, x1W0 + kHalf);
  s->x2.assign(kMix2Vol, x2W0 + kHalf);
  // M_mx1mask0 / M_sm7mask0 (sse.cpp:190-196): piecewise maps of the partial-byte
  // context j. mx1: 0,0,1..31 then pairs up to 47, quads up to 63, octets up to 78.
  for (int j = 0; j < 256; ++j) {
    int v;
    if (j < 2) v = 0;
    else if (j <= 32) v = j - 1;
    else if (j <= 63) v = 31 + (j - 32) / 2;
    else if (j <= 127) v = 47 + (j - 64) / 4;
    else v = 63 + (j - 128) / 8;
    s->mx1mask[j] = (uint8_t)v;
    s->sm7mask[j] = (uint8_t)(j < 2 ? 0 : j - 1);
  }
  return s;
}
void sse_destroy(Sse* s) { delete s; }

static unsigned Estimate(Sse* s, unsigned p) {     // M_Estimate, sse.cpp:243-289
  const StSq& T = stsq();
  unsigned j = s->j, pc = s->pc, ffl = s->ffl, prq = p >> 11;
  unsigned q3 = (prq > 0) + (prq > 14);
  unsigned q4 = (prq > 0) + (prq > 7) + (prq > 14);
  s->sm7x = ((((size_t)q3 << 5) + (ffl & 31)) << 8) + (pc & 255);
  s->sm7x = s->sm7x * 255 + s->sm7mask[j];
  s->mix2 = ((((size_t)q3 << 1) + (ffl & 1)) << 8) + (pc & 255);
  s->mix2 = s->mix2 * 256 + j;
  s->sm6x = ((((size_t)q3 << 7) + (ffl & 127)) << 8) + (pc & 255);
  s->sm6x = s->sm6x * 256 + j;
  s->mix1 = ((((size_t)q4 << 8) + (ffl & 255)) << 3) + ((pc >> 5) & 7);
  s->mix1 = s->mix1 * 79 + s->mx1mask[j];

  unsigned p0 = p;
  unsigned p1 = SsePred(&s->s6[s->sm6x * 7], T.sq[Extrap(T.st[p0], f0C)], s->su6);
  unsigned s0 = Extrap(T.st[p0], f1C);
  unsigned s1 = Extrap(T.st[p1], f2C);
  s->mix1_s0 = s0; s->mix1_s1 = s1;
  unsigned s2 = Mixup(s->x1[s->mix1], s->mix1_s0, s->mix1_s1);
  s2 = Extrap(s2, sm6C1);
  s->mix1_p = T.sq[s2];
  unsigned p2 = SsePred(&s->s7[s->sm7x * 7], T.sq[Extrap(T.st[p0], f3C)], s->su7);
  unsigned s4 = Extrap(T.st[p2], f4C);
  s->mix2_s0 = s2; s->mix2_s1 = s4;
  unsigned s5 = Mixup(s->x2[s->mix2], s->mix2_s0, s->mix2_s1);
  s5 = Extrap(s5, sm7C1);
  s->mix2_p = T.sq[s5];
  return s->mix2_p;
}

static void MixUpdate(int& w, int y, int p0, int p1, int wq, int pm) {   // sse.cpp:177-183
  int py = kScale - (y << kLog);
  int e = py - pm;
  int d = rdiv(e * (p0 - p1), 1 << (kLog - 1), kLog);
  d = rdiv(d * wq, 1 << (kLog - 1), kLog);
  w += d;
}

float sse_predict(Sse* s, float input) {           // sse.cpp:320-324
  int discrete = 1 + (1 - input) * 32766;
  int estimate = Estimate(s, discrete);
  return 1 - ((estimate - 1) / 32766.0);
}

void sse_perceive(Sse* s, int bit) {               // M_Update, sse.cpp:291-305
  SseUpdate(bit, sm6wrB, s->su6);
  MixUpdate(s->x1[s->mix1], bit, s->mix1_s0, s->mix1_s1, x1wr, s->mix1_p);
  SseUpdate(bit, sm7wrB, s->su7);
  MixUpdate(s->x2[s->mix2], bit, s->mix2_s0, s->mix2_s1, x2wr, s->mix2_p);
  s->j += s->j + bit;
  if (s->j >= 256) {
    s->ffl = (uint8_t)(s->ffl * 2 + (s->pc >= 0x40));
    s->pc = (uint8_t)s->j;
    s->j = 1;
  }
}

}  // namespace op

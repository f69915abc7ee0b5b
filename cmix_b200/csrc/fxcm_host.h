// cmix_b200/csrc/fxcm_host.h — host-side construction of the resident FXCM model (fxcm_model.h).
//
// Builds the read-only tables exactly as the reference builds them at start-up (fxcmv1.cpp:4838-4895,
// PredictorInit :3291-3419): the squash/stretch tables go through the same libm calls of the same glibc as the
// oracle (expf/logf overload resolution spelled out), the six bit-history state tables come from the generator of
// fxcmv1.cpp:241-357 restated below, and the per-map input tables from the c_r/c_s/c_s3/c_s4 parameter rows
// (:3213-3216). Memory is obtained through a backend (cudaMalloc + fill kernels in engine.cu, calloc in
// tools/fxcm_check.cpp) so the CPU pinning tool and the device engine share every line of set-up.
#ifndef CMIXB200_FXCM_HOST_H
#define CMIXB200_FXCM_HOST_H

#include <math.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "fxcm_model.h"

namespace cmixb200 {
namespace fx {

// ---- bit-history state table generator (fxcmv1.cpp:241-357) ----
struct StateGen {
  int mdc, b[6];
  unsigned char ns[1024];
  unsigned char t[64][64][2];
  int count(int x, int y) const {
    if (x < y) return count(y, x);
    if (x < 0 || y < 0 || x >= 64 || y >= 64 || y >= 5 || x >= b[y]) return 0;
    return 1 + (y > 0 && x + y < b[5]);
  }
  void discount(int& x) const {
    if (x > 2) { int y = 0; for (int i = 1; i < mdc; ++i) y += x >= i; x = y; }
  }
  void step(int& x, int& y, int bit) const {
    if (x < y) { step(y, x, 1 - bit); return; }
    if (bit) { ++y; discount(x); } else { ++x; discount(y); }
    while (!t[x][y][1]) {
      if (y < 2) --x;
      else { x = (x * (y - 1) + (y / 2)) / y; --y; }
    }
  }
  void build(const int p[7], u8* out) {
    for (int i = 0; i < 6; ++i) b[i] = p[i];
    mdc = p[6];
    memset(ns, 0, sizeof ns);
    memset(t, 0, sizeof t);
    int state = 0;
    for (int i = 0; i < 256; ++i)
      for (int y = 0; y <= i; ++y) {
        const int x = i - y, n = count(x, y);
        if (n) { t[x][y][0] = (unsigned char)state; t[x][y][1] = (unsigned char)n; state += n; }
      }
    state = 0;
    bool stop = false;
    for (int i = 0; i < 64 && !stop; ++i)
      for (int y = 0; y <= i && !stop; ++y) {
        const int x = i - y;
        for (int k = 0; k < t[x][y][1]; ++k) {
          int x0 = x, y0 = y, x1 = x, y1 = y;
          step(x0, y0, 0);
          step(x1, y1, 1);
          ns[state * 4] = t[x0][y0][0];
          ns[state * 4 + 1] = (unsigned char)(t[x1][y1][0] + (t[x1][y1][1] > 1));
          ns[state * 4 + 2] = (unsigned char)x;
          ns[state * 4 + 3] = (unsigned char)y;
          if (state > 0xff || t[x][y][1] == 0 || t[x0][y0][1] == 0 || t[x1][y1][1] == 0) { stop = true; break; }
          ++state;
          if (state > 0xff) { stop = true; break; }
        }
      }
    memcpy(out, ns, 1024);
  }
};

inline int host_sc(int p) { return p > 0 ? p >> 7 : (p + 127) >> 7; }   // fxcmv1.cpp:905-908

// Parse "digits and spaces" into a byte table
inline void fill_digits(u8* dst, int n, const char* hex) {
  int k = 0;
  for (const char* p = hex; *p && k < n; ++p) {
    if (*p == ' ' || *p == '\n') continue;
    dst[k++] = (u8)(*p <= '9' ? *p - '0' : *p - 'a' + 10);
  }
  for (; k < n; ++k) dst[k] = 0;
}

// Fill every field of T except the dictionary pointers.
inline void build_tables(Tables& T) {
  memset(&T, 0, sizeof T);
  T.map = T.map_store; T.st2 = T.st2_store;
  // squash / stretch (fxcmv1.cpp:137-175): float/double mix exactly as written there
  for (int d = -2047; d <= 2047; ++d) {
    float p = 1.0f / (1.0f + exp(-d / 256.0));
    p *= 4096.0;
    u32 pi = (u32)roundf(p);
    if (pi > 4095) pi = 4095;
    if (pi < 1) pi = 1;
    T.sqt[d + 2047] = (short)pi;
  }
  for (int i = 0; i <= 4095; ++i) {
    int p = i == 0 ? 1 : i;
    float f = p / 4096.0f;
    float d = logf(f / (1.0f - f)) * 256.0f;
    int di = (int)roundf(d);
    if (di > 2047) di = 2047;
    if (di < -2047) di = -2047;
    T.strt[i] = (short)di;
  }
  {  // ilog (fxcmv1.cpp:219-227)
    u32 x = 14155776;
    for (int i = 2; i < 257; ++i) { x += 774541002 / (i * 2 - 1); T.ilog[i - 1] = (u8)(x >> 24); }
  }
  { int o = 2; for (int i = 0; i < 1024; ++i) { T.dt[i] = 4096 / o; ++o; } T.dt[1023] = 1; }
  {
    static const int par[6][7] = {{28, 28, 31, 29, 23, 4, 17}, {32, 28, 31, 28, 21, 5, 6}, {31, 27, 30, 27, 24, 4, 27},
                                  {33, 31, 31, 24, 20, 4, 33}, {28, 29, 30, 30, 23, 3, 22}, {28, 29, 33, 23, 23, 6, 14}};
    StateGen* g = new StateGen();
    for (int k = 0; k < 6; ++k) g->build(par[k], T.sta[k]);
    delete g;
  }
  for (int i = 0; i < 4096; ++i) {
    T.st2_store[0][i] = 0;
    T.st2_store[1][i] = (short)clp(host_sc(12 * (i - 2048)));
    T.st2_store[2][i] = (short)clp(host_sc(14 * (i - 2048)));
  }
  for (int r = 0; r < 256; ++r) {   // RunContextMap::Init(m, 6) (fxcmv1.cpp:765-777)
    int c = T.ilog[r] * 8;
    if ((r & 1) == 0) c = c * 6 / 4;
    T.rcm_rc[r + 256] = (short)clp(c);
    T.rcm_rc[r] = (short)clp(-c);
  }
  static const u32 c_r[27] = {3, 4, 6, 4, 6, 6, 2, 3, 3, 3, 6, 4, 3, 4, 5, 6, 2, 6, 4, 4, 4, 4, 4, 4, 4, 4, 4};
  static const u32 c_s[27] = {28, 26, 28, 31, 34, 31, 33, 33, 35, 35, 29, 32, 33, 34, 30, 36, 31, 32, 32, 32, 32, 32, 33, 32, 32, 32, 32};
  static const u32 c_s3[27] = {43, 33, 34, 28, 34, 29, 32, 33, 37, 35, 33, 28, 31, 35, 28, 30, 33, 34, 32, 32, 32, 32, 32, 32, 32, 32, 32};
  static const u32 c_s4[27] = {9, 8, 9, 5, 8, 12, 15, 8, 8, 12, 10, 7, 7, 8, 8, 13, 13, 14, 8, 8, 12, 12, 12, 12, 12, 12, 12};
  for (int id = 0; id < N_MAPS; ++id) {   // ContextMap::Init table part (fxcmv1.cpp:1000-1043)
    const MapSpec sp = kMapSpec[id];
    T.spec[id] = sp;
    MapTab& mt = T.map_store[id];
    const u8* nn = T.sta[sp.sta];
    const int cmul = (int)c_r[sp.par], cms = (int)c_s[sp.par], cms3 = (int)c_s3[sp.par], cms4 = (int)c_s4[sp.par];
    for (int rc = 0; rc < 256; ++rc) {
      int c = T.ilog[rc];
      c = c << (2 + (~rc & 1));
      if ((rc & 1) == 0) c = c * cmul / 4;
      mt.rc1[rc + 256] = (short)clp(c);
      mt.rc1[rc] = (short)clp(-c);
    }
    for (int i = 0; i < 4096; ++i) mt.st1[i] = (short)clp(host_sc(cms * T.strt[i]));
    for (int s = 0; s < 256; ++s) {
      const int n0 = -!nn[s * 4 + 2], n1 = -!nn[s * 4 + 3];
      int r = 0, sp0 = 0;
      if ((n1 - n0) == 1) { sp0 = 0; r = 1; }
      if ((n1 - n0) == -1) { sp0 = 4095; r = 1; }
      if (r) {
        const u32 a0 = nn[s * 4 + 2] * 3 + 1, a1 = nn[s * 4 + 3] * 3 + 1;
        const int pre = (int)((a1 << 12) / (a0 + a1));
        mt.st8[s] = (short)clp(host_sc(cms4 * (pre - sp0)));
        mt.st32[s] = (short)clp(host_sc(cms3 * T.strt[pre]));
        if (s < 8) mt.st32[s] = 0;
      } else { mt.st8[s] = 0; mt.st32[s] = 0; }
    }
  }
  for (int i = 0; i < N_MIX; ++i) { T.mix_m[i] = kMixM[i]; T.mix_shift[i] = kMixShift[i]; T.mix_uperr[i] = kMixUperr[i]; }
  // byte classes wrt_2b / wrt_3b / wrt_4b (fxcmv1.cpp:51-90, :1843-1862)
  fill_digits(T.wrt2, 256,
      "2313301233001333 3333333333303333 3202132133332302 1111111111322322 2200231212222200 2222222230232023"
      "1111111111111111 1111111111111111 1111111111111111 1111111111111111 1111111111111111 1111111111111111"
      "1111111111111111 0000000000000000 0000000000000000 0000000000000000");
  fill_digits(T.wrt3, 256,
      "0020560602043000 0000000000000000 2414474737223531 1111111111053355 0557501545006071 3374557022544746"
      "5555555555555555 5555555555555555 6666666666666666 6666666666666666 6666666666666666 6666666666666666"
      "6666666666666666 7777777777777777 7777777777777777 7777777777777777");
  fill_digits(T.wrt4, 256,
      "60cfcfee53e0fd8d 0000000000000000 d5fbac6c0be11a98 77777777779b6104 9aa45142b8410aa5 47f45d014c01333b"
      "2222222222222222 22222222222380b7 2222222222222222 2222222222222222 2222222222222222 2222222222222222"
      "2222222222222222 2222222222222222 2222222222222222 2222222222222222");
  memset(T.fcy, 0, sizeof T.fcy); memset(T.fcq, 0, sizeof T.fcq);
  T.fcy[34] = 5; T.fcy[39] = 6; T.fcy[40] = 1; T.fcy[76] = 4; T.fcy[80] = 2; T.fcy[91] = 3;
  T.fcq[42] = 6; T.fcq[64] = 1; T.fcq[74] = 3; T.fcq[76] = 4; T.fcq[77] = 5; T.fcq[80] = 2; T.fcq[81] = 7; T.fcq[91] = 2; T.fcq[96] = 2;
  static const u32 primes[14] = {0, 257, 251, 241, 239, 233, 229, 227, 223, 211, 199, 197, 193, 191};
  for (int i = 0; i < 14; ++i) T.primes[i] = primes[i];
  static const int e_l[8] = {1830, 1997, 1973, 1851, 1897, 1690, 1998, 1842};
  for (int i = 0; i < 8; ++i) T.e_l[i] = e_l[i];
}

// The WRT dictionary as flat storage (fxcmv1.cpp:372-410): lines of `path`, each NUL terminated.
struct HostDict {
  std::vector<char> chars;
  std::vector<u32> off;
  bool loaded = false;
  void load(const char* path) {
    if (!path) return;
    FILE* f = fopen(path, "rb");
    if (!f) return;
    std::string line;
    int c;
    bool any = false;
    while ((c = getc(f)) != EOF) {
      any = true;
      if (c == '\n') {
        off.push_back((u32)chars.size());
        chars.insert(chars.end(), line.begin(), line.end());
        chars.push_back(0);
        line.clear();
        any = false;
        if (off.size() >= 44516) break;
      } else line.push_back((char)c);
    }
    if (any && off.size() < 44516) { off.push_back((u32)chars.size()); chars.insert(chars.end(), line.begin(), line.end()); chars.push_back(0); }
    fclose(f);
    loaded = true;
  }
};

// Backend concept:  void* alloc(size_t bytes) -> zeroed memory;  void fill16(void*, size_t n, u16 v);
//                   void fill32(void*, size_t n, u32 v);  void upload(void* dst, const void* src, size_t bytes);
// build_state fills the HOST mirror `S` (pointers refer to backend memory) and the HOST mirror `X` of the text state;
// the caller uploads both (S.text and S.T must then be patched to the backend copies).
template <class B> bool build_state(B& be, const Tables& T, State& S, TextState& X) {
  memset(&S, 0, sizeof S);
  memset(&X, 0, sizeof X);
  S.c0 = 1; S.pr = 2048; S.rate = 6;
  for (int i = 0; i <= N_OUT; ++i) S.codes[i] = 0xFFFF;
  for (int id = 0; id < N_MAPS; ++id) {
    const MapSpec sp = kMapSpec[id];
    MapState& m = S.map[id];
    const int sh = map_shift(sp.kind), A = map_slots(sp.kind);
    u32 mem = sp.mem;
    size_t n_el;
    if (sp.kind == 2) { mem *= 2; m.tmask = (mem >> 7) - 1; n_el = (size_t)(mem >> 7) + 128; }
    else { m.tmask = (mem >> 6) - 1; n_el = (size_t)(mem >> 6) + 64; }
    m.t = (u8*)be.alloc(n_el << sh);
    m.sm = (u32*)be.alloc((size_t)sp.C * 256 * 4);
    if (!m.t || !m.sm) return false;
    std::vector<u32> smi(256);
    const u8* nn = T.sta[sp.sta];
    for (int i = 0; i < 256; ++i) { const u32 n0 = nn[i * 4 + 2] * 3 + 1, n1 = nn[i * 4 + 3] * 3 + 1; smi[i] = ((n1 << 20) / (n0 + n1)) << 12; }
    for (int c = 0; c < sp.C; ++c) be.upload(m.sm + c * 256, smi.data(), 1024);
    for (int c = 0; c < 8; ++c) { m.cp[c] = m.cp0[c] = (u32)(2 * A + 1); m.runp[c] = m.cp[c] + 3; }
    m.mask = (u16)(((1 < sp.C) - 1) * 2);
  }
  static const int scm_bits[7] = {8, 8, 8, 9, 8, 8, 7};
  for (int k = 0; k < 7; ++k) {
    ScmState& c = S.scm[k];
    c.mask = (1 << scm_bits[k]) - 1; c.stride = 255; c.btotal = 8;
    const size_t n = ((size_t)1 << scm_bits[k]) * 255;
    c.data = (u16*)be.alloc(n * 2);
    if (!c.data) return false;
    be.fill16(c.data, n, 0x7FFF);
  }
  static const int sma_bits[3] = {9, 19, 16};
  for (int k = 0; k < 3; ++k) {
    const size_t n = (size_t)1 << sma_bits[k];
    S.sma[k].t = (u32*)be.alloc(n * 4);
    if (!S.sma[k].t) return false;
    be.fill32(S.sma[k].t, n, 1u << 31);
    S.sma[k].mask = (int)n - 1;
  }
  for (int i = 0; i < N_MIX; ++i) {
    MixState& m = S.mix[i];
    const size_t n = (size_t)kMixM[i] * (i < 10 ? N_IN1 : N_IN2);
    m.w = (short*)be.alloc((n + 32) * 2);
    if (!m.w) return false;
    be.fill16(m.w, n, 129);
    m.pr = 2048; m.elim = kMixElim[i];
  }
  static const size_t apm_n[6] = {256, 0x10000, 0x10000, 0x40000, 0x40000, 0x40000};
  {
    std::vector<u16> row(33);
    for (int j = 0; j < 33; ++j) row[j] = (u16)(squash(T, (j - 16) * 128) * 16);
    for (int k = 0; k < 6; ++k) {
      S.apm[k].t = (u16*)be.alloc(apm_n[k] * 33 * 2 + 4);
      if (!S.apm[k].t) return false;
      std::vector<u16> all(apm_n[k] * 33);
      for (size_t i = 0; i < all.size(); ++i) all[i] = row[i % 33];
      be.upload(S.apm[k].t, all.data(), all.size() * 2);
    }
  }
  S.rcm_t = (u8*)be.alloc((size_t)4096 * 4096 + 64);
  S.rcm_n = 4096 * 4096 / 4 - 1;
  S.rcm_cp = 1;
  S.mhash = (u32*)be.alloc(((size_t)MATCH_HASH + 32) * 16);
  S.sm_table = (u32*)be.alloc((size_t)1024 * 1024 * 4);
  for (int i = 0; i < 4; ++i) { S.sm_prev[i] = i - 1; S.sm_next[i] = i + 1; }
  S.sm_next[3] = -1;
  S.buffer = (u8*)be.alloc(BUF_MASK + 1);
  S.ind3 = (u16*)be.alloc((size_t)IND3_SIZE * 2);
  S.t2 = (u32*)be.alloc(0x10000 * 4);
  S.wp = (int*)be.alloc(0x10000 * 4);
  if (!S.rcm_t || !S.mhash || !S.sm_table || !S.buffer || !S.ind3 || !S.t2 || !S.wp) return false;
  // text state (PredictorInit, fxcmv1.cpp:3291-3419)
  X.n3b = X.n2b = 0xffffffffu;
  X.ah2 = 0x765BA55C;
  X.cword = 0; X.pword = 3;
  X.so = X.colonstr = -1;
  static const u16 brackets[8] = {'(', ')', kCurlyOpen, kCurlyClose, '[', ']', kLess, kGreater};
  static const u16 quotes[4] = {kApos, kApos, kQuote, kQuote};
  static const u16 fchar[20] = {kFirstUpper, kLF, kTextData, kLF, kColon, kLF, kLess, kGreater, kEquals, kLF, kSqOpen, kSqClose,
                                kCurlyOpen, kCurlyClose, '*', kLF, kVBar, kLF, kHtLink, kLF};
  static const u16 html[2] = {'&' * 256 + 'L', '&' * 256 + 'N'};
  X.br.init(brackets, 8, 0, 256);
  X.qo.init(quotes, 4, 1, 256);
  X.fcx.init(fchar, 20, 0, 256);
  X.ht.init(html, 2, 0, 0xfff);
  X.cols.init();
  return true;
}

}  // namespace fx
}  // namespace cmixb200
#endif

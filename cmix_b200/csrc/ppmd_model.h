// cmix_b200/csrc/ppmd_model.h - the order-25 PPM byte model as integer-only
// host/device code.
//
// Behavioural restatement of the reference's PPMD model (src/models/ppmd.cpp, "ppmd var.J" with SEE
// and information inheritance as adapted by mod_ppmd): after every coded byte it is updated with that
// byte (ppmd_UpdateByte, ppmd.cpp:1283-1314) and then walks the suffix chain of the new context to
// emit a 256-symbol distribution (ppmd_PrepareByte + ConvertSQ, ppmd.cpp:1256-1281, 1116-1140),
// which PPMD::ByteUpdate floors, masks by the vocabulary and normalises (ppmd.cpp:1328-1338).
//
// What is reproduced exactly: every frequency, escape estimate, SEE / binary-context adaptation,
// successor creation and order bookkeeping, i.e. every number that reaches the distribution.
// What is NOT reproduced: the reference's sub-allocator (12-byte units, free-block gluing) and its
// model cut-off / restart when the 14 GB heap or the text area fills up. Storage here is three flat
// arenas (contexts, state arrays in power-of-two capacity classes, text) addressed by 32-bit
// references; the only observable uses of addresses in the reference are "is this successor a
// raw text position or a real context" and the ordering of text positions, and the reference
// encoding below preserves both (text positions < context references). Running out of arena sets
// `error` instead of triggering the reference's cut-off, which for its 14 GB heap happens only
// after gigabytes of input.
//
// The same header is compiled by g++ (tests: pinned against per-byte dumps of the unmodified
// reference) and by nvcc for the device (one thread per stream): integer arithmetic only, so the
// two agree by construction.
#ifndef CMIXB200_PPMD_MODEL_H
#define CMIXB200_PPMD_MODEL_H

#include <stdint.h>

#if defined(__CUDACC__)
#define PP_HD __host__ __device__ inline
#else
#define PP_HD inline
#endif
// Per-phase cycle accounting on the device (tools/ppmd_prof.py); compiled out on the host.
#if defined(__CUDA_ARCH__)
#define PP_TICK(m, k) do { const long long now_ = clock64(); (m).prof[k] += (unsigned long long)(now_ - (m).tprev); (m).tprev = now_; } while (0)
#else
#define PP_TICK(m, k) do { } while (0)
#endif

namespace cmixb200 {

struct PpmdSt {              // one (symbol, frequency, successor) triple - STATE, ppmd.cpp:406-410
  uint32_t succ;             // 0 none | text position | PPMD_CTX_BASE + context index
  uint8_t sym, freq;
  uint16_t pad;
};

struct PpmdCtx {             // PPM_CONTEXT, ppmd.cpp:425-434; ns = number of states - 1
  uint8_t ns, flags;
  uint16_t summ;
  uint32_t stats;            // index of the state array in the pool (binary contexts: one state there)
  uint32_t suffix;           // context index, 0 = none (the order-0 context)
  uint32_t cap_class;        // capacity class of the state array: capacity = 1 << cap_class
};

struct PpmdSee { uint16_t summ; uint8_t shift, count; };          // SEE2_CONTEXT, ppmd.cpp:463-491
struct PpmdQ { uint16_t sym, freq, total; };                      // qsym, ppmd.cpp:1098-1108

enum : uint32_t { PPMD_CTX_BASE = 0x40000000u, PPMD_NIL = 0xffffffffu };
enum { PPMD_MAX_FREQ = 124, PPMD_MAX_ORDER = 25, PPMD_INTERVAL = 128, PPMD_PERIOD_BITS = 7, PPMD_BIN_SCALE = 1 << 14,
       PPMD_SCALE = 1 << 15 };

struct PpmdModel {
  // arenas (device or host memory)
  PpmdCtx* ctx; uint32_t ctx_cap, ctx_top;            // index 0 unused (0 = "no suffix")
  PpmdSt* pool; uint32_t pool_cap, pool_top;
  uint32_t free_head[9];                               // per capacity class, PPMD_NIL = empty
  uint8_t* text; uint32_t text_cap, text_pos;          // pText - HeapStart
  uint32_t error;                                      // 1: arena exhausted (the reference would cut the model off)
  // model registers
  int32_t order_fall, run_length, init_rl, bsumm, prev_success, num_masked;
  uint32_t esc_count, max_context, found;              // found: pool index of FoundState, PPMD_NIL = none
  uint32_t char_mask[256];
  uint16_t bin_summ[25][64];
  PpmdSee see[23][32], dummy_see;
  uint8_t ns2bs[256], qtable[260];
  // PrepareByte scratch
  PpmdQ sq[1024]; uint32_t sq_n;
  uint32_t sqp[256];
  // cycles per phase (device only): 0 symbol search, 1 model update, 2 suffix walk of PrepareByte, 3 ConvertSQ, 4 emit, 5 bytes
  unsigned long long prof[6]; long long tprev;
};

PP_HD bool pp_is_ctx(uint32_t v) { return v >= PPMD_CTX_BASE; }
PP_HD PpmdCtx& pp_ctx(PpmdModel& m, uint32_t ref) { return m.ctx[ref - PPMD_CTX_BASE]; }

PP_HD uint32_t pp_alloc_states(PpmdModel& m, uint32_t cls) {
  const uint32_t h = m.free_head[cls];
  if (h != PPMD_NIL) { m.free_head[cls] = m.pool[h].succ; return h; }
  const uint32_t n = 1u << cls;
  if (m.pool_top + n > m.pool_cap) { m.error = 1; return PPMD_NIL; }
  const uint32_t r = m.pool_top;
  m.pool_top += n;
  return r;
}
PP_HD void pp_free_states(PpmdModel& m, uint32_t idx, uint32_t cls) { m.pool[idx].succ = m.free_head[cls]; m.free_head[cls] = idx; }
PP_HD uint32_t pp_alloc_ctx(PpmdModel& m) {
  if (m.ctx_top >= m.ctx_cap) { m.error = 1; return 0; }
  return m.ctx_top++;
}
PP_HD uint32_t pp_class_for(uint32_t n) { uint32_t c = 0; while ((1u << c) < n) ++c; return c; }

// SEE2_CONTEXT::update / setShift_rare (ppmd.cpp:478-490)
PP_HD void pp_see_update(PpmdSee& s) {
  if (--s.count == 0) {
    uint32_t i = (uint32_t)(s.summ >> s.shift);
    i = PPMD_PERIOD_BITS - (i > 40) - (i > 280) - (i > 1020);
    if (i < s.shift) { s.summ = (uint16_t)(s.summ >> 1); s.shift--; }
    else if (i > s.shift) { s.summ = (uint16_t)(s.summ << 1); s.shift++; }
    s.count = (uint8_t)(5 << s.shift);
  }
}

// PPMD_STARTUP + StartModelRare (ppmd.cpp:372-402, 645-691) for MaxOrder = 25.
PP_HD void ppmd_init(PpmdModel& m) {
  static const signed char kEscCoef[12] = {16, -10, 1, 51, 14, 89, 23, 35, 64, 26, -42, 43};
  m.error = 0;
  for (int i = 0; i < 6; ++i) m.prof[i] = 0;
  m.tprev = 0;
  m.ns2bs[0] = 0; m.ns2bs[1] = 2; m.ns2bs[2] = 2;
  for (int i = 3; i < 29; ++i) m.ns2bs[i] = 4;
  for (int i = 29; i < 256; ++i) m.ns2bs[i] = 6;
  {
    int i, k, mm, step;
    for (i = 0; i < 5; ++i) m.qtable[i] = (uint8_t)i;
    for (mm = i = 5, k = step = 1; i < 260; ++i) { m.qtable[i] = (uint8_t)mm; if (!--k) { k = ++step; mm++; } }
  }
  for (int i = 0; i < 256; ++i) m.char_mask[i] = 0;
  m.esc_count = 1;
  m.order_fall = PPMD_MAX_ORDER;
  m.init_rl = -13;
  m.run_length = m.init_rl;
  m.bsumm = 0; m.num_masked = 0; m.prev_success = 0; m.found = PPMD_NIL; m.sq_n = 0;
  m.dummy_see.summ = 0; m.dummy_see.shift = 0; m.dummy_see.count = 0;
  for (int c = 0; c < 9; ++c) m.free_head[c] = PPMD_NIL;
  m.ctx_top = 1; m.pool_top = 0; m.text_pos = 0;
  const uint32_t root = pp_alloc_ctx(m);
  PpmdCtx& r = m.ctx[root];
  r.ns = 255; r.summ = 257; r.flags = 0; r.suffix = 0; r.cap_class = 8;
  r.stats = pp_alloc_states(m, 8);
  for (int i = 0; i < 256; ++i) { PpmdSt& s = m.pool[r.stats + i]; s.sym = (uint8_t)i; s.freq = 1; s.succ = 0; s.pad = 0; }
  m.max_context = PPMD_CTX_BASE + root;
  uint8_t i2f[25];
  { int i, k; for (k = i = 0; i < 25; i2f[i++] = (uint8_t)(k + 1)) while (m.qtable[k] == i) k++; }
  for (int k = 0; k < 64; ++k) {
    int s = 0;
    for (int i = 0; i < 6; ++i) s += kEscCoef[2 * i + ((k >> i) & 1)];
    s = s < 32 ? 32 : (s > 224 ? 224 : s);
    s *= 128;
    for (int i = 0; i < 25; ++i) m.bin_summ[i][k] = (uint16_t)(PPMD_BIN_SCALE - s / i2f[i]);
  }
  for (int i = 0; i < 23; ++i)
    for (int k = 0; k < 32; ++k) { m.see[i][k].shift = PPMD_PERIOD_BITS - 4; m.see[i][k].summ = (uint16_t)((8 * i + 5) << (PPMD_PERIOD_BITS - 4)); m.see[i][k].count = 7; }
}

// rescale (ppmd.cpp:493-554): halve the frequencies of context q, keep the order, drop zero states.
// Returns the pool index of the found state after the move to the front.
PP_HD uint32_t pp_rescale(PpmdModel& m, PpmdCtx& q, int order_fall, uint32_t found) {
  PpmdSt* st = m.pool + q.stats;
  q.flags &= 0x14;
  {
    const PpmdSt tmp = m.pool[found];
    for (uint32_t k = found; k != q.stats; --k) m.pool[k] = m.pool[k - 1];
    st[0] = tmp;
  }
  const int of = order_fall != 0;
  int f0 = st[0].freq;
  int sf = q.summ;
  int esc = sf - f0;
  st[0].freq = (uint8_t)((f0 + of) >> 1);
  q.summ = st[0].freq;
  int a;
  int p = 0;
  for (int i = 0; i < q.ns; ++i) {
    ++p;
    a = st[p].freq;
    esc -= a;
    a = (a + of) >> 1;
    st[p].freq = (uint8_t)a;
    q.summ = (uint16_t)(q.summ + a);
    if (a) q.flags |= 0x08 * (st[p].sym >= 0x40);
    if (a > st[p - 1].freq) {
      const PpmdSt tmp = st[p];
      int p1 = p;
      for (; tmp.freq > st[p1 - 1].freq; --p1) st[p1] = st[p1 - 1];
      st[p1] = tmp;
    }
  }
  if (st[p].freq == 0) {
    int i = 0;
    for (; st[p].freq == 0; ++i, --p) {}
    esc += i;
    q.ns = (uint8_t)(q.ns - i);
    if (q.ns == 0) {
      PpmdSt tmp = st[0];
      int v = (2 * tmp.freq + esc - 1) / esc;
      if (v > PPMD_MAX_FREQ / 3) v = PPMD_MAX_FREQ / 3;
      tmp.freq = (uint8_t)v;
      q.flags &= 0x18;
      st[0] = tmp;                              // the single state stays in the (oversized) array
      return q.stats;
    }
  }
  q.summ = (uint16_t)(q.summ + ((esc + 1) >> 1));
  if (order_fall || (q.flags & 0x04) == 0) {
    sf -= esc;
    a = sf - f0;
    const uint32_t v = (uint32_t)((f0 * (int)q.summ - sf * (int)st[0].freq + a - 1) / a);
    a = (int)(v < 2u ? 2u : (v > (uint32_t)(PPMD_MAX_FREQ / 2 - 18) ? (uint32_t)(PPMD_MAX_FREQ / 2 - 18) : v));
  } else {
    a = 2;
  }
  st[0].freq = (uint8_t)(st[0].freq + a);
  q.summ = (uint16_t)(q.summ + a);
  q.flags |= 0x04;
  return q.stats;
}

// The SummFreq word of a binary context is the storage of its single state in the reference
// (PPM_CONTEXT::oneState, ppmd.cpp:433): symbol in the low byte, frequency in the high byte.
PP_HD int pp_summ_read(const PpmdModel& m, const PpmdCtx& q) {
  if (q.ns) return q.summ;
  const PpmdSt& s = m.pool[q.stats];
  return (int)s.sym | ((int)s.freq << 8);
}

// processBinSymbol<0> (ppmd.cpp:961-983)
PP_HD void pp_bin_symbol(PpmdModel& m, PpmdCtx& q, int symbol) {
  PpmdSt& rs = m.pool[q.stats];
  const int i = m.ns2bs[m.ctx[q.suffix].ns] + m.prev_success + q.flags + ((m.run_length >> 26) & 0x20);
  uint16_t& bs = m.bin_summ[m.qtable[rs.freq - 1]][i];
  m.bsumm = bs;
  bs = (uint16_t)(bs - ((m.bsumm + 64) >> PPMD_PERIOD_BITS));
  if (rs.sym != symbol) {
    m.char_mask[rs.sym] = m.esc_count;
    m.num_masked = 0;
    m.prev_success = 0;
    m.found = PPMD_NIL;
  } else {
    bs = (uint16_t)(bs + PPMD_INTERVAL);
    rs.freq = (uint8_t)(rs.freq + (rs.freq < 196));
    m.run_length++;
    m.prev_success = 1;
    m.found = q.stats;
  }
}

// processSymbol1<0> (ppmd.cpp:985-1041)
PP_HD void pp_symbol1(PpmdModel& m, PpmdCtx& q, int symbol) {
  PpmdSt* p = m.pool + q.stats;
  const int cnum = q.ns;
  m.prev_success = 0;
  if (p[0].sym == symbol) {
    p[0].freq = (uint8_t)(p[0].freq + 4);
    q.summ = (uint16_t)(q.summ + 4);
    m.found = q.stats;
  } else {
    int i = 1;
    bool hit = false;
    for (; i <= cnum; ++i) if (p[i].sym == symbol) { hit = true; break; }
    if (hit) {
      p[i].freq = (uint8_t)(p[i].freq + 4);
      q.summ = (uint16_t)(q.summ + 4);
      if (p[i].freq > p[i - 1].freq) { const PpmdSt t = p[i]; p[i] = p[i - 1]; p[i - 1] = t; --i; }
      m.found = q.stats + (uint32_t)i;
    } else {
      m.num_masked = cnum;
      for (i = 0; i <= cnum; ++i) m.char_mask[p[i].sym] = m.esc_count;
      m.found = PPMD_NIL;
    }
  }
  if (m.found != PPMD_NIL && m.pool[m.found].freq > PPMD_MAX_FREQ) m.found = pp_rescale(m, q, m.order_fall, m.found);
}

PP_HD PpmdSee* pp_see_for(PpmdModel& m, const PpmdCtx& q, int* see_freq) {
  const int cnum = q.ns;
  if (cnum != 0xFF) {
    const int col = (q.summ > 10 * (cnum + 1)) + 2 * (2 * cnum < (int)m.ctx[q.suffix].ns + m.num_masked) + q.flags;
    PpmdSee* s = &m.see[m.qtable[cnum + 3] - 4][0] + col;        // may step into the next row exactly as the reference's pointer does
    *see_freq = (int)(s->summ >> s->shift) + 1;
    return s;
  }
  *see_freq = 1;
  return &m.dummy_see;
}

// processSymbol2<0> (ppmd.cpp:1046-1112)
PP_HD void pp_symbol2(PpmdModel& m, PpmdCtx& q, int symbol) {
  PpmdSt* p = m.pool + q.stats;
  const int cnum = q.ns;
  int see_freq;
  PpmdSee* see = pp_see_for(m, q, &see_freq);
  int low = 0, pl = 0, j = 0;
  bool hit = false;
  for (int i = 0; i <= cnum; ++i) {
    const int c = p[i].sym;
    if (m.char_mask[c] != m.esc_count) {
      m.char_mask[c] = m.esc_count;
      low += p[i].freq;
      if (c == symbol) { hit = true; j = i; pl = low; }
    }
  }
  (void)pl;
  const int total = see_freq + low;
  if (hit) {
    if (see_freq > 2) see->summ = (uint16_t)(see->summ - see_freq);
    pp_see_update(*see);
    m.found = q.stats + (uint32_t)j;
    p[j].freq = (uint8_t)(p[j].freq + 4);
    q.summ = (uint16_t)(q.summ + 4);
    if (p[j].freq > PPMD_MAX_FREQ) m.found = pp_rescale(m, q, m.order_fall, m.found);
    m.run_length = m.init_rl;
    m.esc_count++;
  } else {
    m.num_masked = cnum;
    see->summ = (uint16_t)(see->summ + (total - see_freq));
  }
}

// CreateSuccessors (ppmd.cpp:848-917). p_idx: state of `pc_ref` already located by the caller or PPMD_NIL.
PP_HD uint32_t pp_create_successors(PpmdModel& m, bool skip, uint32_t p_idx, uint32_t pc_ref) {
  uint32_t ps[256];
  int n = 0;
  uint8_t sym = m.pool[m.found].sym;
  const uint32_t up_branch = m.pool[m.found].succ;
  uint32_t pc = pc_ref - PPMD_CTX_BASE;
  bool enter_loop = true, at_entry = false;
  if (!skip) {
    ps[n++] = m.found;
    if (!m.ctx[pc].suffix) enter_loop = false;
  }
  if (enter_loop) {
    uint32_t p = p_idx;
    if (p != PPMD_NIL) { pc = m.ctx[pc].suffix; at_entry = true; }
    for (;;) {
      if (!at_entry) {
        pc = m.ctx[pc].suffix;
        PpmdCtx& c = m.ctx[pc];
        if (c.ns) {
          p = c.stats;
          while (m.pool[p].sym != sym) ++p;
          const int t = 2 * (m.pool[p].freq < PPMD_MAX_FREQ - 1);
          m.pool[p].freq = (uint8_t)(m.pool[p].freq + t);
          c.summ = (uint16_t)(c.summ + t);
        } else {
          p = c.stats;
          m.pool[p].freq = (uint8_t)(m.pool[p].freq + ((!m.ctx[c.suffix].ns) & (m.pool[p].freq < 16)));
        }
      }
      at_entry = false;
      if (m.pool[p].succ != up_branch) { pc = m.pool[p].succ - PPMD_CTX_BASE; break; }
      ps[n++] = p;
      if (!m.ctx[pc].suffix) break;
    }
  }
  if (n == 0) return PPMD_CTX_BASE + pc;
  // the new contexts all hold the single symbol that follows the up-branch position in the text
  uint8_t ct_flags = (uint8_t)(0x10 * (sym >= 0x40));
  sym = m.text[up_branch];
  const uint32_t ct_succ = up_branch + 1;
  ct_flags |= (uint8_t)(0x08 * (sym >= 0x40));
  uint8_t ct_freq;
  {
    PpmdCtx& c = m.ctx[pc];
    if (c.ns) {
      uint32_t p = c.stats;
      while (m.pool[p].sym != sym) ++p;
      uint32_t cf = (uint32_t)m.pool[p].freq - 1;
      const uint32_t s0 = (uint32_t)c.summ - c.ns - cf;
      cf = 1 + ((2 * cf < s0) ? (uint32_t)(12 * cf > s0) : 2 + cf / s0);
      ct_freq = (uint8_t)(cf < 7 ? cf : 7);
    } else {
      ct_freq = m.pool[c.stats].freq;
    }
  }
  do {
    const uint32_t nc = pp_alloc_ctx(m);
    if (!nc) return 0;
    const uint32_t sidx = pp_alloc_states(m, 0);
    if (sidx == PPMD_NIL) return 0;
    PpmdCtx& c = m.ctx[nc];
    c.ns = 0; c.flags = ct_flags; c.summ = 0; c.stats = sidx; c.cap_class = 0; c.suffix = pc;
    PpmdSt& s = m.pool[sidx];
    s.sym = sym; s.freq = ct_freq; s.succ = ct_succ; s.pad = 0;
    pc = nc;
    m.pool[ps[--n]].succ = PPMD_CTX_BASE + pc;
  } while (n != 0);
  return PPMD_CTX_BASE + pc;
}

// ReduceOrder (ppmd.cpp:919-965)
PP_HD uint32_t pp_reduce_order(PpmdModel& m, uint32_t p_idx, uint32_t pc_ref) {
  uint32_t pc = pc_ref - PPMD_CTX_BASE;
  const uint32_t pc1 = pc;
  m.pool[m.found].succ = m.text_pos;
  const uint8_t sym = m.pool[m.found].sym;
  const uint32_t up_branch = m.text_pos;
  m.order_fall++;
  uint32_t p = p_idx;
  bool at_entry = false;
  if (p != PPMD_NIL) { pc = m.ctx[pc].suffix; at_entry = true; }
  for (;;) {
    if (!at_entry) {
      if (!m.ctx[pc].suffix) return PPMD_CTX_BASE + pc;
      pc = m.ctx[pc].suffix;
      PpmdCtx& c = m.ctx[pc];
      if (c.ns) {
        p = c.stats;
        while (m.pool[p].sym != sym) ++p;
        const int t = 2 * (m.pool[p].freq < PPMD_MAX_FREQ - 3);
        m.pool[p].freq = (uint8_t)(m.pool[p].freq + t);
        c.summ = (uint16_t)(c.summ + t);
      } else {
        p = c.stats;
        m.pool[p].freq = (uint8_t)(m.pool[p].freq + (m.pool[p].freq < 11));
      }
    }
    at_entry = false;
    if (m.pool[p].succ) break;
    m.pool[p].succ = up_branch;
    m.order_fall++;
  }
  if (m.pool[p].succ <= up_branch) {
    const uint32_t saved = m.found;
    m.found = p;
    const uint32_t r = pp_create_successors(m, false, PPMD_NIL, PPMD_CTX_BASE + pc);
    m.pool[p].succ = r;
    m.found = saved;
  }
  if (m.order_fall == 1 && PPMD_CTX_BASE + pc1 == m.max_context) {
    m.pool[m.found].succ = m.pool[p].succ;
    m.text_pos--;
  }
  return m.pool[p].succ;
}

// UpdateModel (ppmd.cpp:729-846). Returns the new MaxContext reference, 0 when an arena is exhausted.
PP_HD uint32_t pp_update_model(PpmdModel& m, uint32_t min_ref) {
  static const uint8_t kExpEscape[16] = {51, 43, 18, 12, 11, 9, 8, 7, 6, 5, 4, 3, 3, 2, 2, 2};
  const uint32_t minc = min_ref - PPMD_CTX_BASE;
  const uint8_t fsym = m.pool[m.found].sym;
  const uint32_t ffreq = m.pool[m.found].freq;
  uint32_t fsucc = m.pool[m.found].succ;
  uint32_t p = PPMD_NIL;
  if (m.ctx[minc].suffix) {
    PpmdCtx& c = m.ctx[m.ctx[minc].suffix];
    if (c.ns) {
      p = c.stats;
      if (m.pool[p].sym != fsym) {
        for (++p; m.pool[p].sym != fsym; ++p) {}
        if (m.pool[p].freq >= m.pool[p - 1].freq) { const PpmdSt t = m.pool[p]; m.pool[p] = m.pool[p - 1]; m.pool[p - 1] = t; --p; }
      }
      if (m.pool[p].freq < PPMD_MAX_FREQ - 3) {
        const uint32_t cf = 2 + (ffreq < 28);
        m.pool[p].freq = (uint8_t)(m.pool[p].freq + cf);
        c.summ = (uint16_t)(c.summ + cf);
      }
    } else {
      p = c.stats;
      m.pool[p].freq = (uint8_t)(m.pool[p].freq + (m.pool[p].freq < 14));
    }
  }
  if (!m.order_fall && fsucc) {
    const uint32_t r = pp_create_successors(m, true, p, min_ref);
    m.pool[m.found].succ = r;
    if (!r) return 0;
    m.max_context = r;
    return r;
  }
  if (m.text_pos + 1 >= m.text_cap) { m.error = 1; return 0; }
  m.text[m.text_pos++] = fsym;
  uint32_t succ = m.text_pos;
  if (fsucc) {
    if (!pp_is_ctx(fsucc)) fsucc = pp_create_successors(m, false, p, min_ref);
  } else {
    fsucc = pp_reduce_order(m, p, min_ref);
  }
  if (!fsucc) return 0;
  if (!--m.order_fall) {
    succ = fsucc;
    m.text_pos -= (m.max_context != min_ref);
  }
  const uint32_t s0 = (uint32_t)pp_summ_read(m, m.ctx[minc]) - ffreq;
  const uint32_t ns = m.ctx[minc].ns;
  const uint8_t flag = (uint8_t)(0x08 * (fsym >= 0x40));
  for (uint32_t pc = m.max_context - PPMD_CTX_BASE; pc != minc; pc = m.ctx[pc].suffix) {
    PpmdCtx& c = m.ctx[pc];
    const uint32_t ns1 = c.ns;
    if (ns1) {
      if (ns1 + 1 == (1u << c.cap_class)) {                    // array full: move to the next capacity class
        const uint32_t nidx = pp_alloc_states(m, c.cap_class + 1);
        if (nidx == PPMD_NIL) return 0;
        for (uint32_t k = 0; k <= ns1; ++k) m.pool[nidx + k] = m.pool[c.stats + k];
        pp_free_states(m, c.stats, c.cap_class);
        c.stats = nidx; c.cap_class++;
      }
      c.summ = (uint16_t)(c.summ + (m.qtable[ns + 4] >> 3));
    } else {
      const uint32_t nidx = pp_alloc_states(m, 1);
      if (nidx == PPMD_NIL) return 0;
      m.pool[nidx] = m.pool[c.stats];
      pp_free_states(m, c.stats, c.cap_class);
      c.stats = nidx; c.cap_class = 1;
      PpmdSt& s = m.pool[nidx];
      s.freq = (uint8_t)((s.freq <= PPMD_MAX_FREQ / 3) ? (2 * s.freq - 1) : (PPMD_MAX_FREQ - 15));
      c.summ = (uint16_t)(s.freq + (ns > 1) + kExpEscape[m.qtable[m.bsumm >> 8]]);
    }
    uint32_t cf = (ffreq - 1) * (5 + (uint32_t)c.summ);
    const uint32_t sf = s0 + c.summ;
    if (cf <= 3 * sf) {
      cf = 1 + (2 * cf > sf) + (2 * cf > 3 * sf);
      c.summ = (uint16_t)(c.summ + 4);
    } else {
      cf = 5 + (cf > 5 * sf) + (cf > 6 * sf) + (cf > 8 * sf) + (cf > 10 * sf) + (cf > 12 * sf);
      c.summ = (uint16_t)(c.summ + cf);
    }
    c.ns = (uint8_t)(c.ns + 1);
    PpmdSt& s = m.pool[c.stats + c.ns];
    s.succ = succ; s.sym = fsym; s.freq = (uint8_t)cf; s.pad = 0;
    c.flags |= flag;
  }
  m.max_context = fsucc;
  return fsucc;
}

// ppmd_UpdateByte (ppmd.cpp:1283-1314)
PP_HD void ppmd_update_byte(PpmdModel& m, int c) {
  if (m.error) return;                       // arena exhausted earlier: the model is frozen, the engine reports CMIXB200_ERR_CAPACITY
  uint32_t minc = m.max_context;
  if (pp_ctx(m, minc).ns) pp_symbol1(m, pp_ctx(m, minc), c); else pp_bin_symbol(m, pp_ctx(m, minc), c);
  while (m.found == PPMD_NIL) {
    do {
      m.order_fall++;
      minc = PPMD_CTX_BASE + pp_ctx(m, minc).suffix;
    } while (pp_ctx(m, minc).ns == m.num_masked);
    pp_symbol2(m, pp_ctx(m, minc), c);
  }
  PP_TICK(m, 0);
  uint32_t r;
  if (m.order_fall != 0 || !pp_is_ctx(m.pool[m.found].succ)) r = pp_update_model(m, minc);
  else { r = m.pool[m.found].succ; m.max_context = r; }
  if (!r) m.error = 1;                       // the reference would cut the model off here (RestoreModelRare)
  PP_TICK(m, 1);
}

PP_HD void pp_sq_store(PpmdModel& m, uint32_t sym, uint32_t freq, uint32_t total) {
  PpmdQ& q = m.sq[m.sq_n++];
  q.sym = (uint16_t)sym; q.freq = (uint16_t)freq; q.total = (uint16_t)total;
}

// ppmd_PrepareByte + ConvertSQ (ppmd.cpp:1256-1281, 1116-1140): fills m.sqp[256].
PP_HD void ppmd_prepare_byte(PpmdModel& m) {
  if (m.error) return;
  m.sq_n = 0; m.num_masked = 0;
  const int saved_fall = m.order_fall;
  uint32_t minc = m.max_context - PPMD_CTX_BASE;
  {
    PpmdCtx& q = m.ctx[minc];
    PpmdSt* p = m.pool + q.stats;
    if (q.ns) {                                                   // processSymbol1_T
      const int cnum = q.ns, total = q.summ;
      int low = 0;
      for (int i = 0; i <= cnum; ++i) { pp_sq_store(m, p[i].sym, p[i].freq, (uint32_t)total); low += p[i].freq; }
      m.num_masked = cnum;
      for (int i = 0; i <= cnum; ++i) m.char_mask[p[i].sym] = m.esc_count;
      pp_sq_store(m, 256, (uint32_t)(total - low), (uint32_t)total);
    } else {                                                      // processBinSymbol_T
      const int i = m.ns2bs[m.ctx[q.suffix].ns] + m.prev_success + q.flags + ((m.run_length >> 26) & 0x20);
      m.bsumm = m.bin_summ[m.qtable[p[0].freq - 1]][i];
      pp_sq_store(m, p[0].sym, (uint32_t)(m.bsumm + m.bsumm), PPMD_SCALE);
      pp_sq_store(m, 256, (uint32_t)(PPMD_SCALE - m.bsumm - m.bsumm), PPMD_SCALE);
      m.char_mask[p[0].sym] = m.esc_count;
      m.num_masked = 0;
    }
  }
  for (;;) {
    bool done = false;
    do {
      if (!m.ctx[minc].suffix) { done = true; break; }
      m.order_fall++;
      minc = m.ctx[minc].suffix;
    } while (m.ctx[minc].ns == m.num_masked);
    if (done) break;
    PpmdCtx& q = m.ctx[minc];                                     // processSymbol2_T
    PpmdSt* p = m.pool + q.stats;
    const int cnum = q.ns;
    int see_freq;
    (void)pp_see_for(m, q, &see_freq);
    int low = 0;
    for (int i = 0; i <= cnum; ++i) if (m.char_mask[p[i].sym] != m.esc_count) low += p[i].freq;
    const int total = see_freq + low;
    for (int i = 0; i <= cnum; ++i) {
      const int c = p[i].sym;
      if (m.char_mask[c] != m.esc_count) { pp_sq_store(m, (uint32_t)c, p[i].freq, (uint32_t)total); m.char_mask[c] = m.esc_count; }
    }
    pp_sq_store(m, 256, (uint32_t)see_freq, (uint32_t)total);
    m.num_masked = cnum;
  }
  m.esc_count++; m.num_masked = 0; m.order_fall = saved_fall;
  PP_TICK(m, 2);
  uint32_t cum = 0xFFFFFF00u;
  for (int i = 0; i < 256; ++i) m.sqp[i] = 0;
  for (uint32_t i = 0; i < m.sq_n; ++i) {
    const uint32_t c = m.sq[i].sym, freq = m.sq[i].freq, total = m.sq[i].total;
    const uint32_t prob = (uint32_t)(((uint64_t)cum * freq) / total);
    if (c < 256) m.sqp[c] = prob + 1; else cum = prob;
  }
  PP_TICK(m, 3);
}

}  // namespace cmixb200
#endif

// cmix_b200/csrc/paq8.cuh — the resident PAQ8 model on the device.
//
// One CTA of 12 warps per stream evaluates paq8_top.h's `bit()` with the work of a bit spread over lanes:
//  * every context of every context map is a lane (210 lanes for the sixteen 7-slot maps on warps 0-6, 63 for the three
//    history maps on warps 7-8): bucket probe, bit-history step, state maps and the 5 / 7 mixer inputs of a context are
//    independent of the other contexts of its map as long as they touch different 64-byte buckets this bit. That is CHECKED
//    per bit (touched_buckets, including the buckets a deferred history write-back will reach); a map with a clash is
//    evaluated by one lane in the reference's order instead.
//  * the one global coupling of the 7-slot maps, the shared pseudo-random sequence that ages high-count states
//    (paq8.cpp:1075), is resolved between the two passes: pass 1 computes each context's aged state and flags the contexts
//    that draw; warp 0 numbers the flagged contexts in program order (ballot masks + a prefix over the maps) and produces the
//    draws 24 at a time (the generator is a lagged XOR: x[i] = x[i-24] ^ x[i-55]); pass 2 applies them. A bit with a
//    clashing map falls back to one lane walking the maps in order.
//  * the match models, the DMC forest, the run maps and the direct maps run one unit per lane on warps 9-11 beside the
//    context maps; on a bit inside a byte all of this is three phases (probe / number / apply). Byte boundaries add the
//    context hashing of the word / nest / indirect / XML / text / x86 / record models (model per lane, two rounds because
//    the sparse and record models consume what the order-N map and the match model produce in the same bit) and the three
//    OLS predictors (one warp each: rank-1 covariance update with coalesced columns, Cholesky with a row per lane in
//    registers, substitutions in the reference's summation order).
//  * the 28 selected int16 weight sets (1552 weights each) are CACHED in shared memory from the dot product of one bit to
//    the SGD step of the next and written back to HBM only when a selector moves to another set; dot products and SGD use
//    all 384 lanes (integer sums: exact under any association).
// The 55 KB state block, the hot read-only tables (21 KB) and the 87 KB weight cache live in shared memory for the launch;
// the ~10 GB of model memory stays in HBM.
// PAQ8 is a producer like FXCM: it depends on the coded bytes only and writes 1591 codes per bit into the `ext` scratch.
#pragma once
#include "paq8_top.h"
#include "state.h"

namespace cmixb200 {

enum { P8_SEEN = 4096 };

// -DP8_PROF: lane 0 accumulates the cycles between phase boundaries (byte-boundary bits and the others apart)
#ifdef P8_PROF
__device__ unsigned long long g_p8_prof[2][96];
#define P8_T(k) do { if (tid == 0) { const long long now_ = clock64(); atomicAdd(&g_p8_prof[sh.prof_row][k], (unsigned long long)(now_ - sh.prof_t)); sh.prof_t = now_; } } while (0)
#define P8_M0 const long long m0_ = clock64()
#define P8_M(slot) atomicAdd(&g_p8_prof[sh.prof_row][slot], (unsigned long long)(clock64() - m0_))
#define P8_C(slot, n) atomicAdd(&g_p8_prof[sh.prof_row][slot], (unsigned long long)(n))
#else
#define P8_T(k) do { } while (0)
#define P8_M0 do { } while (0)
#define P8_M(slot) do { } while (0)
#define P8_C(slot, n) do { } while (0)
#endif
enum { P8_THREADS = 512, P8_WARPS = 16, P8_MAP_THREADS = 384, P8_MAP_WARPS = 12, P8_SGD_THREADS = 128, P8_N_CM = 16, P8_N_CM2 = 3, P8_CM_LANES = 210, P8_CM2_LANES = 63, P8_N_UNITS = 53,
       P8_CM2_TID0 = 224, P8_TID_PIC = 287, P8_TID_MATCH = 288, P8_TID_W10 = 320, P8_TID_W11 = 352 };

struct P8Shared {
  p8::State S;
  alignas(16) short wc[p8::N_SETS][p8::N_IN];   // the weight sets of the pending prediction (slot i holds set wc_set[i])
  alignas(16) unsigned char tab[p8::TABLES_HOT_BYTES];
  alignas(16) short tx_old[p8::N_IN];           // the inputs of the previous bit: the SGD warps train while the map warps write new ones
  int wc_set[p8::N_SETS];
  int sgd_err[p8::N_SETS], sgd_nx, sgd_ncxt;
  int unit_off[P8_N_UNITS + 1];
  // pass-1 results of the 7-slot maps
  short ns[P8_CM_LANES];
  u32 ids[P8_CM_LANES][5];
  u32 ids2[P8_CM2_LANES][5];
  int clash[P8_N_CM], clash2[P8_N_CM2];
  int order, res2[P8_N_CM2];
  int dot[p8::N_SETS];
  u32 snap_spaces, snap_words, snap_frstchar, snap_spafdo;
  int dmc_st[10];
  u32 flag_mask[8];            // ballot of "this context draws" per warp of map lanes
  int flag_base[P8_N_CM];      // index of a map's first draw in draws[]
  u32 draws[P8_CM_LANES + 6];
  u32 rnd_prev[64]; int rnd_prev_i;   // the generator as the bit found it: flagged contexts compute their own draw from it
  int any_clash, text_pending;
  union {
    unsigned long long seen[P8_SEEN];   // open-addressing set of (map, bucket) pairs touched this bit
    struct { double ch[3][32 * 33]; double pb[3][32]; } ols;   // Cholesky factor rows (padded) and a product buffer; byte boundaries only
  } u;
#ifdef P8_PROF
  long long prof_t; int prof_row;
#endif
};

// program order of the sixteen 7-slot maps, their lane capacity and the number of contexts a full byte sets
__device__ __forceinline__ p8::Cm& p8_cm(p8::State& S, int k) {
  switch (k) {
    case 0: return S.sparse.cm; case 1: return S.sparse1.cm; case 2: return S.distance.cm; case 3: return S.record.cm; case 4: return S.record.cn;
    case 5: return S.record.co; case 6: return S.record.cp; case 7: return S.record1.cm; case 8: return S.record1.cn; case 9: return S.record1.co;
    case 10: return S.record1.cq; case 11: return S.record1.cp; case 12: return S.word.cm; case 13: return S.nest.cm; case 14: return S.indirect.cm;
    default: return S.xml.cm;
  }
}
__device__ __forceinline__ p8::Cm2& p8_cm2(p8::State& S, int k) { return k == 0 ? S.cm : (k == 1 ? S.text.map : S.exe.cm); }
__constant__ unsigned char c_p8_cm_cap[P8_N_CM] = {42, 31, 3, 3, 3, 3, 16, 2, 5, 4, 3, 3, 61, 12, 15, 4};
__constant__ unsigned char c_p8_cm_base[P8_N_CM + 1] = {0, 42, 73, 76, 79, 82, 85, 101, 103, 108, 112, 115, 118, 179, 191, 206, 210};
__constant__ unsigned char c_p8_cm_full[P8_N_CM] = {42, 29, 3, 3, 3, 3, 16, 2, 5, 4, 3, 3, 57, 12, 15, 4};
__constant__ unsigned char c_p8_cm_unit[P8_N_CM] = {9, 10, 18, 20, 21, 22, 23, 36, 37, 38, 39, 40, 41, 42, 43, 45};
__constant__ unsigned char c_p8_cm2_cap[P8_N_CM2] = {10, 33, 20};
__constant__ unsigned char c_p8_cm2_unit[P8_N_CM2] = {3, 46, 47};
// unit u of the mixer-input order (paq8_top.h context_model): >= 0: fixed number of inputs, -1-k: 7-slot map k, -20-k: history map k
__constant__ signed char c_p8_unit_kind[P8_N_UNITS] = {
    1, 1, 1, -20, 1, 1, 1, 17, 11, -1, -2, 2, 2, 2, 2, 2, 2, 2, -3, 3, -4, -5, -6, -7, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2,
    -8, -9, -10, -11, -12, -13, -14, -15, 6, -16, -21, -22, 2, 2, 2, 2, 2};

__device__ __forceinline__ int p8_unit_count(p8::State& S, int u, bool byte_start) {
  const int kind = c_p8_unit_kind[u];
  if (kind >= 0) return kind;
  if (kind <= -20) { const int k = -20 - kind; return 7 * (byte_start ? (int)c_p8_cm2_cap[k] : p8_cm2(S, k).index); }
  const int k = -1 - kind;
  return 5 * (byte_start ? (int)c_p8_cm_full[k] : p8_cm(S, k).cn);
}

__device__ __forceinline__ p8::Out p8_out(P8Shared& sh, int offset) { p8::Out o; o.T = sh.S.T; o.tx = sh.S.m.tx; o.codes = sh.S.codes; o.n = offset; return o; }

__device__ __forceinline__ void p8_copy_words(void* dst, const void* src, size_t bytes, int tid) {
  u32* d = (u32*)dst; const u32* s = (const u32*)src;
  for (size_t i = tid; i < bytes / 4; i += P8_THREADS) d[i] = s[i];
}
static_assert(sizeof(p8::State) % 4 == 0, "state block is copied word by word");

// One OLS predictor of the linear-prediction model at a byte boundary, on one warp: OLS::Update(val) with the byte just
// coded, then Add() of the 32 new taps and Predict() (paq8_top.h ols_update / linear_predict, reference :4476-4502).
// Element-wise steps use any lane mapping; every SUM runs in the reference's order:
//  * covariance: lane = column, 32 coalesced rows (the matrix stays exactly symmetric, so the lane also holds row `lane`);
//  * Cholesky: lane r keeps row r in registers; column c takes row c's finished entries from shared memory;
//  * forward substitution column by column (row i subtracts w[0..i-1] in ascending order), backward substitution row by
//    row with the products gathered in shared memory and summed in ascending order.
__device__ __noinline__ void p8_ols_byte_warp(P8Shared& sh, int k, int lane) {
  using namespace p8;
  State& S = sh.S;
  LinearM& M = S.linear;
  const double lambda = 0.995, nu = 0.001, one_minus = 1.0 - 0.995;
  double* blk = M.ols + (size_t)k * OLS_STRIDE;
  double* x = blk; double* w = blk + 32; double* b = blk + 64; double* cov = blk + 96;
  double* chs = sh.u.ols.ch[k]; double* pb = sh.u.ols.pb[k];
  double* row = chs + lane * 33;
  const unsigned full = 0xffffffffu;
  const double val = (double)(u8)buf(S, 1);
  const double xl = x[lane];
  int km = M.ols_km[k] + 1;
  const bool solve = km >= 4;
  {
    double c[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) c[j] = cov[j * 32 + lane];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const double xj = __shfl_sync(full, xl, j);
      const double v = P8_DADD(P8_DMUL(lambda, c[j]), P8_DMUL(one_minus, P8_DMUL(xj, xl)));
      cov[j * 32 + lane] = v;
      if (solve) row[j] = (j == lane) ? P8_DADD(v, nu) : v;     // the matrix is exactly symmetric: column `lane` is row `lane`
    }
  }
  double bl = b[lane];
  bl = P8_DADD(P8_DMUL(lambda, bl), P8_DMUL(one_minus, P8_DMUL(xl, val)));
  b[lane] = bl;
  double wl = w[lane];
  if (solve) {
    __syncwarp();
    bool fail = false;
    for (int col = 0; col < 32 && !fail; ++col) {
      const double* rc = chs + col * 33;
      double s = row[col];
#pragma unroll 4
      for (int q = 0; q < col; ++q) s = P8_DSUB(s, P8_DMUL(row[q], rc[q]));
      const double d = __shfl_sync(full, s, col);
      if (d > 1E-8) {
        const double dd = P8_DSQRT(d);
        if (lane >= col) row[col] = (lane == col) ? dd : P8_DDIV(s, dd);
      } else fail = true;
      __syncwarp();
    }
    if (!fail) {
      double sum = bl;
      for (int q = 0; q < 32; ++q) {
        const double wq = P8_DDIV(__shfl_sync(full, sum, q), chs[q * 33 + q]);
        if (lane == q) wl = wq;
        sum = P8_DSUB(sum, P8_DMUL(row[q], wq));          // rows below q; the others hold values nobody reads
      }
      const double zl = wl;
      for (int i = 31; i >= 0; --i) {
        pb[lane] = P8_DMUL(row[i], wl);                   // ch[lane][i] * w[lane], read for lane > i only
        __syncwarp();
        double s = __shfl_sync(full, zl, i);
#pragma unroll 4
        for (int j = i + 1; j < 32; ++j) s = P8_DSUB(s, pb[j]);
        const double wi = P8_DDIV(s, chs[i * 33 + i]);
        if (lane == i) wl = wi;
        __syncwarp();
      }
      w[lane] = wl;
    }
    km = 0;
  }
  // Add() the taps of this predictor and Predict()
  const int i1 = lane + 1;
  const double xn = (double)(u8)buf(S, k == 0 ? i1 : (k == 1 ? 2 * i1 - 1 : 2 * i1));
  x[lane] = xn;
  pb[lane] = P8_DMUL(wl, xn);
  __syncwarp();
  if (lane == 0) {
    double sum = 0.;
    for (int i = 0; i < 32; ++i) sum = P8_DADD(sum, pb[i]);
    M.prd[k] = (u8)clip8((int)P8_FLOOR(sum));
    M.ols_km[k] = km;
  }
  __syncwarp();
}

// Insert the (map, bucket) pairs of one context into the per-bit set; returns true when a pair was already there, i.e. another
// context of the same map touches the same 64-byte bucket this bit. A context's own repeats are removed first.
__device__ bool p8_claim(unsigned long long* seen, int map, const u32* ids, int n) {
  bool clash = false;
  for (int a = 0; a < n; ++a) {
    bool dup = false;
    for (int b = 0; b < a; ++b) dup = dup || ids[b] == ids[a];
    if (dup) continue;
    const unsigned long long key = ((unsigned long long)(map + 1) << 32) | ids[a];
    u32 slot = (u32)((key * 0x9E3779B97F4A7C15ull) >> 52) & (P8_SEEN - 1);
    for (;;) {
      const unsigned long long old = atomicCAS(&seen[slot], 0ull, key);
      if (old == 0ull) break;
      if (old == key) { clash = true; break; }
      slot = (slot + 1) & (P8_SEEN - 1);
    }
  }
  return clash;
}

// map (k, i) of a 7-slot-map lane, false if the lane is beyond the maps
__device__ __forceinline__ bool p8_cm_lane(int lane, int& k, int& i) {
  if (lane >= P8_CM_LANES) return false;
  k = 0;
#pragma unroll
  for (int q = 1; q < P8_N_CM; ++q) k += lane >= (int)c_p8_cm_base[q];
  i = lane - c_p8_cm_base[k];
  return true;
}
__device__ __forceinline__ bool p8_cm2_lane(int lane, int& k, int& i) {
  if (lane < 0 || lane >= P8_CM2_LANES) return false;
  k = (lane >= 10) + (lane >= 43);
  i = lane - (k == 0 ? 0 : (k == 1 ? 10 : 43));
  return true;
}
// how many contexts of the lanes [lo, hi) draw this bit
__device__ __forceinline__ int p8_flags_in(const u32* mask, int lo, int hi) {
  int r = 0;
  for (int w = lo >> 5; w <= ((hi - 1) >> 5) && hi > lo; ++w) {
    u32 m = mask[w];
    if (w == (lo >> 5)) m &= ~0u << (lo & 31);
    if (w == (hi >> 5)) m &= (1u << (hi & 31)) - 1u;
    r += __popc(m);
  }
  return r;
}

// ---- the pieces of a bit ----------------------------------------------------------------------------------------------
// probe of the history maps (lanes P8_CM2_TID0 ..): buckets each context touches this bit
__device__ __noinline__ void p8_probe_cm2(P8Shared& sh, int tid, int bpos) {
  int k, i;
  if (!p8_cm2_lane(tid - P8_CM2_TID0, k, i)) return;
  p8::Cm2& m = p8_cm2(sh.S, k);
  u32* ids = sh.ids2[tid - P8_CM2_TID0];
  const int n = i < m.index ? p8::cm2_touched(m, i, bpos, ids) : 0;
  if (n && p8_claim(sh.u.seen, P8_N_CM + k, ids, n)) sh.clash2[k] = 1;
}
// pass 1 of the 7-slot maps (warps 0-6, whole warps): aged state, draw flag, touched buckets
__device__ __noinline__ void p8_probe_cm(P8Shared& sh, int tid, int y, int c0, int bpos) {
  using namespace p8;
  int k, i;
  bool flag = false;
  if (p8_cm_lane(tid, k, i)) {
    Cm& m = p8_cm(sh.S, k);
    int ns = -1;
    if (i < m.cn) {
      ns = cm_next_state(*sh.S.T, m, i, y);
      const int n = cm_touched(m, i, c0, bpos, sh.ids[tid]);
      if (p8_claim(sh.u.seen, k, sh.ids[tid], n)) { sh.clash[k] = 1; sh.any_clash = 1; }
    }
    sh.ns[tid] = (short)ns;
    flag = ns >= 204;
  }
  const u32 mask = __ballot_sync(0xffffffffu, flag);
  if ((tid & 31) == 0) sh.flag_mask[tid >> 5] = mask;
}
// the units that are one lane each (warps 8-11)
__device__ __noinline__ void p8_probe_single(P8Shared& sh, int tid, int y, int c0, int bpos) {
  using namespace p8;
  State& S = sh.S;
  const p8::Tables& T = *S.T;
  if (tid == P8_TID_PIC) pic_core(S);
  else if (tid == P8_TID_MATCH) { Out o = p8_out(sh, sh.unit_off[7]); match_core(S, o); }
  else if (tid == P8_TID_MATCH + 1) { if (bpos != 0) record_pre(S); }   // on a byte boundary it follows record_byte (round 2)
  else if (tid >= P8_TID_W10 && tid < P8_TID_W10 + 10) sh.dmc_st[tid - P8_TID_W10] = dmc_st(T, S.dmc[tid - P8_TID_W10], y);
  else if (tid >= P8_TID_MATCH + 2 && tid < P8_TID_MATCH + 5) {
    const int r = tid - (P8_TID_MATCH + 2);
    Out o = p8_out(sh, sh.unit_off[4 + r]);
    rcm_mix(r == 0 ? S.rcm7 : r == 1 ? S.rcm9 : S.rcm10, o, c0, bpos);
  } else if (tid >= P8_TID_W11 + 4 && tid < P8_TID_W11 + 9) { const int r = tid - (P8_TID_W11 + 4); Out o = p8_out(sh, sh.unit_off[48 + r]); linear_small(S, o, r); }
  else if (tid == P8_TID_W11) { Out o = p8_out(sh, sh.unit_off[8]); smatch_head(S, o); }
  else if (tid == P8_TID_W11 + 1 || tid == P8_TID_W11 + 2) {
    const int r = tid - (P8_TID_W11 + 1);
    Out o = p8_out(sh, sh.unit_off[1 + r]);
    add(o, (stretch(T, sm32_p(T, r == 0 ? S.sm0 : S.sm1, y, r == 0 ? c0 : (c0 | (buf(S, 1) << 8)))) + 1) >> 1);
  } else if (tid == P8_TID_W11 + 3) { Out o = p8_out(sh, sh.unit_off[0]); add(o, 64); }
}
// x[i0 + 1 + r] of the lagged generator x[i] = x[i-24] ^ x[i-55] from the table as it stood at i0
__device__ __forceinline__ u32 p8_draw_at(const u32* prev, int i0, int r) {
  if (r < 24) { const int idx = i0 + 1 + r; return prev[(idx - 24) & 63] ^ prev[(idx - 55) & 63]; }
  u32 t[64], v = 0;
  for (int j = 0; j < 64; ++j) t[j] = prev[j];
  for (int j = 0; j <= r; ++j) { const int idx = i0 + 1 + j; v = t[(idx - 24) & 63] ^ t[(idx - 55) & 63]; t[idx & 63] = v; }
  return v;
}
// advance the generator by the number of flagged contexts (one warp; 24 new values depend on old ones only)
__device__ __forceinline__ void p8_advance_rnd(P8Shared& sh, int lane) {
  p8::Rnd& r = sh.S.rnd;
  const int total = p8_flags_in(sh.flag_mask, 0, P8_CM_LANES);
  const int i0 = r.i;
  __syncwarp();
  for (int s0 = 0; s0 < total; s0 += 24) {
    const int n = min(24, total - s0);
    const int idx = i0 + 1 + s0 + lane;
    u32 v = 0;
    if (lane < n) v = r.table[(idx - 24) & 63] ^ r.table[(idx - 55) & 63];
    __syncwarp();
    if (lane < n) r.table[idx & 63] = v;
    __syncwarp();
  }
  if (lane == 0) r.i = i0 + total;
}
// A bit in which two contexts of one 7-slot map meet in a bucket: one lane walks the maps in order, running the clashing ones
// on the spot and laying out the draws of the others (draws[flag_base[k] + rank in map]).
__device__ __noinline__ void p8_number(P8Shared& sh, int tid, int y, int c0, int bpos) {
  using namespace p8;
  State& S = sh.S;
  if (tid != 0) return;
  const int c1 = buf(S, 1);
  int cur = 0;
  for (int k = 0; k < P8_N_CM; ++k) {
    Cm& m = p8_cm(S, k);
    if (sh.clash[k]) { Out o = p8_out(sh, sh.unit_off[c_p8_cm_unit[k]]); cm_mix(m, o, S.rnd, y, c0, bpos, c1); }
    else {
      const int n = p8_flags_in(sh.flag_mask, c_p8_cm_base[k], c_p8_cm_base[k + 1]);
      sh.flag_base[k] = cur;
      for (int j = 0; j < n; ++j) sh.draws[cur++] = rnd_next(S.rnd);
    }
  }
}
// pass 2
__device__ __noinline__ void p8_apply_cm2(P8Shared& sh, int tid, int y, int bpos) {
  int k, i;
  if (!p8_cm2_lane(tid - P8_CM2_TID0, k, i)) return;
  p8::Cm2& m = p8_cm2(sh.S, k);
  const int off = sh.unit_off[c_p8_cm2_unit[k]];
  if (!sh.clash2[k]) {
    if (i < m.index) { p8::Out o = p8_out(sh, off + 7 * i); if (p8::cm2_step(m, i, o, y, bpos)) atomicAdd(&sh.res2[k], 1); }
  } else if (i == 0) {      // in-order evaluation by one lane (the two loops of ContextMap2::mix)
    p8::Out o = p8_out(sh, off);
    sh.res2[k] = p8::cm2_mix_body(m, o, y, bpos);
  }
}
__device__ __noinline__ void p8_apply_cm(P8Shared& sh, int tid, int y, int c0, int bpos) {
  using namespace p8;
  const bool slow = sh.any_clash != 0;
  if (!slow && tid < 32) p8_advance_rnd(sh, tid);      // flagged lanes read the snapshot, not the live table
  int k, i;
  if (!p8_cm_lane(tid, k, i)) return;
  Cm& m = p8_cm(sh.S, k);
  if (sh.clash[k] || i >= m.cn) return;
  int ns = sh.ns[tid];
  if (ns >= 204) {
    const u32 r = slow ? sh.draws[sh.flag_base[k] + p8_flags_in(sh.flag_mask, c_p8_cm_base[k], tid)]
                       : p8_draw_at(sh.rnd_prev, sh.rnd_prev_i, p8_flags_in(sh.flag_mask, 0, tid));
    if (cm_draw_hits(r, ns)) ns -= 4;
  }
  Out o = p8_out(sh, sh.unit_off[c_p8_cm_unit[k]] + 5 * i);
  cm_step(m, i, o, ns, y, c0, bpos, buf(sh.S, 1));
}
__device__ __noinline__ void p8_apply_small(P8Shared& sh, int tid, int y, int bpos) {
  using namespace p8;
  State& S = sh.S;
  if (tid >= P8_TID_MATCH && tid < P8_TID_MATCH + 9) {   // the nine maps behind the match model
    Out o = p8_out(sh, sh.unit_off[7]);
    match_unit(S, o, tid - P8_TID_MATCH);
  } else if (tid >= P8_TID_W10 && tid < P8_TID_W10 + 12) {   // the record model's 12 direct maps (contexts selected by record_pre)
    Out o = p8_out(sh, sh.unit_off[24 + (tid - P8_TID_W10)]);
    record_small(S, o, tid - P8_TID_W10);
  } else if (tid == P8_TID_W10 + 12) {                     // DMC forest combination and reset (dmcForest::mix)
    Out o = p8_out(sh, sh.unit_off[44]);
    const u32 params[10] = {2, 32, 64, 4, 128, 8, 256, 16, 1024, 1536};
    add(o, sh.dmc_st[9] >> 3);
    add(o, sh.dmc_st[8] >> 3);
    for (int i = 7; i > 0; i -= 2) add(o, (sh.dmc_st[i] + sh.dmc_st[i - 1]) >> 4);
    if (bpos == 0)
      for (int i = 7; i >= 0; --i)
        if ((S.dmc[i].extra >> 7) > S.dmc[i].size) dmc_reset(S.dmc[i], params[i]);
  } else if (tid >= P8_TID_W11 && tid < P8_TID_W11 + 7) {   // sparseModel1's seven stationary maps
    Out o = p8_out(sh, sh.unit_off[11 + (tid - P8_TID_W11)]);
    scm_mix(S.sparse1.scm[tid - P8_TID_W11], o, y);
  } else if (tid >= P8_TID_W11 + 7 && tid < P8_TID_W11 + 11) {
    Out o = p8_out(sh, sh.unit_off[8]);
    smatch_unit(S, o, tid - (P8_TID_W11 + 7));
  } else if (tid >= P8_TID_W11 + 11 && tid < P8_TID_W11 + 14) {
    Out o = p8_out(sh, sh.unit_off[19]);
    pic_unit(S, o, tid - (P8_TID_W11 + 11));
  }
}

// The SSE stage (paq8_top.h sse_stage) on one warp: the APMs of a level side by side.
__device__ void p8_sse_warp(P8Shared& sh, int pr0, int lane) {
  using namespace p8;
  State& S = sh.S;
  const p8::Tables& T = *S.T;
  const unsigned full = 0xffffffffu;
  const int y = S.y, c0 = S.c0, bpos = S.bpos;
  const u32 c4 = S.c4;
  u16* codes = S.codes + S.m.n2 + S.m.ncxt;
  const u32 mlen = umin(3, ilog2(S.st_match_length + 1));
  int p = 0, q = 0, pr, pr1, pr2, pr3, pr0b;
  if (S.st_type == FT_TEXT) {
    const int limit = 0x3FF >> ((S.blpos < 0xFFF) * 2);
    if (lane < 4) {
      int cx;
      if (lane == 0) cx = (c0 << 8) | (S.st_text_mask & 0xF) | (int)((S.st_misses & 0xF) << 4);
      else if (lane == 1) cx = (int)finalize64(hash(sx(bpos), S.st_misses & 3, (u64)(c4 & 0xffff), (u64)(S.st_text_mask >> 4)), 16);
      else if (lane == 2) cx = (int)finalize64(hash(sx(c0), S.st_match_expected, mlen), 16);
      else cx = (int)finalize64(hash(sx(c0), (u64)(c4 & 0xffff), S.st_text_first), 16);
      p = apm_p(T, S.text_apm[lane], y, pr0, cx, limit);
    }
    pr = __shfl_sync(full, p, 0); pr1 = __shfl_sync(full, p, 1); pr2 = __shfl_sync(full, p, 2); pr3 = __shfl_sync(full, p, 3);
    pr0b = (pr0 + pr1 + pr2 + pr3 + 2) >> 2;
    if (lane < 3) {
      int cx;
      if (lane == 0) cx = (int)finalize64(hash(S.st_match_expected, mlen, (u64)(c4 & 0xff)), 16);
      else if (lane == 1) cx = (int)finalize64(hash(sx(c0), (u64)(c4 & 0x00ffffff)), 16);
      else cx = (int)finalize64(hash(sx(c0), (u64)(c4 & 0xffffff00)), 16);
      q = apm1_p(T, S.text_apm1[lane], y, lane == 0 ? pr0b : pr, cx, lane == 0 ? 7 : 6);
    }
  } else {
    const u16 ctx1 = (u16)(c0 | buf(S, 1) << 8);
    const u16 ctx2 = (u16)(c0 ^ finalize64(hash((u64)(c4 & 0xffff)), 16));
    const u16 ctx3 = (u16)(c0 ^ finalize64(hash((u64)(c4 & 0xffffff)), 16));
    if (lane < 4) {
      const int cx = lane == 0 ? (int)((mlen << 11) | ((u32)c0 << 3) | (u32)(S.st_misses & 0x7)) : lane == 1 ? (int)ctx1 : lane == 2 ? (int)ctx2 : (int)ctx3;
      p = apm1_p(T, S.generic_apm1[lane], y, pr0, cx);
    }
    pr = __shfl_sync(full, p, 0); pr1 = __shfl_sync(full, p, 1); pr2 = __shfl_sync(full, p, 2); pr3 = __shfl_sync(full, p, 3);
    pr0b = (pr0 + pr1 + pr2 + pr3 + 2) >> 2;
    if (lane < 3) {
      const int cx = lane == 0 ? ((S.st_match_expected << 8) | buf(S, 1)) : lane == 1 ? (int)ctx2 : (int)ctx3;
      q = apm1_p(T, S.generic_apm1[4 + lane], y, pr, cx);
    }
  }
  const int q1 = __shfl_sync(full, q, 0), q2 = __shfl_sync(full, q, 1), q3 = __shfl_sync(full, q, 2);
  const int prf = (pr + q1 + q2 + q3 + 2) >> 2;
  const int fin = (prf + pr0b + 1) >> 1;
  if (lane == 0) {
    int e = 0;
    codes[e++] = (u16)pr0; codes[e++] = (u16)pr; codes[e++] = (u16)pr1; codes[e++] = (u16)pr2; codes[e++] = (u16)pr3;
    if (S.st_type == FT_TEXT) codes[e++] = (u16)pr0b;
    codes[e++] = (u16)q1; codes[e++] = (u16)q2; codes[e++] = (u16)q3; codes[e++] = (u16)prf; codes[e++] = (u16)fin;
    S.pr = fin;
    S.last_prediction = fin;
  }
}

// weight-set cache: rows move between HBM and shared memory in 16-byte words, past L1
__device__ __forceinline__ void p8_row_load(short* dst, const short* src, int lane) {
  const uint4* s = reinterpret_cast<const uint4*>(src); uint4* d = reinterpret_cast<uint4*>(dst);
  uint4 v[7];
#pragma unroll
  for (int j = 0; j < 7; ++j) { const int q = lane + 32 * j; if (q < p8::N_IN / 8) v[j] = __ldcg(s + q); }
#pragma unroll
  for (int j = 0; j < 7; ++j) { const int q = lane + 32 * j; if (q < p8::N_IN / 8) d[q] = v[j]; }
}
__device__ __forceinline__ void p8_row_store(short* dst, const short* src, int lane) {
  const uint4* s = reinterpret_cast<const uint4*>(src); uint4* d = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int j = 0; j < 7; ++j) { const int q = lane + 32 * j; if (q < p8::N_IN / 8) __stcg(d + q, s[q]); }
}
static_assert(p8::N_IN % 8 == 0 && p8::N_IN / 8 <= 7 * 32, "a weight row is at most 7 16-byte words per lane");

// named barrier 2: the model warps signal "the state the 19 order-independent selector sets read is final" (arrive), the
// first SGD warp waits for it (sync) and computes those sets beside the apply phase
__device__ __forceinline__ void p8_signal_selects() { asm volatile("bar.arrive 2, 416;" ::: "memory"); }
__device__ __forceinline__ void p8_await_selects() { asm volatile("bar.sync 2, 416;" ::: "memory"); }
__device__ __forceinline__ void p8_sync_maps() { asm volatile("bar.sync 1, 384;" ::: "memory"); static_assert(P8_MAP_THREADS == 384, "named barrier width"); }

// Mixer::update for the 28 cached weight sets selected for the previous bit, by the four SGD warps (tid 0..127 of them)
__device__ __forceinline__ void p8_sgd(P8Shared& sh, int t) {
  using namespace p8;
  const int n8 = sh.sgd_nx >> 3, total = sh.sgd_ncxt * n8;
  if (n8 == 0) return;
  int i = t / n8, q = t - i * n8;
  for (int idx = t; idx < total; idx += P8_SGD_THREADS) {
    const int err = sh.sgd_err[i];
    if (err) {
      uint4* wp = reinterpret_cast<uint4*>(&sh.wc[i][q * 8]);
      uint4 wv = *wp;
      const uint4 xv = *reinterpret_cast<const uint4*>(&sh.tx_old[q * 8]);
      short* w = reinterpret_cast<short*>(&wv);
      const short* x = reinterpret_cast<const short*>(&xv);
#pragma unroll
      for (int e = 0; e < 8; ++e) w[e] = train_one(x[e], w[e], err);
      *wp = wv;
    }
    q += P8_SGD_THREADS;
    while (q >= n8) { q -= n8; ++i; }
  }
}

// One bit: PAQ8::Perceive(y). All P8_THREADS lanes call it: warps 0-11 evaluate the models, warps 12-15 train the mixer beside them.
__device__ void p8_bit(P8Shared& sh, int y, int nb, int tid, bool fresh = false) {   // nb: the bit position after this bit, (S.bpos + 1) & 7; fresh: shared memory holds nothing from the previous bit
  using namespace p8;
  State& S = sh.S;
  const p8::Tables& T = *S.T;
  const int warp = tid >> 5, lane = tid & 31;
#ifdef P8_PROF
  if (tid == 0) { sh.prof_t = clock64(); sh.prof_row = (S.bpos == 7) ? 0 : 1; }   // bpos before bit_begin: 7 -> this bit starts a byte
#endif
  // ---- phase 0: bookkeeping; the previous bit's inputs and errors move aside for the SGD warps
  if (tid == 0) {
    sh.sgd_nx = S.m.nx; sh.sgd_ncxt = S.m.ncxt;
    bit_begin(S, y);
    if (S.bpos == 0) block_parse(S);
    sh.snap_spaces = S.spaces; sh.snap_words = S.words; sh.snap_frstchar = S.frstchar; sh.snap_spafdo = S.spafdo;
  }
  if (tid >= 32 && tid < 32 + P8_N_CM) sh.clash[tid - 32] = 0;
  if (tid >= 64 && tid < 64 + P8_N_CM2) { sh.clash2[tid - 64] = 0; sh.res2[tid - 64] = 0; }
  if (tid == 67) { sh.any_clash = 0; sh.rnd_prev_i = S.rnd.i; }
  if (tid >= 320 && tid < 384) sh.rnd_prev[tid - 320] = S.rnd.table[tid - 320];
  if (tid >= 96 && tid < 96 + N_SETS) sh.sgd_err[tid - 96] = ((y << 12) - S.m.pr[tid - 96]) * 7;
  if (tid >= 128 && tid < 128 + N_IN / 8) reinterpret_cast<uint4*>(sh.tx_old)[tid - 128] = reinterpret_cast<const uint4*>(S.m.tx)[tid - 128];
  if (tid >= P8_CM2_TID0 && tid < P8_CM2_TID0 + P8_N_CM2) cm2_begin(p8_cm2(S, tid - P8_CM2_TID0), y, nb);
  if (warp == P8_WARPS - 1 && (nb <= 1 || fresh)) {      // mixer-input offsets of the units (they change on the first two bits of a byte only)
    const unsigned full = 0xffffffffu;
    const int a = p8_unit_count(S, lane, nb == 0);
    const int b = lane + 32 < P8_N_UNITS ? p8_unit_count(S, lane + 32, nb == 0) : 0;
    int ia = a, ib = b;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int va = __shfl_up_sync(full, ia, d), vb = __shfl_up_sync(full, ib, d);
      if (lane >= d) { ia += va; ib += vb; }
    }
    const int ta = __shfl_sync(full, ia, 31);
    sh.unit_off[lane] = ia - a;
    if (lane + 32 <= P8_N_UNITS) sh.unit_off[lane + 32] = ta + ib - b;
  }
  for (int k = tid; k < P8_SEEN; k += P8_THREADS) sh.u.seen[k] = 0ull;
  __syncthreads();
  P8_T(0);
  const int bpos = S.bpos, c0 = S.c0;
  const bool byte_start = bpos == 0;
  if (tid >= P8_MAP_THREADS) {
    p8_sgd(sh, tid - P8_MAP_THREADS);
    if (warp == P8_MAP_WARPS) {            // ModelStats and the 19 selector sets that do not wait for the order-N map
      p8_await_selects();
      if (lane == 0) {
        xml_stats(S);
        smatch_select(S);
        record_select(S);
        text_select(S);
        exe_select(S);
        if (S.m.ncxt != MAIN_SET_FIRST || S.m.base != MAIN_SET_BASE) S.error |= ERR_MIXER_ALIAS;
        S.m.ncxt = N_SETS;
      }
    }
  } else {
    if (tid == 0) { S.m.nx = S.m.base = S.m.ncxt = 0; }
    if (byte_start) {
      // ---- byte boundary, round 1: context computation, one model per lane / warp
      {
        P8_M0;
        if (warp < 3) p8_ols_byte_warp(sh, warp, lane);
        else if (warp == 8) {       // the text model: state on one lane (the stemmers of a completed word follow below)
          if (lane == 0) sh.text_pending = text_update_a(S);
        } else if (warp == 5) {     // the word model: state on one lane, its 57 contexts side by side
          if (lane == 0) word_update(S);
          __syncwarp();
          const int n = word_contexts(S, CtxSel{lane, 32});
          __syncwarp();
          if (lane == 0) { S.word.cm.cn = n; word_finish(S); }
        } else if (lane == 0) {
          switch (warp) {
            case 3: ordern_byte(S); break;
            case 4: distance_byte(S); record1_byte(S); break;
            case 6: nest_byte(S); indirect_byte(S); break;
            case 7: xml_byte(S); break;
            case 9: exe_byte(S); break;
            case 10: {
              const u8 W = (u8)buf(S, 1), WW = (u8)buf(S, 2), WWW = (u8)buf(S, 3);
              S.linear.prd[3] = (u8)clip8(W * 2 - WW);
              S.linear.prd[4] = (u8)clip8(W * 3 - WW * 3 + WWW);
            } break;
          }
        }
        if (lane == 0) P8_M(24 + warp);
      }
      p8_sync_maps();
      if (sh.text_pending) {      // a word ended: its three stemmers on three warps, then the rest of the text model's byte
        const int split = S.text.stem_split;
        if (lane == 0 && (warp == 8 || warp == 10 || warp == 11)) {
          const int i = warp == 8 ? LANG_EN : (warp == 10 ? LANG_FR : LANG_DE);
          if (i >= split) text_stem(S, i);
        }
        p8_sync_maps();
        if (tid == 8 * 32) {
          text_stem_mid(S);
          for (int i = split - 1; i > LANG_UNKNOWN; --i) text_stem(S, i);
          text_update_b(S);
        }
        p8_sync_maps();
      }
      if (warp == 8) {            // the text model's 33 contexts side by side
        const int n = text_contexts(S, CtxSel{lane, 32});
        __syncwarp();
        if (lane == 0) S.text.map.index = n;
      }
      P8_T(2);
      for (int k = tid; k < P8_SEEN; k += P8_MAP_THREADS) sh.u.seen[k] = 0ull;   // the OLS warps used this memory
      p8_sync_maps();
      // ---- the history maps and the single-lane units
      p8_probe_cm2(sh, tid, bpos);
      p8_probe_single(sh, tid, y, c0, bpos);
      p8_sync_maps();
      P8_T(3);
      p8_apply_cm2(sh, tid, y, bpos);
      p8_sync_maps();
      P8_T(4);
      // ---- round 2 of the byte boundary (needs the order-N result and the match model)
      {
        const int ismatch = ilog(T, S.match.length);
        if (tid == 32) sparse_byte(S, ismatch, sh.res2[0]);
        else if (tid == 64) {
          // sparseModel1 runs BEFORE wordModel in the reference: it sees the previous byte's word statistics
          const u32 a = S.spaces, b = S.words, c = S.frstchar, d = S.spafdo;
          S.spaces = sh.snap_spaces; S.words = sh.snap_words; S.frstchar = sh.snap_frstchar; S.spafdo = sh.snap_spafdo;
          sparse1_byte(S, ismatch, sh.res2[0]);
          S.spaces = a; S.words = b; S.frstchar = c; S.spafdo = d;
        } else if (tid == 96) { record_byte(S); record_pre(S); }
      }
      for (int k = tid; k < P8_SEEN; k += P8_MAP_THREADS) sh.u.seen[k] = 0ull;
      p8_sync_maps();
      p8_signal_selects();
      P8_T(5);
      if (warp < 7) p8_probe_cm(sh, tid, y, c0, bpos);
      p8_sync_maps();
      P8_T(6);
      if (sh.any_clash) { p8_number(sh, tid, y, c0, bpos); p8_sync_maps(); }
      P8_T(7);
      p8_apply_cm(sh, tid, y, c0, bpos);
      p8_apply_small(sh, tid, y, bpos);
      p8_sync_maps();
      P8_T(8);
    } else {
      // ---- inside a byte: probe / number / apply, all map families side by side
      {
        P8_M0;
        if (warp < 7) p8_probe_cm(sh, tid, y, c0, bpos);
        else { p8_probe_cm2(sh, tid, bpos); p8_probe_single(sh, tid, y, c0, bpos); }
        __syncwarp();
        if (lane == 0) P8_M(24 + warp);
      }
      p8_sync_maps();
      p8_signal_selects();
      P8_T(3);
      if (sh.any_clash) { p8_number(sh, tid, y, c0, bpos); p8_sync_maps(); }
      P8_T(7);
      {
        P8_M0;
        if (warp < 7) p8_apply_cm(sh, tid, y, c0, bpos);
        else { p8_apply_cm2(sh, tid, y, bpos); p8_apply_small(sh, tid, y, bpos); }
        __syncwarp();
        if (lane == 0) P8_M(48 + warp);
      }
      p8_sync_maps();
      P8_T(8);
    }
    // ---- epilogues, ModelStats, the 28 selector sets in the reference's order
    if (tid == 0) {
      if (bpos == 7) {
        for (int k = 0; k < P8_N_CM; ++k) if (!sh.clash[k]) p8_cm(S, k).cn = 0;
        for (int k = 0; k < P8_N_CM2; ++k) p8_cm2(S, k).index = 0;
      }
      main_select_fixed(S, sh.res2[0]);      // sets 19..27; the first SGD warp writes 0..18 (and the count) beside this
      Mixer& m = S.m;
      m.nx = sh.unit_off[P8_N_UNITS];
      m.n2 = m.nx;
      while (m.nx & 7) m.tx[m.nx++] = 0;
    }
  }
  __syncthreads();     // the SGD warps join
  P8_T(9);
  // ---- final-mixer SGD (32 weights); the 28 dot products over the cached sets (a selector that moved: write back, load)
  {
    Mixer& m = S.m;
    if (warp == P8_WARPS - 1) {
      const int err = ((y << 12) - m.pr2) * 7;
      if (err && lane < m.nx2) m.w2[lane] = train_one(m.tx2[lane], m.w2[lane], err);
    }
    if (warp == P8_WARPS - 2 && lane < m.ncxt) {        // two selectors on one weight set would need the reference's sequential SGD
      bool dup = false;
      for (int j = 0; j < lane; ++j) dup = dup || m.cxt[j] == m.cxt[lane];
      if (dup) S.error |= ERR_MIXER_ALIAS;
    }
    for (int i = warp; i < m.ncxt; i += P8_WARPS) {
      short* row = sh.wc[i];
      const int set = m.cxt[i], old = sh.wc_set[i];
      if (old != set) {
        if (old >= 0) p8_row_store(m.w + (size_t)old * N_IN, row, lane);
        p8_row_load(row, m.w + (size_t)set * N_IN, lane);
        __syncwarp();
        if (lane == 0) sh.wc_set[i] = set;
      }
      int acc = 0;
      const int n8 = m.nx >> 3;
      for (int q = lane; q < n8; q += 32) {
        const uint4 wv = *reinterpret_cast<const uint4*>(row + q * 8);
        const uint4 xv = *reinterpret_cast<const uint4*>(m.tx + q * 8);
        const short* w = reinterpret_cast<const short*>(&wv);
        const short* x = reinterpret_cast<const short*>(&xv);
#pragma unroll
        for (int e = 0; e < 8; e += 2) acc += dot_pair(x + e, w + e);
      }
      acc = __reduce_add_sync(0xffffffffu, acc);
      if (lane == 0) sh.dot[i] = acc;
    }
  }
  __syncthreads();
  P8_T(10);
  // ---- squash, final mixer, SSE stage (one warp)
  if (warp == 0) {
    Mixer& m = S.m;
    const int base = m.n2, n = m.ncxt, nx2 = (n + 7) & ~7;
    int x = 0;
    if (lane < n) {
      const int pr = squash(T, (int)((u32)sh.dot[lane] * 9u) >> 9);
      m.pr[lane] = pr;
      x = stretch(T, pr);
      S.codes[base + lane] = (u16)squash(T, x);
    }
    m.tx2[lane] = (short)x;
    __syncwarp();
    int z = 2 * lane < nx2 ? dot_pair(m.tx2 + 2 * lane, m.w2 + 2 * lane) : 0;
    z = __reduce_add_sync(0xffffffffu, z);
    const int pr2 = squash(T, z >> 9);
    if (lane == 0) { m.nx2 = nx2; m.pr2 = pr2; }
    __syncwarp();
    p8_sse_warp(sh, pr2, lane);
  }
  __syncthreads();
  P8_T(11);
}

// state block, hot tables and the pending weight sets into shared memory
__device__ __forceinline__ const p8::Tables* p8_enter(P8Shared& sh, p8::State* g, int tid) {
  p8_copy_words(&sh.S, g, sizeof(p8::State), tid);
  __syncthreads();
  const p8::Tables* gT = sh.S.T;
  p8_copy_words(sh.tab, gT, p8::TABLES_HOT_BYTES, tid);
  if (tid < p8::N_SETS) sh.wc_set[tid] = -1;
  __syncthreads();
  if (tid == 0) sh.S.T = reinterpret_cast<const p8::Tables*>(sh.tab);
  const int warp = tid >> 5, lane = tid & 31;
  for (int i = warp; i < sh.S.m.ncxt; i += P8_WARPS) {
    p8_row_load(sh.wc[i], sh.S.m.w + (size_t)sh.S.m.cxt[i] * p8::N_IN, lane);
    if (lane == 0) sh.wc_set[i] = sh.S.m.cxt[i];
  }
  __syncthreads();
  return gT;
}
__device__ __forceinline__ void p8_leave(P8Shared& sh, p8::State* g, const p8::Tables* gT, int tid) {
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31;
  for (int i = warp; i < p8::N_SETS; i += P8_WARPS)
    if (sh.wc_set[i] >= 0) p8_row_store(sh.S.m.w + (size_t)sh.wc_set[i] * p8::N_IN, sh.wc[i], lane);
  if (tid == 0) sh.S.T = gT;
  __syncthreads();
  p8_copy_words(g, &sh.S, sizeof(p8::State), tid);
}

// Bulk: CTA b serves stream b of the launch group: writes ext[t][431..2021] for every bit t of the sub-chunk.
__global__ void __launch_bounds__(P8_THREADS, 1) paq8_kernel(const ChunkArgs* __restrict__ args_all) {
  extern __shared__ __align__(16) unsigned char p8_raw[];
  P8Shared& sh = *reinterpret_cast<P8Shared*>(p8_raw);
  const ChunkArgs a = args_all[blockIdx.x];
  if (a.paq8 == nullptr) return;
  const int tid = threadIdx.x;
  p8::State* g = (p8::State*)a.paq8;
  const p8::Tables* gT = p8_enter(sh, g, tid);
  const u32 n_bits = a.n_bytes * 8;
  for (u32 t = 0; t < n_bits; ++t) {
    if (!a.pretrain) {
      u16* out = a.ext_gen + (size_t)t * N_EXT + 431;
      for (int k = tid; k < p8::N_OUT; k += P8_THREADS) out[k] = sh.S.codes[k];
    }
    const int y = (a.bytes[t >> 3] >> (7 - (t & 7))) & 1;
    p8_bit(sh, y, (int)((t + 1) & 7), tid);
  }
  p8_leave(sh, g, gT, tid);
}

// Lock-step: one bit per launch; the codes for the next Predict() land in ext_bit[431..2021].
__global__ void __launch_bounds__(P8_THREADS, 1) paq8_bit_kernel(p8::State* g, int y, u16* ext_bit, const u32* dbit) {
  if (dbit) y = (int)dbit[0];
  extern __shared__ __align__(16) unsigned char p8_raw[];
  P8Shared& sh = *reinterpret_cast<P8Shared*>(p8_raw);
  const int tid = threadIdx.x;
  const int nb = (g->bpos + 1) & 7;       // read from HBM: the shared copy is being updated by lane 0 inside p8_bit
  const p8::Tables* gT = p8_enter(sh, g, tid);
  p8_bit(sh, y, nb, tid, true);
  if (ext_bit) for (int k = tid; k < p8::N_OUT; k += P8_THREADS) ext_bit[431 + k] = sh.S.codes[k];
  p8_leave(sh, g, gT, tid);
}

}  // namespace cmixb200

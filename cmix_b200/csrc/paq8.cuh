// cmix_b200/csrc/paq8.cuh — the resident PAQ8 model on the device.
//
// One CTA of 12 warps per stream evaluates paq8_top.h's `bit()` with the work of a bit spread over lanes:
//  * every context of every context map is a lane (210 lanes for the sixteen 7-slot maps, 63 for the three history maps):
//    bucket probe, bit-history step, state maps and the 5 / 7 mixer inputs of a context are independent of the other
//    contexts of its map as long as they touch different 64-byte buckets this bit. That is CHECKED per bit (touched_buckets,
//    including the buckets a deferred history write-back will reach); a map with a clash is evaluated by one lane in the
//    reference's order instead.
//  * the one global coupling of the 7-slot maps, the shared pseudo-random sequence that ages high-count states
//    (paq8.cpp:1075), is resolved by lane 0 between the two passes: pass 1 computes each context's aged state, lane 0 hands
//    out the draws in program order (running clashing maps on the spot), pass 2 applies them.
//  * byte-boundary work (context hashing of the word / nest / indirect / XML / text / x86 / record models, three stemmers,
//    the OLS updates) runs model-per-lane in two rounds, because the sparse models and the record model consume what the
//    order-N map and the match model produce in the same bit;
//  * the 28 x 1552 int16 dot products and SGD steps use all 384 lanes (integer sums: exact under any association).
// The 55 KB state block lives in shared memory for the launch; tables (90 KB) and the ~10 GB of model memory stay in HBM.
// PAQ8 is a producer like FXCM: it depends on the coded bytes only and writes 1591 codes per bit into the `ext` scratch.
#pragma once
#include "paq8_top.h"
#include "state.h"

namespace cmixb200 {

enum { P8_SEEN = 4096 };
enum { P8_THREADS = 384, P8_WARPS = 12, P8_N_CM = 16, P8_N_CM2 = 3, P8_CM_LANES = 210, P8_CM2_LANES = 63, P8_N_UNITS = 53 };

struct P8Shared {
  p8::State S;
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
  unsigned long long seen[P8_SEEN];   // open-addressing set of (map, bucket) pairs touched this bit
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
__constant__ unsigned char c_p8_cm_full[P8_N_CM] = {42, 29, 3, 3, 3, 3, 16, 2, 5, 4, 3, 3, 57, 12, 15, 4};
__constant__ unsigned char c_p8_cm_unit[P8_N_CM] = {9, 10, 18, 20, 21, 22, 23, 36, 37, 38, 39, 40, 41, 42, 43, 45};
__constant__ unsigned char c_p8_cm2_cap[P8_N_CM2] = {10, 33, 20};
__constant__ unsigned char c_p8_cm2_unit[P8_N_CM2] = {3, 46, 47};

// unit u of the mixer-input order (paq8_top.h context_model): how many inputs it emits this bit
__device__ int p8_unit_count(p8::State& S, int u, bool byte_start) {
  if (u <= 2) return 1;
  if (u == 3) return 7 * (byte_start ? 10 : S.cm.index);
  if (u <= 6) return 1;
  if (u == 7) return 17;
  if (u == 8) return 11;
  if (u >= 11 && u <= 17) return 2;
  if (u == 19) return 3;
  if (u >= 24 && u <= 35) return 2;
  if (u == 44) return 6;
  if (u == 46) return 7 * (byte_start ? 33 : S.text.map.index);
  if (u == 47) return 7 * (byte_start ? 20 : S.exe.cm.index);
  if (u >= 48) return 2;
  for (int k = 0; k < P8_N_CM; ++k) if (c_p8_cm_unit[k] == u) return 5 * (byte_start ? c_p8_cm_full[k] : p8_cm(S, k).cn);
  return 0;
}

__device__ __forceinline__ p8::Out p8_out(P8Shared& sh, int offset) { p8::Out o; o.T = sh.S.T; o.tx = sh.S.m.tx; o.codes = sh.S.codes; o.n = offset; return o; }

__device__ __forceinline__ void p8_copy_words(void* dst, const void* src, size_t bytes, int tid) {
  u32* d = (u32*)dst; const u32* s = (const u32*)src;
  for (size_t i = tid; i < bytes / 4; i += P8_THREADS) d[i] = s[i];
}
static_assert(sizeof(p8::State) % 4 == 0, "state block is copied word by word");

// The three OLS updates of a byte (paq8_top.h ols_update) spread over one warp each: covariance rows in parallel, Cholesky
// column by column with every element summed in the reference's k order, substitutions on lane 0.
__device__ void p8_ols_update_warp(double* blk, int& km, u8 val, int lane) {
  using namespace p8;
  const double lambda = 0.995, nu = 0.001, one_minus = 1.0 - 0.995;
  double* x = blk; double* w = blk + 32; double* b = blk + 64; double* cov = blk + 96; double* ch = blk + 96 + 1024;
  const int j = lane;   // lane = row
  for (int i = 0; i < 32; ++i) cov[j * 32 + i] = P8_DADD(P8_DMUL(lambda, cov[j * 32 + i]), P8_DMUL(one_minus, P8_DMUL(x[j], x[i])));
  b[j] = P8_DADD(P8_DMUL(lambda, b[j]), P8_DMUL(one_minus, P8_DMUL(x[j], (double)val)));
  __syncwarp();
  int k_new = km + 1;
  if (k_new >= 4) {
    for (int i = 0; i < 32; ++i) ch[j * 32 + i] = cov[j * 32 + i];
    ch[j * 32 + j] = P8_DADD(ch[j * 32 + j], nu);
    __syncwarp();
    bool fail = false;
    for (int c = 0; c < 32; ++c) {          // column c: the diagonal first, then every row below it
      if (lane == c) {
        double sum = ch[c * 32 + c];
        for (int k = 0; k < c; ++k) sum = P8_DSUB(sum, P8_DMUL(ch[c * 32 + k], ch[c * 32 + k]));
        if (sum > 1E-8) ch[c * 32 + c] = P8_DSQRT(sum); else ch[c * 32 + c] = -1.0;   // -1 marks Factor()'s failure exit
      }
      __syncwarp();
      if (ch[c * 32 + c] < 0.0) { fail = true; break; }
      if (lane > c) {
        double sum = ch[lane * 32 + c];
        for (int k = 0; k < c; ++k) sum = P8_DSUB(sum, P8_DMUL(ch[lane * 32 + k], ch[c * 32 + k]));
        ch[lane * 32 + c] = P8_DDIV(sum, ch[c * 32 + c]);
      }
      __syncwarp();
    }
    if (!fail && lane == 0) {
      for (int i = 0; i < 32; ++i) {
        double sum = b[i];
        for (int q = 0; q < i; ++q) sum = P8_DSUB(sum, P8_DMUL(ch[i * 32 + q], w[q]));
        w[i] = P8_DDIV(sum, ch[i * 32 + i]);
      }
      for (int i = 31; i >= 0; --i) {
        double sum = w[i];
        for (int q = i + 1; q < 32; ++q) sum = P8_DSUB(sum, P8_DMUL(ch[q * 32 + i], w[q]));
        w[i] = P8_DDIV(sum, ch[i * 32 + i]);
      }
    }
    k_new = 0;
    __syncwarp();
  }
  if (lane == 0) km = k_new;
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

// map (k, i) of a 7-slot-map lane, -1 if the lane is beyond the map's capacity
__device__ __forceinline__ bool p8_cm_lane(int lane, int& k, int& i) {
  int base = 0;
  for (k = 0; k < P8_N_CM; ++k) { const int cap = c_p8_cm_cap[k]; if (lane < base + cap) { i = lane - base; return true; } base += cap; }
  return false;
}
__device__ __forceinline__ bool p8_cm2_lane(int lane, int& k, int& i) {
  int base = 0;
  for (k = 0; k < P8_N_CM2; ++k) { const int cap = c_p8_cm2_cap[k]; if (lane < base + cap) { i = lane - base; return true; } base += cap; }
  return false;
}

// One bit: PAQ8::Perceive(y). All P8_THREADS lanes call it.
__device__ void p8_bit(P8Shared& sh, int y, int tid) {
  using namespace p8;
  State& S = sh.S;
  const p8::Tables& T = *S.T;
  const int warp = tid >> 5, lane = tid & 31;
  // ---- phase 0: bookkeeping
  if (tid == 0) {
    bit_begin(S, y);
    if (S.bpos == 0) block_parse(S);
    sh.snap_spaces = S.spaces; sh.snap_words = S.words; sh.snap_frstchar = S.frstchar; sh.snap_spafdo = S.spafdo;
    for (int k = 0; k < P8_N_CM; ++k) sh.clash[k] = 0;
    for (int k = 0; k < P8_N_CM2; ++k) { sh.clash2[k] = 0; sh.res2[k] = 0; }
  }
  for (int k = tid; k < P8_SEEN; k += P8_THREADS) sh.seen[k] = 0ull;
  __syncthreads();
  const int bpos = S.bpos, c0 = S.c0;
  const bool byte_start = bpos == 0;
  // ---- phase 1: SGD on the 28 weight sets selected for the previous bit (Mixer::update)
  {
    Mixer& m = S.m;
    for (int i = 0; i < m.ncxt; ++i) {
      const int err = ((y << 12) - m.pr[i]) * 7;
      if (!err) continue;
      short* w = m.w + (size_t)m.cxt[i] * N_IN;
      for (int k = tid; k < m.nx; k += P8_THREADS) w[k] = train_one(m.tx[k], w[k], err);
    }
  }
  __syncthreads();
  if (tid == 0) { S.m.nx = S.m.base = S.m.ncxt = 0; }
  // ---- phase 2 (byte boundary): round-1 context computation, one model per lane / warp
  if (byte_start) {
    if (warp < 3) {            // the three OLS predictors, one warp each
      p8_ols_update_warp(S.linear.ols + (size_t)warp * OLS_STRIDE, S.linear.ols_km[warp], (u8)buf(S, 1), lane);
    } else if (lane == 0) {
      switch (warp) {
        case 3: ordern_byte(S); break;
        case 4: distance_byte(S); record1_byte(S); break;
        case 5: word_byte(S); break;
        case 6: nest_byte(S); indirect_byte(S); break;
        case 7: xml_byte(S); break;
        case 8: text_update(S); text_contexts(S); break;
        case 9: exe_byte(S); break;
      }
    }
  }
  __syncthreads();
  if (tid == 0) {
    if (byte_start) linear_predict(S);
    for (int k = 0; k < P8_N_CM2; ++k) cm2_begin(p8_cm2(S, k), y, bpos);
    int n = 0;
    for (int u = 0; u < P8_N_UNITS; ++u) { sh.unit_off[u] = n; n += p8_unit_count(S, u, byte_start); }
    sh.unit_off[P8_N_UNITS] = n;
  }
  __syncthreads();
  // ---- phase 3: buckets the history-map contexts touch; the match models and the small direct units
  if (tid < P8_CM2_LANES) {
    int k, i;
    p8_cm2_lane(tid, k, i);
    Cm2& m = p8_cm2(S, k);
    const int n = i < m.index ? cm2_touched(m, i, bpos, sh.ids2[tid]) : 0;
    if (n && p8_claim(sh.seen, k, sh.ids2[tid], n)) sh.clash2[k] = 1;
  } else if (tid == 64) { Out o = p8_out(sh, sh.unit_off[7]); match_bit(S, o); }
  else if (tid == 96) { Out o = p8_out(sh, sh.unit_off[8]); smatch_core(S, o); }
  else if (tid == 128) {
    Out o = p8_out(sh, sh.unit_off[0]);
    add(o, 64);
    add(o, (stretch(T, sm32_p(T, S.sm0, y, c0)) + 1) >> 1);
    add(o, (stretch(T, sm32_p(T, S.sm1, y, c0 | (buf(S, 1) << 8))) + 1) >> 1);
  } else if (tid >= 129 && tid <= 131) {
    Out o = p8_out(sh, sh.unit_off[4 + (tid - 129)]);
    rcm_mix(tid == 129 ? S.rcm7 : tid == 130 ? S.rcm9 : S.rcm10, o, c0, bpos);
  } else if (tid == 160) { Out o = p8_out(sh, sh.unit_off[19]); pic_bit(S, o); }
  else if (tid >= 192 && tid < 202) sh.dmc_st[tid - 192] = dmc_st(T, S.dmc[tid - 192], y);
  else if (tid >= 224 && tid < 229) { Out o = p8_out(sh, sh.unit_off[48 + (tid - 224)]); linear_small(S, o, tid - 224); }
  __syncthreads();
  // ---- phase 4: DMC forest combination and reset (dmcForest::mix)
  if (tid == 192) {
    Out o = p8_out(sh, sh.unit_off[44]);
    const u32 params[10] = {2, 32, 64, 4, 128, 8, 256, 16, 1024, 1536};
    add(o, sh.dmc_st[9] >> 3);
    add(o, sh.dmc_st[8] >> 3);
    for (int i = 7; i > 0; i -= 2) add(o, (sh.dmc_st[i] + sh.dmc_st[i - 1]) >> 4);
    if (byte_start)
      for (int i = 7; i >= 0; --i)
        if ((S.dmc[i].extra >> 7) > S.dmc[i].size) dmc_reset(S.dmc[i], params[i]);
  }
  __syncthreads();
  // ---- phase 5: the history maps
  if (tid < P8_CM2_LANES) {
    int k, i;
    p8_cm2_lane(tid, k, i);
    Cm2& m = p8_cm2(S, k);
    const int off = sh.unit_off[c_p8_cm2_unit[k]];
    if (!sh.clash2[k]) {
      if (i < m.index) { Out o = p8_out(sh, off + 7 * i); if (cm2_step(m, i, o, y, bpos)) atomicAdd(&sh.res2[k], 1); }
    } else if (i == 0) {      // in-order evaluation by one lane (the two loops of ContextMap2::mix)
      Out o = p8_out(sh, off);
      sh.res2[k] = cm2_mix_body(m, o, y, bpos);
    }
  }
  __syncthreads();
  // ---- phase 6 (byte boundary): round-2 context computation (needs the order-N result and the match model)
  if (tid == 0) sh.order = sh.res2[0];
  if (byte_start) {
    const int ismatch = ilog(T, S.match.length);
    if (tid == 32) sparse_byte(S, ismatch, sh.res2[0]);
    else if (tid == 64) {
      // sparseModel1 runs BEFORE wordModel in the reference: it sees the previous byte's word statistics
      const u32 a = S.spaces, b = S.words, c = S.frstchar, d = S.spafdo;
      S.spaces = sh.snap_spaces; S.words = sh.snap_words; S.frstchar = sh.snap_frstchar; S.spafdo = sh.snap_spafdo;
      sparse1_byte(S, ismatch, sh.res2[0]);
      S.spaces = a; S.words = b; S.frstchar = c; S.spafdo = d;
    } else if (tid == 96) record_byte(S);
  }
  for (int k = tid; k < P8_SEEN; k += P8_THREADS) sh.seen[k] = 0ull;
  __syncthreads();
  if (tid == 0) record_pre(S);
  // ---- phase 7: pass 1 of the 7-slot maps: aged state and touched buckets per context
  if (tid < P8_CM_LANES) {
    int k, i;
    p8_cm_lane(tid, k, i);
    Cm& m = p8_cm(S, k);
    if (i < m.cn) {
      sh.ns[tid] = (short)cm_next_state(T, m, i, y);
      const int n = cm_touched(m, i, c0, bpos, sh.ids[tid]);
      if (p8_claim(sh.seen, k, sh.ids[tid], n)) sh.clash[k] = 1;
    } else sh.ns[tid] = -1;
  }
  __syncthreads();
  // ---- phase 8: lane 0 hands out the random draws in program order; clashing maps are evaluated here, in order
  if (tid == 0) {
    int base = 0;
    const int c1 = buf(S, 1);
    for (int k = 0; k < P8_N_CM; ++k) {
      Cm& m = p8_cm(S, k);
      if (sh.clash[k]) { Out o = p8_out(sh, sh.unit_off[c_p8_cm_unit[k]]); cm_mix(m, o, S.rnd, y, c0, bpos, c1); }
      else
        for (int i = 0; i < m.cn; ++i) {
          const int ns = sh.ns[base + i];
          if (ns >= 204 && cm_draw_hits(rnd_next(S.rnd), ns)) sh.ns[base + i] = (short)(ns - 4);
        }
      base += c_p8_cm_cap[k];
    }
  }
  __syncthreads();
  // ---- phase 9: pass 2 of the 7-slot maps
  if (tid < P8_CM_LANES) {
    int k, i;
    p8_cm_lane(tid, k, i);
    Cm& m = p8_cm(S, k);
    if (!sh.clash[k] && i < m.cn) { Out o = p8_out(sh, sh.unit_off[c_p8_cm_unit[k]] + 5 * i); cm_step(m, i, o, sh.ns[tid], y, c0, bpos, buf(S, 1)); }
  } else if (tid >= 256 && tid < 268) {   // the record model's 12 direct maps (contexts selected by record_pre above)
    Out o = p8_out(sh, sh.unit_off[24 + (tid - 256)]);
    record_small(S, o, tid - 256);
  } else if (tid >= 288 && tid < 295) {   // sparseModel1's seven stationary maps
    Out o = p8_out(sh, sh.unit_off[11 + (tid - 288)]);
    scm_mix(S.sparse1.scm[tid - 288], o, y);
  }
  __syncthreads();
  // ---- phase 10: epilogues, ModelStats, the 28 selector sets in the reference's order
  if (tid == 0) {
    if (bpos == 7) {
      for (int k = 0; k < P8_N_CM; ++k) if (!sh.clash[k]) p8_cm(S, k).cn = 0;
      for (int k = 0; k < P8_N_CM2; ++k) p8_cm2(S, k).index = 0;
    }
    xml_stats(S);
    S.m.nx = sh.unit_off[P8_N_UNITS];
    smatch_select(S);
    record_select(S);
    text_select(S);
    exe_select(S);
    main_select(S, sh.order);
    Mixer& m = S.m;
    m.n2 = m.nx;
    while (m.nx & 7) m.tx[m.nx++] = 0;
  }
  __syncthreads();
  // ---- phase 11: final-mixer SGD (32 weights) and the 28 dot products
  {
    Mixer& m = S.m;
    if (warp == P8_WARPS - 1) {
      const int err = ((y << 12) - m.pr2) * 7;
      if (err && lane < m.nx2) m.w2[lane] = train_one(m.tx2[lane], m.w2[lane], err);
    }
    for (int i = warp; i < m.ncxt; i += P8_WARPS) {
      const short* w = m.w + (size_t)m.cxt[i] * N_IN;
      int acc = 0;
      for (int k = 2 * lane; k < m.nx; k += 64) acc += dot_pair(m.tx + k, w + k);
      acc = __reduce_add_sync(0xffffffffu, acc);
      if (lane == 0) sh.dot[i] = acc;
    }
  }
  __syncthreads();
  // ---- phase 12: squash, final mixer, SSE stage
  if (tid == 0) {
    Mixer& m = S.m;
    const int base = m.n2;
    m.nx2 = 0;
    for (int i = 0; i < m.ncxt; ++i) {
      m.pr[i] = squash(T, (int)((u32)sh.dot[i] * 9u) >> 9);
      const int x = stretch(T, m.pr[i]);
      S.codes[base + i] = (u16)squash(T, x);
      m.tx2[m.nx2++] = (short)x;
    }
    while (m.nx2 & 7) m.tx2[m.nx2++] = 0;
    int z = 0;
    for (int k = 0; k < m.nx2; k += 2) z += dot_pair(m.tx2 + k, m.w2 + k);
    m.pr2 = squash(T, z >> 9);
    sse_stage(S, m.pr2);
  }
  __syncthreads();
}

// Bulk: CTA b serves stream b of the launch group: writes ext[t][431..2021] for every bit t of the sub-chunk.
__global__ void __launch_bounds__(P8_THREADS, 1) paq8_kernel(const ChunkArgs* __restrict__ args_all) {
  extern __shared__ __align__(16) unsigned char p8_raw[];
  P8Shared& sh = *reinterpret_cast<P8Shared*>(p8_raw);
  const ChunkArgs a = args_all[blockIdx.x];
  if (a.paq8 == nullptr) return;
  const int tid = threadIdx.x;
  p8::State* g = (p8::State*)a.paq8;
  p8_copy_words(&sh.S, g, sizeof(p8::State), tid);
  __syncthreads();
  const u32 n_bits = a.n_bytes * 8;
  for (u32 t = 0; t < n_bits; ++t) {
    if (!a.pretrain) {
      u16* out = a.ext_gen + (size_t)t * N_EXT + 431;
      for (int k = tid; k < p8::N_OUT; k += P8_THREADS) out[k] = sh.S.codes[k];
    }
    const int y = (a.bytes[t >> 3] >> (7 - (t & 7))) & 1;
    p8_bit(sh, y, tid);
  }
  __syncthreads();
  p8_copy_words(g, &sh.S, sizeof(p8::State), tid);
}

// Lock-step: one bit per launch; the codes for the next Predict() land in ext_bit[431..2021].
__global__ void __launch_bounds__(P8_THREADS, 1) paq8_bit_kernel(p8::State* g, int y, u16* ext_bit) {
  extern __shared__ __align__(16) unsigned char p8_raw[];
  P8Shared& sh = *reinterpret_cast<P8Shared*>(p8_raw);
  const int tid = threadIdx.x;
  p8_copy_words(&sh.S, g, sizeof(p8::State), tid);
  __syncthreads();
  p8_bit(sh, y, tid);
  if (ext_bit) for (int k = tid; k < p8::N_OUT; k += P8_THREADS) ext_bit[431 + k] = sh.S.codes[k];
  __syncthreads();
  p8_copy_words(g, &sh.S, sizeof(p8::State), tid);
}

}  // namespace cmixb200

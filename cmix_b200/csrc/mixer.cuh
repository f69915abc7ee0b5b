// cmix_b200/csrc/mixer.cuh
//
// Kernel "mix": the three-layer gated logistic mixer + the integer SSE stage
// (reference src/mixer/mixer.cpp:16-72, src/mixer/mixer-input.cpp,
// src/predictor.cpp:388-418,432-437, src/mixer/sse.cpp).
// This is the roofline-defining kernel: per coded bit it touches
// 26*(2078+i) + 20*(29+i) + 49 = 55 172 fp32 weights twice (dot, then SGD).
//
// Layout: one thread-block CLUSTER of 2 CTAs per stream (persistent over the
// whole chunk). CTA r keeps the currently selected weight rows of layer-0
// mixers [13r, 13r+13) resident in shared memory (13 x 8.4 KB): a row is read
// from HBM only when its selector context changes (16 of the 26 selectors
// change once per byte, not per bit) and written back only when it is evicted,
// so steady-state HBM traffic is well under the algorithmic 450 KB/bit.
//
// Parity contract: the reference sums each dot product sequentially in fp32
// (mixer.cpp:41-43, no FMA). A tree/warp-shuffle reduction changes the rounding
// and, through the 15-bit SSE quantisation (sse.cpp:321), moves the coded
// probability by >1e-5 on a fraction of bits. So each dot product is ONE serial
// FADD chain here, and the parallelism is ACROSS the 13 chains of a CTA (13
// lanes of one warp, conflict-free row pitch in shared memory) and across the
// 480 other threads that stage inputs, move rows and apply the SGD update.
#pragma once
#include <cooperative_groups.h>

#include "exact_math.h"
#include "lstm.cuh"
#include "small_models.cuh"
#include "state.h"

namespace cmixb200 {
namespace cg = cooperative_groups;

enum {
  MIX_THREADS = 512,
  MIX_PER_CTA = 13,
  ROW_PITCH_S = 2108,     // shared-memory row pitch: 527 float4 (odd) -> LDS.128 conflict-free across lanes
};

struct MixShared {
  float rows[MIX_PER_CTA][ROW_PITCH_S];
  float x[N_INPUTS + 26 + 8];     // staged inputs followed by the 26 layer-0 "extra inputs"
  float mains[N_L0 + 6];          // main dot products of all 26 layer-0 mixers (CTA0 collects)
  float we[N_L0][N_L0 + 2];       // extra-input weights of all 26 layer-0 mixers (CTA0 collects)
  float upd[N_L0 + 6];            // SGD coefficient `update` per layer-0 mixer
  u32 cur_slot[MIX_PER_CTA + 3];  // resident row per local mixer (0xffffffff = none)
  u32 want_slot[MIX_PER_CTA + 3];
  u32 shrink[MIX_PER_CTA + 3];
  u32 sel[SEL_PITCH];
  float in1[L1_IN + 3], in2[L2_IN + 3];
  float l1row[N_L1][ROW_PITCH_L1];
  float l2row[ROW_PITCH_L2];
  float l1extra[N_L1 + 4];
  float upd1[N_L1 + 4];
  u32 slot1[N_L1 + 4];
  u32 shrink1[N_L1 + 4];
  float mixp[N_MIXERS + 1];
  int bit;
};

// Mixer::GetContextData (mixer.cpp:16-36): first-come row assignment, overflow row after 10 000.
__device__ __forceinline__ u32 resolve_slot(MixerState& m, u32 ctx) {
  u32 s = m.slot_table[ctx];
  if (s == 0) {
    const u32 cap = m.n_rows - 1;                      // = min(table_size, 10000)
    if (m.n_assigned < cap && m.n_assigned < (u32)SLOT_LIMIT) { s = ++m.n_assigned; m.slot_table[ctx] = s; }
    else s = m.n_rows;                                 // shared overflow row (key 0xDEADBEEF)
  }
  return s - 1;
}

// `update` of Mixer::Perceive (mixer.cpp:58-66); also advances the step counters.
__device__ __forceinline__ float mixer_update_coeff(MixerState& m, u32 slot, float decay_base, float p, int bit, u32* shrink) {
  const u64 rs = m.row_steps[slot];
  float decay = decay_base;
  decay = (float)((double)decay * (1.5 - ((1.0 * (double)rs) / (double)m.max_steps)));
  const float update = XM_FMUL(XM_FMUL(decay, m.lr), XM_FSUB(xm_logistic(p), (float)bit));
  const u64 ns = rs + 1;
  m.row_steps[slot] = ns;
  if (ns > m.max_steps) m.max_steps = ns;
  *shrink = ((ns & 1023) == 0) ? 1u : 0u;
  return update;
}

__device__ __forceinline__ float clamp_stretched(const Tables& T, float p) {    // mixer-input.cpp:17-27
  if (p > T.stretch_max) p = T.stretch_max; else if (p < T.stretch_min) p = T.stretch_min;
  return p;
}

// ------------------------------------------------------------------ SSE ----
__device__ __forceinline__ int sse_extrap(int p1, int C) {
  p1 = (((p1 - 16384) * C) >> 13) + 16384;
  if (p1 < 1) p1 = 1;
  if (p1 > 32767) p1 = 32767;
  return p1;
}
__device__ __forceinline__ int sse_rdiv(int x, int a, int d) { return x >= 0 ? (x + a) >> d : -((-x + a) >> d); }
__device__ __forceinline__ int sse_mixup(int w, int s1, int s0) {
  int x = s1 + sse_rdiv((w - 16384) * (s0 - s1), 1 << 14, 15);
  return (x > 0) ? (x < 32768) ? x : 32767 : 1;
}
__device__ __forceinline__ int sse_mask1(int j) {   // M_mx1mask0 (sse.cpp:190)
  if (j < 2) return 0;
  if (j <= 32) return j - 1;
  if (j <= 63) return 31 + (j - 32) / 2;
  if (j <= 127) return 47 + (j - 64) / 4;
  return 63 + (j - 128) / 8;
}
__device__ __forceinline__ int sse_pred(const u16* bucket, int iP, int* sw, int* q, int* P) {
  *q = (6 * iP) >> 15;
  *sw = (6 * iP) & 32767;
  int f = (((32768 - *sw) * (int)bucket[*q] + *sw * (int)bucket[*q + 1]) >> 15) - 8192;
  if (f <= 0) f = 1;
  if (f >= 32768) f = 32767;
  *P = f;
  return f;
}
__device__ __forceinline__ void sse_bucket_update(u16* bucket, int c, int wr0, int sw, int q, int P) {
  P = P * (32768 - wr0) >> 15;
  if (c == 0) P += wr0;
  const int dC = (int)bucket[q] - (int)bucket[q + 1];
  const int sw_dC = (sw * dC + 32767) >> 15;
  bucket[q] = (u16)(P + sw_dC + 8192);
  bucket[q + 1] = (u16)(P - (dC - sw_dC) + 8192);
}
__device__ __forceinline__ void sse_mix_update(int* w, int y, int p0, int p1, int wq, int pm) {
  const int py = 32768 - (y << 15);
  const int e = py - pm;
  int d = sse_rdiv(e * (p0 - p1), 1 << 14, 15);
  d = sse_rdiv(d * wq, 1 << 14, 15);
  *w += d;
}
// SSE::Predict (sse.cpp:320-324 -> M_Estimate :243-289). Single thread.
__device__ float sse_predict(SseState& S, float input) {
  const int discrete = (int)XM_FADD(1.0f, XM_FMUL(XM_FSUB(1.0f, input), 32766.0f));
  const u32 p = (u32)discrete;
  const u32 j = S.j, pc = S.pc, ffl = S.ffl, prq = p >> 11;
  const u32 q3 = (prq > 0) + (prq > 14);
  const u32 q4 = (prq > 0) + (prq > 7) + (prq > 14);
  S.sm7x = ((((q3 << 5) + (ffl & 31)) << 8) + (pc & 255)) * 255 + (j < 2 ? 0 : j - 1);
  S.mix2 = ((((q3 << 1) + (ffl & 1)) << 8) + (pc & 255)) * 256 + j;
  S.sm6x = ((((q3 << 7) + (ffl & 127)) << 8) + (pc & 255)) * 256 + j;
  S.mix1 = ((((q4 << 8) + (ffl & 255)) << 3) + ((pc >> 5) & 7)) * 79 + sse_mask1((int)j);
  const u16* st = S.st; const u16* sq = S.sq;
  const int stp = st[p];
  const int p1 = sse_pred(S.s6 + (size_t)S.sm6x * 8, sq[sse_extrap(stp, 10240)], &S.sw6, &S.q6, &S.P6);
  const int s0 = sse_extrap(stp, 7935);
  const int s1 = sse_extrap(st[p1], 9592);
  S.mix1_s0 = s0; S.mix1_s1 = s1;
  int s2 = sse_mixup(S.x1[S.mix1], s0, s1);
  s2 = sse_extrap(s2, 8092);
  S.mix1_p = sq[s2];
  const int p2 = sse_pred(S.s7 + (size_t)S.sm7x * 8, sq[sse_extrap(stp, 8200)], &S.sw7, &S.q7, &S.P7);
  const int s4 = sse_extrap(st[p2], 7677);
  S.mix2_s0 = s2; S.mix2_s1 = s4;
  int s5 = sse_mixup(S.x2[S.mix2], s2, s4);
  s5 = sse_extrap(s5, 8202);
  S.mix2_p = sq[s5];
  const int estimate = S.mix2_p;
  return (float)(1.0 - ((double)(estimate - 1) / 32766.0));
}
// SSE::Perceive (M_Update, sse.cpp:291-305). Single thread.
__device__ void sse_perceive(SseState& S, int bit) {
  sse_bucket_update(S.s6 + (size_t)S.sm6x * 8, bit, 106, S.sw6, S.q6, S.P6);
  sse_mix_update(&S.x1[S.mix1], bit, S.mix1_s0, S.mix1_s1, 6202, S.mix1_p);
  sse_bucket_update(S.s7 + (size_t)S.sm7x * 8, bit, 127, S.sw7, S.q7, S.P7);
  sse_mix_update(&S.x2[S.mix2], bit, S.mix2_s0, S.mix2_s1, 8320, S.mix2_p);
  S.j += S.j + bit;
  if (S.j >= 256) {
    S.ffl = (u8)(S.ffl * 2 + (S.pc >= 0x40));
    S.pc = (u8)S.j;
    S.j = 1;
  }
}

// --------------------------------------------------------- layer-0 chains --
// 13 independent serial FADD chains, one per lane; weights in shared memory.
__device__ __forceinline__ float chain_l0(const float* __restrict__ x, const float* __restrict__ row) {
  float p = 0.0f;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  const float4* w4 = reinterpret_cast<const float4*>(row);
#pragma unroll 8
  for (int k = 0; k < N_INPUTS / 4; ++k) {            // 519 * 4 = 2076
    const float4 a = x4[k], b = w4[k];
    p = XM_FADD(p, XM_FMUL(a.x, b.x));
    p = XM_FADD(p, XM_FMUL(a.y, b.y));
    p = XM_FADD(p, XM_FMUL(a.z, b.z));
    p = XM_FADD(p, XM_FMUL(a.w, b.w));
  }
  p = XM_FADD(p, XM_FMUL(x[2076], row[2076]));
  p = XM_FADD(p, XM_FMUL(x[2077], row[2077]));
  return p;
}

// Stage the 2078 layer-0 inputs of bit t into shared memory (predictor.cpp:362-387 order).
__device__ __forceinline__ void stage_inputs(float* x, const Tables& T, const u16* ext, const float* small_x,
                                             float lstm_x, int tid, int nthreads) {
  for (int k = tid; k < N_INPUTS; k += nthreads) {
    float v;
    if (k < 3) v = small_x[k];
    else if (k < 3 + N_EXT) {
      const u32 code = ext ? ext[k - 3] : 0xFFFFu;
      v = T.lut12[code == 0xFFFFu ? 4096 : (code > 4095u ? 4095u : code)];
    }
    else if (k < 2076) v = small_x[k - N_EXT];
    else if (k == 2076) v = small_x[N_SMALL];
    else v = lstm_x;
    x[k] = v;
  }
}

// auxiliary_context_ (predictor.cpp:388-393)
__device__ __forceinline__ u32 aux_context(const float* x) {
  float avg = 0.0f;
  avg = XM_FADD(avg, xm_logistic(x[433]));
  avg = XM_FADD(avg, xm_logistic(x[2024]));
  avg = XM_FADD(avg, xm_logistic(x[2077]));
  avg = XM_FDIV(avg, 3.0f);
  return (u32)(unsigned long long)XM_FMUL(avg, 15.0f);
}

// Layers 1 and 2 + SSE for one bit, run by ONE warp. Inputs: sh.mains (26 layer-0 main sums),
// sh.we (their extra weights), sh.l1row/sh.l2row (selected rows). Outputs: sh.x[2078..] extras,
// sh.mixp, sh.in1, sh.in2, sh.l1extra; returns the final probability in lane 0.
__device__ float final_stage(MixShared& sh, const Tables& T, SseState& sse, int lane) {
  // ---- layer 0: forward substitution through the "extra inputs" (mixer.cpp:45-53) ----
  float main = lane < N_L0 ? sh.mains[lane] : 0.0f;
  float e = 0.0f, pfin = 0.0f;
  for (int k = 0; k < N_L0; ++k) {
    if (lane == k) pfin = XM_FADD(main, e);
    const float pk = __shfl_sync(0xffffffffu, pfin, k);
    const float ck = clamp_stretched(T, pk);
    if (lane == k) { sh.mixp[k] = pk; sh.x[N_INPUTS + k] = ck; sh.in1[k] = ck; sh.in2[k] = ck; }
    if (lane > k && lane < N_L0) e = XM_FADD(e, XM_FMUL(ck, sh.we[lane][k]));
  }
  if (lane < N_AUX) {
    const int idx = lane == 0 ? 433 : (lane == 1 ? 2024 : 2077);
    const float c = clamp_stretched(T, sh.x[idx]);
    sh.in1[N_L0 + lane] = c;
    sh.in2[N_L0 + N_L1 + lane] = c;
  }
  __syncwarp();
  // ---- layer 1 ----
  main = 0.0f;
  if (lane < N_L1) {
    const float* w = sh.l1row[lane];
#pragma unroll
    for (int k = 0; k < L1_IN; ++k) main = XM_FADD(main, XM_FMUL(sh.in1[k], w[k]));
  }
  e = 0.0f; pfin = 0.0f;
  for (int k = 0; k < N_L1; ++k) {
    if (lane == k) pfin = XM_FADD(main, e);
    const float pk = __shfl_sync(0xffffffffu, pfin, k);
    const float ck = clamp_stretched(T, pk);
    if (lane == k) { sh.mixp[N_L0 + k] = pk; sh.l1extra[k] = ck; sh.in2[N_L0 + k] = ck; }
    if (lane > k && lane < N_L1) e = XM_FADD(e, XM_FMUL(ck, sh.l1row[lane][L1_IN + k]));
  }
  __syncwarp();
  // ---- layer 2 + squash + SSE ----
  float p = 0.0f;
  if (lane == 0) {
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < L2_IN; ++k) s = XM_FADD(s, XM_FMUL(sh.in2[k], sh.l2row[k]));
    s = XM_FADD(s, 0.0f);                                // p_ += e with no extra inputs
    sh.mixp[N_L0 + N_L1] = s;
    p = sse_predict(sse, xm_logistic(s));
  }
  return p;
}

// (The first bulk kernel, a barrier-per-phase cluster kernel, lived here through round 1; mixer_v3.cuh replaced it. History: DESIGN.md §4.1.)

}  // namespace cmixb200

// ---------------------------------------------------------------------------
// Lock-step halves (host calls Predictor::Predict / Perceive bit by bit, e.g. the
// reference's Decoder, coder/decoder.cpp:20-39). Single CTA, rows stay in HBM.
// Same arithmetic as the bulk kernel (mixer_v3.cuh); intermediate vectors are parked in StreamState.
namespace cmixb200 {

// ---------------------------------------------------------------------------------------------
// Lock-step halves (Predict() / Perceive(bit) one bit at a time: the decoder's order). Nothing can
// stay resident between the two calls, so the work is spread over SMs instead: one CTA per layer-0
// mixer for the 26 serial chains and for the 26 row updates, one CTA for layers 1-2 + SSE.
//
// mix_predict_rows_kernel<<<26, 256>>>: CTA i stages the 2078 inputs, resolves mixer i's row
// (mixer 12's selector is the auxiliary context of the staged inputs, predictor.cpp:388-393),
// copies the row into shared memory and runs its chain on one thread (Mixer::Mix, mixer.cpp:41-43).
struct LockRowShared {
  alignas(16) float x[N_INPUTS + 2];
  alignas(16) float row[ROW_PITCH_L0];
  u32 slot;
};

__global__ void __launch_bounds__(256, 1)
mix_predict_rows_kernel(StreamState* st, Tables T, const u16* ext /* device, N_EXT codes or null */) {
  __shared__ LockRowShared sh;
  const int tid = threadIdx.x, i = blockIdx.x;
  stage_inputs(sh.x, T, ext, st->small_x, st->lstm_x, tid, 256);
  __syncthreads();
  if (tid == 0) {
    const u32 sel = i == 12 ? aux_context(sh.x) : st->sel[i];
    const u32 s = resolve_slot(st->mixer[i], sel);
    st->slot[i] = s;
    sh.slot = s;
  }
  __syncthreads();
  {
    const float4* g = reinterpret_cast<const float4*>(st->mixer[i].rows + (size_t)sh.slot * ROW_PITCH_L0);
    float4* d = reinterpret_cast<float4*>(sh.row);
    for (int k = tid; k < ROW_PITCH_L0 / 4; k += 256) d[k] = g[k];
  }
  if (i == 0) for (int k = tid; k < N_INPUTS; k += 256) st->x[k] = sh.x[k];
  __syncthreads();
  if (tid == 0) st->mains[i] = chain_l0(sh.x, sh.row);
}

// mix_predict_final_kernel<<<1, 512>>>: extra-input substitution, layers 1-2, SSE (mixer.cpp:45-53,
// predictor.cpp:394-418, sse.cpp:243-289) from the 26 main sums.
__global__ void __launch_bounds__(MIX_THREADS, 1)
mix_predict_final_kernel(StreamState* st, Tables T) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  MixShared& sh = *reinterpret_cast<MixShared*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid < N_L0) sh.mains[tid] = st->mains[tid];
  if (tid >= 32 && tid < 35) { const int idx = tid == 32 ? 433 : (tid == 33 ? 2024 : 2077); sh.x[idx] = st->x[idx]; }
  if (tid >= N_L0 && tid < N_MIXERS) { const u32 s = resolve_slot(st->mixer[tid], st->sel[tid]); st->slot[tid] = s; sh.slot1[tid - N_L0] = s; }
  __syncthreads();
  for (int k = tid; k < N_L1 * ROW_PITCH_L1; k += MIX_THREADS) {
    const int i = k / ROW_PITCH_L1, c = k - i * ROW_PITCH_L1;
    sh.l1row[i][c] = st->mixer[N_L0 + i].rows[(size_t)sh.slot1[i] * ROW_PITCH_L1 + c];
  }
  if (tid < ROW_PITCH_L2) sh.l2row[tid] = st->mixer[N_L0 + N_L1].rows[(size_t)sh.slot1[N_L1] * ROW_PITCH_L2 + tid];
  for (int k = tid; k < N_L0 * N_L0; k += MIX_THREADS) {
    const int i = k / N_L0, c = k - i * N_L0;
    sh.we[i][c] = st->mixer[i].rows[(size_t)st->slot[i] * ROW_PITCH_L0 + N_INPUTS + c];
  }
  __syncthreads();
  if (warp == 0) {
    const float p = final_stage(sh, T, st->sse, lane);
    if (lane == 0) st->last_p = st->lstm_override >= 0.0f ? st->lstm_override : p;
  }
  __syncthreads();
  if (tid < N_L0) st->extras0[tid] = sh.x[N_INPUTS + tid];
  if (tid < N_L1) st->extras1[tid] = sh.l1extra[tid];
  if (tid < L2_IN) st->in2[tid] = sh.in2[tid];
  if (tid < N_MIXERS) st->mix_p[tid] = sh.mixp[tid];
}

// mix_perceive_kernel<<<26 + 2, 512>>>: Mixer::Perceive (mixer.cpp:56-72) for every mixer and
// SSE::Perceive (sse.cpp:291-305). CTA i < 26 owns layer-0 mixer i's row; CTA 26 the 21 small rows of
// layers 1-2, the SSE update and the step counter; CTA 27 the LSTM read-out's bit update
// (ByteModel::Perceive, byte-model.cpp:17-30).
__global__ void __launch_bounds__(MIX_THREADS, 1)
mix_perceive_kernel(StreamState* st, int bit, float decay_base, const u32* dbit = nullptr) {
  if (dbit) {                            // decode loop: the bit and the step's decay factor come from the device
    const DecodeState* ds = reinterpret_cast<const DecodeState*>(dbit);
    bit = (int)ds->bit;
    if (ds->decay) decay_base = ds->decay[ds->t - 1];
  }
  __shared__ float upd[N_L1 + 2];
  __shared__ u32 shrink[N_L1 + 2];
  const int tid = threadIdx.x, blk = blockIdx.x;
  if (blk < N_L0) {
    const int i = blk;
    if (tid == 0) upd[0] = mixer_update_coeff(st->mixer[i], st->slot[i], decay_base, st->mix_p[i], bit, &shrink[0]);
    __syncthreads();
    float* row = st->mixer[i].rows + (size_t)st->slot[i] * ROW_PITCH_L0;
    const int n = N_INPUTS + i;
    const float u = upd[0];
    const bool shr = shrink[0] != 0;
    for (int k = tid; k < n; k += MIX_THREADS) {
      const float xin = k < N_INPUTS ? st->x[k] : st->extras0[k - N_INPUTS];
      float w = XM_FSUB(row[k], XM_FMUL(u, xin));
      if (shr) w = XM_FMUL(w, 1.0f - 3.0e-6f);
      row[k] = w;
    }
    return;
  }
  if (blk == N_L0 + 1) {
    if (tid == 0) bm_perceive(st->lstm.bm, bit);
    return;
  }
  if (tid < N_L1 + 1) upd[tid] = mixer_update_coeff(st->mixer[N_L0 + tid], st->slot[N_L0 + tid], decay_base, st->mix_p[N_L0 + tid], bit, &shrink[tid]);
  if (tid == 64) sse_perceive(st->sse, bit);
  __syncthreads();
  for (int k = tid; k < (N_L1 + 1) * ROW_PITCH_L1; k += MIX_THREADS) {
    const int i = k / ROW_PITCH_L1, c = k - i * ROW_PITCH_L1;
    const int mi = N_L0 + i;
    const int n = i < N_L1 ? L1_IN + i : L2_IN;
    if (c < n) {
      float xin;
      if (i < N_L1) {
        if (c < N_L0) xin = st->extras0[c];                       // layer-1 input c = clamp(layer-0 output c)
        else if (c < L1_IN) xin = st->in2[N_L0 + N_L1 + (c - N_L0)];   // the 3 auxiliary inputs
        else xin = st->extras1[c - L1_IN];
      } else xin = st->in2[c];
      float* w = st->mixer[mi].rows + (size_t)st->slot[mi] * ROW_PITCH_L1 + c;
      float v = XM_FSUB(*w, XM_FMUL(upd[i], xin));
      if (shrink[i]) v = XM_FMUL(v, 1.0f - 3.0e-6f);
      *w = v;
    }
  }
  if (tid == 0) st->bits_done += 1;
}

}  // namespace cmixb200

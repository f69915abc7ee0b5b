// cmix_b200/csrc/lstm.cuh
//
// Kernel "lstm": the byte-level LSTM mixer (reference src/mixer/lstm.cpp,
// src/mixer/lstm-layer.cpp, src/mixer/byte-mixer.cpp; SURVEY §8 rows a9-a12,
// Appendix E). 2 layers x 200 cells, 3 gates with RMS-norm, truncated BPTT over
// 100 bytes, Adam.
//
// Parity contract: every floating-point result is bit-identical to the strict-FP
// reference. That fixes the ORDER of every sum (the reference's scalar loops and
// libstdc++'s valarray reductions), so parallelism comes only from the
// independent chains: 600 gate rows per layer in the forward pass, 200 columns in
// the transposed mat-vecs, and ~1M independent weight elements in the
// weight-gradient accumulation, which is restructured from "100 rank-1 updates"
// into one pass where each thread owns one weight and adds its 100 terms in the
// reference's time order (99 -> 0) — same values, ~100x less memory traffic.
// Tensor cores are deliberately not used: tcgen05 kinds round products to
// TF32/BF16 and accumulate in an unspecified order, either of which breaks the
// bit-exact contract.
#pragma once
#include "exact_math.h"
#include "small_models.cuh"
#include "state.h"

namespace cmixb200 {

enum { LSTM_THREADS = 1024 };
#define LC LSTM_CELLS
#define LH LSTM_HORIZON

struct LstmShared {
  float in[2 * 256 + 2 * LC + 8];     // current layer input vector
  float norm[3][LC];
  float e[3][LC];
  float act[3][LC];
  float tmp[3][LC];
  float tmp2[3][LC];
  float vec[LC];
  float out[256];
  float err[256];
  float red[32];
  float scal[8];
  int sym[LH];
};

__device__ __forceinline__ float clipf(float v, float c) { return v < -c ? -c : (v > c ? c : v); }

// LstmLayer::ForwardPass (lstm-layer.cpp:62-99) for one layer; all threads of the CTA.
__device__ void lstm_layer_forward(LayerState& L, int V, int sym, float* hidden_out, LstmShared& sh, int tid) {
  const int e = L.epoch;
  const int in_size = L.in_size;
  const float* in_g = L.input + (size_t)e * in_size;
  for (int j = tid; j < in_size; j += LSTM_THREADS) sh.in[j] = in_g[j];
  __syncthreads();
  if (tid < 3 * LC) {
    const int g = tid / LC, i = tid % LC;
    const float* w = L.gate[g].w;
    float f = w[(size_t)sym * LC + i];
    const float* wc = w + (size_t)V * LC + i;
#pragma unroll 8
    for (int j = 0; j < in_size; ++j) f = XM_FADD(f, XM_FMUL(sh.in[j], wc[(size_t)j * LC]));
    sh.norm[g][i] = f;
  }
  __syncthreads();
  if (tid < 96 && (tid & 31) == 0) {          // one lane per gate: _Expr::sum() adds back to front
    const int g = tid >> 5;
    float ss = XM_FMUL(sh.norm[g][LC - 1], sh.norm[g][LC - 1]);
    for (int i = LC - 2; i >= 0; --i) ss = XM_FADD(ss, XM_FMUL(sh.norm[g][i], sh.norm[g][i]));
    const float iv = XM_FDIV(1.0f, __fsqrt_rn(XM_FADD(XM_FDIV(ss, (float)LC), 1e-5f)));
    L.gate[g].ivar[e] = iv;
    sh.scal[g] = iv;
  }
  __syncthreads();
  if (tid < 3 * LC) {
    const int g = tid / LC, i = tid % LC;
    GateState& G = L.gate[g];
    const float n = XM_FMUL(sh.norm[g][i], sh.scal[g]);
    G.norm[(size_t)e * LC + i] = n;
    float s = XM_FADD(XM_FMUL(n, G.gamma[i]), G.beta[i]);
    s = (g == 1) ? xm_tanhf(s) : xm_logistic(s);
    G.state[(size_t)e * LC + i] = s;
    sh.act[g][i] = s;
  }
  __syncthreads();
  if (tid < LC) {
    const int i = tid;
    const float fs = sh.act[0][i], gs = sh.act[1][i], os = sh.act[2][i];
    float c = L.state[i];
    L.last_state[(size_t)e * LC + i] = c;
    const float ig = XM_FSUB(1.0f, fs);
    L.input_gate_state[(size_t)e * LC + i] = ig;
    c = XM_FMUL(c, fs);
    c = XM_FADD(c, XM_FMUL(gs, ig));
    L.state[i] = c;
    const float ts = xm_tanhf(c);
    L.tanh_state[(size_t)e * LC + i] = ts;
    hidden_out[i] = XM_FMUL(os, ts);
  }
  __syncthreads();
  if (tid == 0) { L.epoch = (e + 1 == LH) ? 0 : e + 1; }
  __syncthreads();
}

// Lstm::Predict (lstm.cpp:120-150)
__device__ void lstm_predict(LstmState& S, unsigned input, LstmShared& sh, int tid) {
  const int V = S.V, e = S.epoch, HW = LSTM_HID;
  for (int l = 0; l < 2; ++l) {
    LayerState& L = S.layer[l];
    float* in = L.input + (size_t)e * L.in_size;
    if (tid < LC) in[V + tid] = S.hidden[l * LC + tid];
    __syncthreads();
    lstm_layer_forward(L, V, (int)input, &S.hidden[l * LC], sh, tid);
    if (l == 0) {
      float* in1 = S.layer[1].input + (size_t)e * S.layer[1].in_size;
      if (tid < LC) in1[V + LC + tid] = S.hidden[tid];
      __syncthreads();
    }
  }
  for (int j = tid; j < HW; j += LSTM_THREADS) sh.in[j] = S.hidden[j];
  __syncthreads();
  float* out = S.output + (size_t)e * V;
  const float* W = S.out_w + (size_t)e * V * HW;
  float sum = 0.0f;
  if (tid < V) {
    const float* wr = W + (size_t)tid * HW;
#pragma unroll 4
    for (int j = 0; j < HW; ++j) sum = XM_FADD(sum, XM_FMUL(sh.in[j], wr[j]));
    sh.out[tid] = sum;
  }
  // max_out = max(0, max_i sum_i): order-independent
  float mx = (tid < V) ? sum : 0.0f;
  mx = fmaxf(mx, 0.0f);
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((tid & 31) == 0) sh.red[tid >> 5] = mx;
  __syncthreads();
  if (tid < 32) {
    float m2 = sh.red[tid];
    for (int o = 16; o > 0; o >>= 1) m2 = fmaxf(m2, __shfl_xor_sync(0xffffffffu, m2, o));
    if (tid == 0) sh.scal[0] = m2;
  }
  __syncthreads();
  const float max_out = sh.scal[0];
  if (tid < V) sh.out[tid] = xm_expf(XM_FSUB(sh.out[tid], max_out));
  __syncthreads();
  if (tid == 0) {                      // valarray::sum(): front to back
    float total = sh.out[0];
    for (int i = 1; i < V; ++i) total = XM_FADD(total, sh.out[i]);
    sh.scal[1] = total;
  }
  __syncthreads();
  if (tid < V) { const float o = XM_FDIV(sh.out[tid], sh.scal[1]); sh.out[tid] = o; out[tid] = o; }
  __syncthreads();
  if (tid == 0) S.epoch = (e + 1 == LH) ? 0 : e + 1;
  __syncthreads();
}

// One (epoch, layer) step of the error recursion (LstmLayer::BackwardPass, lstm-layer.cpp:108-197)
// WITHOUT the weight-gradient accumulation and Adam, which lstm_apply_updates() does afterwards.
__device__ void lstm_layer_backward(LstmState& S, int l, int ep, LstmShared& sh, int tid,
                                    float* gamma_u, float* beta_u) {
  LayerState& L = S.layer[l];
  const int V = S.V;
  const float kClip = 10.0f;
  float he = 0.0f, stored = 0.0f, se = 0.0f;
  if (tid < LC) {
    const int i = tid;
    const size_t o = (size_t)ep * LC + i;
    const float ts = L.tanh_state[o], os = L.gate[2].state[o], gs = L.gate[1].state[o], fs = L.gate[0].state[o];
    const float ig = L.input_gate_state[o], ls = L.last_state[o];
    he = S.hidden_error[i];
    if (ep == LH - 1) { stored = he; se = 0.0f; }
    else { stored = XM_FADD(L.stored_error[i], he); se = L.state_error[i]; }
    sh.e[2][i] = XM_FMUL(XM_FMUL(XM_FMUL(ts, stored), os), XM_FSUB(1.0f, os));
    se = XM_FADD(se, XM_FMUL(XM_FMUL(stored, os), XM_FSUB(1.0f, XM_FMUL(ts, ts))));
    sh.e[1][i] = XM_FMUL(XM_FMUL(se, ig), XM_FSUB(1.0f, XM_FMUL(gs, gs)));
    sh.e[0][i] = XM_FMUL(XM_FMUL(XM_FMUL(XM_FSUB(ls, gs), se), fs), ig);
    he = 0.0f;
    if (ep > 0) { se = XM_FMUL(se, fs); stored = 0.0f; }
  }
  if (tid == 0 && ep == 0) { if (L.update_steps < 3000) ++L.update_steps; }
  __syncthreads();
  // per gate: beta_u/gamma_u accumulation, RMS-norm backward
  if (tid < 3 * LC) {
    const int g = tid / LC, i = tid % LC;
    GateState& G = L.gate[g];
    const float n = G.norm[(size_t)ep * LC + i];
    float e = sh.e[g][i];
    if (ep == LH - 1) { *gamma_u = 0.0f; *beta_u = 0.0f; }
    *beta_u = XM_FADD(*beta_u, e);
    *gamma_u = XM_FADD(*gamma_u, XM_FMUL(e, n));
    e = XM_FMUL(e, XM_FMUL(G.gamma[i], G.ivar[ep]));
    sh.e[g][i] = e;
    sh.norm[g][i] = n;
  }
  __syncthreads();
  if (tid < 96 && (tid & 31) == 0) {
    const int g = tid >> 5;
    float s = XM_FMUL(sh.e[g][LC - 1], sh.norm[g][LC - 1]);
    for (int i = LC - 2; i >= 0; --i) s = XM_FADD(s, XM_FMUL(sh.e[g][i], sh.norm[g][i]));
    sh.scal[g] = XM_FDIV(s, (float)LC);
  }
  __syncthreads();
  if (tid < 3 * LC) {
    const int g = tid / LC, i = tid % LC;
    const float e = XM_FSUB(sh.e[g][i], XM_FMUL(sh.scal[g], sh.norm[g][i]));
    sh.act[g][i] = e;                                   // final gate error for this step
    L.gate[g].err[(size_t)ep * LC + i] = e;
  }
  __syncthreads();
  // transposed mat-vecs: hidden_error (to the layer below) and stored_error (to the previous step)
  if (tid < 3 * LC) {
    const int g = tid / LC, i = tid % LC;
    const float* w = L.gate[g].w;
    float f1 = 0.0f, f2 = 0.0f;
    if (l > 0) {
      const float* col = w + (size_t)(2 * V + LC + i) * LC;
#pragma unroll 8
      for (int j = 0; j < LC; ++j) f1 = XM_FADD(f1, XM_FMUL(sh.act[g][j], col[j]));
    }
    if (ep > 0) {
      const float* col = w + (size_t)(2 * V + i) * LC;
#pragma unroll 8
      for (int j = 0; j < LC; ++j) f2 = XM_FADD(f2, XM_FMUL(sh.act[g][j], col[j]));
    }
    sh.tmp[g][i] = f1;
    sh.tmp2[g][i] = f2;
  }
  __syncthreads();
  if (tid < LC) {
    const int i = tid;
    if (l > 0) { he = XM_FADD(he, sh.tmp[0][i]); he = XM_FADD(he, sh.tmp[1][i]); he = XM_FADD(he, sh.tmp[2][i]); }
    if (ep > 0) { stored = XM_FADD(stored, sh.tmp2[0][i]); stored = XM_FADD(stored, sh.tmp2[1][i]); stored = XM_FADD(stored, sh.tmp2[2][i]); }
    L.state_error[i] = clipf(se, kClip);
    L.stored_error[i] = clipf(stored, kClip);
    S.hidden_error[i] = clipf(he, kClip);
  }
  __syncthreads();
}

// Weight-gradient accumulation in the reference's time order + Adam (lstm-layer.cpp:11-32,182-196).
__device__ void lstm_apply_updates(LstmState& S, LstmShared& sh, int tid, float gamma_u[2], float beta_u[2]) {
  const int V = S.V;
  const float beta1 = 0.025f, beta2 = 0.9999f, eps = 1e-6f;
  for (int l = 0; l < 2; ++l) {
    LayerState& L = S.layer[l];
    const float* ad = S.adam + 4 * L.update_steps;
    const float alpha = ad[0], bc1 = ad[1], bc2 = ad[2];
    const int in_size = L.in_size;
    for (int g = 0; g < 3; ++g) {
      GateState& G = L.gate[g];
      const size_t n = (size_t)G.row * LC;
      for (size_t idx = tid; idx < n; idx += LSTM_THREADS) {
        const int col = (int)(idx / LC), i = (int)(idx % LC);
        float acc = 0.0f;
        if (col >= V) {
          const float* in = L.input + (col - V);
          for (int ep = LH - 1; ep >= 0; --ep)
            acc = XM_FADD(acc, XM_FMUL(G.err[(size_t)ep * LC + i], in[(size_t)ep * in_size]));
        } else {
          for (int ep = LH - 1; ep >= 0; --ep)
            if (sh.sym[ep] == col) acc = XM_FADD(acc, G.err[(size_t)ep * LC + i]);
        }
        float m = G.m[idx], v = G.v[idx], w = G.w[idx];
        m = XM_FMUL(m, beta1); m = XM_FADD(m, XM_FMUL(1.0f - beta1, acc));
        v = XM_FMUL(v, beta2); v = XM_FADD(v, XM_FMUL(XM_FMUL(1.0f - beta2, acc), acc));
        w = XM_FSUB(w, XM_FMUL(alpha, XM_FDIV(XM_FDIV(m, bc1), __fsqrt_rn(XM_FADD(XM_FDIV(v, bc2), eps)))));
        G.m[idx] = m; G.v[idx] = v; G.w[idx] = w;
      }
    }
    if (tid < 3 * LC) {
      const int g = tid / LC, i = tid % LC;
      GateState& G = L.gate[g];
      {
        const float acc = gamma_u[l];
        float m = G.gamma_m[i], v = G.gamma_v[i], w = G.gamma[i];
        m = XM_FMUL(m, beta1); m = XM_FADD(m, XM_FMUL(1.0f - beta1, acc));
        v = XM_FMUL(v, beta2); v = XM_FADD(v, XM_FMUL(XM_FMUL(1.0f - beta2, acc), acc));
        w = XM_FSUB(w, XM_FMUL(alpha, XM_FDIV(XM_FDIV(m, bc1), __fsqrt_rn(XM_FADD(XM_FDIV(v, bc2), eps)))));
        G.gamma_m[i] = m; G.gamma_v[i] = v; G.gamma[i] = w;
      }
      {
        const float acc = beta_u[l];
        float m = G.beta_m[i], v = G.beta_v[i], w = G.beta[i];
        m = XM_FMUL(m, beta1); m = XM_FADD(m, XM_FMUL(1.0f - beta1, acc));
        v = XM_FMUL(v, beta2); v = XM_FADD(v, XM_FMUL(XM_FMUL(1.0f - beta2, acc), acc));
        w = XM_FSUB(w, XM_FMUL(alpha, XM_FDIV(XM_FDIV(m, bc1), __fsqrt_rn(XM_FADD(XM_FDIV(v, bc2), eps)))));
        G.beta_m[i] = m; G.beta_v[i] = v; G.beta[i] = w;
      }
    }
    __syncthreads();
  }
}

// ByteMixer::ByteUpdate -> Lstm::SetInput + Lstm::Perceive + Lstm::Predict (byte-mixer.cpp:22-38,
// lstm.cpp:80-150). `ppmd` = 256-entry PPMD distribution after this byte (or null), `byte` = the
// byte just completed. Leaves the new 256-entry distribution in S.bm.probs.
__device__ void lstm_byte_update(LstmState& S, const float* ppmd, u32 byte, LstmShared& sh, int tid) {
  const int V = S.V, HW = LSTM_HID;
  const unsigned input = (unsigned)S.byte_map[byte];
  // SetInput: aux[k] = 2 * ppmd[k-th vocabulary byte] into both layers' input at epoch_
  if (tid < 256 && S.vocab[tid]) {
    const float p = ppmd ? ppmd[tid] : (float)(1. / 256);
    const float a = XM_FMUL(XM_FADD(0.0f, p), 2.0f);
    const int k = S.byte_map[tid];
    S.layer[0].input[(size_t)S.epoch * S.layer[0].in_size + k] = a;
    S.layer[1].input[(size_t)S.epoch * S.layer[1].in_size + k] = a;
  }
  __syncthreads();
  const int epoch = S.epoch;
  const int last_epoch = epoch == 0 ? LH - 1 : epoch - 1;
  const int old_input = (int)S.input_history[last_epoch];
  __syncthreads();
  if (tid == 0) S.input_history[last_epoch] = input;
  __syncthreads();
  if (epoch == 0) {
    // input symbol of step ep is input_history[ep-1]; for ep == 0 the value that was just overwritten
    if (tid < LH) sh.sym[tid] = tid == 0 ? old_input : (int)S.input_history[tid - 1];
    float gamma_u[2] = {0.0f, 0.0f}, beta_u[2] = {0.0f, 0.0f};
    __syncthreads();
    for (int ep = LH - 1; ep >= 0; --ep) {
      const float* out = S.output + (size_t)ep * V;
      const float* W = S.out_w + (size_t)ep * V * HW;
      if (tid < V) sh.err[tid] = ((unsigned)tid == S.input_history[ep]) ? XM_FSUB(out[tid], 1.0f) : out[tid];
      __syncthreads();
      for (int l = 1; l >= 0; --l) {
        if (tid < LC) {
          float he = S.hidden_error[tid];
          const float* wc = W + l * LC + tid;
#pragma unroll 4
          for (int i = 0; i < V; ++i) he = XM_FADD(he, XM_FMUL(wc[(size_t)i * HW], sh.err[i]));
          S.hidden_error[tid] = he;
        }
        __syncthreads();
        lstm_layer_backward(S, l, ep, sh, tid, &gamma_u[l], &beta_u[l]);
      }
    }
    lstm_apply_updates(S, sh, tid, gamma_u, beta_u);
  }
  // output layer SGD (lstm.cpp:112-116): copy W_o[last_epoch] to W_o[epoch] with the step applied
  {
    const float* out = S.output + (size_t)last_epoch * V;
    const float* Wl = S.out_w + (size_t)last_epoch * V * HW;
    float* We = S.out_w + (size_t)epoch * V * HW;
    for (int j = tid; j < HW; j += LSTM_THREADS) sh.in[j] = S.hidden[j];
    if (tid < V) sh.err[tid] = XM_FMUL(0.03f, ((unsigned)tid == input) ? XM_FSUB(out[tid], 1.0f) : out[tid]);
    __syncthreads();
    const int n = V * HW;
    for (int idx = tid; idx < n; idx += LSTM_THREADS) {
      const int i = idx / HW, j = idx - i * HW;
      We[idx] = XM_FSUB(Wl[idx], XM_FMUL(sh.err[i], sh.in[j]));
    }
    __syncthreads();
  }
  lstm_predict(S, input, sh, tid);
  // ByteMixer: scatter back to 256 bytes; ByteModel::ByteUpdate resets the range
  if (tid < 256) S.bm.probs[tid] = S.vocab[tid] ? sh.out[S.byte_map[tid]] : 0.0f;
  if (tid == 0) { S.bm.top = 255; S.bm.bot = 0; }
  __syncthreads();
}

// Bit-level read-out of the byte distribution (ByteModel::Predict + the override test of
// predictor.cpp:378-387) followed by ByteModel::Perceive. Single thread.
__device__ void lstm_readout(LstmState& S, const Tables& T, float* x_out, float* override_out) {
  const float p = bytemodel_predict(S.bm.probs, S.bm.bot, S.bm.top, &S.bm.ex);
  *override_out = (p == 0.0f || p == 1.0f) ? p : -1.0f;
  *x_out = stretch(T, p);
}
__device__ void bm_perceive(ByteModelState& b, int bit) {
  b.mid = b.bot + ((b.top - b.bot) / 2);
  if (bit) b.bot = b.mid + 1; else b.top = b.mid;
}

__global__ void __launch_bounds__(LSTM_THREADS, 1)
lstm_kernel(const ChunkArgs* __restrict__ args_all, Tables T) {
  const ChunkArgs a = args_all[blockIdx.x];
  LstmState& S = a.st->lstm;
  extern __shared__ unsigned char smem_raw[];
  LstmShared& sh = *reinterpret_cast<LstmShared*>(smem_raw);
  const int tid = threadIdx.x;
  for (u32 pos = 0; pos < a.n_bytes; ++pos) {
    const u32 byte = a.bytes[pos];
    if (tid == 0) {
      for (int j = 7; j >= 0; --j) {
        const u64 t = (u64)pos * 8 + (7 - j);
        lstm_readout(S, T, &a.lstm_x[2 * t], &a.lstm_x[2 * t + 1]);
        bm_perceive(S.bm, (byte >> j) & 1);
      }
    }
    __syncthreads();
    lstm_byte_update(S, a.ppmd ? a.ppmd + (u64)pos * 256 : nullptr, byte, sh, tid);
  }
}

// Lock-step halves.
__global__ void lstm_predict_kernel(StreamState* st, Tables T) {
  if (threadIdx.x == 0) lstm_readout(st->lstm, T, &st->lstm_x, &st->lstm_override);
}
__global__ void __launch_bounds__(LSTM_THREADS, 1)
lstm_perceive_kernel(StreamState* st, int bit, int byte_done, u32 byte, const float* ppmd) {
  extern __shared__ unsigned char smem_raw[];
  LstmShared& sh = *reinterpret_cast<LstmShared*>(smem_raw);
  if (threadIdx.x == 0) bm_perceive(st->lstm.bm, bit);
  __syncthreads();
  if (byte_done) lstm_byte_update(st->lstm, ppmd, byte, sh, threadIdx.x);
}

#undef LC
#undef LH
}  // namespace cmixb200

// oracle/port/lstm.cpp — TEST INFRASTRUCTURE.
// Restatement of the byte-level LSTM mixer (reference src/mixer/lstm.cpp,
// src/mixer/lstm-layer.cpp) with flat
// arrays. Every floating-point expression keeps the reference's operand
// order, including the two different summation directions libstdc++ uses:
// std::valarray<T>::sum() adds front-to-back, while the expression-template
// _Expr::sum() (e.g. `(a*b).sum()`) adds back-to-front.
#include "internal.h"

#include <math.h>

namespace op {

namespace {

const int C = 200;        // cells
const int H = 100;        // BPTT horizon
const float kLr = 0.03f, kClip = 10.0f;
const int kUpdateLimit = 3000;

struct Gate {             // NeuronLayer (lstm-layer.h:9-30)
  int row;                // weights per cell
  std::vector<float> w, update, m, v;            // [C][row]
  std::vector<float> transpose;                  // [row - 2V][C]
  std::vector<float> state, norm;                // [H][C]
  float ivar[H];
  float error[C], gamma[C], gamma_u[C], gamma_m[C], gamma_v[C], beta[C], beta_u[C], beta_m[C], beta_v[C];
  void Init(int row_, int V) {
    row = row_;
    w.assign((size_t)C * row, 0); update = m = v = w;
    transpose.assign((size_t)(row - 2 * V) * C, 0);
    state.assign(H * C, 0); norm = state;
    for (int i = 0; i < C; ++i) {
      error[i] = gamma_u[i] = gamma_m[i] = gamma_v[i] = 0;
      beta[i] = beta_u[i] = beta_m[i] = beta_v[i] = 0;
      gamma[i] = 1.0f;
    }
    for (float& x : ivar) x = 0;
  }
};

struct AdamStep { float alpha, bc1, bc2; };

// lstm-layer.cpp:11-32, one vector.
void Adam(float* g, float* m, float* v, float* w, int n, const AdamStep& s) {
  const float beta1 = 0.025, beta2 = 0.9999, eps = 1e-6f;
  for (int i = 0; i < n; ++i) {
    m[i] *= beta1;
    m[i] += (1.0f - beta1) * g[i];
    v[i] *= beta2;
    v[i] += (1.0f - beta2) * g[i] * g[i];
    w[i] -= s.alpha * ((m[i] / s.bc1) / (sqrtf(v[i] / s.bc2 + eps)));
  }
}

AdamStep MakeAdamStep(float learning_rate, float t, unsigned long long update_limit) {
  const float beta1 = 0.025, beta2 = 0.9999;
  AdamStep s;
  if (t < update_limit) {
    s.alpha = learning_rate * 0.1f / sqrt(5e-5f * t + 1.0f);
    s.bc1 = (float)(1.0f - pow(beta1, t));
    s.bc2 = (float)(1.0f - pow(beta2, t));
  } else {
    s.alpha = learning_rate * 0.1f / sqrt(5e-5f * update_limit + 1.0f);
    s.bc1 = (float)(1.0f - pow(beta1, update_limit));
    s.bc2 = (float)(1.0f - pow(beta2, update_limit));
  }
  return s;
}

struct Layer {            // LstmLayer
  int V, in_size;         // in_size = length of the dense input vector (V+201 / V+401)
  float state[C], state_error[C], stored_error[C];
  std::vector<float> tanh_state, input_gate_state, last_state;   // [H][C]
  int epoch = 0;
  unsigned long long update_steps = 0;
  Gate forget, inode, ogate;

  void Init(int V_, int in_size_, GlibcRand& rng) {
    V = V_; in_size = in_size_;
    int row = in_size + V;
    forget.Init(row, V); inode.Init(row, V); ogate.Init(row, V);
    for (int i = 0; i < C; ++i) state[i] = state_error[i] = stored_error[i] = 0;
    tanh_state.assign(H * C, 0); input_gate_state = last_state = tanh_state;
    float val = sqrt(6.0f / float(V + V));                // lstm-layer.cpp:49
    float low = -val;
    float range = 2 * val;
    auto rnd = [&]() { return static_cast<float>(rng.next()) / static_cast<float>(RAND_MAX); };
    for (int i = 0; i < C; ++i) {
      for (int j = 0; j < row; ++j) {
        forget.w[(size_t)i * row + j] = low + rnd() * range;
        inode.w[(size_t)i * row + j] = low + rnd() * range;
        ogate.w[(size_t)i * row + j] = low + rnd() * range;
      }
      forget.w[(size_t)i * row + row - 1] = 1;
    }
  }

  void GateForward(Gate& g, const float* input, int sym) {        // lstm-layer.cpp:85-99
    float* norm = &g.norm[epoch * C];
    for (int i = 0; i < C; ++i) {
      const float* wr = &g.w[(size_t)i * g.row];
      float f = wr[sym];
      for (int j = 0; j < in_size; ++j) f += input[j] * wr[V + j];
      norm[i] = f;
    }
    float ss = norm[C - 1] * norm[C - 1];                          // _Expr::sum(): back to front
    for (int i = C - 2; i >= 0; --i) ss += norm[i] * norm[i];
    g.ivar[epoch] = 1.0f / sqrt((ss / C) + 1e-5f);
    float* st = &g.state[epoch * C];
    for (int i = 0; i < C; ++i) norm[i] *= g.ivar[epoch];
    for (int i = 0; i < C; ++i) st[i] = norm[i] * g.gamma[i] + g.beta[i];
  }

  void Forward(const float* input, int sym, float* hidden_out) {   // lstm-layer.cpp:62-83
    float* fs = &forget.state[epoch * C];
    float* gs = &inode.state[epoch * C];
    float* os = &ogate.state[epoch * C];
    for (int i = 0; i < C; ++i) last_state[epoch * C + i] = state[i];
    GateForward(forget, input, sym);
    GateForward(inode, input, sym);
    GateForward(ogate, input, sym);
    for (int i = 0; i < C; ++i) {
      fs[i] = Logistic(fs[i]);
      gs[i] = tanh(gs[i]);
      os[i] = Logistic(os[i]);
    }
    float* ig = &input_gate_state[epoch * C];
    float* ts = &tanh_state[epoch * C];
    for (int i = 0; i < C; ++i) ig[i] = 1.0f - fs[i];
    for (int i = 0; i < C; ++i) state[i] *= fs[i];
    for (int i = 0; i < C; ++i) state[i] += gs[i] * ig[i];
    for (int i = 0; i < C; ++i) ts[i] = tanh(state[i]);
    for (int i = 0; i < C; ++i) hidden_out[i] = os[i] * ts[i];
    if (++epoch == H) epoch = 0;
  }

  void Clip(float* a) {
    for (int i = 0; i < C; ++i) { if (a[i] < -kClip) a[i] = -kClip; else if (a[i] > kClip) a[i] = kClip; }
  }

  void GateBackward(Gate& g, const float* input, int ep, int layer, int sym, float* hidden_error) {
    const int row = g.row;                                         // lstm-layer.cpp:145-197
    if (ep == H - 1) {
      for (int i = 0; i < C; ++i) g.gamma_u[i] = g.beta_u[i] = 0;
      for (int i = 0; i < C; ++i) {
        for (int j = 0; j < row; ++j) g.update[(size_t)i * row + j] = 0;
        int offset = 2 * V;
        for (int j = 0; j < row - offset; ++j) g.transpose[(size_t)j * C + i] = g.w[(size_t)i * row + j + offset];
      }
    }
    const float* norm = &g.norm[ep * C];
    for (int i = 0; i < C; ++i) g.beta_u[i] += g.error[i];
    for (int i = 0; i < C; ++i) g.gamma_u[i] += g.error[i] * norm[i];
    for (int i = 0; i < C; ++i) g.error[i] *= g.gamma[i] * g.ivar[ep];
    float s = g.error[C - 1] * norm[C - 1];
    for (int i = C - 2; i >= 0; --i) s += g.error[i] * norm[i];
    s = s / C;
    for (int i = 0; i < C; ++i) g.error[i] -= s * norm[i];
    if (layer > 0) {
      for (int i = 0; i < C; ++i) {
        float f = 0;
        const float* tr = &g.transpose[(size_t)(C + i) * C];
        for (int j = 0; j < C; ++j) f += g.error[j] * tr[j];
        hidden_error[i] += f;
      }
    }
    if (ep > 0) {
      for (int i = 0; i < C; ++i) {
        float f = 0;
        const float* tr = &g.transpose[(size_t)i * C];
        for (int j = 0; j < C; ++j) f += g.error[j] * tr[j];
        stored_error[i] += f;
      }
    }
    for (int i = 0; i < C; ++i) {
      float* u = &g.update[(size_t)i * row];
      for (int j = 0; j < in_size; ++j) u[V + j] += g.error[i] * input[j];
      u[sym] += g.error[i];
    }
    if (ep == 0) {
      AdamStep st = MakeAdamStep(kLr, update_steps, kUpdateLimit);
      for (int i = 0; i < C; ++i)
        Adam(&g.update[(size_t)i * row], &g.m[(size_t)i * row], &g.v[(size_t)i * row], &g.w[(size_t)i * row], row, st);
      Adam(g.gamma_u, g.gamma_m, g.gamma_v, g.gamma, C, st);
      Adam(g.beta_u, g.beta_m, g.beta_v, g.beta, C, st);
    }
  }

  void Backward(const float* input, int ep, int layer, int sym, float* hidden_error) {   // :108-143
    const float* ts = &tanh_state[ep * C];
    const float* os = &ogate.state[ep * C];
    const float* gs = &inode.state[ep * C];
    const float* fs = &forget.state[ep * C];
    const float* ig = &input_gate_state[ep * C];
    const float* ls = &last_state[ep * C];
    if (ep == H - 1) {
      for (int i = 0; i < C; ++i) { stored_error[i] = hidden_error[i]; state_error[i] = 0; }
    } else {
      for (int i = 0; i < C; ++i) stored_error[i] += hidden_error[i];
    }
    for (int i = 0; i < C; ++i) ogate.error[i] = ts[i] * stored_error[i] * os[i] * (1.0f - os[i]);
    for (int i = 0; i < C; ++i) state_error[i] += stored_error[i] * os[i] * (1.0f - (ts[i] * ts[i]));
    for (int i = 0; i < C; ++i) inode.error[i] = state_error[i] * ig[i] * (1.0f - (gs[i] * gs[i]));
    for (int i = 0; i < C; ++i) forget.error[i] = (ls[i] - gs[i]) * state_error[i] * fs[i] * ig[i];
    for (int i = 0; i < C; ++i) hidden_error[i] = 0;
    if (ep > 0) {
      for (int i = 0; i < C; ++i) { state_error[i] *= fs[i]; stored_error[i] = 0; }
    } else {
      if (update_steps < (unsigned long long)kUpdateLimit) ++update_steps;
    }
    GateBackward(forget, input, ep, layer, sym, hidden_error);
    GateBackward(inode, input, ep, layer, sym, hidden_error);
    GateBackward(ogate, input, ep, layer, sym, hidden_error);
    Clip(state_error); Clip(stored_error); Clip(hidden_error);
  }
};

}  // namespace

struct Lstm {             // Lstm (lstm.cpp) with 2 layers
  int V, epoch = 0;
  Layer layer[2];
  int in_size[2];
  std::vector<float> layer_input[2];   // [H][in_size]
  std::vector<float> out_w;            // [H][V][2C+1]
  std::vector<float> output;           // [H][V]
  float hidden[2 * C + 1], hidden_error[C];
  unsigned input_history[H];
};

Lstm* lstm_create(int V, GlibcRand& rng) {
  Lstm* L = new Lstm();
  L->V = V;
  L->in_size[0] = V + C + 1;
  L->in_size[1] = V + 2 * C + 1;
  for (int l = 0; l < 2; ++l) {
    L->layer_input[l].assign((size_t)H * L->in_size[l], 0.0f);
    for (int e = 0; e < H; ++e) L->layer_input[l][(size_t)e * L->in_size[l] + L->in_size[l] - 1] = 1;
  }
  L->out_w.assign((size_t)H * V * (2 * C + 1), 0.0f);
  L->output.assign((size_t)H * V, 1.0 / V);
  for (float& x : L->hidden) x = 0;
  L->hidden[2 * C] = 1;
  for (float& x : L->hidden_error) x = 0;
  for (unsigned& x : L->input_history) x = 0;
  for (int l = 0; l < 2; ++l) L->layer[l].Init(V, L->in_size[l], rng);
  return L;
}
void lstm_destroy(Lstm* L) { delete L; }

static const float* Predict(Lstm* L, unsigned input) {              // lstm.cpp:120-150
  const int V = L->V, e = L->epoch, HW = 2 * C + 1;
  for (int l = 0; l < 2; ++l) {
    float* in = &L->layer_input[l][(size_t)e * L->in_size[l]];
    for (int i = 0; i < C; ++i) in[V + i] = L->hidden[l * C + i];
    L->layer[l].Forward(in, input, &L->hidden[l * C]);
    if (l == 0) {
      float* in1 = &L->layer_input[1][(size_t)e * L->in_size[1]];
      for (int i = 0; i < C; ++i) in1[V + C + i] = L->hidden[i];
    }
  }
  float* out = &L->output[(size_t)e * V];
  const float* W = &L->out_w[(size_t)e * V * HW];
  float max_out = 0;
  for (int i = 0; i < V; ++i) {
    float sum = 0;
    for (int j = 0; j < HW; ++j) sum += L->hidden[j] * W[(size_t)i * HW + j];
    out[i] = sum;
    max_out = std::max(sum, max_out);
  }
  for (int i = 0; i < V; ++i) out[i] = exp(out[i] - max_out);
  float total = out[0];                                              // valarray::sum(): front to back
  for (int i = 1; i < V; ++i) total += out[i];
  for (int i = 0; i < V; ++i) out[i] /= total;
  if (++L->epoch == H) L->epoch = 0;
  return out;
}

const float* lstm_byte_update(Lstm* L, const float* aux, int symbol) {
  const int V = L->V, HW = 2 * C + 1;
  // Lstm::SetInput (lstm.cpp:80-85)
  for (int l = 0; l < 2; ++l)
    memcpy(&L->layer_input[l][(size_t)L->epoch * L->in_size[l]], aux, V * sizeof(float));
  // Lstm::Perceive (lstm.cpp:87-118)
  const unsigned input = symbol;
  int last_epoch = L->epoch - 1;
  if (last_epoch == -1) last_epoch = H - 1;
  int old_input = L->input_history[last_epoch];
  L->input_history[last_epoch] = input;
  if (L->epoch == 0) {
    for (int ep = H - 1; ep >= 0; --ep) {
      for (int l = 1; l >= 0; --l) {
        int offset = l * C;
        const float* out = &L->output[(size_t)ep * V];
        const float* W = &L->out_w[(size_t)ep * V * HW];
        for (int i = 0; i < V; ++i) {
          float error = ((unsigned)i == L->input_history[ep]) ? (out[i] - 1) : out[i];
          for (int j = 0; j < C; ++j) L->hidden_error[j] += W[(size_t)i * HW + j + offset] * error;
        }
        int prev = ep - 1;
        if (prev == -1) prev = H - 1;
        int sym = L->input_history[prev];
        if (ep == 0) sym = old_input;
        L->layer[l].Backward(&L->layer_input[l][(size_t)ep * L->in_size[l]], ep, l, sym, L->hidden_error);
      }
    }
  }
  {
    const float* out = &L->output[(size_t)last_epoch * V];
    const float* Wl = &L->out_w[(size_t)last_epoch * V * HW];
    float* We = &L->out_w[(size_t)L->epoch * V * HW];
    for (int i = 0; i < V; ++i) {
      float error = ((unsigned)i == input) ? (out[i] - 1) : out[i];
      float k = kLr * error;
      for (int j = 0; j < HW; ++j) {
        float base = Wl[(size_t)i * HW + j];
        We[(size_t)i * HW + j] = base - k * L->hidden[j];
      }
    }
  }
  return Predict(L, input);
}

}  // namespace op

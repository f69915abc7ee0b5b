// oracle/port/predictor.cpp — TEST INFRASTRUCTURE.
// Orchestration of Predict/Perceive/Pretrain (reference src/predictor.cpp:361-487)
// and the three-layer gated mixer (reference src/mixer/mixer.cpp,
// src/mixer/mixer-input.cpp), SURVEY §8 rows a1-a6.
#include "internal.h"

#include <math.h>
#include <stdio.h>

using namespace op;

namespace {

enum { N_IN = OP_N_INPUTS, N_L0 = 26, N_L1 = 20, N_AUX = 3 };
const int kAuxIndex[3] = {433, 2024, 2077};   // predictor.cpp:80,87,191 (AddAuxiliary)

// Selector ids for the mixers (predictor.cpp:201-356).
enum Sel { S_ZERO, S_LONGBIT, S_RB0, S_RB1, S_RB2, S_RB3, S_LINEBREAK, S_LONGEST, S_WRT, S_AUX,
           S_IV0, S_IV1, S_IV2, S_IV3, S_IV4, S_IV6, S_IVH,
           S_BC0, S_BC1, S_BC2, S_BC3, S_BC_ALNUM, S_BC_W2, S_BC_W3, S_BC_RB1, S_COMB0, S_COMB1 };

struct MixSpec { Sel sel; float lr; };
const MixSpec kL0[N_L0] = {
    {S_BC0, 0.005}, {S_BC0, 0.0005}, {S_BC1, 0.005}, {S_BC1, 0.0005}, {S_BC2, 0.005}, {S_BC3, 0.002},
    {S_RB2, 0.002}, {S_RB3, 0.005}, {S_ZERO, 0.00005}, {S_LINEBREAK, 0.0007}, {S_LONGEST, 0.0005},
    {S_WRT, 0.002}, {S_AUX, 0.0005}, {S_IV0, 0.001}, {S_IV1, 0.001}, {S_IV2, 0.001},
    {S_BC_ALNUM, 0.005}, {S_IV3, 0.001}, {S_IV4, 0.001}, {S_BC_W2, 0.005}, {S_IV6, 0.001},
    {S_IVH, 0.001}, {S_BC_W3, 0.005}, {S_BC_RB1, 0.005}, {S_COMB0, 0.005}, {S_COMB1, 0.003}};
const MixSpec kL1[N_L1] = {
    {S_ZERO, 0.005}, {S_ZERO, 0.0005}, {S_LONGBIT, 0.005}, {S_LONGBIT, 0.0005}, {S_LONGBIT, 0.00001},
    {S_RB0, 0.005}, {S_RB1, 0.005}, {S_RB2, 0.005}, {S_LONGEST, 0.0005}, {S_WRT, 0.002},
    {S_IV0, 0.001}, {S_IV1, 0.001}, {S_IV2, 0.001}, {S_IV3, 0.001}, {S_IV4, 0.001}, {S_IV6, 0.001},
    {S_IVH, 0.001}, {S_BC_W2, 0.001}, {S_BC_ALNUM, 0.001}, {S_BC_W3, 0.001}};
const MixSpec kL2 = {S_ZERO, 0.0003};

u64 SelectorValue(const Contexts& c, Sel s) {
  switch (s) {
    case S_ZERO: return c.zero_context;
    case S_LONGBIT: return c.long_bit_context;
    case S_RB0: return c.recent_bytes[0];
    case S_RB1: return c.recent_bytes[1];
    case S_RB2: return c.recent_bytes[2];
    case S_RB3: return c.recent_bytes[3];
    case S_LINEBREAK: return c.line_break;
    case S_LONGEST: return c.longest_match;
    case S_WRT: return c.wrt_context;
    case S_AUX: return c.auxiliary_context;
    case S_IV0: return c.interval[0];
    case S_IV1: return c.interval[1];
    case S_IV2: return c.interval[2];
    case S_IV3: return c.interval[3];
    case S_IV4: return c.interval[4];
    case S_IV6: return c.interval[6];
    case S_IVH: return c.ivh_ctx;
    case S_BC0: return c.bitctx[0];
    case S_BC1: return c.bitctx[1];
    case S_BC2: return c.bitctx[2];
    case S_BC3: return c.bitctx[3];
    case S_BC_ALNUM: return c.bitctx[4];
    case S_BC_W2: return c.bitctx[5];
    case S_BC_W3: return c.bitctx[6];
    case S_BC_RB1: return c.bitctx[7];
    case S_COMB0: return c.combined[0];
    case S_COMB1: return c.combined[1];
  }
  return 0;
}

}  // namespace

namespace op {

WeightSet* MixerUnit::Select(u64 ctx64) {
  const u32 ctx = (u32)ctx64;            // key type is unsigned int (mixer.h:33)
  const size_t limit = 10000;
  u32 key = ctx;
  if (sets.size() >= limit && sets.find(ctx) == sets.end()) key = 0xDEADBEEF;
  std::unique_ptr<WeightSet>& slot = sets[key];
  if (!slot) {
    slot.reset(new WeightSet());
    slot->w.assign(n_in, 0.0f);
    slot->we.assign(n_extra, 0.0f);
  }
  return slot.get();
}

}  // namespace op

struct op_predictor {
  u8 vocab[256];
  int vocab_size = 0;
  int byte_map[256];
  GlibcRand rng;
  Contexts ctx;

  // small models, in models_ order (predictor.cpp:28-35)
  BracketModel bracket;
  DirectModel direct_bracket;                 // idx 1
  IndirectModel ind_bracket;                  // idx 2
  IndirectModel ind_word[18];                 // idx 2025..2042
  MatchModel match_word[6];                   // word matches
  IndirectModel ind_run;                      // idx 2045
  DirectModel dhash_word;                     // idx 2046
  DirectModel direct_o[3];                    // idx 2051..2053
  DirectModel dhash_o3;                       // idx 2054
  MatchModel match_o[10];                     // idx 2055..2064
  IndirectModel ind_double[11];               // idx 2065..2075
  ByteModelState ppmd;                        // idx 2076 (distribution replayed)
  ByteModelState lstm_bm;                     // idx 2077
  Lstm* lstm = nullptr;
  std::vector<float> lstm_aux;

  // mixer network
  float in0[N_IN];                            // stretched layer-0 inputs
  float in1[N_L0 + N_AUX];
  float in2[N_L0 + N_L1 + N_AUX];
  float stretched_min, stretched_max;
  MixerUnit l0[N_L0], l1[N_L1], l2;
  u64 sel_used[OP_N_MIXERS];
  Sse* sse = nullptr;

  float small_out[54];

  explicit op_predictor(const uint8_t* v) : rng(0xDEADBEEF) {
    for (int i = 0; i < 256; ++i) {
      vocab[i] = v[i] ? 1 : 0;
      byte_map[i] = vocab_size;
      if (vocab[i]) ++vocab_size;
    }
    const Tables& T = tables();
    stretched_min = T.Logit(0);
    stretched_max = T.Logit(1);
    // --- construction order == rand() order (predictor.cpp:28-36) ---
    bracket.Init();
    direct_bracket.Init(30, 0, 257 * 256, false);
    ind_bracket.Init(false, 300, rng);
    for (int i = 0; i < 18; ++i) ind_word[i].Init(false, 200, rng);
    for (int i = 0; i < 6; ++i) {
      match_word[i].Init(200, 0.5, 10000000);
      if (i == 1) { ind_run.Init(true, 200, rng); dhash_word.Init(30, 0, 500000, true); }
    }
    direct_o[0].Init(30, 0, 1, false);
    direct_o[1].Init(30, 0, 256, false);
    direct_o[2].Init(30, 0, 65536, false);
    dhash_o3.Init(30, 0, 100000, true);
    static const u64 msize[10] = {1, 256, 65536, 20000000, 20000000, 20000000, 20000000,
                                  20000000, 1048576, 20000000};
    for (int i = 0; i < 10; ++i) match_o[i].Init(200, 0.5, msize[i]);
    for (int i = 0; i < 11; ++i) ind_double[i].Init(false, 400, rng);
    for (int i = 0; i < 256; ++i) ppmd.probs[i] = lstm_bm.probs[i] = 1.0 / 256;
    lstm = lstm_create(vocab_size, rng);
    lstm_aux.assign(vocab_size, 0.0f);

    for (int i = 0; i < N_IN; ++i) in0[i] = 0.5f;   // MixerInput::SetNumModels (mixer-input.cpp:7-9)
    for (float& x : in1) x = 0.5f;
    for (float& x : in2) x = 0.5f;
    for (int i = 0; i < N_L0; ++i) { l0[i].lr = kL0[i].lr; l0[i].n_in = N_IN; l0[i].n_extra = i; l0[i].extra_snapshot.assign(i, 0); }
    for (int i = 0; i < N_L1; ++i) { l1[i].lr = kL1[i].lr; l1[i].n_in = N_L0 + N_AUX; l1[i].n_extra = i; l1[i].extra_snapshot.assign(i, 0); }
    l2.lr = kL2.lr; l2.n_in = N_L0 + N_L1 + N_AUX; l2.n_extra = 0;
    sse = sse_create();
    for (auto& s : sel_used) s = 0;
  }
  ~op_predictor() { lstm_destroy(lstm); sse_destroy(sse); }

  // word-model context for match_word[i] (predictor.cpp:117: {0},{1},{7},{1,3},{1,2,3},{7,2})
  u64 MatchWordCtx(int i) const {
    static const int idx[6] = {0, 4, 3, 7, 6, 2};
    return ctx.sparse[idx[i]];
  }
  u64 MatchOCtx(int i) const {
    static const int idx[10] = {0, 1, 2, 4, 5, 6, 7, 8, 9, 10};
    return ctx.chash[idx[i]];
  }

  // ----- per-model primitives sharing the manager state -----
  float IndPredict(IndirectModel& m) {
    m.map_index += ctx.bit_context;                                // indirect.cpp:17
    return m.pred[ctx.shared_map[m.map_index]];
  }
  void IndPerceive(IndirectModel& m, int bit) {
    const Tables& T = tables();
    int state = ctx.shared_map[m.map_index];
    m.pred[state] += (bit - m.pred[state]) * m.divisor;
    ctx.shared_map[m.map_index] = m.run ? T.runmap[state * 2 + bit] : T.nonstat[state][bit];
    m.map_index -= ctx.bit_context;
  }
  void IndByteUpdate(IndirectModel& m, u64 byte_ctx) {
    m.map_index = (257 * byte_ctx + m.map_offset) % (2048000000ull - 257);
  }
  float MatchPredict(const MatchModel& m) const {
    if (m.cur_byte & m.bit_pos) return m.pred[m.match_length];
    return 1 - m.pred[m.match_length];
  }
  void MatchPerceive(MatchModel& m, u64 byte_ctx, int bit) {
    int match = (bit == ((m.cur_byte & m.bit_pos) != 0)) ? 1 : 0;
    m.bit_pos /= 2;
    float d = m.divisor;
    if (m.count[m.match_length] < m.limit) {
      ++m.count[m.match_length];
      d = 1.0 / (m.count[m.match_length] + m.delta);
    }
    m.pred[m.match_length] += (match - m.pred[m.match_length]) * d;
    if (match) { if (m.match_length < 255) ++m.match_length; } else m.match_length = 0;
    if (ctx.bit_context >= 128) {
      m.map[byte_ctx % m.map.size()] = (u32)m.history_pos;
      ++m.history_pos;
    }
  }
  void MatchByteUpdate(MatchModel& m, u64 byte_ctx) {
    if (m.match_length < 8) m.cur_match = m.map[byte_ctx % m.map.size()];
    else ++m.cur_match;
    m.cur_match %= ctx.history.size();
    m.cur_byte = ctx.history[m.cur_match];
    m.bit_pos = 128;
    u64 mc = m.match_length / 32;
    if (mc > ctx.longest_match) ctx.longest_match = mc;
  }

  // The 54 small-model outputs in models_ order, skipping FXCM/PAQ8.
  void SmallPredict() {
    float* o = small_out;
    int k = 0;
    o[k++] = bracket.bm.Predict();
    o[k++] = direct_bracket.Predict(ctx.bracket_ctx, ctx.bit_context);
    o[k++] = IndPredict(ind_bracket);
    for (int i = 0; i < 18; ++i) o[k++] = IndPredict(ind_word[i]);
    for (int i = 0; i < 6; ++i) {
      o[k++] = MatchPredict(match_word[i]);
      if (i == 1) { o[k++] = IndPredict(ind_run); o[k++] = dhash_word.Predict(0, ctx.bit_context); }
    }
    for (int i = 0; i < 3; ++i) o[k++] = direct_o[i].Predict(ctx.chash[i], ctx.bit_context);
    o[k++] = dhash_o3.Predict(0, ctx.bit_context);
    for (int i = 0; i < 10; ++i) o[k++] = MatchPredict(match_o[i]);
    for (int i = 0; i < 11; ++i) o[k++] = IndPredict(ind_double[i]);
  }
  void SmallPerceive(int bit) {
    bracket.bm.Perceive(bit);
    direct_bracket.Perceive(ctx.bracket_ctx, ctx.bit_context, bit);
    IndPerceive(ind_bracket, bit);
    for (int i = 0; i < 18; ++i) IndPerceive(ind_word[i], bit);
    for (int i = 0; i < 6; ++i) {
      MatchPerceive(match_word[i], MatchWordCtx(i), bit);
      if (i == 1) { IndPerceive(ind_run, bit); dhash_word.Perceive(0, ctx.bit_context, bit); }
    }
    for (int i = 0; i < 3; ++i) direct_o[i].Perceive(ctx.chash[i], ctx.bit_context, bit);
    dhash_o3.Perceive(0, ctx.bit_context, bit);
    for (int i = 0; i < 10; ++i) MatchPerceive(match_o[i], MatchOCtx(i), bit);
    for (int i = 0; i < 11; ++i) IndPerceive(ind_double[i], bit);
  }
  void SmallByteUpdate() {
    bracket.ByteUpdate(ctx.bit_context, vocab);
    IndByteUpdate(ind_bracket, ctx.bracket_ctx);
    for (int i = 0; i < 18; ++i) IndByteUpdate(ind_word[i], ctx.sparse[i]);
    for (int i = 0; i < 6; ++i) {
      MatchByteUpdate(match_word[i], MatchWordCtx(i));
      if (i == 1) { IndByteUpdate(ind_run, ctx.sparse[4]); dhash_word.ByteUpdate(ctx.sparse[4]); }
    }
    dhash_o3.ByteUpdate(ctx.chash[3]);
    for (int i = 0; i < 10; ++i) MatchByteUpdate(match_o[i], MatchOCtx(i));
    for (int i = 0; i < 11; ++i) IndByteUpdate(ind_double[i], ctx.ihash[i].ctx);
  }

  // mixer-input.cpp:11-27
  void SetInput(int index, float p) {
    if (p < 1.0e-4f) p = 1.0e-4f; else if (p > 1 - 1.0e-4f) p = 1 - 1.0e-4f;
    in0[index] = tables().Logit(p);
  }
  float ClampStretched(float p) const {
    if (p > stretched_max) p = stretched_max; else if (p < stretched_min) p = stretched_min;
    return p;
  }

  // mixer.cpp:38-54
  float Mix(MixerUnit& m, u64 sel, const float* in, const std::vector<float>& extras) {
    WeightSet* d = m.Select(sel);
    float p = 0;
    for (int i = 0; i < m.n_in; ++i) p += in[i] * d->w[i];
    m.p = p;
    for (int i = 0; i < m.n_extra; ++i) m.extra_snapshot[i] = extras[i];
    float e = 0;
    for (int i = 0; i < m.n_extra; ++i) e += m.extra_snapshot[i] * d->we[i];
    m.p += e;
    return m.p;
  }
  // mixer.cpp:56-72
  void Train(MixerUnit& m, u64 sel, const float* in, int bit) {
    WeightSet* d = m.Select(sel);
    float decay = 0.9 / pow(0.0000001 * m.steps + 0.8, 0.8);
    decay *= 1.5 - ((1.0 * d->steps) / m.max_steps);
    float update = decay * m.lr * (Logistic(m.p) - bit);
    ++m.steps;
    ++d->steps;
    if (d->steps > m.max_steps) m.max_steps = d->steps;
    for (int i = 0; i < m.n_in; ++i) d->w[i] -= update * in[i];
    for (int i = 0; i < m.n_extra; ++i) d->we[i] -= update * m.extra_snapshot[i];
    if ((d->steps & 1023) == 0) {
      for (int i = 0; i < m.n_in; ++i) d->w[i] *= 1.0f - 3.0e-6f;
      for (int i = 0; i < m.n_extra; ++i) d->we[i] *= 1.0f - 3.0e-6f;
    }
  }

  float Predict(const uint16_t* ext) {
    const float cf = 1.0 / 4095;                       // paq8.cpp:499, fxcmv1.cpp:99
    SmallPredict();
    for (int i = 0; i < 3; ++i) SetInput(i, small_out[i]);
    for (int i = 0; i < OP_N_EXT; ++i) SetInput(3 + i, ext[i] == 0xFFFF ? 0.5f : ext[i] * cf);
    for (int i = 3; i < 54; ++i) SetInput(2025 - 3 + i, small_out[i]);
    SetInput(2076, ppmd.Predict());
    float override_p = -1;
    {
      float p = lstm_bm.Predict();
      if (p == 0 || p == 1) override_p = p;
      SetInput(2077, p);
    }
    float avg = 0;
    for (int i = 0; i < N_AUX; ++i) avg += Logistic(in0[kAuxIndex[i]]);
    avg /= N_AUX;
    ctx.auxiliary_context = avg * 15;

    std::vector<float> extras;
    for (int i = 0; i < N_L0; ++i) {
      u64 s = SelectorValue(ctx, kL0[i].sel);
      sel_used[i] = s;
      float p = Mix(l0[i], s, in0, extras);
      extras.push_back(ClampStretched(p));
      in1[i] = in2[i] = ClampStretched(p);
    }
    extras.clear();
    for (int i = 0; i < N_AUX; ++i) {
      float p = in0[kAuxIndex[i]];
      in1[N_L0 + i] = ClampStretched(p);
      in2[N_L0 + N_L1 + i] = ClampStretched(p);
    }
    for (int i = 0; i < N_L1; ++i) {
      u64 s = SelectorValue(ctx, kL1[i].sel);
      sel_used[N_L0 + i] = s;
      float p = Mix(l1[i], s, in1, extras);
      extras.push_back(ClampStretched(p));
      in2[N_L0 + i] = ClampStretched(p);
    }
    extras.clear();
    u64 s2 = SelectorValue(ctx, kL2.sel);
    sel_used[N_L0 + N_L1] = s2;
    float p = Logistic(Mix(l2, s2, in2, extras));
    p = sse_predict(sse, p);
    if (override_p >= 0) return override_p;
    return p;
  }

  void Perceive(int bit, const float* ppmd_after_byte) {
    SmallPerceive(bit);
    ppmd.Perceive(bit);
    lstm_bm.Perceive(bit);
    // Mixer::GetContextData() is re-evaluated with the still-unchanged contexts.
    for (int i = 0; i < N_L0; ++i) Train(l0[i], SelectorValue(ctx, kL0[i].sel), in0, bit);
    for (int i = 0; i < N_L1; ++i) Train(l1[i], SelectorValue(ctx, kL1[i].sel), in1, bit);
    Train(l2, SelectorValue(ctx, kL2.sel), in2, bit);
    sse_perceive(sse, bit);
    bool byte_update = ctx.bit_context >= 128;
    ctx.Update(bit);
    if (byte_update) {
      SmallByteUpdate();
      // PPMD::ByteUpdate: distribution replayed (already floor-1/vocab-masked/normalised)
      for (int i = 0; i < 256; ++i) ppmd.probs[i] = ppmd_after_byte[i];
      ppmd.top = 255; ppmd.bot = 0;
      // ByteMixer::SetInput x256 + ByteUpdate (byte-mixer.cpp:15-38)
      int k = 0;
      for (int i = 0; i < 256; ++i) if (vocab[i]) lstm_aux[k++] = 0.0f + ppmd.probs[i];
      for (int i = 0; i < vocab_size; ++i) lstm_aux[i] *= 2;   // 2 / num_models_ (=1), integer division
      const float* out = lstm_byte_update(lstm, lstm_aux.data(), byte_map[ctx.bit_context]);
      k = 0;
      for (int i = 0; i < 256; ++i) lstm_bm.probs[i] = vocab[i] ? out[k++] : 0;
      lstm_bm.ByteUpdate(vocab);
      ctx.bit_context = 1;
    }
  }

  void Pretrain(int bit) {
    SmallPredict();
    SmallPerceive(bit);
    bool byte_update = ctx.bit_context >= 128;
    ctx.Update(bit);
    if (byte_update) { SmallByteUpdate(); ctx.bit_context = 1; }
  }
};

extern "C" {

op_predictor* op_create(const uint8_t vocab[256]) { return new op_predictor(vocab); }
void op_destroy(op_predictor* p) { delete p; }
float op_predict(op_predictor* p, const uint16_t* ext) { return p->Predict(ext); }
void op_perceive(op_predictor* p, int bit, const float* ppmd) { p->Perceive(bit, ppmd); }
void op_pretrain(op_predictor* p, int bit) { p->Pretrain(bit); }
void op_get_inputs(const op_predictor* p, float* out) { memcpy(out, p->in0, sizeof(p->in0)); }
void op_get_mixer_outputs(const op_predictor* p, float* out) {
  int k = 0;
  for (int i = 0; i < N_L0; ++i) out[k++] = p->l0[i].p;
  for (int i = 0; i < N_L1; ++i) out[k++] = p->l1[i].p;
  out[k++] = p->l2.p;
}
void op_get_mixer_contexts(const op_predictor* p, uint32_t* out) {
  for (int i = 0; i < OP_N_MIXERS; ++i) out[i] = (uint32_t)p->sel_used[i];
}
void op_get_lstm_probs(const op_predictor* p, float* out) { memcpy(out, p->lstm_bm.probs, 256 * sizeof(float)); }

void op_run(op_predictor* p, const uint8_t* stream, size_t n_bytes, const uint16_t* ext,
            const float* ppmd, float* p_out) {
  size_t t = 0;
  for (size_t pos = 0; pos < n_bytes; ++pos) {
    for (int j = 7; j >= 0; --j, ++t) {
      int bit = (stream[pos] >> j) & 1;
      p_out[t] = p->Predict(ext + t * OP_N_EXT);
      p->Perceive(bit, ppmd + pos * 256);
    }
  }
}

float op_libm_expf(float x) { return expf(x); }
float op_libm_tanhf(float x) { return tanhf(x); }
float op_logistic(float x) { return op::Logistic(x); }

}  // extern "C"

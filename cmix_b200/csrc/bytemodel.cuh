// cmix_b200/csrc/bytemodel.cuh — ByteModel::Predict, shared by the small models, the LSTM read-out and the FXCM feedback.
#pragma once
#include "exact_math.h"

namespace cmixb200 {

// ByteModel::Predict (byte-model.cpp:8-24): sequential range sums, first-max argmax.
__device__ inline float bytemodel_predict(const float* probs, int bot, int top, int* ex_out) {
  const int m = bot + ((top - bot) / 2);
  float num = 0.0f;
#pragma unroll 8
  for (int i = m + 1; i <= top; ++i) num = XM_FADD(num, probs[i]);
  float denom = num;
#pragma unroll 8
  for (int i = bot; i <= m; ++i) denom = XM_FADD(denom, probs[i]);
  int ex = bot; float best = probs[bot];
  for (int i = bot + 1; i <= top; ++i) if (probs[i] > best) { best = probs[i]; ex = i; }
  if (ex_out) *ex_out = ex;
  if (denom == 0) return 0.5f;
  return XM_FDIV(num, denom);
}

}  // namespace cmixb200

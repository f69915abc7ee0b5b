// cmix_b200/shim/predictor.h — the reference's `class Predictor` surface (reference
// src/predictor.h:17-22) re-declared on top of the C-ABI in include/cmixb200.h.
//
// Drop-in: compile the reference's runner.cpp, coder/*.cpp and preprocess/*.cpp UNCHANGED with
//   g++ -include cmix_b200/shim/predictor.h -I<reference>/src ...
// This header uses the reference header's own include guard, so when it is force-included first
// every `#include "predictor.h"` / `#include "../predictor.h"` in the reference sources becomes a
// no-op and they bind to this class instead (see INTEGRATION.md and cmix_b200/shim/Makefile).
#ifndef PREDICTOR_H
#define PREDICTOR_H

// the reference's predictor.h transitively provides these to runner/coder/preprocess
#include <cstdint>
#include <memory>
#include <set>
#include <valarray>
#include <vector>

struct cmixb200_predictor;

class Predictor {
 public:
  Predictor(const std::vector<bool>& vocab);   // predictor.h:19
  ~Predictor();
  float Predict();                             // predictor.h:20
  void Perceive(int bit);                      // predictor.h:21
  void Pretrain(int bit);                      // predictor.h:22

 private:
  Predictor(const Predictor&);
  Predictor& operator=(const Predictor&);
  void FlushPretrain();
  cmixb200_predictor* impl_;
  // Pretrain() arrives bit by bit (preprocessor.cpp:52,66) but only before the first Predict():
  // whole bytes are buffered and trained through the bulk entry point, one launch set per 64 KiB.
  std::vector<unsigned char> pre_bytes_;
  unsigned pre_acc_, pre_nbits_;
};

#endif

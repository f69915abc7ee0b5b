// oracle/port/internal.h — TEST INFRASTRUCTURE. Internal types of the CPU port.
#ifndef ORACLE_PORT_INTERNAL_H
#define ORACLE_PORT_INTERNAL_H

#include "port.h"

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <unordered_map>
#include <vector>

namespace op {

typedef uint64_t u64;
typedef uint32_t u32;
typedef uint8_t u8;

// ---------------------------------------------------------------- tables ---
struct Tables {
  std::vector<float> logit;        // sigmoid.cpp:5-10, 100001 entries
  u8 nonstat[256][2];              // states/nonstationary.cpp:3
  u8 runmap[512];                  // states/run-map.cpp:3-20
  Tables();
  float Logit(float p) const;      // sigmoid.cpp:12-17
};
const Tables& tables();
float Logistic(float x);           // sigmoid.cpp:19-21
// glibc rand() stream private to one predictor (predictor.cpp:26 srand(0xDEADBEEF)).
struct GlibcRand {
  char statebuf[128];
  struct random_data rd;
  explicit GlibcRand(unsigned seed);
  int next();
};

// -------------------------------------------------------------- contexts ---
// context-manager.cpp + contexts/*.cpp, flattened.
struct Contexts {
  // manager scalars (context-manager.h:23-27)
  u32 bit_context = 1, wrt_state = 0;
  u64 long_bit_context = 1, zero_context = 0, history_pos = 0, line_break = 0,
      longest_match = 0, auxiliary_context = 0, wrt_context = 0;
  std::vector<u8> history;   // 100 000 000
  u8* shared_map;            // 2 048 000 000 (calloc: pages touched lazily)
  u64 words[8], recent_bytes[8];

  // BracketContext(256, 15)  (contexts/bracket-context.cpp)
  std::vector<u32> br_active, br_distance;
  u64 bracket_ctx = 0;
  // Sparse word contexts (contexts/sparse.cpp), 18 of them
  u64 sparse[18];
  // ContextHash (order, bits): (0,8) (1,8) (2,8) (3,8) (7,4) (11,3) (13,2) (15,2) (17,2) (20,1) (25,1) (2,4) (3,2)
  u64 chash[13];
  // IndirectHash x11 (contexts/indirect-hash.cpp)
  struct IH { u64 ctx1 = 0, ctx = 0; u32 size1; u64 size; u32 h1, h2; std::vector<u64> hashes; };
  IH ihash[11];
  // Interval contexts: map1/8, map2/8, alnum/7, wrt2b/10, wrt2b/15, wrt2b/7, wrt3b/9, wrt3b/7
  u64 interval[8];
  // IntervalHash(wrt3b, 8, 7, 2)
  u64 ivh_ctx = 0; u32 ivh_interval = 0;
  // BitContexts (contexts/bit-context.cpp): over chash(0,8),(1,8),(2,4),(3,2), alnum/7, wrt2b/7, wrt3b/7, recent_bytes[1]
  u64 bitctx[8];
  // CombinedContext x2
  u64 combined[2];

  Contexts();
  ~Contexts();
  void Update(int bit);       // context-manager.cpp:69-94
};

// ---------------------------------------------------------- small models ---
struct ByteModelState {        // models/byte-model.cpp
  int top = 255, mid = 0, bot = 0, ex = 0;
  float probs[256];
  float Predict();             // :8-24
  void Perceive(int bit);      // :30-37
  void ByteUpdate(const u8* vocab);  // :39-45
};

struct DirectModel {           // models/direct.cpp, models/direct-hash.cpp
  int limit; float delta, divisor;
  u64 rows; bool hashed; u64 index = 0;
  std::vector<float> pred;     // rows*256
  std::vector<u8> count;
  std::vector<u64> checksum;
  void Init(int limit, float delta, u64 rows, bool hashed);
  float Predict(u64 byte_ctx, u32 bit_ctx) const;
  void Perceive(u64 byte_ctx, u32 bit_ctx, int bit);
  void ByteUpdate(u64 byte_ctx);
};

struct IndirectModel {         // models/indirect.cpp
  u64 map_index = 0, map_offset = 0; float divisor; bool run;
  float pred[256];
  void Init(bool run_map, float delta, GlibcRand& rng);
};

struct MatchModel {            // models/match.cpp
  u64 history_pos = 0, cur_match = 0;
  u8 cur_byte = 0, bit_pos = 128, match_length = 0;
  int limit; float delta, divisor;
  std::vector<u32> map;
  float pred[256]; int count[256];
  void Init(int limit, float delta, u64 map_size);
};

struct BracketModel {          // models/bracket.cpp
  ByteModelState bm;
  std::vector<u32> active, distance;
  std::vector<u32> first, second;   // stats_[256][200]
  void Init();
  void ByteUpdate(u32 byte, const u8* vocab);
};

// ------------------------------------------------------------- mixer net ---
struct WeightSet { u64 steps = 0; std::vector<float> w, we; };
struct MixerUnit {             // mixer/mixer.cpp
  float lr; int n_in, n_extra;
  float p = 0.5f;
  u64 max_steps = 1, steps = 0;
  std::vector<float> extra_snapshot;
  std::unordered_map<u32, std::unique_ptr<WeightSet>> sets;
  WeightSet* Select(u64 ctx);  // :16-36
};

struct Sse;                    // sse.cpp
Sse* sse_create();
void sse_destroy(Sse*);
float sse_predict(Sse*, float p);   // mixer/sse.cpp:320-324
void sse_perceive(Sse*, int bit);   // :326-328

struct Lstm;                   // lstm.cpp
Lstm* lstm_create(int vocab_size, GlibcRand& rng);
void lstm_destroy(Lstm*);
// ByteMixer::ByteUpdate (byte-mixer.cpp:22-38) minus the 256<->vocab mapping:
// `aux` = 2*PPMD probs over the vocab; returns V probabilities.
const float* lstm_byte_update(Lstm*, const float* aux, int symbol);

}  // namespace op
#endif

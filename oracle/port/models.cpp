// oracle/port/models.cpp — TEST INFRASTRUCTURE.
// Tables, shared context state and the small cmix bit models
// (SURVEY §8 rows a3, a8, a16, a17, a18), restated as flat structs.
#include "internal.h"

#include <math.h>
#include <stdio.h>

namespace op {

// ------------------------------------------------------------------ tables
static int hexval(char c) { return c <= '9' ? c - '0' : c - 'a' + 10; }

// states/nonstationary.cpp:3 — the 256x2 next-state table, as hex bytes
// (state-major, [state][bit]).
static const char kNonstatHex[] =
    "020c90050705190e3605101d1a8ca9051018540410250a620ad42d25c9001cca"
    "2017b78806d9bc2608251545083d380d16251b42558c23296bc7085b66d96411"
    "68175e05602f2bb02e3016276b331833fbf8190136223594463a5617163d4617"
    "243f3d3f07053e415ad84369442296056049462f484a494b2e3f4a311459244f"
    "148a3e51aa0e5245542215e960480ad42c3b2c3b3a4c3a4c4b3d3c3116274d59"
    "13413e5b5ad996b95e225c7f6017a7cf92934d635d660b65616967f968221341"
    "6a397a7b1627306d00cb0b6f03536c70722219f57422132771e63077a6bc3e79"
    "8c60751b7c04c2e72a228d7fa4b33f81ebcc6b837d7f6ed98604848512f5c90e"
    "3eaf308b80b70b97891e23cc87049004af958ed93eb73033c901bdab91855fb2"
    "962198995fc3a7b3234240ac9a096b9b9c9d1bef9e09269fa0a143c2a221508f"
    "a3b9c9f9a609a53400e7a88c6ebc64cfadb0b121aebc150097c0b5451b9afc8a"
    "4221bd09e48f1f28be5ab8b9bc793fbbba948d34c15380bf72b6cddcc53480c3"
    "92d4c43464d782c7c685ecefcee9fccb64cb034292fe5bcfb9c2e501d1e65fd3"
    "6411d6d8641182d740e1c9ccdaef92dba6e6f901deccc8df8601a7d7e2ae83e3"
    "ddf740ede84557e71341eedc230cc811b4ad43d91c65c5d5f0f1fcfdf2dc6ca4"
    "f3f4c9a4f62978d5c9349a017e377376737619d826c3a7d7148a400f1c580e11";

Tables::Tables() : logit(100001) {
  const int n = 100001;
  for (int i = 0; i < n; ++i) {
    float p = (i + 0.5f) / n;           // sigmoid.cpp:8
    logit[i] = log(p / (1 - p));        // sigmoid.cpp:23-25 (float -> logf)
  }
  for (int s = 0; s < 256; ++s)
    for (int b = 0; b < 2; ++b) {
      const char* h = kNonstatHex + 2 * (2 * s + b);
      nonstat[s][b] = (u8)(hexval(h[0]) * 16 + hexval(h[1]));
    }
  for (int i = 0; i < 512; ++i) {       // run-map.cpp:3-20
    int state = i / 2;
    if (i % 2 == 0) {
      if (state < 127) ++state; else if (state >= 128) state = 0;
    } else {
      if (state < 128) state = 128; else if (state < 255) ++state;
    }
    runmap[i] = (u8)state;
  }
}

float Tables::Logit(float p) const {
  int index = p * 100001;
  if (index >= 100001) index = 100000; else if (index < 0) index = 0;
  return logit[index];
}

const Tables& tables() { static Tables t; return t; }

float Logistic(float x) { return 1 / (1 + exp(-x)); }

GlibcRand::GlibcRand(unsigned seed) {
  memset(&rd, 0, sizeof(rd));
  memset(statebuf, 0, sizeof(statebuf));
  initstate_r(seed, statebuf, sizeof(statebuf), &rd);   // TYPE_3, same as rand()
}
int GlibcRand::next() { int32_t r; random_r(&rd, &r); return r; }

// ---------------------------------------------------------------- contexts
namespace {

// predictor.cpp:223-304 interval maps.
int IntervalMap(int which, int c) {
  static const int t1[] = {1, 32, 64, 128, 255, 142, 138, 140, 137, 97};
  static const int t2[] = {41, 92, 124, 58, 11, 46, 36, 47, 64, 4, 61, 97, 125, 45, 48};
  static const char m4[] =   // predictor.cpp:255-271, bytes 0..95; 96..207 -> 1, 208.. -> 0
      "2313301233001333" "3333333333303333" "3202132133332302"
      "1111111111322322" "2200231212222200" "2222222230232023";
  static const char m6[] =   // predictor.cpp:285-301, bytes 0..95; 96..127 -> 5, 128..207 -> 6, 208.. -> 7
      "0020560602043000" "0000000000000000" "2414474737223531"
      "1111111111053355" "0557501545006071" "3374557022544746";
  int v = 0;
  switch (which) {
    case 0: for (int t : t1) v += c < t; return v;
    case 1: for (int t : t2) v += c < t; return v;
    case 2: return ((c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || (c >= '0' && c <= '9') || c >= 0x80) ? 1 : 0;
    case 3: return c < 96 ? m4[c] - '0' : (c < 208 ? 1 : 0);
    default: return c < 96 ? m6[c] - '0' : (c < 128 ? 5 : (c < 208 ? 6 : 7));
  }
}

struct IntervalSpec { int map, bits, shift; };
const IntervalSpec kIntervals[8] = {
    {0, 8, 4}, {1, 8, 4}, {2, 7, 1}, {3, 10, 2}, {3, 15, 2}, {3, 7, 2}, {4, 9, 3}, {4, 7, 3}};

struct CHashSpec { int order, bits; };
const CHashSpec kCHash[13] = {{0, 8}, {1, 8}, {2, 8}, {3, 8}, {7, 4}, {11, 3}, {13, 2},
                              {15, 2}, {17, 2}, {20, 1}, {25, 1}, {2, 4}, {3, 2}};

const int kIHashSpec[11][4] = {{1, 8, 1, 8}, {2, 8, 1, 8}, {1, 8, 2, 8}, {2, 8, 2, 8},
                               {1, 8, 3, 8}, {3, 8, 1, 8}, {4, 6, 4, 8}, {5, 5, 5, 5},
                               {1, 8, 4, 8}, {1, 8, 5, 6}, {6, 4, 6, 4}};

// predictor.cpp:104-108 word-model orders
const int kSparseOrders[18][7] = {{1, 0}, {2, 0, 1}, {2, 7, 2}, {1, 7}, {1, 1}, {2, 1, 2},
    {3, 1, 2, 3}, {2, 1, 3}, {2, 1, 4}, {2, 1, 5}, {2, 2, 3}, {2, 3, 4}, {3, 1, 2, 4},
    {4, 1, 2, 3, 4}, {3, 2, 3, 4}, {1, 2}, {5, 1, 2, 3, 4, 5}, {6, 1, 2, 3, 4, 5, 6}};

int closing_bracket(u32 c, bool quotes) {   // bracket-context.cpp:6, bracket.cpp:10-11
  switch (c) {
    case '(': return ')'; case '{': return '}'; case '[': return ']'; case '<': return '>';
    case '\'': return quotes ? '\'' : -1; case '"': return quotes ? '"' : -1;
  }
  return -1;
}

}  // namespace

Contexts::Contexts() : history(100000000, 0) {
  shared_map = (u8*)calloc(256ull * 8000000ull, 1);
  if (!shared_map) { fprintf(stderr, "oracle port: cannot allocate shared map\n"); abort(); }
  for (int i = 0; i < 8; ++i) words[i] = recent_bytes[i] = 0;
  for (auto& s : sparse) s = 0;
  for (auto& c : chash) c = 0;
  for (auto& c : interval) c = 0;
  for (auto& c : bitctx) c = 0;
  combined[0] = combined[1] = 0;
  for (int i = 0; i < 11; ++i) {
    IH& h = ihash[i];
    h.h1 = kIHashSpec[i][1]; h.h2 = kIHashSpec[i][3];
    h.size1 = (u32)(1ull << (kIHashSpec[i][1] * kIHashSpec[i][0]));
    h.size = 1ull << (kIHashSpec[i][3] * kIHashSpec[i][2]);
    h.hashes.assign(h.size1, 0);
  }
}
Contexts::~Contexts() { free(shared_map); }

void Contexts::Update(int bit) {
  bit_context += bit_context + bit;
  long_bit_context = bit_context;
  if (bit_context >= 256) {
    bit_context -= 256;
    long_bit_context = 1;
    longest_match = 0;
    const u32 c = bit_context;
    if (c == '\n') line_break = 0; else if (line_break < 99) ++line_break;
    // UpdateHistory (context-manager.cpp:23-27)
    history[history_pos] = (u8)c;
    if (++history_pos == history.size()) history_pos = 0;
    // UpdateWords (:29-48)
    {
      u8 w = (u8)c;
      if ((w >= 'a' && w <= 'z') || (w >= 'A' && w <= 'Z') || w >= 0x80) words[7] = words[7] * 997 * 16 + w;
      else words[7] = 0;
      if (w >= 'A' && w <= 'Z') w += 'a' - 'A';
      if ((w >= 'a' && w <= 'z') || (w >= '0' && w <= '9') || w == 8 || w == 6 || w >= 0x80) {
        words[0] = words[0] * 997 * 16 + w;
        words[0] &= 0xfffffff;
        words[1] = words[1] * 263 * 32 + w;
      } else {
        for (int i = 6; i >= 2; --i) words[i] = words[i - 1];
        words[1] = 0;
      }
    }
    // UpdateRecentBytes (:50-55)
    for (int i = 7; i >= 1; --i) recent_bytes[i] = recent_bytes[i - 1];
    recent_bytes[0] = c;
    // UpdateWRTContext (:57-67)
    if (c < 0x80) {
      wrt_state = 0;
    } else {
      if (wrt_state == 0) wrt_context = 0;
      wrt_state = 1;
      wrt_context <<= 8;
      wrt_context += c;
      if (wrt_context > 0xFFEFCF) wrt_context = 0;
    }
    // --- every byte-level Context::Update() ---
    // BracketContext (bracket-context.cpp:11-35); stack never trimmed (quirk 15)
    if (!br_active.empty()) {
      if (closing_bracket(br_active.back(), false) == (int)c || br_distance.back() >= 256 - 1) {
        br_active.pop_back(); br_distance.pop_back();
      } else {
        ++br_distance.back();
      }
    }
    if (closing_bracket(c, false) >= 0) { br_active.push_back(c); br_distance.push_back(0); }
    bracket_ctx = br_active.empty() ? 0 : 256ull * (br_active.back() + 1) + br_distance.back();
    // Sparse (sparse.cpp:17-22)
    static const u64 factors[6] = {1, 256, 29 * 31, 29 * 31 * 37, 29 * 31 * 37 * 41, 29ull * 31 * 37 * 41 * 43};
    for (int i = 0; i < 18; ++i) {
      const int* o = kSparseOrders[i];
      u64 v = words[o[1]];
      for (int k = 1; k < o[0]; ++k) v += (u64)(u32)factors[k] * words[o[1 + k]];
      sparse[i] = v;
    }
    // ContextHash (context-hash.cpp:9-11)
    for (int i = 0; i < 13; ++i) {
      u64 size = 1ull << (kCHash[i].bits * kCHash[i].order);
      chash[i] = (chash[i] * (u64)(1 << kCHash[i].bits) + c) % size;
    }
    // IndirectHash (indirect-hash.cpp:13-17)
    for (int i = 0; i < 11; ++i) {
      IH& h = ihash[i];
      h.hashes[h.ctx1] = (h.ctx * (u64)(1 << h.h2) + c) % h.size;
      h.ctx1 = (h.ctx1 * (u64)(1 << h.h1) + c) % h.size1;
      h.ctx = h.hashes[h.ctx1];
    }
    // Interval (interval.cpp:17-19)
    for (int i = 0; i < 8; ++i) {
      const IntervalSpec& s = kIntervals[i];
      u64 mask = (1ull << s.bits) - 1;
      interval[i] = mask & ((interval[i] << s.shift) + IntervalMap(s.map, c));
    }
    // IntervalHash(wrt3b, 8, 7, 2) (interval-hash.cpp:18-21)
    ivh_interval = 255 & ((ivh_interval << 3) + IntervalMap(4, c));
    ivh_ctx = (ivh_ctx * 4 + ivh_interval) % 16384;
    // CombinedContext (combined-context.cpp:13-15)
    combined[0] = (recent_bytes[0] << 8) + recent_bytes[1];
    combined[1] = (recent_bytes[1] << 8) + recent_bytes[2];
  }
  // BitContext::Update every bit (bit-context.cpp:11-13)
  const u64 byte_ctx[8] = {chash[0], chash[1], chash[11], chash[12], interval[2],
                           interval[5], interval[7], recent_bytes[1]};
  for (int i = 0; i < 8; ++i) bitctx[i] = (byte_ctx[i] << 8) + long_bit_context;
}

// Sparse factor note: sparse.cpp stores factors_ as `unsigned int`, so
// 29*31*37*41*43 (=58 642 669... fits in 32 bits) is used as a u32 value.

// ------------------------------------------------------------- ByteModel
float ByteModelState::Predict() {
  int m = bot + ((top - bot) / 2);
  float num = 0.0f;
  for (int i = m + 1; i <= top; ++i) num += probs[i];
  float denom = num;
  for (int i = bot; i <= m; ++i) denom += probs[i];
  ex = bot;
  float best = probs[bot];
  for (int i = bot + 1; i <= top; ++i)
    if (probs[i] > best) { best = probs[i]; ex = i; }
  if (denom == 0) return 0.5f;
  return num / denom;
}
void ByteModelState::Perceive(int bit) {
  mid = bot + ((top - bot) / 2);
  if (bit) bot = mid + 1; else top = mid;
}
void ByteModelState::ByteUpdate(const u8* vocab) {
  top = 255; bot = 0;
  for (int i = 0; i < 256; ++i) if (!vocab[i]) probs[i] = 0;
}

// ---------------------------------------------------------------- Direct
void DirectModel::Init(int limit_, float delta_, u64 rows_, bool hashed_) {
  limit = limit_; delta = delta_; divisor = 1.0 / (limit_ + delta_);
  rows = rows_; hashed = hashed_;
  pred.assign(rows * 256, 0.5f);
  count.assign(rows * 256, 0);
  if (hashed) checksum.assign(rows, 0);
}
float DirectModel::Predict(u64 byte_ctx, u32 bit_ctx) const {
  u64 row = hashed ? index : byte_ctx;
  return pred[row * 256 + bit_ctx];
}
void DirectModel::Perceive(u64 byte_ctx, u32 bit_ctx, int bit) {
  u64 i = (hashed ? index : byte_ctx) * 256 + bit_ctx;
  float d = divisor;
  if (count[i] < limit) {
    ++count[i];
    d = 1.0 / (count[i] + delta);
  }
  pred[i] += (bit - pred[i]) * d;
}
void DirectModel::ByteUpdate(u64 byte_ctx) {   // direct-hash.cpp:31-48
  if (!hashed) return;
  index = byte_ctx % rows;
  for (int i = 0; i < 20; ++i) {
    if (checksum[index] == 0) { checksum[index] = byte_ctx; break; }
    if (checksum[index] == byte_ctx) break;
    if (i == 19) {
      for (int k = 0; k < 256; ++k) { pred[index * 256 + k] = 0.5f; count[index * 256 + k] = 0; }
      checksum[index] = byte_ctx;
      break;
    }
    if (++index == rows) index = 0;
  }
}

// -------------------------------------------------------------- Indirect
void IndirectModel::Init(bool run_map, float delta, GlibcRand& rng) {
  run = run_map;
  divisor = 1.0 / delta;
  map_offset = (u64)rng.next() % (2048000000ull - 257);   // indirect.cpp:10
  for (int i = 0; i < 256; ++i) {
    if (!run) pred[i] = 0.5f;
    else pred[i] = i < 128 ? (128.0 - i) / 256 : i / 256.0;  // run-map.cpp:13-16
  }
}

// ----------------------------------------------------------------- Match
void MatchModel::Init(int limit_, float delta_, u64 map_size) {
  limit = limit_; delta = delta_; divisor = 1.0 / (limit_ + delta_);
  map.assign(map_size, 0);
  for (int i = 0; i < 256; ++i) { pred[i] = 0.5 + (i + 0.5) / 512; count[i] = 0; }
}

// --------------------------------------------------------------- Bracket
void BracketModel::Init() {
  first.assign(256 * 200, 1);
  second.assign(256 * 200, 256);
  for (int i = 0; i < 256; ++i) bm.probs[i] = 1.0 / 256;
}
void BracketModel::ByteUpdate(u32 byte, const u8* vocab) {   // bracket.cpp:13-60
  const u32 kDistLimit = 200, kStackLimit = 10, kStatsLimit = 100000;
  float* probs = bm.probs;
  for (int i = 0; i < 256; ++i) probs[i] = 1. / 256;
  auto fill = [&](float p, int hot) {
    float rest = (1 - p) / 255;
    for (int i = 0; i < 256; ++i) probs[i] = rest;
    probs[hot] = p;
  };
  int close = closing_bracket(byte, true);
  if (active.empty() || (close >= 0 && !(active.back() == byte && (u32)close == byte))) {
    if (close >= 0) {
      active.push_back(byte); distance.push_back(0);
      if (active.size() > kStackLimit) { active.erase(active.begin()); distance.erase(distance.begin()); }
      float p = (1. * first[byte * 200]) / second[byte * 200];
      fill(p, close);
    }
  } else {
    u32 a = active.back(), d = distance.back();
    ++second[a * 200 + d];
    if (closing_bracket(a, true) == (int)byte) ++first[a * 200 + d];
    if (second[a * 200 + d] > kStatsLimit) { first[a * 200 + d] /= 2; second[a * 200 + d] /= 2; }
    if (closing_bracket(a, true) == (int)byte || d >= kDistLimit - 1) {
      active.pop_back(); distance.pop_back();
      if (!active.empty()) {
        u32 a2 = active.back(), d2 = distance.back();
        float p = (1. * first[a2 * 200 + d2]) / second[a2 * 200 + d2];
        fill(p, closing_bracket(a2, true));
      }
    } else {
      ++distance.back(); ++d;
      float p = (1. * first[a * 200 + d]) / second[a * 200 + d];
      fill(p, closing_bracket(a, true));
    }
  }
  bm.ByteUpdate(vocab);
}

}  // namespace op

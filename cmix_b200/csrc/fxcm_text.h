// cmix_b200/csrc/fxcm_text.h — byte-level text analysis of the resident FXCM model.
//
// FXCM (reference src/models/fxcmv1.cpp) derives ~80 hashed contexts per input byte from a hand-written
// wiki/XML-aware text analyser: bracket/quote/first-char stacks (fxcmv1.cpp:1932-2020), a row/column and
// wiki-table tracker (:2022-2170), sentence/paragraph word lists (:2180-2300), an English affix stemmer with
// word-type tags (:2302-3205) and the WRT code-word decoder (:378-460). All of it is a pure function of the
// coded byte stream (plus the dictionary), runs ONCE PER BYTE and is scalar by nature; it is restated here as
// integer-only host/device code over one flat state block so that one lane of the FXCM CTA (fxcm.cuh) can run
// it, ahead of the bit loop. The CPU build of the same code is pinned bit-for-bit against the unmodified
// reference through the exported 12-bit codes (tools/fxcm_check.cpp, tests/test_fxcm_model.py).
#ifndef CMIXB200_FXCM_TEXT_H
#define CMIXB200_FXCM_TEXT_H

#include <stdint.h>

#if defined(__CUDACC__)
#define FX_HD __host__ __device__
#else
#define FX_HD
#endif

namespace cmixb200 {
namespace fx {

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;

FX_HD inline int imin(int a, int b) { return a < b ? a : b; }
FX_HD inline int imax(int a, int b) { return a < b ? b : a; }
FX_HD inline int cstrlen(const char* s) { int n = 0; while (s[n]) ++n; return n; }
FX_HD inline bool bytes_eq(const u8* a, const char* b, int n) { for (int i = 0; i < n; ++i) if (a[i] != (u8)b[i]) return false; return true; }

// WRT-transformed punctuation (fxcmv1.cpp:1853-1880): the preprocessor swaps some ASCII ranges
enum : int { kColon = 'J', kSemicolon = 'K', kLess = 'L', kEquals = 'M', kGreater = 'N', kQuestion = 'O', kFirstUpper = 64,
             kSqOpen = 91, kSqClose = 93, kCurlyOpen = 'P', kVBar = 'Q', kCurlyClose = 'R', kApos = 39, kQuote = 34, kSpace = 32,
             kHtLink = 31, kHtml = 30, kLF = 10, kEscape = 12, kUpper = 7, kTextData = 96, kWikiTable = '-' };

// word-type flags (fxcmv1.cpp:2375-2396)
enum : u32 { T_Verb = 1u << 0, T_Noun = 1u << 1, T_Adjective = 1u << 2, T_Plural = 1u << 3, T_PastTense = (1u << 5) | T_Verb,
             T_PresentParticiple = (1u << 4) | T_Verb, T_AdjSuperlative = (1u << 5) | T_Adjective, T_AdjWithout = (1u << 6) | T_Adjective,
             T_AdjFull = (1u << 7) | T_Adjective, T_AdverbOfManner = 1u << 8, T_Suffix = 1u << 9, T_Prefix = 1u << 10, T_Male = 1u << 11,
             T_Female = 1u << 13, T_Article = 1u << 14, T_Conjunction = 1u << 15, T_Adposition = 1u << 16, T_Number = 1u << 17,
             T_ConjAdverb = 1u << 19 };
enum : u32 { P_Negation = 1, P_Irr = 2 | 1, P_Over = 4, P_Under = 8, P_Unn = 16 | 1, P_Non = 32 | 1, P_Anti = 64 | 1, P_Dis = 128 | 1 };
enum : u32 { X_NESS = 1, X_ITY = 2 | T_Noun, X_Capable = 4, X_NCE = 8, X_NT = 16, X_ION = 32, X_AL = 64 | T_Adjective, X_IC = 128 | T_Adjective,
             X_IVE = 256, X_OUS = 512 | T_Adjective };

// ---------------------------------------------------------------- small ring "vectors" (fxcmv1.cpp:1884-1930)
// size wraps at the capacity; pop clears the slot behind the new end and never goes below 0.
template <class T, int S> struct Ring {
  T v[S];
  int n;
  FX_HD void push(T e) { v[n++] = e; n &= S - 1; }
  FX_HD void pop() { if (n > 0) { v[n] = 0; --n; } }
  FX_HD void reset() { v[0] = 0; n = 0; }
  FX_HD T prev() const { return n > 1 ? v[n - 2] : (T)0; }
  FX_HD T back() const { return v[n - 1]; }
};

// Open-bracket stack with distance since opening (fxcmv1.cpp:1932-2020). W = 8 or 16: width of the stored symbol.
template <int W> struct Nest {
  u32 context;
  Ring<int, 512> open, dist;
  u32 cxt, dst;          // truncated to W bits on every store, as the reference's T members
  u16 pair[20];          // (opening, closing) pairs
  int n_pair;            // number of table entries (2 per pair)
  int do_pop, limit;
  FX_HD void init(const u16* p, int n, int pop, int lim) { for (int i = 0; i < n; ++i) pair[i] = p[i]; n_pair = n; do_pop = pop; limit = lim; }
  FX_HD void clear() { open.reset(); dist.reset(); context = cxt = dst = 0; }
  FX_HD bool opens(int b) const { for (int i = 0; i < n_pair; i += 2) if (pair[i] == b) return true; return false; }
  FX_HD bool closes(int b, int c) const { bool f = false; for (int i = 0; i < n_pair; i += 2) if (pair[i] == b && pair[i + 1] == c) f = true; return f; }
  FX_HD int last() const { return open.prev(); }
  FX_HD void update(int byte) {
    bool popped = false;
    if (open.n != 0) {
      if (closes(open.back(), byte) || dist.back() >= limit) { open.pop(); dist.pop(); popped = do_pop != 0; }
      else dist.v[dist.n - 1]++;
    }
    if (!popped && opens(byte)) { open.push(byte); dist.push(0); }
    if (open.n != 0) {
      const u32 mask = (1u << W) - 1;
      cxt = (u32)open.back() & mask;
      dst = (u32)imin(dist.back(), (1 << W) - 1) & mask;
      context = (1u << W) * cxt + dst;
    } else context = cxt = dst = 0;
  }
};

// Row / column / wiki-table tracker (fxcmv1.cpp:2022-2170)
struct Columns {
  struct Row { u32 linepos; u8 fc; Ring<u8, 2048> bytes; };
  Row row[4];
  Ring<u32, 32> cell[4];
  int rows, cell_count, cells, above, above1;
  int nl, is_temp, limit;
  u8 nl_char;
  FX_HD void init() { limit = 31; nl_char = kLF; }
  FX_HD u8 lastfc(int i = 0) const { return row[(rows - i) & 3].fc; }
  FX_HD int collen(int i = 0, int l = 0) const { return imin(l ? l : limit, row[(rows - i) & 3].bytes.n + 1); }
  FX_HD u32 nlpos(int i) const { return row[(rows - i) & 3].linepos; }
  FX_HD u8 colb(int i, int j, int l = 0) const {
    if (collen(0, l) < collen(i, l)) return row[(rows - i) & 3].bytes.v[collen() - (1 + j)];
    return 0;
  }
  FX_HD int cells_in(int r = 1) const { return cell[(cells - r) & 3].n; }
  FX_HD int cell_pos(int id, int r = 1) const { int t = cells_in(r) - 1; t = imin(t, id); return (int)cell[(cells - r) & 3].v[t]; }
  FX_HD void reset_cells() { for (int i = 0; i < 4; ++i) cell[i].reset(); }
  FX_HD void new_row_cells(u32 blpos) { cells = (cells + 1) & 3; cell[cells].reset(); cell[cells].push(blpos); cell_count = above = above1 = 0; }
  FX_HD void step_above(bool newcell) {
    if (above) { ++above; if (above > above1) above = above1 = 0; }
    if (newcell && cells_in() > 0) { above = cell_pos(cell_count - 1); above1 = cell_pos(cell_count); }
  }
  FX_HD void update(int byte, u32 b2, u32 blpos, bool is_pre) {
    if (b2 == (u32)((kCurlyOpen << 16) + (kCurlyOpen << 8) + kVBar)) nl_char = kWikiTable;
    else if (b2 == (u32)((kVBar << 16) + (kCurlyClose << 8) + kCurlyClose)) { nl_char = kLF; reset_cells(); }
    if (byte != kCurlyOpen && (b2 & 0xff00) == (u32)(kCurlyOpen << 8) && (b2 & 0xff0000) != (u32)(kCurlyOpen << 16)) is_temp = 1;
    else if (is_temp && byte == kCurlyClose) is_temp = 0;
    nl = 0;
    if (byte == kLF) {
      row[rows].bytes.push((u8)byte);
      rows = (rows + 1) & 3;
      row[rows].bytes.reset();
      row[rows].fc = 0;
      row[rows].linepos = blpos - 1;
    } else {
      row[rows].bytes.push((u8)byte);
      if (collen() == 2) {
        row[rows].fc = (u8)imin(byte, kTextData);
        nl = 1;
        if (row[rows].fc == kGreater && !is_pre) nl_char = kGreater;
        if (row[rows].fc == kSqOpen && nl_char == kGreater) nl_char = kLF;
      }
    }
    if (nl_char == kWikiTable) {
      if ((b2 & 0xffff) == (u32)(kWikiTable + kVBar * 256)) new_row_cells(blpos);
      bool newcell = false;
      if ((b2 & 0xffff) == (u32)(kVBar + kVBar * 256) || (b2 & 0xffff00) == (u32)((kVBar + kLF * 256) * 256)) {
        cell[cells].push(blpos); ++cell_count; newcell = true;
      }
      step_above(newcell);
    }
    if (nl_char == kGreater) {
      if ((b2 & 0xffff) == (u32)(kGreater + kLF * 256)) new_row_cells(blpos);
      else {
        bool newcell = false;
        if ((b2 & 0xff) == (u32)kGreater) { cell[cells].push(blpos); ++cell_count; newcell = true; }
        step_above(newcell);
      }
    }
  }
};

// Word list of the current sentence / paragraph / stream (fxcmv1.cpp:2180-2300)
struct WordList {
  Ring<u16, 256> sbytes;
  Ring<u32, 256> type, stem;
  Ring<u8, 256> capital;
  u32 fword, ftype;
  u8 pbyte;
  int wordcount, upper, ref;
  FX_HD void clear() { sbytes.reset(); type.reset(); stem.reset(); capital.reset(); fword = 0; pbyte = 0; wordcount = upper = 0; ftype = 0; ref = 0; }
  FX_HD void set(u8 b, int a = 0) { pbyte = b; upper = a; }
  FX_HD void add(u32 w, u8 b, u32 t, u32 s) {
    if (fword == 0) fword = w;
    sbytes.push((u16)(pbyte * 256 + b)); type.push(t); stem.push(s); capital.push((u8)upper);
    pbyte = 0; ++wordcount;
    if (ftype == 0 && t) ftype = t;
  }
  FX_HD void remove() { if (stem.n) { sbytes.pop(); type.pop(); stem.pop(); capital.pop(); --wordcount; } }
  FX_HD u32 word(int i = 1) const { return stem.n >= i ? stem.v[stem.n - i] : 0; }
  FX_HD u16 sb(int i = 1) const { return sbytes.n >= i ? sbytes.v[sbytes.n - i] : 0; }
  FX_HD u32 typ(int i = 1) const { return type.n >= i ? type.v[type.n - i] : 0; }
  FX_HD u8 cap(int i = 1) const { return capital.n >= i ? capital.v[capital.n - i] : 0; }
  FX_HD u32 last(int j, u32 t) const {
    if (t == 0) return word(j);
    if (type.n >= j) for (int i = j; i < type.n; ++i) if (typ(i) & t) return word(i);
    return word(j);
  }
  FX_HD u32 last_if(int j, u32 t) const {
    if (t == 0) return word(j);
    if (type.n >= j) for (int i = j; i < type.n; ++i) if (typ(i) & t) return word(i);
    return 0;
  }
  FX_HD void drop_left(int len, u8 c, u8 d, bool f = true) {
    if ((sb(1) & 0xff) == d)
      for (int i = 1; i < len; ++i)
        if ((sb(i) >> 8) == c) { while ((sb(1) >> 8) != c) remove(); if (f) remove(); break; }
  }
  FX_HD void drop_right(int len, u8 c, u8 d, bool f = true) {
    if ((sb(1) & 0xff) == d)
      for (int i = 1; i < len; ++i)
        if ((sb(i) & 0xff) == c) { while ((sb(1) & 0xff) != c) remove(); if (f) remove(); break; }
  }
};

// ---------------------------------------------------------------- word + English affix stemmer (fxcmv1.cpp:2302-3205)
struct Word {
  u8 L[64];
  u8 s, e;               // first / last letter
  u32 hash, type, suffix, prefix;
  FX_HD void clear() { for (int i = 0; i < 64; ++i) L[i] = 0; s = e = 0; hash = type = suffix = prefix = 0; }
  FX_HD u32 len() const { return L[s] != 0 ? (u32)(e - s + 1) : 0u; }
  FX_HD u8 at(int i) const { return (e - s >= i) ? L[(u8)(s + i)] : 0; }          // from the front
  FX_HD u8 rat(int i) const { return (e - s >= i) ? L[(u8)(e - i)] : 0; }         // from the back
  FX_HD void append(int c) { if (c > 0 && c < 128 && e < 63) { e = (u8)(e + (L[e] > 0)); L[e] = (u8)c; } }
  FX_HD bool is(const char* w) const { const int n = cstrlen(w); return (int)(e - s + (L[s] != 0)) == n && bytes_eq(&L[s], w, n); }
  FX_HD bool ends(const char* w) const { const u32 n = cstrlen(w); return len() > n && bytes_eq(&L[e - n + 1], w, (int)n); }
  FX_HD bool starts(const char* w) const { const u32 n = cstrlen(w); return len() > n && bytes_eq(&L[s], w, (int)n); }
  FX_HD bool swap_suffix(const char* from, const char* to) {
    const u32 n = cstrlen(from);
    if (len() > n && bytes_eq(&L[e - n + 1], from, (int)n)) {
      const int m = cstrlen(to);
      if (m > 0) {
        const int cnt = imin(63, e + m) - e;
        for (int i = 0; i < cnt; ++i) L[e - n + 1 + i] = (u8)to[i];
        e = (u8)imin(63, (int)e - (int)n + m);
      } else e = (u8)(e - n);
      return true;
    }
    return false;
  }
  // `list` = words separated by '|'
  FX_HD bool any_of(const char* list) const {
    const int n = (int)len();
    for (const char* p = list; *p;) {
      int k = 0;
      while (p[k] && p[k] != '|') ++k;
      if (k == n && bytes_eq(&L[s], p, n)) return true;
      p += k;
      if (*p == '|') ++p;
    }
    return false;
  }
};

FX_HD inline bool in_set(int c, const char* set) { for (; *set; ++set) if ((u8)*set == (u8)c) return true; return false; }
FX_HD inline bool vowel(int c) { return in_set(c, "aeiouy"); }

struct Stemmer {
  FX_HD static void rehash(Word& w) { u32 h = 0xb0a710adu; for (int i = w.s; i <= w.e; ++i) h = h * 263u * 32u + w.L[i]; w.hash = h; }
  FX_HD static u32 region(const Word& w, u32 from) {
    bool seen = false;
    for (int i = w.s + (int)from; i <= w.e; ++i) {
      if (vowel(w.L[i])) { seen = true; continue; }
      if (seen) return (u32)(i - w.s + 1);
    }
    return w.len();
  }
  FX_HD static u32 region1(const Word& w) {
    if (w.starts("gener")) return 5;
    if (w.starts("arsen")) return 5;
    if (w.starts("commun")) return 6;
    return region(w, 0);
  }
  FX_HD static bool in_rn(const Word& w, u32 rn, int suffix_len) { return w.s != w.e && (u64)rn <= (u64)w.len() - (u64)suffix_len; }
  FX_HD static bool short_syllable(const Word& w) {
    if (w.e == w.s) return false;
    if (w.e == w.s + 1) return vowel(w.rat(1)) && !vowel(w.rat(0));
    return !vowel(w.rat(2)) && vowel(w.rat(1)) && !vowel(w.rat(0)) && !in_set(w.rat(0), "wxY");
  }
  FX_HD static bool short_word(const Word& w) { return short_syllable(w) && region1(w) == w.len(); }
  FX_HD static bool has_vowel(const Word& w) { for (int i = w.s; i <= w.e; ++i) if (vowel(w.L[i])) return true; return false; }

  FX_HD static bool trim_apostrophes(Word& w) {
    bool r = false;
    int cnt = 0;
    while (w.s != w.e && w.at(0) == kApos) { r = true; ++w.s; ++cnt; }
    while (w.s != w.e && w.rat(0) == kApos) { if (cnt == 0) break; --w.e; --cnt; }
    if (w.rat(0) == '-') --w.e;
    return r;
  }
  FX_HD static void mark_y(Word& w) {
    if (w.at(0) == 'y') w.L[w.s] = 'Y';
    for (int i = w.s + 1; i <= w.e; ++i) if (vowel(w.L[i - 1]) && w.L[i] == 'y') w.L[i] = 'Y';
  }
  FX_HD static bool prefixes(Word& w) {
    if (w.starts("irr") && w.len() > 5 && (w.at(3) == 'a' || w.at(3) == 'e')) { w.s += 2; w.type |= T_Prefix; w.prefix |= P_Irr; }
    else if (w.starts("over") && w.len() > 5) { w.s += 4; w.type |= T_Prefix; w.prefix |= P_Over; }
    else if (w.starts("under") && w.len() > 6) { w.s += 5; w.type |= T_Prefix; w.prefix |= P_Under; }
    else if (w.starts("unn") && w.len() > 5) { w.s += 2; w.type |= T_Prefix; w.prefix |= P_Unn; }
    else if (w.starts("non") && w.len() > (u32)(5 + (w.at(3) == '-'))) { w.s += 2 + (w.at(3) == '-'); w.type |= T_Prefix; w.prefix |= P_Non; }
    else if (w.starts("anti") && w.len() > 6 && w.at(4) == '-') { w.s += 4 + (w.at(4) == '-'); w.type |= T_Prefix; w.prefix |= P_Anti; }
    else if (w.starts("dis") && w.len() > 5 && w.at(3) == '-') { w.s += 2 + (w.at(3) == '-'); w.type |= T_Prefix; w.prefix |= P_Dis; }
    else return false;
    return true;
  }
  FX_HD static bool superlatives(Word& w) {
    if (w.ends("est") && w.len() > 4) {
      const u8 keep = w.e;
      w.e -= 3;
      w.type |= T_AdjSuperlative;
      if (w.rat(0) == w.rat(1) && w.rat(0) != 'r' && !(w.len() >= 4 && bytes_eq(&w.L[w.e - 3], "sugg", 4))) {
        w.e -= (((w.rat(0) != 'f' && w.rat(0) != 'l' && w.rat(0) != 's') ||
                 (w.len() > 4 && w.rat(1) == 'l' && (w.rat(2) == 'u' || w.rat(3) == 'u' || w.rat(3) == 'v'))) &&
                (!(w.len() == 3 && w.rat(1) == 'd' && w.rat(2) == 'o')));
        if (w.len() == 2 && (w.at(0) != 'i' || w.at(1) != 'n')) { w.e = keep; w.type &= ~T_AdjSuperlative; }
      } else {
        switch (w.rat(0)) {
          case 'd': case 'k': case 'm': case 'y': break;
          case 'g':
            if (!(w.len() > 3 && (w.rat(1) == 'n' || w.rat(1) == 'r') && !bytes_eq(&w.L[w.e - 3], "cong", 4))) { w.e = keep; w.type &= ~T_AdjSuperlative; }
            else w.e += (w.rat(2) == 'a');
            break;
          case 'i': w.L[w.e] = 'y'; break;
          case 'l':
            if (w.e == w.s + 1 || bytes_eq(&w.L[w.e - 2], "mo", 2)) { w.e = keep; w.type &= ~T_AdjSuperlative; }
            else w.e += !vowel(w.rat(1));
            break;
          case 'n':
            if (w.len() < 3 || !vowel(w.rat(1)) || !vowel(w.rat(2))) { w.e = keep; w.type &= ~T_AdjSuperlative; }
            break;
          case 'r':
            if (w.len() > 3 && vowel(w.rat(1)) && vowel(w.rat(2))) w.e += (w.rat(2) == 'u') && (w.rat(1) == 'a' || w.rat(1) == 'i');
            else { w.e = keep; w.type &= ~T_AdjSuperlative; }
            break;
          case 's': ++w.e; break;
          case 'w':
            if (!(w.len() > 2 && vowel(w.rat(1)))) { w.e = keep; w.type &= ~T_AdjSuperlative; }
            break;
          case 'h':
            if (!(w.len() > 2 && !vowel(w.rat(1)))) { w.e = keep; w.type &= ~T_AdjSuperlative; }
            break;
          default: w.e += 3; w.type &= ~T_AdjSuperlative;
        }
      }
    }
    return (w.type & T_AdjSuperlative) > 0;
  }
  FX_HD static bool step0(Word& w) {
    if (w.ends("'s'")) { w.e -= 3; w.type |= T_Plural; return true; }
    if (w.ends("'s")) { w.e -= 2; w.type |= T_Plural; return true; }
    if (w.ends("'")) { w.e -= 1; w.type |= T_Plural; return true; }
    return false;
  }
  FX_HD static bool step1a(Word& w) {
    if (w.ends("sses")) { w.e -= 2; w.type |= T_Plural; return true; }
    if (w.ends("ied") || w.ends("ies")) { w.type |= (w.rat(0) == 'd') ? T_PastTense : T_Plural; w.e -= 1 + (w.len() > 4); return true; }
    if (w.ends("us") || w.ends("ss")) return false;
    if (w.rat(0) == 's' && w.len() > 2)
      for (int i = w.s; i <= w.e - 2; ++i)
        if (vowel(w.L[i])) { --w.e; w.type |= T_Plural; return true; }
    if (w.ends("n't") && w.len() > 4) {
      switch (w.rat(3)) {
        case 'a': if (w.rat(4) == 'c') w.e -= 2; else w.swap_suffix("n't", "ll"); break;
        case 'i': w.swap_suffix("in't", "m"); break;
        case 'o': if (w.rat(4) == 'w') w.swap_suffix("on't", "ill"); else w.e -= 3; break;
        default: w.e -= 3;
      }
      w.type |= T_Prefix; w.prefix |= P_Negation;
      return true;
    }
    if (w.ends("hood") && w.len() > 7) { w.e -= 4; return true; }
    return false;
  }
  FX_HD static bool step1b(Word& w, u32 r1) {
    const char* suf[6] = {"eedly", "eed", "ed", "edly", "ing", "ingly"};
    const u32 typ[6] = {T_AdverbOfManner, 0, T_PastTense, T_AdverbOfManner | T_PastTense, T_PresentParticiple, T_AdverbOfManner | T_PresentParticiple};
    for (int i = 0; i < 6; ++i) {
      if (!w.ends(suf[i])) continue;
      if (i < 2) {
        if (in_rn(w, r1, cstrlen(suf[i]))) w.e -= 1 + i * 2;
      } else {
        const u8 keep = w.e;
        w.e -= cstrlen(suf[i]);
        if (!has_vowel(w)) { w.e = keep; return false; }
        if (w.ends("at") || w.ends("bl") || w.ends("iz") || short_word(w)) w.append('e');
        else if (w.len() > 2) {
          if (w.rat(0) == w.rat(1) && in_set(w.rat(0), "bdfgmnprt")) --w.e;
          else if (i == 2 || i == 3) {
            switch (w.rat(0)) {
              case 'c': case 's': case 'v': w.e += !(w.ends("ss") || w.ends("ias")); break;
              case 'd': w.e += vowel(w.rat(1)) && !in_set(w.rat(2), "aeio"); break;
              case 'k': w.e += w.ends("uak"); break;
              case 'l': w.e += in_set(w.rat(1), "bcdfgkptyz") || (in_set(w.rat(1), "aiou") && !vowel(w.rat(2))); break;
            }
          } else if (i >= 4) {
            switch (w.rat(0)) {
              case 'd': if (vowel(w.rat(1)) && w.rat(2) != 'a' && w.rat(2) != 'e' && w.rat(2) != 'o') w.append('e'); break;
              case 'g':
                if (in_set(w.rat(1), "adeilru") ||
                    (w.rat(1) == 'n' && (w.rat(2) == 'e' || (w.rat(2) == 'u' && w.rat(3) != 'b' && w.rat(3) != 'd') ||
                                         (w.rat(2) == 'a' && (w.rat(3) == 'r' || (w.rat(3) == 'h' && w.rat(4) == 'c'))) ||
                                         (w.ends("ring") && (w.rat(4) == 'c' || w.rat(4) == 'f')))))
                  w.append('e');
                break;
              case 'l':
                if (!(w.rat(1) == 'l' || w.rat(1) == 'r' || w.rat(1) == 'w' || (vowel(w.rat(1)) && vowel(w.rat(2))))) w.append('e');
                if (w.ends("uell") && w.len() > 4 && w.rat(4) != 'q') --w.e;
                break;
              case 'r':
                if (((w.rat(1) == 'i' && w.rat(2) != 'a' && w.rat(2) != 'e' && w.rat(2) != 'o') ||
                     (w.rat(1) == 'a' && !(w.rat(2) == 'e' || w.rat(2) == 'o' || (w.rat(2) == 'l' && w.rat(3) == 'l'))) ||
                     (w.rat(1) == 'o' && !(w.rat(2) == 'o' || (w.rat(2) == 't' && w.rat(3) != 's'))) ||
                     w.rat(1) == 'c' || w.rat(1) == 't') && !w.ends("str"))
                  w.append('e');
                break;
              case 't': if (w.rat(1) == 'o' && w.rat(2) != 'g' && w.rat(2) != 'l' && w.rat(2) != 'i' && w.rat(2) != 'o') w.append('e'); break;
              case 'u': if (!(w.len() > 3 && vowel(w.rat(1)) && vowel(w.rat(2)))) w.append('e'); break;
              case 'z':
                if (w.ends("izz") && w.len() > 3 && (w.rat(3) == 'h' || w.rat(3) == 'u')) --w.e;
                else if (w.rat(1) != 't' && w.rat(1) != 'z') w.append('e');
                break;
              case 'k': if (w.ends("uak")) w.append('e'); break;
              case 'b': case 'c': case 's': case 'v':
                if (!((w.rat(0) == 'b' && (w.rat(1) == 'm' || w.rat(1) == 'r')) || w.ends("ss") || w.ends("ias") || w.is("zinc"))) w.append('e');
                break;
            }
          }
        }
      }
      w.type |= typ[i];
      return true;
    }
    return false;
  }
  FX_HD static bool step1c(Word& w) {
    if (w.len() > 2 && w.rat(0) == 'y' && !vowel(w.rat(1))) { w.L[w.e] = 'i'; return true; }
    return false;
  }
  FX_HD static bool step2(Word& w, u32 r1) {
    const char* from[22] = {"ization", "ational", "ousness", "iveness", "fulness", "tional", "lessli", "biliti", "entli", "ation", "alism",
                            "aliti", "fulli", "ousli", "iviti", "enci", "anci", "abli", "izer", "ator", "alli", "bli"};
    const char* to[22] = {"ize", "ate", "ous", "ive", "ful", "tion", "less", "ble", "ent", "ate", "al", "al", "ful", "ous", "ive", "ence",
                          "ance", "able", "ize", "ate", "al", "ble"};
    const u32 typ[22] = {T_Suffix, T_Suffix | T_Adjective, T_Suffix, T_Suffix, T_Suffix, T_Suffix | T_Adjective, T_AdverbOfManner,
                         T_AdverbOfManner | T_Noun | T_Suffix, T_AdverbOfManner, T_Suffix, 0, T_Noun | T_Suffix, T_AdverbOfManner, T_AdverbOfManner,
                         T_Noun | T_Suffix, 0, 0, T_AdverbOfManner, 0, 0, T_AdverbOfManner, T_AdverbOfManner};
    const u32 sfx[22] = {X_ION, X_ION | X_AL, X_NESS, X_NESS, X_NESS, X_ION | X_AL, 0, X_ITY, 0, X_ION, 0, X_ITY, 0, 0, X_ITY, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 22; ++i)
      if (w.ends(from[i]) && in_rn(w, r1, cstrlen(from[i]))) { w.swap_suffix(from[i], to[i]); w.type |= typ[i]; w.suffix |= sfx[i]; return true; }
    if (w.ends("logi") && in_rn(w, r1, 3)) { --w.e; return true; }
    if (w.ends("li")) {
      if (in_rn(w, r1, 2) && in_set(w.rat(2), "cdeghkmnrt")) { w.e -= 2; w.type |= T_AdverbOfManner; return true; }
      if (w.len() > 3) {
        switch (w.rat(2)) {
          case 'b': w.L[w.e] = 'e'; w.type |= T_AdverbOfManner; return true;
          case 'i': if (w.len() > 4) { w.e -= 2; w.type |= T_AdverbOfManner; return true; } break;
          case 'l': if (w.len() > 5 && (w.rat(3) == 'a' || w.rat(3) == 'u')) { w.e -= 2; w.type |= T_AdverbOfManner; return true; } break;
          case 's': w.e -= 2; w.type |= T_AdverbOfManner; return true;
          case 'e': case 'g': case 'm': case 'n': case 'r': case 'w':
            if (w.len() > (u32)(4 + (w.rat(2) == 'r'))) { w.e -= 2; w.type |= T_AdverbOfManner; return true; }
        }
      }
    }
    return false;
  }
  FX_HD static bool step3(Word& w, u32 r1, u32 r2) {
    const char* from[8] = {"ational", "tional", "alize", "icate", "iciti", "ical", "ful", "ness"};
    const char* to[8] = {"ate", "tion", "al", "ic", "ic", "ic", "", ""};
    const u32 typ[8] = {T_Suffix | T_Adjective, T_Suffix | T_Adjective, 0, 0, T_Noun | T_Suffix, T_Suffix | T_Adjective, T_AdjFull, T_Suffix};
    const u32 sfx[8] = {X_ION | X_AL, X_ION | X_AL, 0, 0, X_ITY, X_AL, 0, X_NESS};
    bool r = false;
    for (int i = 0; i < 8; ++i)
      if (w.ends(from[i]) && in_rn(w, r1, cstrlen(from[i]))) { w.swap_suffix(from[i], to[i]); w.type |= typ[i]; w.suffix |= sfx[i]; r = true; break; }
    if (w.ends("ative") && in_rn(w, r2, 5)) { w.e -= 5; w.type |= T_Suffix; w.suffix |= X_IVE; return true; }
    if (w.len() > 5 && w.ends("less")) { w.e -= 4; w.type |= T_AdjWithout; return true; }
    return r;
  }
  FX_HD static bool step4(Word& w, u32 r2) {
    const char* suf[20] = {"al", "ance", "ence", "er", "ic", "able", "ible", "ant", "ement", "ment", "ent", "ou", "ism", "ate", "iti", "ous",
                           "ive", "ize", "sion", "tion"};
    const u32 typ[20] = {T_Suffix | T_Adjective, T_Suffix, T_Suffix, 0, T_Suffix | T_Adjective, T_Suffix, T_Suffix, T_Suffix, 0, 0, T_Suffix, 0, 0, 0,
                         T_Suffix | T_Noun, T_Suffix | T_Adjective, T_Suffix, 0, T_Suffix, T_Suffix};
    const u32 sfx[20] = {X_AL, X_NCE, X_NCE, 0, X_IC, X_Capable, X_Capable, X_NT, 0, 0, X_NT, 0, 0, 0, X_ITY, X_OUS, X_IVE, 0, X_ION, X_ION};
    bool r = false;
    for (int i = 0; i < 20; ++i) {
      if (w.ends(suf[i]) && in_rn(w, r2, cstrlen(suf[i]))) {
        w.e -= cstrlen(suf[i]) - (i > 17);
        if (i != 10 || w.rat(0) != 'm') { w.type |= typ[i]; w.suffix |= sfx[i]; }
        if (i == 0 && w.ends("nti")) { --w.e; r = true; continue; }
        return true;
      }
    }
    return r;
  }
  FX_HD static bool step5(Word& w, u32 r1, u32 r2) {
    if (w.rat(0) == 'e' && !w.is("here")) {
      if (in_rn(w, r2, 1)) --w.e;
      else if (in_rn(w, r1, 1)) { --w.e; w.e += short_syllable(w); }
      else return false;
      return true;
    }
    if (w.len() > 1 && w.rat(0) == 'l' && in_rn(w, r2, 1) && w.rat(1) == 'l') { --w.e; return true; }
    return false;
  }
  FX_HD static void set_letters(Word& w, const char* t) { const int n = cstrlen(t); for (int i = 0; i < n; ++i) w.L[w.s + i] = (u8)t[i]; w.e = (u8)(w.s + n - 1); }

  FX_HD static bool stem(Word& w, u32 blpos) {
    bool r = trim_apostrophes(w);
    if (prefixes(w)) r = true;
    if (superlatives(w)) r = true;
    {
      const char* a[19] = {"skis", "skies", "dying", "lying", "tying", "idly", "gently", "ugly", "early", "only", "singly", "sky", "news",
                           "howe", "atlas", "cosmos", "bias", "andes", "texas"};
      const char* b[11] = {"ski", "sky", "die", "lie", "tie", "idle", "gentle", "ugli", "earli", "onli", "singl"};
      const u32 t[19] = {T_Noun | T_Plural, T_Noun | T_Plural, T_PresentParticiple, T_PresentParticiple, T_PresentParticiple, T_AdverbOfManner,
                         T_AdverbOfManner, T_Adjective, T_Adjective | T_AdverbOfManner, 0, T_AdverbOfManner, T_Noun, T_Noun, 0, T_Noun, T_Noun,
                         T_Noun, T_Noun | T_Plural, T_Noun};
      for (int i = 0; i < 19; ++i)
        if (w.is(a[i])) {
          if (i < 11) set_letters(w, b[i]);
          rehash(w);
          w.type |= t[i];
          return i < 11;
        }
    }
    mark_y(w);
    const u32 r1 = region1(w), r2 = region(w, r1);
    if (step0(w)) r = true;
    if (step1a(w)) r = true;
    {
      const char* a[8] = {"inning", "outing", "canning", "herring", "earring", "proceed", "exceed", "succeed"};
      for (int i = 0; i < 8; ++i)
        if (w.is(a[i])) { rehash(w); w.type |= i < 5 ? T_Noun : T_Verb; return r; }
    }
    if (step1b(w, r1)) r = true;
    if (step1c(w)) r = true;
    if (step2(w, r1)) r = true;
    if (step3(w, r1, r2)) r = true;
    if (step4(w, r2)) r = true;
    if (step5(w, r1, r2)) r = true;
    for (u8 i = w.s; i <= w.e; ++i) if (w.L[i] == 'Y') w.L[i] = 'y';
    if (!w.type || w.type == T_Plural) {
      if (w.any_of("he|him|his|himself|man|men|boy|husband|actor")) { r = true; w.type |= T_Male; }
      else if (w.any_of("she|her|herself|woman|women|girl|wife|actress")) { r = true; w.type |= T_Female; }
      else if (w.any_of("a|an|the")) { r = true; w.type |= T_Article; }
      else if (w.any_of("for|and|nor|but|or|yet|so|than|as|that|if|when|because|while|where|after|though|whether|before|although|like|once|unless|now|except")) { r = true; w.type |= T_Conjunction; }
      else if (w.any_of("in|during|at|on|since|until|above|across|against|along|among|around|behind|below|beneath|beside|between|by|down|from|into|near|of|off|to|toward|under|upon|with|within")) { r = true; w.type |= T_Adposition; }
      else if (w.any_of("also|thus")) { r = true; w.type |= T_ConjAdverb; }
      else if (blpos < 451531986u && w.any_of("has|had|have|was|were|may|might|must|shall|should|can|could|will|would|is|am|are|be|being|been|do|does|did")) { r = true; w.type |= T_Verb; }
      else if (w.any_of("one|two|three|four|five|six|seven|eight|nine|ten|twenty|thirty|forty|fifty|sixty|seventy|eighty|ninety|hundred|thousand|million")) { r = true; w.type |= T_Number; }
    }
    rehash(w);
    return r;
  }
};

// 4-bit class of a word type (fxcmv1.cpp:3736-3752)
FX_HD inline int word_class(u32 t) {
  if (t & T_Verb) return 1;
  if (t & T_Noun) return 2;
  if (t & T_Adjective) return 3;
  if (t & T_Male) return 4;
  if (t & T_Female) return 5;
  if (t & T_Article) return 6;
  if (t & T_Conjunction) return 7;
  if (t & T_Adposition) return 8;
  if (t & T_ConjAdverb) return 9;
  if (t & T_AdverbOfManner) return 11;
  if (t & T_Suffix) return 12;
  if (t & T_Prefix) return 13;
  if (t & T_Plural) return 10;
  if (t) return 14;
  return 15;
}

FX_HD inline int char_swap(int c) {   // fxcmv1.cpp:2291-2298
  if (c >= '{' && c < 127) c += 'P' - '{';
  else if (c >= 'P' && c < 'T') c -= 'P' - '{';
  else if ((c >= ':' && c <= '?') || (c >= 'J' && c <= 'O')) c ^= 0x70;
  if (c == 'X' || c == '`') c ^= 'X' ^ '`';
  return c;
}

FX_HD inline u32 hash3(u32 a, u32 b, u32 c = 0xffffffffu) {   // fxcmv1.cpp:2285-2288
  const u32 h = a * 110002499u + b * 30005491u + c * 50004239u;
  return h ^ h >> 9 ^ a >> 3 ^ b >> 3 ^ c >> 4;
}

}  // namespace fx
}  // namespace cmixb200
#endif

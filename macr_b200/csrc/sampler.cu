// sampler.cu -- host-side (no GPU work) native twins of the reference's batch samplers:
//   MF       Data.sample()  macr_mf/load_data.py:543-566        (CPython `random` only)
//   LightGCN Data.sample()  macr_lightgcn/utility/load_data.py:174-212
//            (`random.sample` for the users, legacy numpy `np.random.randint(size=1)` for items)
// The reference samples in pure Python at ~3.6 us per triple -- two orders of magnitude slower
// than the GPU step it feeds.  These functions consume the SAME Mersenne-Twister streams word for
// word (the caller hands in random.getstate() / np.random.get_state() and puts the advanced
// states back), so the triples are bit-identical to the reference's and the interpreter-side
// generators stay in step for whatever draws from them next (the evaluation pass).
//
// Restated algorithms (CPython 3.x Lib/random.py, numpy legacy RandomState):
//   genrand_uint32         MT19937, tempering 11 / 7,0x9d2c5680 / 15,0xefc60000 / 18
//   getrandbits(k<=32)     genrand_uint32() >> (32 - k)
//   _randbelow(n)          k = n.bit_length(); r = getrandbits(k); while r >= n: r = getrandbits(k)
//   choice(seq)            seq[_randbelow(len(seq))]
//   sample(pop, k)         setsize = 21 (+ 4**ceil(log(3k, 4)) if k > 5); n <= setsize: partial
//                          shuffle of a pool copy; else rejection against a set of chosen indices
//   np randint(0, n, 1)    rng = n - 1; rng == 0: no draw; else mask = 2^bits(rng) - 1 and
//                          32-bit draws `genrand & mask` until <= rng   (masked rejection)
#include <emmintrin.h>
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>

#include "common.cuh"

namespace macr {
namespace {

struct MT {
  uint32_t *mt;  // 624 state words
  int idx;       // position, 624 = regenerate before the next draw

  void regen() {
    const uint32_t UPPER = 0x80000000u, LOWER = 0x7fffffffu, MATRIX_A = 0x9908b0dfu;
    int kk;
    uint32_t y;
    for (kk = 0; kk < 624 - 397; kk++) {
      y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER);
      mt[kk] = mt[kk + 397] ^ (y >> 1) ^ ((y & 1u) ? MATRIX_A : 0u);
    }
    for (; kk < 623; kk++) {
      y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER);
      mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? MATRIX_A : 0u);
    }
    y = (mt[623] & UPPER) | (mt[0] & LOWER);
    mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1u) ? MATRIX_A : 0u);
    idx = 0;
  }
  uint32_t next() {
    if (idx >= 624) regen();
    uint32_t y = mt[idx++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
  }
};

inline int bit_length(uint32_t n) { return n ? 32 - __builtin_clz(n) : 0; }

// random.Random._randbelow_with_getrandbits (n >= 1, n < 2^31 here)
template <class G>
inline uint32_t py_randbelow(G &g, uint32_t n) {
  const int k = bit_length(n);
  uint32_t r = g.next() >> (32 - k);
  while (r >= n) r = g.next() >> (32 - k);
  return r;
}

// numpy legacy RandomState.randint(0, n, size=1)[0]
template <class G>
inline uint32_t np_randint(G &g, uint32_t n) {
  const uint32_t rng = n - 1;
  if (rng == 0) return 0;
  uint32_t mask = rng;
  mask |= mask >> 1, mask |= mask >> 2, mask |= mask >> 4, mask |= mask >> 8, mask |= mask >> 16;
  uint32_t v;
  do v = g.next() & mask;
  while (v > rng);
  return v;
}

// random.sample(pop, k) / [random.choice(pop) for _ in range(k)] as the samplers use them
struct PickScratch {
  std::vector<int32_t> pool;
  std::vector<uint8_t> chosen;
};

template <class G>
void py_pick_users(G &g, const int32_t *pop, int n, int B, int n_users_flag, int32_t *out,
                   PickScratch *scratch = nullptr) {
  PickScratch local;
  if (!scratch) scratch = &local;
  if (B <= n_users_flag) {  // rd.sample(pop, B)
    long long setsize = 21;
    if (B > 5) setsize += (long long)llround(pow(4.0, ceil(log((double)B * 3.0) / log(4.0))));
    if (n <= setsize) {
      std::vector<int32_t> &pool = scratch->pool;
      pool.assign(pop, pop + n);
      for (int i = 0; i < B; ++i) {
        const uint32_t j = py_randbelow(g, (uint32_t)(n - i));
        out[i] = pool[j];
        pool[j] = pool[n - i - 1];
      }
    } else {
      std::vector<uint8_t> &chosen = scratch->chosen;
      chosen.assign((size_t)n, 0);
      for (int i = 0; i < B; ++i) {
        uint32_t j = py_randbelow(g, (uint32_t)n);
        while (chosen[j]) j = py_randbelow(g, (uint32_t)n);
        chosen[j] = 1;
        out[i] = pop[j];
      }
    }
  } else {
    for (int i = 0; i < B; ++i) out[i] = pop[py_randbelow(g, (uint32_t)n)];
  }
}

// membership in an ascending id list: short lists (the common case: tens of train items) are
// scanned without branches (vectorisable), long ones binary-searched
inline bool in_sorted(const int32_t *b, const int32_t *e, int32_t v) {
  const long n = e - b;
  if (n <= 64) {
    int hit = 0;
    for (long k = 0; k < n; ++k) hit |= (b[k] == v);
    return hit != 0;
  }
  return std::binary_search(b, e, v);
}

}  // namespace
}  // namespace macr

using namespace macr;

// lists: CSR over user ids -- `order` keeps the reference's list order (choice indexes into it),
// `sorted` the same ids ascending (membership test of the rejection loop).
extern "C" int macr_sample_mf(uint32_t *py_state /*[625]: 624 words + index*/,
                              const int32_t *users_pop, int n_pop, int n_users, int n_items,
                              const int64_t *rowptr, const int32_t *order, const int32_t *sorted,
                              int B, int32_t *users, int32_t *pos, int32_t *neg) {
  MACR_CHECK_ARG(py_state && users_pop && rowptr && order && sorted && users && pos && neg,
                 "macr_sample_mf: null pointer");
  MACR_CHECK_ARG(n_pop > 0 && n_items > 0 && B > 0, "macr_sample_mf: empty population");
  MACR_CHECK_ARG(B > n_users || B <= n_pop, "macr_sample_mf: sample larger than population");
  MT g{py_state, (int)py_state[624]};
  py_pick_users(g, users_pop, n_pop, B, n_users, users);
  for (int i = 0; i < B; ++i) {
    // the batch's users are known up front: pull their list heads into cache ahead of use
    if (i + 16 < B) __builtin_prefetch(rowptr + users[i + 16]);
    if (i + 8 < B) {
      const int64_t l8 = rowptr[users[i + 8]], h8 = rowptr[users[i + 8] + 1];
      const int64_t e8 = h8 < l8 + 128 ? h8 : l8 + 128;  // up to 8 cache lines of each list
      for (int64_t k = l8; k < e8; k += 16) {
        __builtin_prefetch(order + k);
        __builtin_prefetch(sorted + k);
      }
    }
    const int32_t u = users[i];
    const int64_t lo = rowptr[u], hi = rowptr[u + 1];
    pos[i] = hi > lo ? order[lo + py_randbelow(g, (uint32_t)(hi - lo))] : 0;
    for (;;) {
      const int32_t c = (int32_t)py_randbelow(g, (uint32_t)n_items);  // rd.choice(range list)
      if (!in_sorted(sorted + lo, sorted + hi, c)) {
        neg[i] = c;
        break;
      }
    }
  }
  py_state[624] = (uint32_t)g.idx;
  return MACR_OK;
}

// LightGCN: users from the Python stream; positives index `order` (pos lists, file order) with a
// numpy draw, negatives are numpy draws rejected while in `banned` (sorted ids, own CSR: the
// train list for sample(), train + test for sample_test()).
extern "C" int macr_sample_lgcn(uint32_t *py_state /*[625]*/, uint32_t *np_state /*[625]*/,
                                const int32_t *users_pop, int n_pop, int n_users, int n_items,
                                const int64_t *pos_rowptr, const int32_t *pos_order,
                                const int64_t *ban_rowptr, const int32_t *ban_sorted, int B,
                                int32_t *users, int32_t *pos, int32_t *neg) {
  MACR_CHECK_ARG(py_state && np_state && users_pop && pos_rowptr && pos_order && ban_rowptr &&
                     ban_sorted && users && pos && neg,
                 "macr_sample_lgcn: null pointer");
  MACR_CHECK_ARG(n_pop > 0 && n_items > 0 && B > 0, "macr_sample_lgcn: empty population");
  MACR_CHECK_ARG(B > n_users || B <= n_pop, "macr_sample_lgcn: sample larger than population");
  MT gp{py_state, (int)py_state[624]};
  MT gn{np_state, (int)np_state[624]};
  py_pick_users(gp, users_pop, n_pop, B, n_users, users);
  for (int i = 0; i < B; ++i) {
    if (i + 16 < B) {
      __builtin_prefetch(pos_rowptr + users[i + 16]);
      __builtin_prefetch(ban_rowptr + users[i + 16]);
    }
    if (i + 8 < B) {
      const int32_t u8 = users[i + 8];
      const int64_t pl = pos_rowptr[u8], ph = pos_rowptr[u8 + 1];
      const int64_t bl = ban_rowptr[u8], bh = ban_rowptr[u8 + 1];
      for (int64_t k = pl; k < (ph < pl + 128 ? ph : pl + 128); k += 16) __builtin_prefetch(pos_order + k);
      for (int64_t k = bl; k < (bh < bl + 128 ? bh : bl + 128); k += 16) __builtin_prefetch(ban_sorted + k);
    }
    const int32_t u = users[i];
    const int64_t lo = pos_rowptr[u], hi = pos_rowptr[u + 1];
    MACR_CHECK_ARG(hi > lo, "macr_sample_lgcn: user %d has no positive item", (int)u);
    pos[i] = pos_order[lo + np_randint(gn, (uint32_t)(hi - lo))];
    const int32_t *bb = ban_sorted + ban_rowptr[u], *be = ban_sorted + ban_rowptr[u + 1];
    for (;;) {
      const int32_t c = (int32_t)np_randint(gn, (uint32_t)n_items);
      if (!in_sorted(bb, be, c)) {
        neg[i] = c;
        break;
      }
    }
  }
  py_state[624] = (uint32_t)gp.idx;
  np_state[624] = (uint32_t)gn.idx;
  return MACR_OK;
}

// =============================================================================================
// Epoch samplers.  Same streams, same triples, but the sequential part no longer waits for
// memory nor for a branch predictor:
//  * the per-batch functions above spend most of their ~50-80 ns per triple on the lists (two
//    dependent cache misses and a scan per triple) although only ~0.1-1 % of the negative draws
//    are ever rejected by them.  Here a chunk of triples is drawn SPECULATIVELY -- every candidate
//    negative assumed "not a train item" -- touching nothing but the word stream and the list
//    lengths; a verification pass then reads the positives and tests the candidates against a
//    hashed pair set with many independent accesses in flight (on a second thread when the lists
//    are sparse).  A candidate that IS in the user's list invalidates what was drawn after it: the
//    word cursor goes back to just behind that draw, the rejection loop is finished the literal
//    way, and the chunk is redrawn from the next triple;
//  * the draws themselves are driven by the stream WORDS (draw_chunk_words, pick_users_words): one
//    word per iteration, accepted or not, with conditional moves instead of rejection loops;
//  * the words come from block-wise (vectorisable) MT19937 generation into a buffer that can be
//    re-read from the chunk's start.
// =============================================================================================
namespace macr {
namespace {

inline uint32_t mt_temper(uint32_t y) {
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  return y;
}

// inverse of mt_temper: the raw state word behind an output word
inline uint32_t mt_untemper(uint32_t y) {
  y ^= y >> 18;
  y ^= (y << 15) & 0xefc60000u;
  uint32_t x = y;
  for (int k = 0; k < 5; ++k) x = y ^ ((x << 7) & 0x9d2c5680u);
  y = x;
  for (int k = 0; k < 3; ++k) x = y ^ (x >> 11);
  return x;
}

// next 624 raw words in place; three ranges so that every loop only reads words that are final
// for it (dependence distance 227: the loops vectorise)
void mt_next_block(uint32_t *__restrict mt) {
  const uint32_t UPPER = 0x80000000u, LOWER = 0x7fffffffu, MATRIX_A = 0x9908b0dfu;
  for (int kk = 0; kk < 227; ++kk) {
    const uint32_t y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER);
    mt[kk] = mt[kk + 397] ^ (y >> 1) ^ ((0u - (y & 1u)) & MATRIX_A);
  }
  for (int kk = 227; kk < 623; ++kk) {
    const uint32_t y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER);
    mt[kk] = mt[kk - 227] ^ (y >> 1) ^ ((0u - (y & 1u)) & MATRIX_A);
  }
  const uint32_t y = (mt[623] & UPPER) | (mt[0] & LOWER);
  mt[623] = mt[396] ^ (y >> 1) ^ ((0u - (y & 1u)) & MATRIX_A);
}

// Output words of one generator, addressed by their absolute position in the stream (position 0
// = first word of the block the caller's state sits in).  `cur` may be moved back to any
// position not yet released.
class WordStream {
 public:
  size_t cur;

  explicit WordStream(uint32_t *state) : cur(state[624]), state_(state), base_(0), end_(0) {
    memcpy(raw_, state, sizeof(raw_));
    append_tempered();
  }
  inline uint32_t next() {
    if (__builtin_expect(cur >= end_, 0)) grow();
    return p_[cur++];
  }
  const uint32_t *words() const { return p_; }  // indexed by absolute position
  size_t end() const { return end_; }
  void ensure() { grow(); }  // make position `cur` readable
  // positions before `mark` will not be revisited (one block of slack is kept for store_state)
  void release_before(size_t mark) {
    const size_t keep = (mark ? (mark - 1) / 624 : 0) * 624;
    if (keep >= base_ + 64 * 624) {
      w_.erase(w_.begin(), w_.begin() + (keep - base_));
      base_ = keep;
      p_ = w_.data() - base_;
    }
  }
  // the generator state a draw-by-draw consumer would hold at `cur` (CPython / numpy regenerate
  // lazily: after the last word of a block the position is 624, not 0 of the next block)
  void store_state() const {
    size_t blk = cur / 624, idx = cur % 624;
    if (cur > 0 && idx == 0) blk -= 1, idx = 624;
    if ((blk + 1) * 624 == end_) {
      memcpy(state_, raw_, sizeof(raw_));
    } else {
      for (int k = 0; k < 624; ++k) state_[k] = mt_untemper(p_[blk * 624 + k]);
    }
    state_[624] = (uint32_t)idx;
  }

 private:
  void append_tempered() {
    const size_t n = w_.size();
    w_.resize(n + 624);
    uint32_t *dst = w_.data() + n;
    for (int k = 0; k < 624; ++k) dst[k] = mt_temper(raw_[k]);
    end_ += 624;
    p_ = w_.data() - base_;
  }
  __attribute__((noinline)) void grow() {
    while (cur >= end_) {
      mt_next_block(raw_);
      append_tempered();
    }
  }
  uint32_t *state_;
  uint32_t raw_[624];  // raw state of the newest generated block
  std::vector<uint32_t> w_;
  const uint32_t *p_;  // w_.data() - base_
  size_t base_, end_;
};

// ---- hashed (row, id) pair set -------------------------------------------------------------
// Buckets of eight 16-bit tags (one 16-byte load, one SSE2 compare, no probe loop); tag 0 = empty
// slot, a bucket whose eighth slot is taken may have lost pairs and answers "maybe" to everything.
// "Maybe" (a true member, or one query in ~10^4 by tag collision) is settled on the exact list.
inline uint64_t pair_hash(uint32_t row, uint32_t id) {
  return (((uint64_t)row << 32) | id) * 0x9E3779B97F4A7C15ull;
}
inline uint16_t pair_tag(uint64_t h) { return (uint16_t)((h >> 8) | 1u); }

struct PairSet {
  const uint16_t *tags;   // [buckets][8]
  int shift;              // 64 - log2(buckets)
  const int64_t *rowptr;  // exact lists behind the tags (ascending ids per row)
  const int32_t *sorted;

  inline const uint16_t *bucket_of(uint64_t h) const { return tags + ((h >> shift) << 3); }
  inline bool maybe(uint64_t h) const {
    const __m128i b = _mm_load_si128(reinterpret_cast<const __m128i *>(bucket_of(h)));
    const __m128i eq = _mm_cmpeq_epi16(b, _mm_set1_epi16((short)pair_tag(h)));
    // byte mask: tag matches anywhere, or slot 7 (bytes 14, 15) not empty
    const int m = _mm_movemask_epi8(eq);
    const int full = _mm_movemask_epi8(_mm_cmpeq_epi16(b, _mm_setzero_si128())) & 0xc000;
    return (m | (full ^ 0xc000)) != 0;
  }
  inline bool exact(uint32_t row, int32_t id) const {
    return in_sorted(sorted + rowptr[row], sorted + rowptr[row + 1], id);
  }
  inline bool contains(uint32_t row, int32_t id) const {
    return maybe(pair_hash(row, (uint32_t)id)) && exact(row, id);
  }
};

// Triples drawn ahead of their verification: a rejected candidate costs half a chunk of redrawn
// triples, so the chunk shrinks with the density of the lists (pairs / (rows * ids)).
inline int chunk_for(double pairs, double rows, double ids) {
  const double p = pairs / (rows * ids > 0 ? rows * ids : 1);
  int c = 256;
  while (c > 16 && c * p > 0.25) c >>= 1;
  return c;
}

}  // namespace
}  // namespace macr

extern "C" int macr_pairset_build(const int64_t *rowptr, const int32_t *ids, int n_rows,
                                  uint16_t *tags, int log2_buckets) {
  MACR_CHECK_ARG(rowptr && ids && tags, "macr_pairset_build: null pointer");
  MACR_CHECK_ARG(n_rows >= 0 && log2_buckets >= 0 && log2_buckets <= 40, "macr_pairset_build: bad size");
  MACR_CHECK_ARG(((uintptr_t)tags & 15) == 0, "macr_pairset_build: tags must be 16-byte aligned");
  const uint64_t buckets = 1ull << log2_buckets;
  memset(tags, 0, buckets * 16);
  for (int r = 0; r < n_rows; ++r)
    for (int64_t k = rowptr[r]; k < rowptr[r + 1]; ++k) {
      const uint64_t h = pair_hash((uint32_t)r, (uint32_t)ids[k]);
      uint16_t *b = tags + ((log2_buckets ? h >> (64 - log2_buckets) : 0) << 3);
      int s = 0;
      while (s < 7 && b[s]) ++s;
      b[s] = pair_tag(h);  // slot 7 taken = "bucket may have lost pairs"
    }
  return MACR_OK;
}

namespace macr {
namespace {

// Register-resident view of a WordStream for the draw loops (the stream object itself is only
// touched when a block runs out).
struct Reader {
  WordStream &g;
  const uint32_t *p;
  size_t cur, end;
  explicit Reader(WordStream &s) : g(s), p(s.words()), cur(s.cur), end(s.end()) {}
  ~Reader() { g.cur = cur; }
  inline uint32_t next() {
    if (__builtin_expect(cur >= end, 0)) refill();
    return p[cur++];
  }
  __attribute__((noinline)) void refill() {
    g.cur = cur;
    g.ensure();
    p = g.words();
    end = g.end();
  }
};

// A bounded draw as one test per stream word: value = op(word, hi32(param)), accepted when
// value <= lo32(param); a rejected word is simply followed by the next one.
struct PyDraw {  // random.choice / _randbelow: getrandbits(k) = word >> (32 - k), accepted below n
  static inline uint32_t below(Reader &r, uint32_t n) { return py_randbelow(r, n); }
  static constexpr uint32_t kMinN = 1;  // every n >= 1 consumes at least one word
  static inline uint64_t param(uint32_t n) { return ((uint64_t)(32 - bit_length(n)) << 32) | (n - 1); }
  static inline uint32_t value(uint32_t w, uint32_t hi) { return w >> hi; }
};
struct NpDraw {  // np.random.randint(0, n, size=1)[0]: word & mask, accepted up to n - 1
  static inline uint32_t below(Reader &r, uint32_t n) { return np_randint(r, n); }
  static constexpr uint32_t kMinN = 2;  // n == 1 returns 0 without a draw
  static inline uint64_t param(uint32_t n) {
    uint32_t mask = n - 1;
    mask |= mask >> 1, mask |= mask >> 2, mask |= mask >> 4, mask |= mask >> 8, mask |= mask >> 16;
    return ((uint64_t)mask << 32) | (n - 1);
  }
  static inline uint32_t value(uint32_t w, uint32_t hi) { return w & hi; }
};

constexpr int kMaxChunk = 256;

// Pass A over triples [i0, i1): positions of the positives, candidate negatives, and the
// cursor behind every triple (relative to `cur0`).  Reads the stream, lo/len -- nothing else.
template <class D>
__attribute__((noinline)) void draw_chunk(WordStream &g, int i0, int i1, const int64_t *lo,
                                          const uint32_t *len, uint32_t n_items, int64_t *pos_at,
                                          int32_t *neg, uint32_t *cur_after, size_t cur0) {
  Reader r(g);
  for (int i = i0; i < i1; ++i) {
    const uint32_t n = len[i];
    pos_at[i] = n ? lo[i] + (int64_t)D::below(r, n) : -1;
    neg[i] = (int32_t)D::below(r, n_items);
    cur_after[i] = (uint32_t)(r.cur - cur0);
  }
}

// The same pass driven by the WORDS instead of the draws.  A rejection loop is a branch the
// predictor cannot learn (about one miss per triple: most of the old pass).  Here every
// iteration takes the next word -- the position advances by one, unconditionally -- tests it
// against the parameters of the current draw (draw 2 j = positive of triple j, draw 2 j + 1 = its
// candidate negative), stores the value into the draw's slot and moves to the next draw only if
// the word was accepted; a rejected value is overwritten by the following word.  No data-dependent
// branch is left; the loop-carried state is the draw index and its parameter word.
// Needs every list of the chunk to take a draw (len >= D::kMinN), checked by the caller.
template <class D>
__attribute__((noinline)) void draw_chunk_words(WordStream &g, int i0, int i1, const int64_t *lo,
                                                const uint32_t *len, uint32_t n_items,
                                                int64_t *pos_at, int32_t *neg, uint32_t *cur_after,
                                                size_t cur0) {
  // parameters of draw q, packed: low half = shift count / mask, high half = largest accepted value
  uint64_t par[2 * kMaxChunk + 4];
  uint64_t rec[2 * kMaxChunk + 1];  // per draw: value | cursor behind the word << 32 (one store per word)
  const int nt = i1 - i0, qend = 2 * nt;
  auto pack = [](uint64_t pp) { return (pp << 32) | (pp >> 32); };
  const uint64_t neg_par = pack(D::param(n_items));
  for (int j = 0; j < nt; ++j) {
    par[2 * j] = pack(D::param(len[i0 + j]));
    par[2 * j + 1] = neg_par;
  }
  for (int k = qend; k < qend + 4; ++k) par[k] = neg_par;  // read ahead, never used
  size_t p = g.cur;
  long q = 0;
  // the current draw's parameters and those of the next two live in registers: on an accepted
  // word they shift down by conditional moves and the slot two ahead is loaded -- its value is
  // not needed before the accept after next, so no load sits in the loop-carried chain
  uint64_t c0 = par[0], c1 = par[1], c2 = par[2];
  while (q < qend) {
    if (p >= g.end()) {
      g.cur = p;
      g.ensure();
    }
    const uint32_t *W = g.words();
    // qend - q more draws need at least that many words and cannot accept more than they get:
    // a pass of exactly that many words (or what the block still holds) has ONE loop condition
    // and can never run past the last draw; ~2/3 of a pass is accepted, so passes shrink fast
    const size_t left = g.end() - p, want = (size_t)(qend - q);
    const size_t pstop = p + (left < want ? left : want);
    while (p < pstop) {
      const uint32_t v = D::value(W[p], (uint32_t)c0);
      ++p;
      rec[q] = (uint64_t)v | ((uint64_t)(p - cur0) << 32);
#if defined(__x86_64__)
      // accepted (v <= bound): c0 <- c1, c1 <- c2, ++q -- conditional moves off ONE compare (the
      // compiler turns plain selects back into a branch, which is what this loop exists to avoid)
      const uint64_t bound = c0 >> 32;
      asm("cmp %[v], %[b]\n\t"
          "cmovae %[c1], %[c0]\n\t"
          "cmovae %[c2], %[c1]\n\t"
          "sbb $-1, %[q]"
          : [c0] "+r"(c0), [c1] "+r"(c1), [q] "+r"(q)
          : [v] "r"((uint64_t)v), [b] "r"(bound), [c2] "r"(c2)
          : "cc");
#else
      const uint64_t m = 0ull - (uint64_t)(v <= (uint32_t)(c0 >> 32));
      q -= (long)m;
      c0 = (c1 & m) | (c0 & ~m);
      c1 = (c2 & m) | (c1 & ~m);
#endif
      c2 = par[q + 2];
    }
  }
  g.cur = p;
  for (int j = 0; j < nt; ++j) {
    pos_at[i0 + j] = lo[i0 + j] + (int64_t)(uint32_t)rec[2 * j];
    neg[i0 + j] = (int32_t)(uint32_t)rec[2 * j + 1];
    cur_after[i0 + j] = (uint32_t)(rec[2 * j + 1] >> 32);
  }
}

// random.sample(pop, B) / [random.choice(pop) ...] driven by the words: a word is rejected when it
// is out of range or (sample) already chosen -- both "draw again" in CPython -- so with the
// out-of-range values marked as taken from the start, one table look-up decides, and again no
// data-dependent branch is left.  The small-population branch of random.sample (partial shuffle
// of a pool copy, n <= setsize) keeps the draw-by-draw code: returns false.
bool pick_users_words(WordStream &g, const int32_t *pop, int n, int B, int n_users_flag,
                      int32_t *out, PickScratch *scratch) {
  const bool sample = B <= n_users_flag;
  if (sample) {
    long long setsize = 21;
    if (B > 5) setsize += (long long)llround(pow(4.0, ceil(log((double)B * 3.0) / log(4.0))));
    if (n <= setsize) return false;
  }
  const int k = bit_length((uint32_t)n), shift = 32 - k;
  std::vector<uint8_t> &taken = scratch->chosen;
  taken.assign((size_t)1 << k, 0);
  std::fill(taken.begin() + n, taken.end(), (uint8_t)1);
  std::vector<int32_t> &idx = scratch->pool;
  idx.resize((size_t)B + 1);
  uint8_t *tk = taken.data();
  int32_t *ix = idx.data();
  size_t p = g.cur;
  int q = 0;
  while (q < B) {
    if (p >= g.end()) {
      g.cur = p;
      g.ensure();
    }
    const uint32_t *W = g.words();
    const size_t pend = g.end();
    if (sample) {
      while (q < B && p < pend) {
        const uint32_t v = W[p] >> shift;
        ++p;
        const uint8_t t = tk[v];
        ix[q] = (int32_t)v;
        tk[v] = 1;  // taken from now on (it already was if t != 0)
        q += t == 0;
      }
    } else {  // with replacement: only the range test rejects
      while (q < B && p < pend) {
        const uint32_t v = W[p] >> shift;
        ++p;
        ix[q] = (int32_t)v;
        q += v < (uint32_t)n;
      }
    }
  }
  g.cur = p;
  for (int i = 0; i < B; ++i) out[i] = pop[ix[i]];
  return true;
}

// The chunk loop shared by both samplers.  `g` is the stream the item draws come from.
template <class D>
void sample_items(WordStream &g, const PairSet &set, int chunk, int B, const int32_t *users,
                  const int64_t *lo, const uint32_t *len, const int32_t *order, uint32_t n_items,
                  int32_t *pos, int32_t *neg, int64_t *pos_at, uint32_t *cur_after) {
  for (int c0 = 0; c0 < B; c0 += chunk) {
    const int c1 = c0 + chunk < B ? c0 + chunk : B;
    const size_t cur0 = g.cur;
    int i0 = c0;  // first triple not yet final
    while (i0 < c1) {
      // speculative draws: every candidate assumed "not in the list"
      bool all_draw = n_items >= D::kMinN && c1 - i0 <= kMaxChunk;
      for (int i = i0; i < c1; ++i) all_draw = all_draw && len[i] >= D::kMinN;
      if (all_draw) draw_chunk_words<D>(g, i0, c1, lo, len, n_items, pos_at, neg, cur_after, cur0);
      else draw_chunk<D>(g, i0, c1, lo, len, n_items, pos_at, neg, cur_after, cur0);
      int hit = -1;
      for (int i = i0; i < c1; ++i) {  // verification: independent accesses, misses overlap
        pos[i] = pos_at[i] >= 0 ? order[pos_at[i]] : 0;
        if (set.maybe(pair_hash((uint32_t)users[i], (uint32_t)neg[i])) && hit < 0 &&
            set.exact((uint32_t)users[i], neg[i]))
          hit = i;
      }
      if (hit < 0) break;
      // the candidate of triple `hit` is in the list: finish its rejection loop the literal way
      // and redraw everything after it
      g.cur = cur0 + cur_after[hit];
      {
        Reader r(g);
        for (;;) {
          const int32_t c = (int32_t)D::below(r, n_items);
          if (!set.contains((uint32_t)users[hit], c)) {
            neg[hit] = c;
            break;
          }
        }
      }
      i0 = hit + 1;
    }
  }
}

// ---- the verification pass on a second thread ------------------------------------------------
// The draws are one dependent chain, but the verification (two cache misses per triple) is not
// part of it: with sparse lists a second thread reads the positives and tests the candidates
// BEHIND the drawing thread, which keeps drawing speculatively.  A candidate found in its list
// is reported back; the drawing thread then returns to that triple, finishes its rejection loop
// and redraws what it had drawn beyond it (with one rejected candidate per ~1400 triples on
// gowalla that is a chunk or two).  Hand-over is by three atomics; the arrays of a batch are
// written by one side and read by the other strictly in that order.
inline void spin_wait(int &spins) {
  if (++spins < 4096) _mm_pause();
  else std::this_thread::yield();  // oversubscribed machine: let the other side run
}

struct Verifier {
  // the batch in flight (written by the drawing thread before `batch` is bumped)
  const PairSet *set = nullptr;
  const int32_t *users = nullptr, *order = nullptr, *neg = nullptr;
  const int64_t *pos_at = nullptr;
  int32_t *pos = nullptr;
  int B = 0;
  std::atomic<int> batch{0};     // bumped when a batch's arrays are ready (drawn == 0)
  std::atomic<int> drawn{0};     // triples of the batch drawn so far
  std::atomic<int> hit{-1};      // first triple whose candidate IS in its list; -1: none pending
  std::atomic<int> verified{0};  // == B once the whole batch is verified
  std::atomic<bool> quit{false};
  std::thread th;

  ~Verifier() { stop(); }  // also on the error returns of the entry points
  bool start() {  // false: no thread to be had -- the caller stays single-threaded
    try {
      th = std::thread([this] { run(); });
    } catch (...) {
      return false;
    }
    return true;
  }
  void stop() {
    quit.store(true, std::memory_order_release);
    if (th.joinable()) th.join();
  }
  void run() {
    int seen = 0;
    for (;;) {
      int spins = 0;
      while (batch.load(std::memory_order_acquire) == seen) {
        if (quit.load(std::memory_order_acquire)) return;
        spin_wait(spins);
      }
      seen = batch.load(std::memory_order_acquire);
      const int n = B;
      int v = 0;
      spins = 0;
      while (v < n) {
        const int d = drawn.load(std::memory_order_acquire);
        if (d <= v) {
          spin_wait(spins);
          continue;
        }
        spins = 0;
        int h = -1;
        for (int i = v; i < d; ++i) {
          pos[i] = pos_at[i] >= 0 ? order[pos_at[i]] : 0;
          if (set->maybe(pair_hash((uint32_t)users[i], (uint32_t)neg[i])) &&
              set->exact((uint32_t)users[i], neg[i])) {
            h = i;
            break;
          }
        }
        if (h < 0) {
          v = d;
          continue;
        }
        hit.store(h, std::memory_order_release);  // everything drawn after h is void
        while (hit.load(std::memory_order_acquire) >= 0) spin_wait(spins);
        v = h + 1;  // the drawing thread has settled triple h and reset `drawn` to h + 1
      }
      verified.store(n, std::memory_order_release);
    }
  }
};

// sample_items with the verification on `vf`'s thread
template <class D>
void sample_items_mt(Verifier &vf, WordStream &g, const PairSet &set, int chunk, int B,
                     const int32_t *users, const int64_t *lo, const uint32_t *len,
                     const int32_t *order, uint32_t n_items, int32_t *pos, int32_t *neg,
                     int64_t *pos_at, uint32_t *cur_after) {
  vf.set = &set, vf.users = users, vf.order = order, vf.neg = neg, vf.pos_at = pos_at, vf.pos = pos, vf.B = B;
  vf.drawn.store(0, std::memory_order_relaxed);
  vf.verified.store(0, std::memory_order_relaxed);
  vf.hit.store(-1, std::memory_order_relaxed);
  vf.batch.fetch_add(1, std::memory_order_release);
  const size_t cur0 = g.cur;
  // triple h's candidate is in its list: finish its rejection loop the literal way; -> next triple
  auto settle = [&](int h) {
    g.cur = cur0 + cur_after[h];
    {
      Reader r(g);
      for (;;) {
        const int32_t c = (int32_t)D::below(r, n_items);
        if (!set.contains((uint32_t)users[h], c)) {
          neg[h] = c;
          break;
        }
      }
    }
    cur_after[h] = (uint32_t)(g.cur - cur0);
    vf.drawn.store(h + 1, std::memory_order_release);  // before the verifier is let go
    vf.hit.store(-1, std::memory_order_release);
    return h + 1;
  };
  int i = 0;
  for (;;) {
    while (i < B) {
      const int c1 = i + chunk < B ? i + chunk : B;
      bool all_draw = n_items >= D::kMinN && c1 - i <= kMaxChunk;
      for (int k = i; k < c1; ++k) all_draw = all_draw && len[k] >= D::kMinN;
      if (all_draw) draw_chunk_words<D>(g, i, c1, lo, len, n_items, pos_at, neg, cur_after, cur0);
      else draw_chunk<D>(g, i, c1, lo, len, n_items, pos_at, neg, cur_after, cur0);
      vf.drawn.store(c1, std::memory_order_release);
      i = c1;
      const int h = vf.hit.load(std::memory_order_acquire);
      if (h >= 0) i = settle(h);
    }
    int spins = 0;
    for (;;) {  // everything is drawn: wait for the verifier's verdict on the rest
      const int h = vf.hit.load(std::memory_order_acquire);
      if (h >= 0) {
        i = settle(h);
        break;
      }
      if (vf.verified.load(std::memory_order_acquire) == B) return;
      spin_wait(spins);
    }
  }
}

// 0: single thread; 1: verification on a second thread.  Sparse lists only (a rejected candidate
// voids what was drawn behind it: with dense lists the second thread would mostly wait);
// MACR_SAMPLER_THREADS=1 / =2 forces the choice.
inline bool use_verifier_thread(int chunk, long long triples) {
  const char *e = getenv("MACR_SAMPLER_THREADS");
  if (e && e[0] == '1') return false;
  if (e && e[0] == '2') return true;
  return chunk >= 64 && triples >= 16384 && std::thread::hardware_concurrency() >= 2;
}

struct EpochScratch {
  std::vector<int64_t> lo, pos_at;
  std::vector<uint32_t> len, cur_after;
  PickScratch pick;
  explicit EpochScratch(int B) : lo((size_t)B), pos_at((size_t)B), len((size_t)B), cur_after((size_t)B) {}
};

}  // namespace
}  // namespace macr

// n_batches consecutive macr_sample_mf calls; out = int32 [n_batches][3][B] (users, pos, neg).
extern "C" int macr_sample_mf_epoch(uint32_t *py_state, const int32_t *users_pop, int n_pop,
                                    int n_users, int n_items, const int64_t *rowptr,
                                    const int32_t *order, const int32_t *sorted,
                                    const uint16_t *tags, int log2_buckets, int B, int n_batches,
                                    int32_t *out) {
  MACR_CHECK_ARG(py_state && users_pop && rowptr && order && sorted && tags && out,
                 "macr_sample_mf_epoch: null pointer");
  MACR_CHECK_ARG(n_pop > 0 && n_items > 0 && B > 0 && n_batches >= 0, "macr_sample_mf_epoch: empty population");
  MACR_CHECK_ARG(B > n_users || B <= n_pop, "macr_sample_mf_epoch: sample larger than population");
  MACR_CHECK_ARG(py_state[624] <= 624 && log2_buckets >= 1 && log2_buckets <= 40 && ((uintptr_t)tags & 15) == 0,
                 "macr_sample_mf_epoch: bad state / table");
  WordStream g(py_state);
  const PairSet set{tags, 64 - log2_buckets, rowptr, sorted};
  const int chunk = chunk_for((double)rowptr[n_users], n_users, n_items);
  EpochScratch s(B);
  Verifier vf;
  const bool mt = n_batches > 0 && use_verifier_thread(chunk, (long long)B * n_batches) && vf.start();
  for (int b = 0; b < n_batches; ++b) {
    int32_t *users = out + (size_t)b * 3 * B, *pos = users + B, *neg = pos + B;
    g.release_before(g.cur);
    if (!pick_users_words(g, users_pop, n_pop, B, n_users, users, &s.pick)) {
      Reader r(g);
      py_pick_users(r, users_pop, n_pop, B, n_users, users, &s.pick);
    }
    for (int i = 0; i < B; ++i) {  // independent loads: the misses overlap
      const int64_t l = rowptr[users[i]];
      s.lo[i] = l;
      s.len[i] = (uint32_t)(rowptr[users[i] + 1] - l);
    }
    if (mt)
      sample_items_mt<PyDraw>(vf, g, set, chunk, B, users, s.lo.data(), s.len.data(), order,
                              (uint32_t)n_items, pos, neg, s.pos_at.data(), s.cur_after.data());
    else
      sample_items<PyDraw>(g, set, chunk, B, users, s.lo.data(), s.len.data(), order,
                           (uint32_t)n_items, pos, neg, s.pos_at.data(), s.cur_after.data());
  }
  if (mt) vf.stop();
  g.store_state();
  return MACR_OK;
}

// n_batches consecutive macr_sample_lgcn calls; out = int32 [n_batches][3][B].  ban_tags: pair
// set of (user, banned id) over ban_rowptr / ban_sorted.
extern "C" int macr_sample_lgcn_epoch(uint32_t *py_state, uint32_t *np_state,
                                      const int32_t *users_pop, int n_pop, int n_users,
                                      int n_items, const int64_t *pos_rowptr,
                                      const int32_t *pos_order, const int64_t *ban_rowptr,
                                      const int32_t *ban_sorted, const uint16_t *ban_tags,
                                      int log2_buckets, int B, int n_batches, int32_t *out) {
  MACR_CHECK_ARG(py_state && np_state && users_pop && pos_rowptr && pos_order && ban_rowptr &&
                     ban_sorted && ban_tags && out,
                 "macr_sample_lgcn_epoch: null pointer");
  MACR_CHECK_ARG(n_pop > 0 && n_items > 0 && B > 0 && n_batches >= 0, "macr_sample_lgcn_epoch: empty population");
  MACR_CHECK_ARG(B > n_users || B <= n_pop, "macr_sample_lgcn_epoch: sample larger than population");
  MACR_CHECK_ARG(py_state[624] <= 624 && np_state[624] <= 624 && log2_buckets >= 1 && log2_buckets <= 40 &&
                     ((uintptr_t)ban_tags & 15) == 0,
                 "macr_sample_lgcn_epoch: bad state / table");
  WordStream gp(py_state), gn(np_state);
  const PairSet set{ban_tags, 64 - log2_buckets, ban_rowptr, ban_sorted};
  const int chunk = chunk_for((double)ban_rowptr[n_users], n_users, n_items);
  EpochScratch s(B);
  Verifier vf;
  const bool mt = n_batches > 0 && use_verifier_thread(chunk, (long long)B * n_batches) && vf.start();
  for (int b = 0; b < n_batches; ++b) {
    int32_t *users = out + (size_t)b * 3 * B, *pos = users + B, *neg = pos + B;
    gp.release_before(gp.cur);
    gn.release_before(gn.cur);
    if (!pick_users_words(gp, users_pop, n_pop, B, n_users, users, &s.pick)) {
      Reader r(gp);
      py_pick_users(r, users_pop, n_pop, B, n_users, users, &s.pick);
    }
    for (int i = 0; i < B; ++i) {
      const int64_t l = pos_rowptr[users[i]];
      s.lo[i] = l;
      s.len[i] = (uint32_t)(pos_rowptr[users[i] + 1] - l);
      MACR_CHECK_ARG(s.len[i] > 0, "macr_sample_lgcn_epoch: user %d has no positive item", (int)users[i]);
    }
    if (mt)
      sample_items_mt<NpDraw>(vf, gn, set, chunk, B, users, s.lo.data(), s.len.data(), pos_order,
                              (uint32_t)n_items, pos, neg, s.pos_at.data(), s.cur_after.data());
    else
      sample_items<NpDraw>(gn, set, chunk, B, users, s.lo.data(), s.len.data(), pos_order,
                           (uint32_t)n_items, pos, neg, s.pos_at.data(), s.cur_after.data());
  }
  if (mt) vf.stop();
  gp.store_state();
  gn.store_state();
  return MACR_OK;
}

// Developer / test hook (not in the public header): the word stream alone.  Consumes `advance`
// words from `state`, moves the cursor `back` words back (<= advance) and writes the generator
// state of that position -- what a draw-by-draw consumer holds after advance - back words, also
// when blocks beyond the final position were already generated (the state is then rebuilt from
// the buffered output words).
extern "C" int macr_sampler_stream_selftest(uint32_t *state /*[625]*/, int64_t advance, int64_t back,
                                            uint32_t *xor_of_words) {
  MACR_CHECK_ARG(state && state[624] <= 624 && advance >= 0 && back >= 0 && back <= advance,
                 "macr_sampler_stream_selftest: bad arguments");
  WordStream g(state);
  uint32_t x = 0;
  for (int64_t k = 0; k < advance; ++k) x ^= g.next();
  g.cur -= (size_t)back;
  g.store_state();
  if (xor_of_words) *xor_of_words = x;
  return MACR_OK;
}

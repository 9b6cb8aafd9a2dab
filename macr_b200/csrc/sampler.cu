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
#include <math.h>

#include <algorithm>
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
inline uint32_t py_randbelow(MT &g, uint32_t n) {
  const int k = bit_length(n);
  uint32_t r = g.next() >> (32 - k);
  while (r >= n) r = g.next() >> (32 - k);
  return r;
}

// numpy legacy RandomState.randint(0, n, size=1)[0]
inline uint32_t np_randint(MT &g, uint32_t n) {
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
void py_pick_users(MT &g, const int32_t *pop, int n, int B, int n_users_flag, int32_t *out) {
  if (B <= n_users_flag) {  // rd.sample(pop, B)
    long long setsize = 21;
    if (B > 5) setsize += (long long)llround(pow(4.0, ceil(log((double)B * 3.0) / log(4.0))));
    if (n <= setsize) {
      std::vector<int32_t> pool(pop, pop + n);
      for (int i = 0; i < B; ++i) {
        const uint32_t j = py_randbelow(g, (uint32_t)(n - i));
        out[i] = pool[j];
        pool[j] = pool[n - i - 1];
      }
    } else {
      std::vector<uint8_t> chosen((size_t)n, 0);
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

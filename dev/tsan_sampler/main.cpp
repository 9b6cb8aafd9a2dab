#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <algorithm>
extern "C" int macr_pairset_build(const int64_t *, const int32_t *, int, uint16_t *, int);
extern "C" int macr_sample_mf_epoch(uint32_t *, const int32_t *, int, int, int, const int64_t *, const int32_t *, const int32_t *, const uint16_t *, int, int, int, int32_t *);
int main() {
  const int n_users = 3000, n_items = 500;
  std::vector<int64_t> rowptr(n_users + 1, 0);
  std::vector<int32_t> order, sorted;
  uint32_t x = 12345;
  auto rnd = [&]() { x = x * 1664525u + 1013904223u; return x >> 8; };
  for (int u = 0; u < n_users; ++u) {
    int len = 1 + rnd() % 40;
    std::vector<int32_t> l;
    while ((int)l.size() < len) { int v = rnd() % n_items; if (std::find(l.begin(), l.end(), v) == l.end()) l.push_back(v); }
    order.insert(order.end(), l.begin(), l.end());
    std::sort(l.begin(), l.end());
    sorted.insert(sorted.end(), l.begin(), l.end());
    rowptr[u + 1] = (int64_t)order.size();
  }
  int log2b = 13;
  uint16_t *tags = (uint16_t *)aligned_alloc(16, (8 << log2b) * 2);
  macr_pairset_build(rowptr.data(), sorted.data(), n_users, tags, log2b);
  std::vector<int32_t> pop(n_users);
  for (int i = 0; i < n_users; ++i) pop[i] = i;
  const int B = 1024, nb = 40;
  std::vector<int32_t> out1((size_t)nb * 3 * B), out2((size_t)nb * 3 * B);
  uint32_t st1[625], st2[625];
  for (int i = 0; i < 624; ++i) st1[i] = st2[i] = rnd() * 2654435761u + i;
  st1[624] = st2[624] = 624;
  setenv("MACR_SAMPLER_THREADS", "1", 1);
  macr_sample_mf_epoch(st1, pop.data(), n_users, n_users, n_items, rowptr.data(), order.data(), sorted.data(), tags, log2b, B, nb, out1.data());
  setenv("MACR_SAMPLER_THREADS", "2", 1);
  macr_sample_mf_epoch(st2, pop.data(), n_users, n_users, n_items, rowptr.data(), order.data(), sorted.data(), tags, log2b, B, nb, out2.data());
  printf("equal triples: %d  equal state: %d\n", out1 == out2, memcmp(st1, st2, sizeof(st1)) == 0);
  return !(out1 == out2 && memcmp(st1, st2, sizeof(st1)) == 0);
}

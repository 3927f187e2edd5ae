// Exhaustive check of pyg2p_b200/csrc/g2p_sincosf.h against the C library's sinf / cosf: all 2^32 float32 bit
// patterns (or every `stride`-th), both glibc build variants.  Prints the number of differing results per variant.
//   g++ -O2 -ffp-contract=off -mfma -pthread -I pyg2p_b200/csrc tools/sincosf_check.cpp -o /tmp/sincosf_check && /tmp/sincosf_check [stride]
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "g2p_sincosf.h"

static inline uint32_t bits(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
}
static inline bool same(float a, float b) { return bits(a) == bits(b) || (a != a && b != b); }

int main(int argc, char** argv) {
  const uint64_t stride = argc > 1 ? strtoull(argv[1], nullptr, 10) : 1;
  const unsigned nt = std::thread::hardware_concurrency() ? std::thread::hardware_concurrency() : 8;
  std::atomic<uint64_t> bad[4] = {{0}, {0}, {0}, {0}};
  std::atomic<uint64_t> total{0};
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nt; ++t)
    th.emplace_back([&, t] {
      uint64_t b[4] = {0, 0, 0, 0}, n = 0;
      for (uint64_t u = t * stride; u < (1ull << 32); u += nt * stride) {
        float y;
        const uint32_t w = (uint32_t)u;
        memcpy(&y, &w, 4);
        const float s = sinf(y), c = cosf(y);
        b[0] += !same(g2p_scf::sincosf_glibc<true, false>(y), s);
        b[1] += !same(g2p_scf::sincosf_glibc<true, true>(y), c);
        b[2] += !same(g2p_scf::sincosf_glibc<false, false>(y), s);
        b[3] += !same(g2p_scf::sincosf_glibc<false, true>(y), c);
        ++n;
      }
      for (int i = 0; i < 4; ++i) bad[i] += b[i];
      total += n;
    });
  for (auto& x : th) x.join();
  printf("arguments checked: %llu\n", (unsigned long long)total.load());
  printf("fma variant    : sinf differs %llu, cosf differs %llu\n", (unsigned long long)bad[0].load(), (unsigned long long)bad[1].load());
  printf("non-fma variant: sinf differs %llu, cosf differs %llu\n", (unsigned long long)bad[2].load(), (unsigned long long)bad[3].load());
  return 0;
}

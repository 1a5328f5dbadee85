// oracle/philox.hpp — TEST INFRASTRUCTURE ONLY.
//
// Philox4x32-10 (Salmon, Moraes, Dror, Shaw, "Parallel random numbers: as easy as 1, 2, 3",
// SC'11) written from the published round function; checked against the paper's known-answer
// vectors in tests/test_oracle_rng.py.  The reference draws from Julia's global RNG
// (MainRunner.jl:365-372), whose stream cannot be reproduced outside Julia; this project
// defines its own counter-based streams:
//   key     = (seed_lo, seed_hi)
//   counter = (block, attempt, event_lo, event_hi | domain << 28)
// domain 0 = surface sampler (5 blocks → 10 uniforms per attempt), 1 = tree Monte-Carlo rand().
#pragma once
#include <cstdint>

namespace orc {

inline void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// 53-bit uniform in [0,1) from two words
inline double u53(uint32_t a, uint32_t b) {
  return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}

// two uniforms of block `block` of (domain, event, attempt)
inline void philox_pair(uint64_t seed, uint32_t domain, uint64_t event, uint32_t attempt,
                        uint32_t block, double* ua, double* ub) {
  uint32_t ctr[4] = {block, attempt, (uint32_t)event, (uint32_t)((event >> 32) & 0x0FFFFFFFu) | (domain << 28)};
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  uint32_t o[4];
  philox4x32_10(ctr, key, o);
  *ua = u53(o[0], o[1]);
  *ub = u53(o[2], o[3]);
}

// the 10 uniforms of one sampler attempt (find_samples_new's draw order, RT.jl:1486-1531,1623)
inline void sampler_uniforms(uint64_t seed, uint64_t event, uint32_t attempt, double U[10]) {
  for (uint32_t b = 0; b < 5; ++b) philox_pair(seed, 0, event, attempt, b, &U[2 * b], &U[2 * b + 1]);
}

// k-th rand() of the tree Monte-Carlo fallback of event `event` (MainRunner.jl:282)
inline double tree_uniform(uint64_t seed, uint64_t event, uint32_t k) {
  double a, b;
  philox_pair(seed, 1, event, k, 0, &a, &b);
  return a;
}

}  // namespace orc

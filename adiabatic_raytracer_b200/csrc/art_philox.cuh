// art_philox.cuh — per-event counter-based RNG streams (Philox4x32-10, Salmon et al., SC'11).
//
// The reference seeds Julia's global RNG once (MainRunner.jl:365-372) and consumes it
// sequentially; that stream neither parallelises nor exists outside Julia.  Here every random
// number is a pure function of (seed, event, attempt, block):
//   key     = (seed_lo, seed_hi)
//   counter = (block, attempt, event_lo, event_hi[27:0] | domain << 28)
// domain 0: surface sampler — 5 blocks × 2 doubles = the 10 uniforms find_samples_new draws per
//           attempt (RT.jl:1486-1497, 1531, 1623);  domain 1: get_tree's Monte-Carlo rand()
//           (MainRunner.jl:282), attempt = draw index.
// Results therefore do not depend on how events are spread over threads, launches or GPUs.
#pragma once
#include <stdint.h>

#include "art_physics.cuh"

namespace art {

ART_HD void mulhilo32(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
#if defined(__CUDA_ARCH__)
  lo = a * b;
  hi = __umulhi(a, b);
#else
  const uint64_t p = (uint64_t)a * (uint64_t)b;
  lo = (uint32_t)p;
  hi = (uint32_t)(p >> 32);
#endif
}

ART_HD void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                          uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t h0, l0, h1, l1;
    mulhilo32(0xD2511F53u, c0, h0, l0);
    mulhilo32(0xCD9E8D57u, c2, h1, l1);
    const uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
    c0 = n0; c1 = l1; c2 = n2; c3 = l0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// 53-bit uniform in [0,1)
ART_HD double uniform53(uint32_t a, uint32_t b) {
  return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}

ART_HD void philox_pair(uint64_t seed, uint32_t domain, uint64_t event, uint32_t attempt, uint32_t block,
                        double& ua, double& ub) {
  uint32_t o[4];
  philox4x32_10(block, attempt, (uint32_t)event, ((uint32_t)(event >> 32) & 0x0FFFFFFFu) | (domain << 28),
                (uint32_t)seed, (uint32_t)(seed >> 32), o);
  ua = uniform53(o[0], o[1]);
  ub = uniform53(o[2], o[3]);
}

ART_HD void sampler_uniforms(uint64_t seed, uint64_t event, uint32_t attempt, double U[10]) {
#pragma unroll
  for (uint32_t b = 0; b < 5; ++b) philox_pair(seed, 0u, event, attempt, b, U[2 * b], U[2 * b + 1]);
}

ART_HD double tree_uniform(uint64_t seed, uint64_t event, uint32_t k) {
  double a, b;
  philox_pair(seed, 1u, event, k, 0u, a, b);
  return a;
}

}  // namespace art

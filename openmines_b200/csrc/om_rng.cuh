// Counter-based Philox4x32-10 (Salmon et al., SC'11) and the draw transforms of the engine.
//
// Keying: key = 64-bit seed of the instance; counter = (epoch, entity, site, stream).
// `site` enumerates the reference's draw sites: every np.random.* / random.*
// call on the hot path has its own site id, so draws are order-independent and reproducible
// for a (seed, instance) regardless of which GPU or warp runs it.
#pragma once
#include <stdint.h>

#define OM_S_LOAD_JITTER 0   // truck.py:230  np.random.uniform(-0.1, 0.1)
#define OM_S_MAINT_GAP 1     // truck.py:358  np.random.exponential(1440)
#define OM_S_MAINT_DUR 2     // truck.py:362  np.random.normal(45, 5)
#define OM_S_BREAK_GAP 3     // truck.py:528  np.random.exponential(360)
#define OM_S_BREAK_DUR 4     // truck.py:532  np.random.normal(10, 3)
#define OM_S_UNREPAIRABLE 5  // truck.py:545  random.expovariate(1/70560)
#define OM_S_ROAD_GAP 6      // road.py:253   random.expovariate(lambda_repair)
#define OM_S_ROAD_DUR 7      // road.py:256   np.random.normal(mu, sigma)
#define OM_S_DETOUR 8        // road.py:329   np.random.normal(0.4, 0.1)
#define OM_S_JAM_COIN 9      // road.py:179   random.random()
#define OM_S_JAM_POS 10      // road.py:203   np.random.choice(1000, p)
#define OM_S_JAM_DUR 11      // road.py:212   np.random.weibull(2)
#define OM_S_DISPATCH 12     // random_dispatcher.py:16,22,29  RandomState.randint

struct Philox4 {
  uint32_t x, y, z, w;
};

// UNROLL: trips of the 10-round loop per loop iteration.  The engine is bound by instruction fetch and dependent-issue
// latency, so this is a measured choice per kernel family (same-call A/B): the fully unrolled body
// (~110 instructions, 14 cache lines) costs the north_pit episode kernel 4 %, two rounds per trip gain it 0.5 %; the
// split kernels of the big mines (7 warps per SM, little fetch pressure) gain 2.7 % from the full unroll.
template <int UNROLL = 2>
__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                 uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll UNROLL
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0;
    uint32_t n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += W0; k1 += W1;
  }
  Philox4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3;
  return o;
}

// uniform in (0, 1], 24 bits
__device__ __forceinline__ float om_u01(uint32_t x) { return (float)((x >> 8) + 1u) * 5.9604644775390625e-08f; }

// -ln(u), u in (0,1]
__device__ __forceinline__ float om_exp1(uint32_t x) { return -__logf(om_u01(x)); }

// standard normal from two 32-bit words (Box-Muller on the MUFU log2 / rsqrt / cos units; live-mode draws are
// distribution-level only): ~12 instructions inline
__device__ __forceinline__ float om_normal_fast_inl(uint32_t a, uint32_t b) {
  float x = -2.0f * __logf(om_u01(a));
  float r = x * rsqrtf(fmaxf(x, 1e-30f));
  return r * __cosf(6.283185307179586f * om_u01(b) - 3.14159265358979f);   // cos(2 pi u - pi) = -cos(2 pi u): same law
}
// one out-of-line copy (the kernels choose per family which of the two they use, see Sim::normal_fast)
__device__ __noinline__ float om_normal_fast(uint32_t a, uint32_t b) { return om_normal_fast_inl(a, b); }

// bounded integer in [0, n): Lemire multiply-shift (no rejection); mirrored by oracle/philox.py
__device__ __forceinline__ uint32_t om_randint(uint32_t x, uint32_t n) { return __umulhi(x, n); }

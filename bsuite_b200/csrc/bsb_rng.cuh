// Per-lane random streams: a bit source (Philox4x64-10 or MT19937) under
// numpy's *legacy* RandomState distribution algorithms.
//
// The reference draws every random number from numpy.random.RandomState
// (deep_sea.py:77, catch.py:58, cartpole.py:91, mountain_car.py:55,
// memory_chain.py:45, umbrella_chain.py:52, mnist.py:53, wrappers.py:267,330).
// numpy is a third-party dependency of the reference (setup.py:85, unpinned;
// 2.3.5 in this image); the algorithms restated here are the ones SURVEY.md 8a
// "RNG draw table" lists:
//   rand()            next_double of the bit generator
//   uniform(lo, hi)   lo + (hi - lo) * next_double
//   binomial(1, .5)   inversion, which for n=1, p=.5 reduces to (u > 0.5)
//   randint(n)        masked rejection on next_uint32; randint(1) draws nothing
//   randn()           Marsaglia polar with the second variate cached
// Parity of these restatements is pinned by tests against numpy itself.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define BSB_HD __host__ __device__ __forceinline__
#else
#define BSB_HD inline
#endif

namespace bsb {

typedef uint64_t u64;
typedef uint32_t u32;

BSB_HD u64 mulhi64(u64 a, u64 b) {
#if defined(__CUDA_ARCH__)
  return __umul64hi(a, b);
#else
  return (u64)(((unsigned __int128)a * (unsigned __int128)b) >> 64);
#endif
}

// ---------------------------------------------------------------------------
// Philox4x64-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3",
// SC'11), with the constants and word order of numpy.random.Philox.
// ---------------------------------------------------------------------------
struct PhiloxBlock { u64 v0, v1, v2, v3; };

BSB_HD PhiloxBlock philox4x64_10(u64 c0, u64 c1, u64 c2, u64 c3, u64 k0, u64 k1) {
  const u64 M0 = 0xD2E7470EE14C6C93ull, M1 = 0xCA5A826395121157ull;
  const u64 W0 = 0x9E3779B97F4A7C15ull, W1 = 0xBB67AE8584CAA73Bull;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int r = 0; r < 10; ++r) {
    const u64 hi0 = mulhi64(M0, c0), lo0 = M0 * c0;
    const u64 hi1 = mulhi64(M1, c2), lo1 = M1 * c2;
    const u64 n0 = hi1 ^ c1 ^ k0;
    const u64 n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += W0; k1 += W1;
  }
  PhiloxBlock b; b.v0 = c0; b.v1 = c1; b.v2 = c2; b.v3 = c3;
  return b;
}

// Stream ids carried in counter word 3.
enum : u64 { STREAM_ENV = 0, STREAM_WRAPPER = 1, STREAM_ACTIONS = 2 };

// On-device uniform random actions (the workload of baselines/random/agent.py:35-37).
// The action stream of a lane is the Philox stream (key = (action_seed, global lane), counter word 3 =
// STREAM_ACTIONS) read as 32-bit chunks: step s uses chunk (s & 7) of block (s >> 3); a chunk r maps to
// floor(r * n / 2^32) (multiply-shift, no rejection).  One block serves 8 consecutive steps of a lane.
struct ActionStream {
  u64 blk_index;          // block currently cached (~0 = none)
  PhiloxBlock blk;
  BSB_HD void open() { blk_index = ~0ull; blk.v0 = blk.v1 = blk.v2 = blk.v3 = 0; }
  BSB_HD int32_t sample(u64 action_seed, u64 global_lane, u64 step, int32_t n) {
    const u64 want = step >> 3;
    if (want != blk_index) { blk = philox4x64_10(want, 0, 0, STREAM_ACTIONS, action_seed, global_lane); blk_index = want; }
    const u32 c = (u32)(step & 7);
    const u64 w = (c >> 1) == 0 ? blk.v0 : ((c >> 1) == 1 ? blk.v1 : ((c >> 1) == 2 ? blk.v2 : blk.v3));
    const u32 r = (c & 1) ? (u32)(w >> 32) : (u32)w;
    return (int32_t)(((u64)r * (u64)(u32)n) >> 32);
  }
};

BSB_HD int32_t sample_action(u64 action_seed, u64 global_lane, u64 step, int32_t n) {
  ActionStream s; s.open();
  return s.sample(action_seed, global_lane, step, n);
}

// ---------------------------------------------------------------------------
// Bit source 1: Philox stream with numpy's buffering semantics.
//   numpy keeps a 4-word buffer; the counter is incremented BEFORE a block is
//   generated, so word w of the stream is word (w & 3) of the block at counter
//   (w >> 2) + 1.  next_uint32 returns the low half of a fresh 64-bit word and
//   keeps the high half for the next call.
// Persistent state per lane is ONE u64:
//   bits  0..53  words consumed so far
//   bits 54..61  d: a saved high half is pending from word (pos - d); 0 = none.
//                (next_uint64 calls do not disturb numpy's saved half, so it can
//                lag behind pos: by at most num_bits + 1 <= 65 words here.)
//   bit  62      a cached gaussian is pending (value kept in a separate array)
// ---------------------------------------------------------------------------
static const u64 RNG_HASGAUSS = 1ull << 62;
static const u64 RNG_POSMASK = (1ull << 54) - 1;
static const int RNG_LAG_SHIFT = 54;
static const u64 RNG_LAG_MAX = 255;

struct PhiloxSrc {
  u64 k0, k1, stream;
  u64 pos;          // 64-bit words consumed
  u64 pend;         // 1 + index of the word whose high half is saved (0 = none)
  u64 blk;          // counter value of the buffered block (0 = none)
  u64 b0, b1, b2, b3;

  BSB_HD void open(u64 seed, u64 global_lane, u64 stream_id, u64 packed) {
    k0 = seed; k1 = global_lane; stream = stream_id;
    pos = packed & RNG_POSMASK;
    const u64 lag = (packed >> RNG_LAG_SHIFT) & RNG_LAG_MAX;
    pend = lag ? (pos - lag + 1) : 0;
    blk = 0; b0 = b1 = b2 = b3 = 0;
  }
  BSB_HD u64 packed() const {
    u64 lag = pend ? (pos - (pend - 1)) : 0;
    if (lag > RNG_LAG_MAX) lag = 0;   // unreachable for the supported families (see header)
    return pos | (lag << RNG_LAG_SHIFT);
  }
  static BSB_HD u64 pick(const PhiloxBlock& b, u32 i) { return i == 0 ? b.v0 : (i == 1 ? b.v1 : (i == 2 ? b.v2 : b.v3)); }

  BSB_HD u64 next64() {
    const u64 want = (pos >> 2) + 1;
    if (want != blk) {
      const PhiloxBlock b = philox4x64_10(want, 0, 0, stream, k0, k1);
      b0 = b.v0; b1 = b.v1; b2 = b.v2; b3 = b.v3; blk = want;
    }
    const u32 i = (u32)(pos & 3);
    ++pos;
    return i == 0 ? b0 : (i == 1 ? b1 : (i == 2 ? b2 : b3));
  }
  BSB_HD u32 next32() {
    if (pend) {
      const u64 w = pend - 1;
      pend = 0;
      const u64 want = (w >> 2) + 1;
      if (want == blk) { const u32 i = (u32)(w & 3); return (u32)((i == 0 ? b0 : (i == 1 ? b1 : (i == 2 ? b2 : b3))) >> 32); }
      return (u32)(pick(philox4x64_10(want, 0, 0, stream, k0, k1), (u32)(w & 3)) >> 32);
    }
    const u64 v = next64();
    pend = pos;   // word index pos - 1, stored + 1
    return (u32)v;
  }
  BSB_HD double next_double() { return (double)(next64() >> 11) * (1.0 / 9007199254740992.0); }
  // next_double() > 0.5 without floating point: (w >> 11) * 2^-53 > 1/2  <=>  (w >> 11) > 2^52  <=>  w >= 2^63 + 2^11
  BSB_HD bool next_above_half() { return next64() >= 0x8000000000000800ull; }

  // The next `n` (<= 64) Bernoulli(1/2) draws as bits of the result (draw k -> bit k): binomial(1, .5, size=n).
  // Once the position is block-aligned, TWO Philox blocks (8 draws) are computed per iteration; the ten-round
  // chains are independent, so the scheduler interleaves them and the integer-multiply latency that bounds a
  // single chain is overlapped (umbrella_chain draws up to 100 of these per lane-step, memory_chain 40 per reset).
  // Measured: memory_size/16 23.0 -> 11.0 us/step, umbrella_distract/22 36.9 -> 29.3; four chains at once gained
  // nothing more and cost registers.
  static BSB_HD u64 nibble(const PhiloxBlock& b, u64 T) {
    return (u64)(b.v0 >= T) | ((u64)(b.v1 >= T) << 1) | ((u64)(b.v2 >= T) << 2) | ((u64)(b.v3 >= T) << 3);
  }
  BSB_HD u64 next_half_bits(int n) {
    const u64 T = 0x8000000000000800ull;
    u64 bits = 0;
    int k = 0;
    while (k < n && (pos & 3) != 0) { bits |= (u64)next_above_half() << k; ++k; }
    while (n - k >= 8) {
      const u64 c = (pos >> 2) + 1;
      const PhiloxBlock x = philox4x64_10(c, 0, 0, stream, k0, k1);
      const PhiloxBlock y = philox4x64_10(c + 1, 0, 0, stream, k0, k1);
      bits |= (nibble(x, T) | (nibble(y, T) << 4)) << k;
      k += 8; pos += 8;
    }
    while (k < n) { bits |= (u64)next_above_half() << k; ++k; }
    return bits;
  }
};

// ---------------------------------------------------------------------------
// Bit source 2: MT19937 with numpy's legacy integer seeding.  The 624-word key
// lives in caller memory with element stride `stride` (device: [624][B]
// lane-minor so that lanes at the same index coalesce; host: stride 1).
// ---------------------------------------------------------------------------
struct MtSrc {
  u32* key; int64_t stride; int32_t idx;

  BSB_HD void open(u32* key_, int64_t stride_, int32_t idx_) { key = key_; stride = stride_; idx = idx_; }
  BSB_HD u32& at(int i) { return key[(int64_t)i * stride]; }

  static BSB_HD u32 twist(u32 cur, u32 nxt, u32 far) {
    const u32 y = (cur & 0x80000000u) | (nxt & 0x7fffffffu);
    return far ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
  }
  BSB_HD void regenerate() {
    for (int k = 0; k < 624 - 397; ++k) at(k) = twist(at(k), at(k + 1), at(k + 397));
    for (int k = 624 - 397; k < 623; ++k) at(k) = twist(at(k), at(k + 1), at(k - (624 - 397)));
    at(623) = twist(at(623), at(0), at(396));
    idx = 0;
  }
  BSB_HD u32 next32() {
    if (idx >= 624) regenerate();
    u32 y = at(idx++);
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
  }
  BSB_HD double next_double() {
    const u32 a = next32() >> 5, b = next32() >> 6;
    return ((double)a * 67108864.0 + (double)b) / 9007199254740992.0;
  }
  // next_double() > 0.5: the numerator a * 2^26 + b is an exact 53-bit integer; compare it with 2^52
  BSB_HD bool next_above_half() {
    const u32 a = next32() >> 5, b = next32() >> 6;
    return (((u64)a << 26) | (u64)b) > (1ull << 52);
  }
  BSB_HD u64 next_half_bits(int n) {
    u64 bits = 0;
    for (int k = 0; k < n; ++k) bits |= (u64)next_above_half() << k;
    return bits;
  }
};

// numpy's legacy seeding of MT19937 from one 32-bit integer.
inline void mt19937_seed_host(u32* key, int64_t stride, u32 seed) {
  for (int i = 0; i < 624; ++i) {
    key[(int64_t)i * stride] = seed;
    seed = 1812433253u * (seed ^ (seed >> 30)) + (u32)i + 1u;
  }
}

// ---------------------------------------------------------------------------
// numpy legacy distributions over either bit source.  `Gauss` is the cached
// second variate of the polar method (RandomState's has_gauss / gauss).
// ---------------------------------------------------------------------------
struct GaussCache { int has; double value; };

template <class Src>
struct LegacyRng {
  Src src;
  GaussCache g;

  BSB_HD double rand() { return src.next_double(); }

  BSB_HD double uniform(double low, double high) {
    const double range = high - low;
    return low + range * src.next_double();
  }
  // binomial(n=1, p=0.5): inversion with qn = exp(log(0.5)) = 0.5, bound = 1.
  BSB_HD int binomial_half() { return src.next_above_half() ? 1 : 0; }
  // binomial(1, 0.5, size=n), n <= 64, draw k in bit k
  BSB_HD u64 binomial_half_bits(int n) { return src.next_half_bits(n); }

  // randint(n) for 1 <= n <= 2^32.
  BSB_HD u32 randint(u32 n) {
    const u32 rng = n - 1u;
    if (rng == 0u) return 0u;
    u32 mask = rng;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
    u32 v;
    do { v = src.next32() & mask; } while (v > rng);
    return v;
  }
  BSB_HD double randn() {
    if (g.has) { g.has = 0; const double t = g.value; g.value = 0.0; return t; }
    double x1, x2, r2;
    do {
      x1 = 2.0 * src.next_double() - 1.0;
      x2 = 2.0 * src.next_double() - 1.0;
      r2 = x1 * x1 + x2 * x2;
    } while (r2 >= 1.0 || r2 == 0.0);
    const double f = sqrt(-2.0 * log(r2) / r2);
    g.value = f * x1; g.has = 1;
    return f * x2;
  }
};

}  // namespace bsb

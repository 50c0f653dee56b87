// Fused transition kernels: one per environment family (template F), each
// specialised on the bit source (Philox / MT19937) and on whether the
// RewardNoise wrapper stream is live.
//
// Thread = lane for the scalar transition (state word(s), action, reward,
// discount, step_type are all coalesced 4/8-byte accesses).  Observations are
// dense float32 tensors that must be written fresh every step (the reference
// allocates a new array per step: deep_sea.py:104, catch.py:114), and they are
// the HBM traffic that bounds the kernel, so they are emitted WARP-
// COOPERATIVELY with 16-byte streaming stores:
//   * tile families (deep_sea N x N one-hot, mnist 28 x 28): the warp walks its
//     32 lanes; for each lane all 32 threads write that lane's contiguous tile,
//     the hot cell (or the gathered image) decided per float4 from a descriptor
//     broadcast with __shfl_sync.
//   * catch: the warp's 32 boards are one contiguous span; each float4 is
//     rendered from the (ball, paddle) cells of the lane(s) it overlaps.
//   * row families ((1,k) vectors): each thread renders its row into a per-warp
//     shared-memory stage, the warp then streams the contiguous [32, k] block.
// A launch covers T consecutive steps with lane state held in registers
// (T = 1 for bsb_step); actions come from the caller or from the on-device
// Philox action stream.
#pragma once
#include "bsb_families.cuh"

namespace bsb {

struct LaunchArgs {
  const int32_t* actions;   // [T,B] or null (sample on device)
  int32_t* actions_out;     // [T,B] or null
  float* obs;               // [T,B,K]
  float* reward;            // [T,B] or null
  double* reward_f64;       // [T,B] or null
  float* discount;          // [T,B] or null
  int32_t* step_type;       // [T,B] or null
  int64_t T;
  int64_t step0;            // global index of the first step of this launch
  uint64_t action_seed;
  int32_t mode;             // 0 = step, 1 = reset every lane, 2 = constructor init
  int32_t obs_vec_ok;       // obs base and per-step stride are 16-byte aligned
};

enum { MODE_STEP = 0, MODE_RESET = 1, MODE_INIT = 2 };

// ----- RNG plumbing ---------------------------------------------------------
template <int RK> struct RngOf;
template <> struct RngOf<0> { typedef LegacyRng<PhiloxSrc> type; };
template <> struct RngOf<1> { typedef LegacyRng<MtSrc> type; };

BSB_HD void rng_open(LegacyRng<PhiloxSrc>& r, const EnvParams& p, int64_t i, bool wrapper) {
  const uint64_t packed = wrapper ? p.wrng_pos[i] : p.rng_pos[i];
  r.src.open(p.seed, p.lane_offset + (uint64_t)i, wrapper ? STREAM_WRAPPER : STREAM_ENV, packed);
  r.g.has = (packed & RNG_HASGAUSS) ? 1 : 0;
  const double* gz = wrapper ? p.wrng_gauss : p.rng_gauss;
  r.g.value = (r.g.has && gz) ? gz[i] : 0.0;
}
BSB_HD void rng_close(const LegacyRng<PhiloxSrc>& r, const EnvParams& p, int64_t i, bool wrapper) {
  const uint64_t packed = r.src.packed() | (r.g.has ? RNG_HASGAUSS : 0ull);
  if (wrapper) p.wrng_pos[i] = packed; else p.rng_pos[i] = packed;
  double* gz = wrapper ? p.wrng_gauss : p.rng_gauss;
  if (gz && r.g.has) gz[i] = r.g.value;
}
BSB_HD void rng_open(LegacyRng<MtSrc>& r, const EnvParams& p, int64_t i, bool wrapper) {
  const int64_t stride = (p.batch > 0) ? p.batch : 1;
  r.src.open((wrapper ? p.wmt_key : p.mt_key) + i, stride, (wrapper ? p.wmt_idx : p.mt_idx)[i]);
  const uint64_t packed = wrapper ? p.wrng_pos[i] : p.rng_pos[i];
  r.g.has = (packed & RNG_HASGAUSS) ? 1 : 0;
  const double* gz = wrapper ? p.wrng_gauss : p.rng_gauss;
  r.g.value = (r.g.has && gz) ? gz[i] : 0.0;
}
BSB_HD void rng_close(const LegacyRng<MtSrc>& r, const EnvParams& p, int64_t i, bool wrapper) {
  (wrapper ? p.wmt_idx : p.mt_idx)[i] = r.src.idx;
  const uint64_t packed = r.g.has ? RNG_HASGAUSS : 0ull;
  if (wrapper) p.wrng_pos[i] = packed; else p.rng_pos[i] = packed;
  double* gz = wrapper ? p.wrng_gauss : p.rng_gauss;
  if (gz && r.g.has) gz[i] = r.g.value;
}

// ----- the per-lane call sequence of base.Environment.step (base.py:59-65) --
template <class F, class R, class WR>
BSB_HD StepOut lane_transition(const EnvParams& p, int64_t i, typename F::Lane& L, R& rng, WR& wrng,
                               int32_t action, int32_t mode, bool noise) {
  StepOut o;
  if (mode == MODE_RESET || L.nr) {        // `if self._reset_next_step: return self.reset()`
    o = F::reset(p, i, L, rng);
    L.nr = 0;
  } else {
    o = F::step(p, i, L, action, rng);
    L.nr = (o.step_type == LAST) ? 1u : 0u;
    if (noise) { if (o.step_type != FIRST) o.reward = o.reward + p.noise_scale * wrng.randn(); }
    else if (p.wrapper == 2) { o.reward = o.reward * p.reward_scale; }
  }
  return o;
}

#if defined(__CUDACC__)

__device__ __forceinline__ void st_stream(float4* dst, float4 v) { __stcs(dst, v); }
__device__ __forceinline__ void st_stream(float* dst, float v) { __stcs(dst, v); }

// ----- observation emitters -------------------------------------------------
static const int EMIT_ROWS = 0, EMIT_ONEHOT = 1, EMIT_TWOHOT = 2, EMIT_IMAGE = 3;
template <class F> struct EmitKind { static const int value = EMIT_ROWS; };
template <> struct EmitKind<DeepSea> { static const int value = EMIT_ONEHOT; };
template <> struct EmitKind<Catch> { static const int value = EMIT_TWOHOT; };
template <> struct EmitKind<Mnist> { static const int value = EMIT_IMAGE; };

// One-hot tiles: `hot` is the flat index of the single 1.0 (or -1: all zeros).
__device__ __forceinline__ void emit_onehot(float* obs_t, int64_t warp_base, int64_t B, int K, int hot, bool vec) {
  const int tid = threadIdx.x & 31;
  const int n_lanes = (B - warp_base) < 32 ? (int)(B - warp_base) : 32;
  if (vec) {
    const int K4 = K >> 2;
    for (int j = 0; j < n_lanes; ++j) {
      const int h = __shfl_sync(0xffffffffu, hot, j);
      const int hq = h >> 2, hc = h & 3;
      float4* dst = reinterpret_cast<float4*>(obs_t + (warp_base + j) * (int64_t)K);
#pragma unroll 8
      for (int q = tid; q < K4; q += 32) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q == hq) { if (hc == 0) v.x = 1.f; else if (hc == 1) v.y = 1.f; else if (hc == 2) v.z = 1.f; else v.w = 1.f; }
        st_stream(dst + q, v);
      }
    }
  } else {
    for (int j = 0; j < n_lanes; ++j) {
      const int h = __shfl_sync(0xffffffffu, hot, j);
      float* dst = obs_t + (warp_base + j) * (int64_t)K;
      for (int e = tid; e < K; e += 32) st_stream(dst + e, e == h ? 1.f : 0.f);
    }
  }
}

// Boards with up to two hot cells; the warp's boards form one contiguous span.
__device__ __forceinline__ void emit_twohot(float* obs_t, int64_t warp_base, int64_t B, int K, int hot_a, int hot_b, bool vec) {
  const int tid = threadIdx.x & 31;
  const int n_lanes = (B - warp_base) < 32 ? (int)(B - warp_base) : 32;
  const int total = n_lanes * K;
  float* dst = obs_t + warp_base * (int64_t)K;
  if (vec && (total & 3) == 0) {
    const int total4 = total >> 2;
    for (int q0 = 0; q0 < total4; q0 += 32) {
      const int q = q0 + tid;
      const int e0 = (q < total4 ? q : 0) << 2;
      int j = e0 / K, c = e0 - j * K;
      float v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int jj = j < 32 ? j : 31;
        const int a = __shfl_sync(0xffffffffu, hot_a, jj);
        const int b = __shfl_sync(0xffffffffu, hot_b, jj);
        v[k] = (c == a || c == b) ? 1.f : 0.f;
        if (++c >= K) { c = 0; ++j; }
      }
      if (q < total4) st_stream(reinterpret_cast<float4*>(dst) + q, make_float4(v[0], v[1], v[2], v[3]));
    }
  } else {
    for (int e0 = 0; e0 < total; e0 += 32) {
      const int e = e0 + tid;
      const int ee = e < total ? e : 0;
      const int j = ee / K, c = ee - j * K;
      const int a = __shfl_sync(0xffffffffu, hot_a, j);
      const int b = __shfl_sync(0xffffffffu, hot_b, j);
      if (e < total) st_stream(dst + e, (c == a || c == b) ? 1.f : 0.f);
    }
  }
}

// Image tiles gathered from the int8 dataset (`image` < 0: zeros).
__device__ __forceinline__ void emit_image(const EnvParams& p, float* obs_t, int64_t warp_base, int64_t B, int K, int image, bool vec) {
  const int tid = threadIdx.x & 31;
  const int n_lanes = (B - warp_base) < 32 ? (int)(B - warp_base) : 32;
  for (int j = 0; j < n_lanes; ++j) {
    const int img = __shfl_sync(0xffffffffu, image, j);
    float* dst = obs_t + (warp_base + j) * (int64_t)K;
    const int8_t* src = p.images + (int64_t)(img < 0 ? 0 : img) * K;
    if (vec) {
      const int K4 = K >> 2;
      for (int q = tid; q < K4; q += 32) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (img >= 0) {
          const char4 c = __ldg(reinterpret_cast<const char4*>(src) + q);
          v = make_float4(Mnist::pixel(c.x), Mnist::pixel(c.y), Mnist::pixel(c.z), Mnist::pixel(c.w));
        }
        st_stream(reinterpret_cast<float4*>(dst) + q, v);
      }
    } else {
      for (int e = tid; e < K; e += 32) st_stream(dst + e, img >= 0 ? Mnist::pixel(src[e]) : 0.f);
    }
  }
}

// Stream the warp's staged [n_lanes, K] block.
__device__ __forceinline__ void flush_rows(const float* stage, float* obs_t, int64_t warp_base, int64_t B, int K, bool vec) {
  const int tid = threadIdx.x & 31;
  const int n_lanes = (B - warp_base) < 32 ? (int)(B - warp_base) : 32;
  const int total = n_lanes * K;
  float* dst = obs_t + warp_base * (int64_t)K;
  if (vec && (total & 3) == 0) {
    const float4* s4 = reinterpret_cast<const float4*>(stage);
    for (int q = tid; q < (total >> 2); q += 32) st_stream(reinterpret_cast<float4*>(dst) + q, s4[q]);
  } else {
    for (int e = tid; e < total; e += 32) st_stream(dst + e, stage[e]);
  }
}

template <class F, class R>
__device__ __forceinline__ void render_row(const EnvParams& p, const typename F::Lane& L, R&, float* dst) { F::row(p, L, dst, 1); }
template <>
__device__ __forceinline__ void render_row<UmbrellaChain, LegacyRng<PhiloxSrc> >(const EnvParams& p, const UmbrellaChain::Lane& L, LegacyRng<PhiloxSrc>& r, float* dst) { UmbrellaChain::row(p, L, r, dst, 1); }
template <>
__device__ __forceinline__ void render_row<UmbrellaChain, LegacyRng<MtSrc> >(const EnvParams& p, const UmbrellaChain::Lane& L, LegacyRng<MtSrc>& r, float* dst) { UmbrellaChain::row(p, L, r, dst, 1); }
// Families without a row() never reach render_row (EmitKind != ROWS); give them a stub.
template <class F> struct HasRow { enum { value = 1 }; };
template <> struct HasRow<DeepSea> { enum { value = 0 }; };
template <> struct HasRow<Catch> { enum { value = 0 }; };
template <> struct HasRow<Mnist> { enum { value = 0 }; };

template <class F, class R, bool kHas> struct RowRenderer {
  static __device__ __forceinline__ void run(const EnvParams& p, const typename F::Lane& L, R& r, float* dst) { render_row<F, R>(p, L, r, dst); }
};
template <class F, class R> struct RowRenderer<F, R, false> {
  static __device__ __forceinline__ void run(const EnvParams&, const typename F::Lane&, R&, float*) {}
};

template <class F> struct Descriptor {
  static __device__ __forceinline__ int a(const typename F::Lane&) { return -1; }
  static __device__ __forceinline__ int b(const typename F::Lane&) { return -1; }
};
template <> struct Descriptor<DeepSea> {
  static __device__ __forceinline__ int a(const DeepSea::Lane& L) { return L.hot; }
  static __device__ __forceinline__ int b(const DeepSea::Lane&) { return -1; }
};
template <> struct Descriptor<Catch> {
  static __device__ __forceinline__ int a(const Catch::Lane& L) { return L.hot_a; }
  static __device__ __forceinline__ int b(const Catch::Lane& L) { return L.hot_b; }
};
template <> struct Descriptor<Mnist> {
  static __device__ __forceinline__ int a(const Mnist::Lane& L) { return L.image; }
  static __device__ __forceinline__ int b(const Mnist::Lane&) { return -1; }
};

// ----- the fused transition kernel -------------------------------------------
// Grid: ceil(B / blockDim.x) CTAs of kWarps warps; warp w of the grid owns lanes
// [32 w, 32 w + 32).  Small CTAs keep the per-SM share of the 2048 warp-tasks of
// a 65 536-lane batch within ~1% of even on 148 SMs.
template <class F, int RK, bool kNoise>
__global__ void __launch_bounds__(128) transition_kernel(const EnvParams p, const LaunchArgs a) {
  typedef typename RngOf<RK>::type R;
  extern __shared__ float4 smem_raw[];
  const int64_t B = p.batch;
  const int64_t lane = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t warp_base = lane - (threadIdx.x & 31);
  if (warp_base >= B) return;                    // whole warp out of range
  const bool active = lane < B;
  const int K = p.obs_numel;
  float* stage = reinterpret_cast<float*>(smem_raw) + (size_t)(threadIdx.x >> 5) * 32 * (size_t)K;

  typename F::Lane L;
  R rng, wrng;
  EpisodeStats ep;
  const bool has_rng = p.rng_pos != nullptr;
  const bool track = p.ep != nullptr;
  if (active) {
    if (a.mode == MODE_INIT) F::init(p, L); else F::load(p, lane, L);
    if (has_rng) rng_open(rng, p, lane, false);
    if (kNoise) rng_open(wrng, p, lane, true);
    if (track) ep.load(p, lane);
  } else {
    F::init(p, L);
  }

  if (a.mode == MODE_INIT) {
    if (active) {
      F::ctor_draws(p, L, rng);
      F::store(p, lane, L);
      if (has_rng) rng_close(rng, p, lane, false);
    }
    return;
  }

  for (int64_t t = 0; t < a.T; ++t) {
    const int64_t off = t * B + lane;
    if (active) {
      int32_t action = 0;
      if (a.mode == MODE_STEP) {
        action = a.actions ? a.actions[off]
                           : sample_action(a.action_seed, p.lane_offset + (uint64_t)lane, (uint64_t)(a.step0 + t), p.num_actions);
        if (a.actions_out) a.actions_out[off] = action;
      }
      const StepOut o = lane_transition<F, R, R>(p, lane, L, rng, wrng, action, a.mode, kNoise);
      if (track) ep.track(o);
      if (a.reward) a.reward[off] = (float)o.reward;
      if (a.reward_f64) a.reward_f64[off] = o.reward;
      if (a.discount) a.discount[off] = o.discount;
      if (a.step_type) a.step_type[off] = o.step_type;
    }
    float* obs_t = a.obs + t * B * (int64_t)K;
    const bool vec = a.obs_vec_ok && ((K & 3) == 0 || EmitKind<F>::value == EMIT_TWOHOT || EmitKind<F>::value == EMIT_ROWS);
    if (EmitKind<F>::value == EMIT_ONEHOT) {
      emit_onehot(obs_t, warp_base, B, K, Descriptor<F>::a(L), vec);
    } else if (EmitKind<F>::value == EMIT_TWOHOT) {
      emit_twohot(obs_t, warp_base, B, K, Descriptor<F>::a(L), Descriptor<F>::b(L), vec);
    } else if (EmitKind<F>::value == EMIT_IMAGE) {
      emit_image(p, obs_t, warp_base, B, K, Descriptor<F>::a(L), vec);
    } else {
      if (active) RowRenderer<F, R, HasRow<F>::value != 0>::run(p, L, rng, stage + (threadIdx.x & 31) * K);
      __syncwarp();
      flush_rows(stage, obs_t, warp_base, B, K, vec);
      __syncwarp();
    }
  }

  if (active) {
    F::store(p, lane, L);
    if (has_rng) rng_close(rng, p, lane, false);
    if (kNoise) rng_close(wrng, p, lane, true);
    if (track) ep.store(p, lane);
  }
}

#endif  // __CUDACC__

}  // namespace bsb

// smct_common.cuh — shared device helpers: warp/block reductions, Philox4x32-7 + Box-Muller,
// LayerNorm (Keras non-fused form).  sm_100a only.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define SMCT_FULL_MASK 0xffffffffu

namespace smct {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(SMCT_FULL_MASK, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(SMCT_FULL_MASK, v, o));
  return v;
}
// reduction over aligned groups of G lanes (G power of two <= 32)
template <int G>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(SMCT_FULL_MASK, v, o);
  return v;
}

// caller-supplied indices (token ids, class ids, forced ancestor indices) are clamped into range on the device: an
// out-of-range value can then never touch memory outside the tensor it indexes (the Python binding rejects such
// inputs with a ValueError before the launch, like TensorFlow's InvalidArgument on CPU)
__device__ __forceinline__ int clamp_index(int v, int n) { return min(max(v, 0), n - 1); }

// block-wide reductions through a 32-float scratch (all threads get the result)
__device__ __forceinline__ float block_sum(float v, float* scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  float r = (lane < nw) ? scratch[lane] : 0.f;
  return warp_sum(r);
}
__device__ __forceinline__ float block_max(float v, float* scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  float r = (lane < nw) ? scratch[lane] : -INFINITY;
  return warp_max(r);
}

// ------------------------------------------------------------------------------------------
// fp16 (h, m) split of a pair of fp32 values (the operand format of every tensor-core product of the fused forward:
// x = h + m to ~22 bits, smct_tc.cuh).  One packed conversion per pair (F2FP.SATFINITE.F16.F32.PACK_AB): the saturating
// form keeps out-of-range values finite (+-65504) without separate clamps, and is the same round-to-nearest conversion
// in range.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_half2_sat(float lo, float hi) {
  uint32_t h;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(hi), "f"(lo));
  return h;
}
__device__ __forceinline__ void split_half2(float x, float y, uint32_t& h, uint32_t& m) {
  h = pack_half2_sat(x, y);
  const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&h));
  m = pack_half2_sat(x - hf.x, y - hf.y);
}

// ------------------------------------------------------------------------------------------
// Philox4x32-R (Salmon et al. SC'11).  Stream layout documented in oracle/philox.py:
//   key = (seed_lo, seed_hi); counter = (group_lo, group_hi, stream, t)
// R = 7 rounds: the smallest round count the Random123 authors report as passing BigCrush ("Philox4x32-7"); the noise
// generation is 256 normals per particle-step and was 15 % of the fused forward with the 10-round default.
// oracle/philox.py restates the same generator (ROUNDS must match).
// ------------------------------------------------------------------------------------------
struct u32x4 { uint32_t x, y, z, w; };
constexpr int SMCT_PHILOX_ROUNDS = 7;

__device__ __forceinline__ u32x4 philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                            uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < SMCT_PHILOX_ROUNDS; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return u32x4{c0, c1, c2, c3};
}

// Box-Muller on two 32-bit words.  u1 = fl(x) 2^-32 + 2^-33 in (0, 1] (the int -> float conversion rounds to 24 bits, so
// no shift is needed; u1 = 1 gives r = 0), angle = 2 pi (u2 - 1/2) as one FMA of the converted word.  The
// special-function unit is used directly (lg2 / sqrt / sin / cos .approx): u1 >= 2^-33 is never denormal and
// -2 ln u1 is in [0, 45.8], so the denormal scaling of __logf and the range check + Newton step + slow-path call of
// the IEEE sqrtf are dead weight here — 13 instead of 30 SASS instructions per pair and no divergence barrier, at
// 256 normals per particle-step.  oracle/philox.py restates the same arithmetic.
__device__ __forceinline__ void box_muller(uint32_t xa, uint32_t xb, float& n0, float& n1) {
  const float u1 = fmaf((float)xa, 2.3283064365386963e-10f, 1.1641532182693481e-10f);   // 2^-32, 2^-33
  const float th = fmaf((float)xb, 1.4629180792671596e-09f, -3.1415926528583237f);     // 2 pi 2^-32, -pi (1 - 2^-32)
  float l2, r, s, c;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(u1));
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(l2 * -1.3862943611198906f));   // sqrt(-2 ln u1)
  asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(th));
  asm("cos.approx.ftz.f32 %0, %1;" : "=f"(c) : "f"(th));
  n0 = r * c; n1 = r * s;
}

// four standard normals of group `grp` (64-bit) in stream `which` at timestep t
__device__ __forceinline__ void philox_normal4(uint64_t seed, uint32_t which, uint32_t t, uint64_t grp,
                                               float n[4]) {
  const u32x4 x = philox4x32((uint32_t)grp, (uint32_t)(grp >> 32), which, t, (uint32_t)seed,
                                (uint32_t)(seed >> 32));
  box_muller(x.x, x.y, n[0], n[1]);
  box_muller(x.z, x.w, n[2], n[3]);
}

// out-of-line variant for call sites that are unrolled many times (keeps the instruction footprint small: the fused
// forward is bound by instruction-cache misses when every normal carries its own copy of the generator)
__device__ __noinline__ float4 philox_normal4_call(uint32_t seed_lo, uint32_t seed_hi, uint32_t which, uint32_t t,
                                                   uint32_t grp_lo, uint32_t grp_hi) {
  const u32x4 x = philox4x32(grp_lo, grp_hi, which, t, seed_lo, seed_hi);
  float4 n;
  box_muller(x.x, x.y, n.x, n.y);
  box_muller(x.z, x.w, n.z, n.w);
  return n;
}

// one standard normal: element d of particle bp (global particle index) with GD = ceil(D/4)
__device__ __forceinline__ float philox_normal1(uint64_t seed, uint32_t which, uint32_t t, uint64_t bp,
                                                int GD, int d) {
  float n[4];
  philox_normal4(seed, which, t, bp * (uint64_t)GD + (uint64_t)(d >> 2), n);
  const int e = d & 3;
  return e == 0 ? n[0] : (e == 1 ? n[1] : (e == 2 ? n[2] : n[3]));
}

__device__ __forceinline__ double philox_uniform(uint64_t seed, uint32_t t, uint64_t bp) {
  const u32x4 x = philox4x32((uint32_t)bp, (uint32_t)(bp >> 32), 4u, t, (uint32_t)seed,
                                (uint32_t)(seed >> 32));
  return ((double)(x.x >> 5) * 67108864.0 + (double)(x.y >> 6)) * 1.1102230246251565e-16;  // 2^-53
}

}  // namespace smct

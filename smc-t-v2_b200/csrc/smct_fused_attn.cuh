// smct_fused_attn.cuh — attention over the particle genealogy on warp-level tensor-core MMAs (mma.sync m16n8k16,
// fp16 operands, fp32 accumulators), used by the fused forward (smct_fused_tc.cuh).
//
// Why: particles are kept in genealogy order, so the 16 particles of a warp share almost all of their history rows
// (measured: ~1.2 distinct rows per 8 particles and slot).  The CUDA-core loop moved 512 B per particle and slot into
// registers and was bound by the LSU data path.  Here a warp gathers only the DISTINCT (slot, row) pairs of its 16
// particles ("columns"), 16 columns per group, and evaluates
//     S = Q · Kcols^T   (16 particles x 16 columns),   masked online softmax,   Z += P · Vcols
// with MMAs; a particle owns exactly one column per slot (mask = the contiguous range of particles [lo,hi) that
// descend from the column's row — contiguous because the ancestry A[q][s] is non-decreasing in the position q).
// Operands are fp16 (h,m) splits of power-of-two-scaled fp32 values, three MMA terms per product (~22 mantissa bits).
// K and V rows are stored pre-split once by the particle that creates them (Ks, Vs: 256 B per row):
//   K row: 16-byte unit (part, half, c) at unit index part*8 + half*4 + c holds the 8 fp16 of dims 16c + 8 half .. +7
//          (part 0 = h, 1 = m)  -> lane (r,c) of the MMA loads four contiguous-per-quad units = its B fragments of
//          the 4 k-steps (k-step j, fragment position kk <-> dim 16c + 4j + {0,1 | 2,3}).
//   V row: h part (64 fp16, natural order, 128 B) then m part; lane (r,c) loads dims 8r..8r+7 of four column rows and
//          pairs them with PRMT into B fragments of the 8 n-tiles (n-tile i, column n <-> dim 8n + i).
#pragma once
#include <cuda_fp16.h>
#include "smct_common.cuh"

namespace smct {
namespace attn {

__device__ __forceinline__ void mma16816(float d[4], const uint32_t a[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// (x, y) -> packed fp16 pair h (x in the low half) and the packed remainder m
__device__ __forceinline__ void split2(float x, float y, uint32_t& h, uint32_t& m) { split_half2(x, y, h, m); }
// values beyond the fp16 range are kept finite by the saturating conversion inside split_half2 (smct_common.cuh)
__device__ __forceinline__ float sat16(float x) { return x; }

// 8 consecutive (scaled) values -> one 16-byte h unit and one m unit
__device__ __forceinline__ void split8(const float* x, uint4& h, uint4& m) {
  split2(sat16(x[0]), sat16(x[1]), h.x, m.x);
  split2(sat16(x[2]), sat16(x[3]), h.y, m.y);
  split2(sat16(x[4]), sat16(x[5]), h.z, m.z);
  split2(sat16(x[6]), sat16(x[7]), h.w, m.w);
}

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// L2 prefetch of the split K and V rows of one column (2 x 256 B), issued by the lane that opens the column
__device__ __forceinline__ void prefetch_column(const unsigned char* Ks, const unsigned char* Vs, int slot, int row, int P) {
  const uint32_t off = ((uint32_t)slot * (uint32_t)P + (uint32_t)row) << 8;   // < 2^32: S P 256 with S <= 4096, P <= 1024
  asm volatile("prefetch.global.L2 [%0];" ::"l"(Ks + off));
  asm volatile("prefetch.global.L2 [%0];" ::"l"(Ks + off + 128));
  asm volatile("prefetch.global.L2 [%0];" ::"l"(Vs + off));
  asm volatile("prefetch.global.L2 [%0];" ::"l"(Vs + off + 128));
}

// per-warp state of the running softmax over the columns processed so far; rows R0 = lane>>2 and R1 = R0 + 8
struct State {
  float Z[8][4];                 // accumulators of the 8 n-tiles: [i][0,1] row R0, [i][2,3] row R1
  float m0, m1, l0, l1;
};

// One group of up to 16 columns.  cd (this warp's 16 descriptors in shared memory): x = slot << 16 | row,
// y = hi << 8 | lo (particles lo..hi-1 of the warp own the column).  Ks/Vs: split rows of this sequence.
// Qf: the group's Q fragments in shared memory, unit k (16 bytes per lane) at Qf[k * 32 + lane]: k < 4 the h part of
// k-step k, k >= 4 the m part of k-step k - 4.
__device__ __forceinline__ void process_group(State& st, const uint2* cd, int ncols, const uint4* __restrict__ Qf,
                                              const unsigned char* __restrict__ Ks, const unsigned char* __restrict__ Vs,
                                              int P, float inv_qk, int lane) {
  const int r = lane >> 2, c = lane & 3;
  uint2 d = cd[lane & 15];
  const uint2 d0 = cd[0];
  if ((lane & 15) >= ncols) { d.x = d0.x; d.y = 0u; }   // unused column: a valid address, an empty owner range
  __syncwarp();
  // byte offset of the column's split rows inside this sequence's tables, computed once by the lane that holds the
  // descriptor and passed around as a 32-bit word (S P 256 < 2^32): the 6 row pointers of a lane are base + offset
  const uint32_t coff = ((d.x >> 16) * (uint32_t)P + (d.x & 0xFFFFu)) << 8;
  // V rows of columns 2c, 2c+1, 2c+8, 2c+9 (dims 8r .. 8r+7)
  const unsigned char* vp[4];
#pragma unroll
  for (int k = 0; k < 4; ++k)
    vp[k] = Vs + (__shfl_sync(SMCT_FULL_MASK, coff, 2 * c + (k & 1) + 8 * (k >> 1)) + 16u * (uint32_t)r);
  float S[2][4];
#pragma unroll
  for (int nt = 0; nt < 2; ++nt) {
    const unsigned char* kp = Ks + (__shfl_sync(SMCT_FULL_MASK, coff, 8 * nt + r) + 16u * (uint32_t)c);
    const uint4 h0 = *reinterpret_cast<const uint4*>(kp);
    const uint4 h1 = *reinterpret_cast<const uint4*>(kp + 64);
    const uint4 m0 = *reinterpret_cast<const uint4*>(kp + 128);
    const uint4 m1 = *reinterpret_cast<const uint4*>(kp + 192);
    const uint32_t kh[4][2] = {{h0.x, h0.y}, {h0.z, h0.w}, {h1.x, h1.y}, {h1.z, h1.w}};
    const uint32_t km[4][2] = {{m0.x, m0.y}, {m0.z, m0.w}, {m1.x, m1.y}, {m1.z, m1.w}};
    S[nt][0] = S[nt][1] = S[nt][2] = S[nt][3] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {      // the Q fragments of one k-step at a time keep the live range short
      const uint4 qh4 = Qf[j * 32 + lane], qm4 = Qf[(4 + j) * 32 + lane];
      const uint32_t qh[4] = {qh4.x, qh4.y, qh4.z, qh4.w}, qm[4] = {qm4.x, qm4.y, qm4.z, qm4.w};
      mma16816(S[nt], qh, kh[j][0], kh[j][1]);
      mma16816(S[nt], qm, kh[j][0], kh[j][1]);
      mma16816(S[nt], qh, km[j][0], km[j][1]);
    }
  }
  // mask: entry (row R, column j) counts iff lo_j <= R < hi_j
  float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
  for (int nt = 0; nt < 2; ++nt) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const uint32_t ly = __shfl_sync(SMCT_FULL_MASK, d.y, 8 * nt + 2 * c + e);
      const int lo = (int)(ly & 0xFFu), hi = (int)(ly >> 8);
      const float s0 = (r >= lo && r < hi) ? S[nt][e] * inv_qk : -INFINITY;
      const float s1 = (r + 8 >= lo && r + 8 < hi) ? S[nt][2 + e] * inv_qk : -INFINITY;
      S[nt][e] = s0; S[nt][2 + e] = s1;
      mx0 = fmaxf(mx0, s0); mx1 = fmaxf(mx1, s1);
    }
  }
  mx0 = fmaxf(mx0, __shfl_xor_sync(SMCT_FULL_MASK, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(SMCT_FULL_MASK, mx0, 2));
  mx1 = fmaxf(mx1, __shfl_xor_sync(SMCT_FULL_MASK, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(SMCT_FULL_MASK, mx1, 2));
  const float mn0 = fmaxf(st.m0, mx0), mn1 = fmaxf(st.m1, mx1);
  const float corr0 = ex2f(st.m0 - mn0), corr1 = ex2f(st.m1 - mn1);
  float ps0 = 0.f, ps1 = 0.f;
#pragma unroll
  for (int nt = 0; nt < 2; ++nt) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      S[nt][e] = ex2f(S[nt][e] - mn0); S[nt][2 + e] = ex2f(S[nt][2 + e] - mn1);
      ps0 += S[nt][e]; ps1 += S[nt][2 + e];
    }
  }
  ps0 += __shfl_xor_sync(SMCT_FULL_MASK, ps0, 1); ps0 += __shfl_xor_sync(SMCT_FULL_MASK, ps0, 2);
  ps1 += __shfl_xor_sync(SMCT_FULL_MASK, ps1, 1); ps1 += __shfl_xor_sync(SMCT_FULL_MASK, ps1, 2);
  st.l0 = fmaf(st.l0, corr0, ps0); st.l1 = fmaf(st.l1, corr1, ps1);
  st.m0 = mn0; st.m1 = mn1;
#pragma unroll
  for (int i = 0; i < 8; ++i) { st.Z[i][0] *= corr0; st.Z[i][1] *= corr0; st.Z[i][2] *= corr1; st.Z[i][3] *= corr1; }
  // P as the A operand of the second product (k = column): a0 = (R0, cols 2c,2c+1) a1 = (R1, same) a2/a3 = cols +8
  uint32_t Ph[4], Pm[4];
  split2(S[0][0], S[0][1], Ph[0], Pm[0]);
  split2(S[0][2], S[0][3], Ph[1], Pm[1]);
  split2(S[1][0], S[1][1], Ph[2], Pm[2]);
  split2(S[1][2], S[1][3], Ph[3], Pm[3]);
#pragma unroll
  for (int part = 0; part < 2; ++part) {
    const uint4 va = *reinterpret_cast<const uint4*>(vp[0] + 128 * part);
    const uint4 vb = *reinterpret_cast<const uint4*>(vp[1] + 128 * part);
    const uint4 vc = *reinterpret_cast<const uint4*>(vp[2] + 128 * part);
    const uint4 vd = *reinterpret_cast<const uint4*>(vp[3] + 128 * part);
    const uint32_t wa[4] = {va.x, va.y, va.z, va.w}, wb[4] = {vb.x, vb.y, vb.z, vb.w};
    const uint32_t wc[4] = {vc.x, vc.y, vc.z, vc.w}, wd[4] = {vd.x, vd.y, vd.z, vd.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint32_t sel = (i & 1) ? 0x7632u : 0x5410u;
      const uint32_t b0 = __byte_perm(wa[i >> 1], wb[i >> 1], sel);
      const uint32_t b1 = __byte_perm(wc[i >> 1], wd[i >> 1], sel);
      mma16816(st.Z[i], Ph, b0, b1);
      if (part == 0) mma16816(st.Z[i], Pm, b0, b1);
    }
  }
}

}  // namespace attn
}  // namespace smct

// smct_tc.cuh — tcgen05 (5th-gen tensor core) building blocks for the dense part of the fused forward:
// shared-memory operand layout, descriptors, TMEM allocation, MMA issue / commit, TMEM loads.
//
// Precision: the parity bar is 1e-4 relative in fp32, which single-pass bf16/tf32 operands cannot meet.
// Every fp32 operand x is split into two fp16 terms  h = fp16(x), m = fp16(x - h)  (2 x 11 mantissa bits) and
// a product A·B is issued as three kind::f16 MMAs into the same fp32 TMEM accumulator:
//     A_h·B_h + A_m·B_h + A_h·B_m          (the dropped A_m·B_m term is O(2^-22) relative).
// fp16 has a narrow exponent range, so every operand is pre-multiplied by an exact power of two that puts
// its largest magnitude near 2^10 (weights: per matrix, computed by tc_weight_prep_kernel; activations: from
// analytic bounds or the row maximum) and the accumulator is multiplied back in the epilogue — exact.
// (A bf16 split needs no scaling but keeps only 16 bits: measured error 1.5e-5 of the tensor scale.)
//
// Operand layout (K-major, no swizzle; cute "INTERLEAVE" canonical layout ((8,n),2):((1,SBO),LBO) in 16-byte
// units): a tile of R rows x 64 k-values of fp16 is stored as 8x8 "core matrices" of 128 bytes
// (8 rows x 16 bytes); element (r,k) lives at byte
//     (r/8)*1024 + (k/8)*128 + (r%8)*16 + (k%8)*2        -> LBO = 128 B (next core along K), SBO = 1024 B.
// One tcgen05.mma consumes K = 16 (two cores along K): K-step j starts at byte offset 256*j.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace smct {
namespace tc {

constexpr uint32_t LBO_BYTES = 128;
constexpr uint32_t SBO_BYTES = 1024;
constexpr uint32_t A_PART_BYTES = 128 * 64 * 2;   // one fp16 part (h or m) of a 128x64 A tile: 16 KB
constexpr uint32_t B_PART_BYTES = 64 * 64 * 2;    // one fp16 part of a 64(N)x64(K) B block: 8 KB

__device__ __forceinline__ uint32_t operand_offset(int r, int k) {
  return (uint32_t)((r >> 3) * (int)SBO_BYTES + (k >> 3) * (int)LBO_BYTES + (r & 7) * 16 + (k & 7) * 2);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// 64-bit shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout type 0 (no swizzle) [61,64).
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((LBO_BYTES >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((SBO_BYTES >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// instruction descriptor (cute::UMMA::InstrDescriptor) for kind::f16: D=F32 (bit 4), A=B=F16 (format 0),
// K-major both, N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// exact power-of-two scale that brings `bound` (> 0) to [2^9, 2^10]; returns the exponent e with scale = 2^e
__device__ __forceinline__ int scale_exponent(float bound) {
  return (bound > 0.f && isfinite(bound)) ? (10 - (ilogbf(bound) + 1)) : 0;
}

__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// D[128 x N] (+)= A[128 x 64] * B[N x 64]^T with the three-term fp16 split.  One thread issues.
// a_base / b_base: shared addresses of the h part; the m part follows at +A_PART_BYTES / +b_part_bytes.
// The start-address field of a descriptor is (address >> 4) in bits [0,14): with < 256 KB of shared memory an offset
// can simply be added to a base descriptor (no carry into the other fields), which keeps the single issuing thread short.
__device__ __forceinline__ void mma_block_f16x3(uint32_t tmem_d, uint32_t a_base, uint32_t b_base, uint32_t b_part_bytes,
                                                 uint32_t idesc, bool accumulate) {
  const uint64_t da0 = make_desc(a_base), db0 = make_desc(b_base);
  const uint32_t a_off[3] = {0u, A_PART_BYTES, 0u};
  const uint32_t b_off[3] = {0u, 0u, b_part_bytes};
#pragma unroll
  for (int term = 0; term < 3; ++term) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const uint64_t da = da0 + (uint64_t)((a_off[term] + 256u * ks) >> 4);
      const uint64_t db = db0 + (uint64_t)((b_off[term] + 256u * ks) >> 4);
      mma_f16(tmem_d, da, db, idesc, (accumulate || term > 0 || ks > 0) ? 1u : 0u);
    }
  }
}

// one lane of a converged warp (elect.sync): the MMA issue sites branch on the WARP and elect inside, so the compiler
// sees a uniform branch instead of a divergent `tid == 0`
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void commit(uint64_t* mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(mbar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// bounded wait: traps instead of hanging the GPU if the MMA never commits.  try_wait carries a suspend-time hint
// (20 us): the warp sleeps in hardware until the phase completes instead of coming back every few hundred cycles — the
// spin loop was 17.6 % of all executed warp-instructions of the fused forward (profiles/r02_fused2_summary.md), issue
// slots taken from the co-resident CTA.
__device__ __forceinline__ void mbar_wait(uint64_t* mbar, uint32_t parity) {
  const uint32_t addr = smem_u32(mbar);
  uint32_t done = 0;
#pragma unroll 1   // (the compiler otherwise unrolls the spin loop 32 times at every wait site)
  for (uint32_t spin = 0; spin < (1u << 20); ++spin) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity), "r"(20000u)
        : "memory");
    if (done) return;
  }
  __trap();
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// TMEM allocation: executed by ONE full warp; the base address is written to *slot (shared memory)
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// 32 consecutive fp32 columns of this thread's TMEM lane (warp w reads lanes 32*(w%4)..+31; lane i of the warp = row)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float v[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// split 4 consecutive (already scaled) fp32 values into packed fp16 h and m parts (8 bytes each)
__device__ __forceinline__ void split4(const float x[4], uint2& h, uint2& m) {
  split_half2(x[0], x[1], h.x, m.x);
  split_half2(x[2], x[3], h.y, m.y);
}

// write 4 consecutive k-values (k0 % 4 == 0) of row r into an A operand tile (h part at base, m part at base + A_PART_BYTES)
__device__ __forceinline__ void store_a4(unsigned char* base, int r, int k0, const float x[4]) {
  uint2 h, m;
  split4(x, h, m);
  const uint32_t off = operand_offset(r, k0);
  *reinterpret_cast<uint2*>(base + off) = h;
  *reinterpret_cast<uint2*>(base + A_PART_BYTES + off) = m;
}

// write 8 consecutive k-values (k0 % 8 == 0: one 16-byte row of a core matrix) of row r into an A operand tile.  A warp
// of 32 consecutive rows then covers 32 banks per 8 rows with its 16-byte stores (store_a4's 8-byte stores hit every
// bank four times).
__device__ __forceinline__ void store_a8(unsigned char* base, int r, int k0, const float x[8]) {
  uint2 h0, m0, h1, m1;
  split4(x, h0, m0);
  split4(x + 4, h1, m1);
  const uint32_t off = operand_offset(r, k0);
  *reinterpret_cast<uint4*>(base + off) = make_uint4(h0.x, h0.y, h1.x, h1.y);
  *reinterpret_cast<uint4*>(base + A_PART_BYTES + off) = make_uint4(m0.x, m0.y, m1.x, m1.y);
}

}  // namespace tc

// ------------------------------------------------------------------------------------------
// weight preparation (one CTA): fp32 Keras kernels -> scaled fp16 (h,m) blocks in the operand layout.
// Block b (16 KB: h part 8 KB then m part 8 KB) holds B[n][k] for n,k < 64:
//   b = 0        : Wz[k][n]   * 2^eWz
//   b = 1 + 2c   : W1[k][64c + n] * 2^eW1
//   b = 2 + 2c   : W2[64c + k][n] * 2^eW2      (c = 0..dff/64-1)
// followed by a header of floats (TcHeader) with the inverse scales and the activation scales.
// ------------------------------------------------------------------------------------------
struct TcHeader {
  float inv_wz, inv_w1, inv_w2;   // 2^-eW*
  float s_o, inv_o;               // scale of o = LN1(.) (bounded by 8 max|gamma1| + max|beta1|) and its inverse
  float s_h, inv_h;               // scale of h = relu(o W1 + b1) (bounded analytically) and its inverse
  float pad;
  // attention operands (warp-level fp16 split MMAs): exact power-of-two scales of q (incl. log2(e)/sqrt(d_k)), k, v
  // from analytic bounds |LN1(x) W + b| + 16 sigma, and the inverse scales of the scores and of z_
  float s_q, s_k, s_v, inv_qk, inv_v;
  float pad2[3];
};

struct TcAttnPrep {     // optional inputs for the attention operand scales and the noise-scale table
  const float *Wq, *bq, *Wk, *bk, *Wv, *bv, *logvar_q, *logvar_k, *logvar_v, *logvar_z;
  int logvar_dim;
};

// four consecutive per-column constants (16-byte aligned) with one 128-bit load
__device__ __forceinline__ void ld4(const float* p, float v[4]) {
  const float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}

__global__ void __launch_bounds__(1024) tc_weight_prep_kernel(const float* Wz, const float* W1, const float* W2,
                                                              const float* b1, const float* ln_g, const float* ln_b,
                                                              int dff, unsigned char* out, TcAttnPrep ap = TcAttnPrep{}) {
  __shared__ float red[32];
  __shared__ float sc[8];
  const int tid = threadIdx.x;
  const int nblk = 1 + 2 * (dff / 64);
  auto bmax = [&](float v) {
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    float r = red[tid & 31];
    for (int o = 16; o > 0; o >>= 1) r = fmaxf(r, __shfl_xor_sync(0xffffffffu, r, o));
    return r;
  };
  float m = 0.f;
  for (int i = tid; i < 64 * 64; i += 1024) m = fmaxf(m, fabsf(Wz[i]));
  const float mz = bmax(m);
  float m1 = 0.f, m2 = 0.f, mb1 = 0.f, mg = 0.f, mbt = 0.f, colsum = 0.f;
  if (dff > 0) {
    for (int i = tid; i < 64 * dff; i += 1024) { m1 = fmaxf(m1, fabsf(W1[i])); m2 = fmaxf(m2, fabsf(W2[i])); }
    for (int i = tid; i < dff; i += 1024) {
      mb1 = fmaxf(mb1, fabsf(b1[i]));
      float cs = 0.f;
      for (int k = 0; k < 64; ++k) cs += fabsf(W1[k * dff + i]);
      colsum = fmaxf(colsum, cs);
    }
    for (int i = tid; i < 64; i += 1024) { mg = fmaxf(mg, fabsf(ln_g[i])); mbt = fmaxf(mbt, fabsf(ln_b[i])); }
  }
  m1 = bmax(m1); m2 = bmax(m2); mb1 = bmax(mb1); mg = bmax(mg); mbt = bmax(mbt); colsum = bmax(colsum);
  float bq = 0.f, bk = 0.f, bv = 0.f;   // bounds of |q|, |k|, |v| (projection of LN1(x) plus 16 standard deviations of noise)
  if (ap.Wq) {
    float cq = 0.f, ck = 0.f, cv = 0.f, aq = 0.f, ak = 0.f, av = 0.f, sq = 0.f, sk = 0.f, sv = 0.f;
    for (int j = tid; j < 64; j += 1024) {
      float s1 = 0.f, s2 = 0.f, s3 = 0.f;
      for (int k = 0; k < 64; ++k) { s1 += fabsf(ap.Wq[k * 64 + j]); s2 += fabsf(ap.Wk[k * 64 + j]); s3 += fabsf(ap.Wv[k * 64 + j]); }
      cq = s1; ck = s2; cv = s3;
      aq = fabsf(ap.bq[j]); ak = fabsf(ap.bk[j]); av = fabsf(ap.bv[j]);
      const int li = ap.logvar_dim > 1 ? j : 0;
      sq = expf(0.5f * ap.logvar_q[li]); sk = expf(0.5f * ap.logvar_k[li]); sv = expf(0.5f * ap.logvar_v[li]);
    }
    cq = bmax(cq); ck = bmax(ck); cv = bmax(cv); aq = bmax(aq); ak = bmax(ak); av = bmax(av);
    sq = bmax(sq); sk = bmax(sk); sv = bmax(sv);
    const float bo = 8.0f * mg + mbt;
    bq = bo * cq + aq + 16.f * sq; bk = bo * ck + ak + 16.f * sk; bv = bo * cv + av + 16.f * sv;
  }
  if (tid == 0) {
    const float bound_o = 8.0f * mg + mbt;                 // |LN(x)_n| <= sqrt(63) |gamma_n| + |beta_n|
    const float bound_h = bound_o * colsum + mb1;          // |relu(o W1 + b1)_j| <= ||o||_inf sum_k |W1[k][j]| + |b1_j|
    const int ez = tc::scale_exponent(mz), e1 = tc::scale_exponent(m1), e2 = tc::scale_exponent(m2);
    const int eo = tc::scale_exponent(bound_o), eh = tc::scale_exponent(bound_h);
    sc[0] = ldexpf(1.f, ez); sc[1] = ldexpf(1.f, e1); sc[2] = ldexpf(1.f, e2);
    TcHeader* h = reinterpret_cast<TcHeader*>(out + (size_t)nblk * 2 * tc::B_PART_BYTES);
    h->inv_wz = ldexpf(1.f, -ez); h->inv_w1 = ldexpf(1.f, -e1); h->inv_w2 = ldexpf(1.f, -e2);
    h->s_o = ldexpf(1.f, eo); h->inv_o = ldexpf(1.f, -eo); h->s_h = ldexpf(1.f, eh); h->inv_h = ldexpf(1.f, -eh);
    h->pad = 0.f;
    // attention operands: bound -> [2^13, 2^14] (fp16 max 65504); typical values sit 10-20x below the bound
    const float qs = 0.125f * 1.4426950408889634f;   // log2(e) / sqrt(d_k), d_k = 64
    const int eq = ap.Wq ? tc::scale_exponent(bq * qs) + 4 : 0, ek = ap.Wq ? tc::scale_exponent(bk) + 4 : 0,
              ev = ap.Wq ? tc::scale_exponent(bv) + 4 : 0;
    h->s_q = ldexpf(qs, eq); h->s_k = ldexpf(1.f, ek); h->s_v = ldexpf(1.f, ev);
    h->inv_qk = ldexpf(1.f, -(eq + ek)); h->inv_v = ldexpf(1.f, -ev);
    h->pad2[0] = h->pad2[1] = h->pad2[2] = 0.f;
  }
  if (ap.Wq && ap.logvar_z && tid < 4 * 64) {
    // noise-scale table behind the header (+256 B): [which][d] = exp(logvar/2) for k,q,v,z, then exp(-logvar) for k,q,v,z
    float* tab = reinterpret_cast<float*>(out + (size_t)nblk * 2 * tc::B_PART_BYTES + 256);
    const int which = tid >> 6, d = tid & 63;
    const float* lv = which == 0 ? ap.logvar_k : (which == 1 ? ap.logvar_q : (which == 2 ? ap.logvar_v : ap.logvar_z));
    const float v = lv[ap.logvar_dim > 1 ? d : 0];
    tab[which * 64 + d] = expf(0.5f * v);
    tab[(4 + which) * 64 + d] = expf(-v);
  }
  __syncthreads();
  for (int idx = tid; idx < nblk * 4096; idx += 1024) {
    const int b = idx >> 12, n = (idx >> 6) & 63, k = idx & 63;
    float w;
    if (b == 0) w = Wz[k * 64 + n] * sc[0];
    else if (b & 1) w = W1[k * dff + ((b - 1) >> 1) * 64 + n] * sc[1];
    else w = W2[(((b - 2) >> 1) * 64 + k) * 64 + n] * sc[2];
    const __half h = __float2half_rn(w);
    const __half mm = __float2half_rn(w - __half2float(h));
    unsigned char* blk = out + (size_t)b * 2 * tc::B_PART_BYTES;
    const uint32_t off = tc::operand_offset(n, k);
    *reinterpret_cast<__half*>(blk + off) = h;
    *reinterpret_cast<__half*>(blk + tc::B_PART_BYTES + off) = mm;
  }
}

// ------------------------------------------------------------------------------------------
// self-test: D[128x64] = A[128x64] * W[64x64] (W in Keras layout [k][n]) through tcgen05 + TMEM.
// One CTA of 128 threads.  Validates descriptors, operand layout, TMEM lane mapping and fences.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) tc_selftest_kernel(const float* A, const unsigned char* wblk, float* D) {
  extern __shared__ __align__(1024) unsigned char tsm[];
  unsigned char* As = tsm;                             // 32 KB (h,m)
  unsigned char* Bs = tsm + 2 * tc::A_PART_BYTES;      // 16 KB (h,m)
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, wid = tid >> 5;
  if (wid == 0) tc::tmem_alloc(&tmem_slot, 64);
  if (tid == 0) { tc::mbar_init(&mbar, 1); tc::fence_mbar_init(); }
  // A tile: thread = row, scaled by an exact power of two from the row maximum
  float rmax = 0.f;
  for (int k = 0; k < 64; ++k) rmax = fmaxf(rmax, fabsf(A[tid * 64 + k]));
  const int ea = tc::scale_exponent(rmax);
  const float sa = ldexpf(1.f, ea), inv_a = ldexpf(1.f, -ea);
  for (int k0 = 0; k0 < 64; k0 += 4) {
    float x[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) x[e] = A[tid * 64 + k0 + e] * sa;
    tc::store_a4(As, tid, k0, x);
  }
  for (int i = tid; i < (int)(2 * tc::B_PART_BYTES / 16); i += 128)
    reinterpret_cast<uint4*>(Bs)[i] = reinterpret_cast<const uint4*>(wblk)[i];
  tc::fence_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    tc::mma_block_f16x3(tmem, tc::smem_u32(As), tc::smem_u32(Bs), tc::B_PART_BYTES, tc::make_idesc(128, 64), false);
    tc::commit(&mbar);
  }
  tc::mbar_wait(&mbar, 0);
  tc::fence_after_sync();
  for (int half = 0; half < 2; ++half) {
    float v[32];
    tc::tmem_ld32(tmem + ((uint32_t)(32 * wid) << 16) + 32 * half, v);
    const float inv = inv_a * reinterpret_cast<const TcHeader*>(wblk + 2 * tc::B_PART_BYTES)->inv_wz;
#pragma unroll
    for (int i = 0; i < 32; ++i) D[tid * 64 + 32 * half + i] = v[i] * inv;
  }
  tc::fence_before_sync();
  __syncthreads();
  if (wid == 0) tc::tmem_dealloc(tmem, 64);
}

}  // namespace smct

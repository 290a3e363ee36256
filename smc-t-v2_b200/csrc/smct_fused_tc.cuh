// smct_fused_tc.cuh — fused persistent forward, tensor-core variant (default SMCT_ALGO_FUSED).
// Same structure as smct_fused.cuh (one CTA per sequence, path storage, genealogy order), but the dense
// part of every 128-particle tile runs on the 5th-gen tensor cores:
//     z  = z_·Wz            -> tcgen05.mma, fp32 accumulator in TMEM (cols   0.. 63)
//     h  = o·W1[:, chunk]   -> tcgen05.mma                           (cols  64..127), 4 chunks of 64 hidden units
//     r' = h·W2[chunk, :]   -> tcgen05.mma accumulating over chunks  (cols 128..191)
// Operands are fp16 (h,m) splits of power-of-two-scaled fp32 values (three MMA terms, see smct_tc.cuh) so that
// the result keeps ~22 mantissa bits and the 1e-4 parity bar holds with margin.  Epilogues (bias, noise, residual, LayerNorm, relu,
// re-split to bf16, output layer, log-weight) are done by all 16 warps straight from TMEM (tcgen05.ld):
// warp w owns TMEM lanes 32*(w%4).. (rows) and columns 16*(w/4).. of the 64-wide accumulator.
// Weight blocks (bf16 h,m, pre-arranged in the operand layout by tc_weight_prep_kernel) stream from L2 into
// shared memory with cp.async, double-buffered per FFN chunk.
#pragma once
#include "smct_fused.cuh"
#include "smct_tc.cuh"

namespace smct {

struct FusedTcSmem {
  FusedCommon c;
  float bz[FD], b1[FF], b2[FD], ln1g[FD], ln1b[FD], ln2g[FD], ln2b[FD], Wo[FD * 16];
  float psum[2][4][FTILE];
  float rs[FTILE];                                    // per-row inverse power-of-two scale of the z_ operand
  TcHeader hdr;                                       // weight / activation scales (tc_weight_prep_kernel)
  float Os[FTILE * FLD];                              // fp32 tile: z_ -> LN1 input -> o -> LN2 input -> r
  __align__(128) unsigned char ZH[2 * tc::A_PART_BYTES];   // A operand: z_ (h,m), later the FFN hidden chunk
  __align__(128) unsigned char Ob[2 * tc::A_PART_BYTES];   // A operand: o = LN1(z + x) (h,m)
  __align__(128) unsigned char Wzb[2 * tc::B_PART_BYTES];  // B operand: Wz (h,m)
  __align__(128) unsigned char Wc[2][4 * tc::B_PART_BYTES];// B operands: W1 chunk (h,m) + W2 chunk (h,m), double-buffered
  __align__(8) uint64_t mbar;
  uint32_t tmem_slot;
};
static_assert(sizeof(FusedTcSmem) <= 227 * 1024, "fused tc kernel shared memory exceeds 227 KB");

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(tc::smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void copy_block_async(unsigned char* dst, const unsigned char* src, int bytes) {
  for (int i = threadIdx.x; i < bytes / 16; i += FNT) cp_async16(dst + 16 * i, src + 16 * i);
}
__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
// lane l4 of a particle prefetches one 256-byte row (two lines) of the slot pair starting at s: lanes 0,1 -> K,V of
// slot s, lanes 2,3 -> K,V of slot s+1
__device__ __forceinline__ void prefetch_rows(const float* Kb, const float* Vb, int slot, int row, int P, int l4) {
  const float* base = ((l4 & 1) ? Vb : Kb) + (size_t)((uint32_t)(slot * P + row) * FD);
  prefetch_l2(base);
  prefetch_l2(base + 32);
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float v[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// one online-softmax update over a block of FSB slots (scores already in the log2 domain)
__device__ __forceinline__ void softmax_block(const float sc[FSB], const float4 vr[FSB][4], float& m, float& l,
                                              float4 zacc[4]) {
  float mn = m;
#pragma unroll
  for (int j = 0; j < FSB; ++j) mn = fmaxf(mn, sc[j]);
  const float corr = ex2_approx(m - mn);
  float pj[FSB], psum = 0.f;
#pragma unroll
  for (int j = 0; j < FSB; ++j) { pj[j] = ex2_approx(sc[j] - mn); psum += pj[j]; }
  l = fmaf(l, corr, psum);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float4 zz = make_float4(zacc[i].x * corr, zacc[i].y * corr, zacc[i].z * corr, zacc[i].w * corr);
#pragma unroll
    for (int j = 0; j < FSB; ++j) {
      zz.x = fmaf(pj[j], vr[j][i].x, zz.x); zz.y = fmaf(pj[j], vr[j][i].y, zz.y);
      zz.z = fmaf(pj[j], vr[j][i].z, zz.z); zz.w = fmaf(pj[j], vr[j][i].w, zz.w);
    }
    zacc[i] = zz;
  }
  m = mn;
}

__global__ void __launch_bounds__(FNT, 1) fused_tc_forward_kernel(const FusedArgs a) {
  extern __shared__ unsigned char smem_raw_tc[];
  FusedTcSmem& sm = *reinterpret_cast<FusedTcSmem*>(
      smem_raw_tc + ((128u - (tc::smem_u32(smem_raw_tc) & 127u)) & 127u));   // 128-byte aligned operand buffers
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int b = blockIdx.x;
  const int P = a.P, S = a.S, Sa = a.Sa, Fy = a.Fy;
  const int ntiles = (P + FTILE - 1) / FTILE;
  const long long bP = (long long)b * P;
  unsigned short* Acur = a.A0 + bP * Sa;
  unsigned short* Anxt = a.A1 + bP * Sa;
  int cur_is_A0 = 1;
  float* const Kb = a.Kp + (long long)b * S * P * FD;
  float* const Vb = a.Vp + (long long)b * S * P * FD;
  float* const Rb = a.Rp + (long long)b * S * P * FD;
  const int pl = tid >> 2, l4 = tid & 3;          // attention: particle-in-tile, lane-in-particle
  const int er = (wid & 3) * 32 + lane;           // epilogue: row of the tile (= TMEM lane)
  const int ecq = wid >> 2;                       // epilogue: column quarter, cols [16*ecq, 16*ecq+16)
  const uint32_t tlane = ((uint32_t)((wid & 3) * 32)) << 16;
  const uint32_t idesc = tc::make_idesc(128, 64);
  double dloss_q = 0.0, dloss_z = 0.0;            // thread 0 only
  uint32_t mphase = 0;

  for (int i = tid; i < P; i += FNT) { sm.c.Lq[i] = (unsigned short)i; sm.c.pos[i] = (unsigned short)i; }
  if (tid < 4 * FD) {
    const int which = tid >> 6, d = tid & 63;
    const float* lv = which == 0 ? a.p.logvar_k : (which == 1 ? a.p.logvar_q : (which == 2 ? a.p.logvar_v : a.p.logvar_z));
    const float v = lv[a.logvar_dim > 1 ? d : 0];
    sm.c.stdtab[which][d] = expf(0.5f * v);
    sm.c.ivar[which][d] = expf(-v);
  }
  for (int i = tid; i < FD; i += FNT) {
    sm.bz[i] = a.p.bz[i]; sm.b2[i] = a.p.b2[i];
    sm.ln1g[i] = a.p.ln1_g[i]; sm.ln1b[i] = a.p.ln1_b[i]; sm.ln2g[i] = a.p.ln2_g[i]; sm.ln2b[i] = a.p.ln2_b[i];
  }
  for (int i = tid; i < FF; i += FNT) sm.b1[i] = a.p.b1[i];
  for (int i = tid; i < FD * Fy; i += FNT) sm.Wo[i] = a.p.Wo[i];
  if (tid == 0) sm.hdr = *reinterpret_cast<const TcHeader*>(a.wtc + (size_t)(1 + 2 * (FF / FD)) * 2 * tc::B_PART_BYTES);
  if (wid == 0) tc::tmem_alloc(&sm.tmem_slot, 256);
  if (tid == 0) { tc::mbar_init(&sm.mbar, 1); tc::fence_mbar_init(); }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = sm.tmem_slot;
  const uint32_t tD_z = tmem, tD_h = tmem + 64, tD_r = tmem + 128;

  for (int t = 0; t < S; ++t) {
    const int s_lo = (a.attn_window >= 0 && t > a.attn_window) ? t - a.attn_window : 0;
    if (tid < 4 * FD) {
      const int which = tid >> 6, d = tid & 63;
      const float* src = which == 0 ? a.xln : (which == 1 ? a.q_ : (which == 2 ? a.k_ : a.v_));
      sm.c.rowbuf[which][d] = src[((long long)b * S + t) * FD + d];
    }
    __syncthreads();
    float loss_q = 0.f, loss_z = 0.f;

    for (int g = 0; g < ntiles; ++g) {
      const int q0 = g * FTILE;
      const int np = min(FTILE, P - q0);
      // weight blocks for this tile: Wz and FFN chunk 0 travel while the tile does attention
      copy_block_async(sm.Wzb, a.wtc, 2 * tc::B_PART_BYTES);
      copy_block_async(sm.Wc[0], a.wtc + 2 * tc::B_PART_BYTES, 4 * tc::B_PART_BYTES);
      cp_async_commit();

      // =========================== phase A: noise, slot-t rows, attention ===========================
      {
        const int q = q0 + pl;
        const bool act = pl < np;
        const int qc = act ? q : (P - 1);
        const int L = sm.c.Lq[qc];
        const unsigned long long gbp = (unsigned long long)(a.b_global0 + b) * (unsigned long long)P + L;
        const long long lrow = ((bP + L) * S + t) * FD;
        const uint32_t prow = (uint32_t)(t * P + qc) * FD;
        const unsigned short* Arow = Acur + (long long)qc * Sa;
        const int nfull = (t - s_lo) / FSB;   // blocks made only of slots < t
        {  // L2 prefetch of the first history blocks: they travel while the noise is generated
          const int npro = min(a.pf_pro, nfull);
          for (int pb = 0; pb < npro; ++pb) {
            const int sp = s_lo + pb * FSB + (l4 >> 1);
            prefetch_rows(Kb, Vb, sp, (int)Arow[sp], P, l4);
          }
        }
        float4 qv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c = l4 + 4 * i, d0 = 4 * c;
          const float4 q_0 = *reinterpret_cast<const float4*>(&sm.c.rowbuf[1][d0]);
          const float4 k_0 = *reinterpret_cast<const float4*>(&sm.c.rowbuf[2][d0]);
          const float4 v_0 = *reinterpret_cast<const float4*>(&sm.c.rowbuf[3][d0]);
          float ek[4], eq[4], ev[4];
          if (a.eps) {
            const long long BPD = (long long)a.B * P * FD;
            const long long e0 = ((long long)t * a.B + b) * P * FD + (long long)L * FD + d0;
            const float4 t0 = *reinterpret_cast<const float4*>(a.eps + (0LL * S) * BPD + e0);
            const float4 t1 = *reinterpret_cast<const float4*>(a.eps + (1LL * S) * BPD + e0);
            const float4 t2 = *reinterpret_cast<const float4*>(a.eps + (2LL * S) * BPD + e0);
            ek[0] = t0.x; ek[1] = t0.y; ek[2] = t0.z; ek[3] = t0.w;
            eq[0] = t1.x; eq[1] = t1.y; eq[2] = t1.z; eq[3] = t1.w;
            ev[0] = t2.x; ev[1] = t2.y; ev[2] = t2.z; ev[3] = t2.w;
          } else {
            philox_normal4(a.seed, 0u, (uint32_t)t, gbp * 16ull + c, ek);
            philox_normal4(a.seed, 1u, (uint32_t)t, gbp * 16ull + c, eq);
            philox_normal4(a.seed, 2u, (uint32_t)t, gbp * 16ull + c, ev);
          }
          const float kin[4] = {k_0.x, k_0.y, k_0.z, k_0.w};
          const float qin[4] = {q_0.x, q_0.y, q_0.z, q_0.w};
          const float vin[4] = {v_0.x, v_0.y, v_0.z, v_0.w};
          float ko[4], qo[4], vo[4], nq[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            ko[e] = __fadd_rn(kin[e], __fmul_rn(sm.c.stdtab[0][d0 + e], ek[e]));
            qo[e] = __fadd_rn(qin[e], __fmul_rn(sm.c.stdtab[1][d0 + e], eq[e]));
            vo[e] = __fadd_rn(vin[e], __fmul_rn(sm.c.stdtab[2][d0 + e], ev[e]));
            nq[e] = qo[e] - qin[e];
            loss_q = act ? fmaf(nq[e] * nq[e], sm.c.ivar[1][d0 + e], loss_q) : loss_q;
          }
          // scores are formed in the log2 domain: q is pre-scaled by log2(e)/sqrt(d_k)
          const float qs = 0.125f * 1.4426950408889634f;
          qv[i] = make_float4(qo[0] * qs, qo[1] * qs, qo[2] * qs, qo[3] * qs);
          if (act) {
            *reinterpret_cast<float4*>(Kb + prow + d0) = make_float4(ko[0], ko[1], ko[2], ko[3]);
            *reinterpret_cast<float4*>(Vb + prow + d0) = make_float4(vo[0], vo[1], vo[2], vo[3]);
            if (a.noise_q) __stcs(reinterpret_cast<float4*>(a.noise_q + lrow + d0), make_float4(nq[0], nq[1], nq[2], nq[3]));
          }
        }
        // causal attention over the genealogy: slot s lives in physical row A[q][s] (s < t) or q (s == t)
        float m = -INFINITY, l = 0.f;
        float4 zacc[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) zacc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        int s = s_lo;
        const int pfd = a.pf_dist;
        int rpf = (pfd > 0 && pfd < nfull) ? (int)Arow[s_lo + pfd * FSB + (l4 >> 1)] : 0;
        int rn[FSB];
#pragma unroll
        for (int j = 0; j < FSB; ++j) rn[j] = (nfull > 0) ? (int)Arow[s_lo + j] : 0;
        for (int blk = 0; blk < nfull; ++blk, s += FSB) {
          uint32_t off[FSB];
#pragma unroll
          for (int j = 0; j < FSB; ++j) off[j] = (uint32_t)((s + j) * P + rn[j]) * FD + 4 * l4;
          if (blk + 1 < nfull) {
#pragma unroll
            for (int j = 0; j < FSB; ++j) rn[j] = (int)Arow[s + FSB + j];
          }
          if (pfd > 0 && blk + pfd < nfull) {
            prefetch_rows(Kb, Vb, s + pfd * FSB + (l4 >> 1), rpf, P, l4);
            if (blk + pfd + 1 < nfull) rpf = (int)Arow[s + (pfd + 1) * FSB + (l4 >> 1)];
          }
          float sc[FSB];
          float4 vr[FSB][4];
#pragma unroll
          for (int j = 0; j < FSB; ++j) {
            float4 kr[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              kr[i] = *reinterpret_cast<const float4*>(Kb + off[j] + 16 * i);
              vr[j][i] = *reinterpret_cast<const float4*>(Vb + off[j] + 16 * i);
            }
            float part = dot4(qv[1], kr[1], dot4(qv[0], kr[0], 0.f)) + dot4(qv[3], kr[3], dot4(qv[2], kr[2], 0.f));
            part += __shfl_xor_sync(SMCT_FULL_MASK, part, 1);
            part += __shfl_xor_sync(SMCT_FULL_MASK, part, 2);
            sc[j] = part;
          }
          softmax_block(sc, vr, m, l, zacc);
        }
        {  // tail block: the remaining slots < t (fewer than FSB) and the current slot t
          float sc[FSB];
          float4 vr[FSB][4];
#pragma unroll
          for (int j = 0; j < FSB; ++j) {
            const int sj = s + j;
            const bool val = sj <= t;
            const int sjc = val ? sj : t;
            const int row = (sjc < t) ? (int)Arow[sjc] : qc;
            const uint32_t off = (uint32_t)(sjc * P + row) * FD + 4 * l4;
            float part = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 kr = *reinterpret_cast<const float4*>(Kb + off + 16 * i);
              vr[j][i] = *reinterpret_cast<const float4*>(Vb + off + 16 * i);
              part = dot4(qv[i], kr, part);
            }
            part += __shfl_xor_sync(SMCT_FULL_MASK, part, 1);
            part += __shfl_xor_sync(SMCT_FULL_MASK, part, 2);
            sc[j] = val ? part : -INFINITY;
          }
          softmax_block(sc, vr, m, l, zacc);
        }
        const float inv_l = act ? 1.0f / l : 0.f;
        float rmax = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          zacc[i] = make_float4(zacc[i].x * inv_l, zacc[i].y * inv_l, zacc[i].z * inv_l, zacc[i].w * inv_l);
          rmax = fmaxf(rmax, fmaxf(fmaxf(fabsf(zacc[i].x), fabsf(zacc[i].y)), fmaxf(fabsf(zacc[i].z), fabsf(zacc[i].w))));
        }
        rmax = fmaxf(rmax, __shfl_xor_sync(SMCT_FULL_MASK, rmax, 1));
        rmax = fmaxf(rmax, __shfl_xor_sync(SMCT_FULL_MASK, rmax, 2));
        const int ez = tc::scale_exponent(rmax);            // exact power-of-two scale of this row's fp16 operand
        const float zs = ldexpf(1.f, ez);
        if (l4 == 0) sm.rs[pl] = ldexpf(1.f, -ez);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int d0 = 4 * (l4 + 4 * i);
          *reinterpret_cast<float4*>(&sm.Os[pl * FLD + d0]) = zacc[i];
          const float zz[4] = {zacc[i].x * zs, zacc[i].y * zs, zacc[i].z * zs, zacc[i].w * zs};
          tc::store_a4(sm.ZH, pl, d0, zz);
        }
      }
      cp_async_wait_all();
      tc::fence_async_smem();
      tc::fence_before_sync();
      __syncthreads();
      tc::fence_after_sync();

      // =========================== phase B: dense tile on the tensor cores ===========================
      if (tid == 0) {
        tc::mma_block_bf16x3(tD_z, tc::smem_u32(sm.ZH), tc::smem_u32(sm.Wzb), tc::B_PART_BYTES, idesc, false);
        tc::commit(&sm.mbar);
      }
      tc::mbar_wait(&sm.mbar, mphase);
      mphase ^= 1;
      tc::fence_after_sync();
      const bool eact = er < np;
      const int eL = sm.c.Lq[eact ? (q0 + er) : (P - 1)];
      const int n0 = 16 * ecq;
      float val[16];
      {  // ---- z = dense(z_) + noise ; LN1(z + x) ----
        tmem_ld16(tD_z + tlane + n0, val);
        const float zscale = sm.rs[er] * sm.hdr.inv_wz;
        const unsigned long long gbp = (unsigned long long)(a.b_global0 + b) * (unsigned long long)P + eL;
        float s1 = 0.f;
#pragma unroll
        for (int gq = 0; gq < 4; ++gq) {
          const int n = n0 + 4 * gq;
          float ez[4];
          if (a.eps) {
            const float4 t3 = *reinterpret_cast<const float4*>(a.eps + (3LL * S) * ((long long)a.B * P * FD) +
                                                               ((long long)t * a.B + b) * P * FD + (long long)eL * FD + n);
            ez[0] = t3.x; ez[1] = t3.y; ez[2] = t3.z; ez[3] = t3.w;
          } else {
            philox_normal4(a.seed, 3u, (uint32_t)t, gbp * 16ull + (n >> 2), ez);
          }
          const float4 zold = *reinterpret_cast<const float4*>(&sm.Os[er * FLD + n]);
          const float z_[4] = {zold.x, zold.y, zold.z, zold.w};
          float nzv[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float zp = fmaf(val[4 * gq + e], zscale, sm.bz[n + e]);   // scale is a power of two: exact
            const float z = __fadd_rn(zp, __fmul_rn(sm.c.stdtab[3][n + e], ez[e]));
            nzv[e] = (a.noise_z_mode == SMCT_NOISE_Z_CHECKOUT) ? (z - z_[e]) : (z - zp);
            loss_z = eact ? fmaf(nzv[e] * nzv[e], sm.c.ivar[3][n + e], loss_z) : loss_z;
            val[4 * gq + e] = z + sm.c.rowbuf[0][n + e];
            s1 += val[4 * gq + e];
          }
          if (eact && a.noise_z)
            __stcs(reinterpret_cast<float4*>(a.noise_z + ((bP + eL) * S + t) * FD + n), make_float4(nzv[0], nzv[1], nzv[2], nzv[3]));
        }
        sm.psum[0][ecq][er] = s1;
        __syncthreads();
        const float mean = (((sm.psum[0][0][er] + sm.psum[0][1][er]) + sm.psum[0][2][er]) + sm.psum[0][3][er]) * (1.0f / FD);
        float s2 = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) { const float cdev = val[j] - mean; s2 = fmaf(cdev, cdev, s2); }
        sm.psum[1][ecq][er] = s2;
        __syncthreads();
        const float var = (((sm.psum[1][0][er] + sm.psum[1][1][er]) + sm.psum[1][2][er]) + sm.psum[1][3][er]) * (1.0f / FD);
        const float rstd = 1.0f / sqrtf(var + a.ln_eps);
#pragma unroll
        for (int gq = 0; gq < 4; ++gq) {
          const int n = n0 + 4 * gq;
          float o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float inv = rstd * sm.ln1g[n + e];
            o[e] = val[4 * gq + e] * inv + (sm.ln1b[n + e] - mean * inv);
          }
          *reinterpret_cast<float4*>(&sm.Os[er * FLD + n]) = make_float4(o[0], o[1], o[2], o[3]);
          const float os[4] = {o[0] * sm.hdr.s_o, o[1] * sm.hdr.s_o, o[2] * sm.hdr.s_o, o[3] * sm.hdr.s_o};
          tc::store_a4(sm.Ob, er, n, os);
        }
      }
      tc::fence_async_smem();
      tc::fence_before_sync();
      __syncthreads();
      tc::fence_after_sync();
      // ---- FFN: 4 chunks of 64 hidden units ----
      for (int c = 0; c < FF / FD; ++c) {
        unsigned char* wb = sm.Wc[c & 1];
        if (tid == 0) {
          tc::mma_block_bf16x3(tD_h, tc::smem_u32(sm.Ob), tc::smem_u32(wb), tc::B_PART_BYTES, idesc, false);
          tc::commit(&sm.mbar);
        }
        tc::mbar_wait(&sm.mbar, mphase);   // also covers the previous chunk's second GEMM: ZH and Wc[(c+1)&1] are free
        mphase ^= 1;
        tc::fence_after_sync();
        if (c + 1 < FF / FD) {
          copy_block_async(sm.Wc[(c + 1) & 1], a.wtc + 2 * tc::B_PART_BYTES + (size_t)(c + 1) * 4 * tc::B_PART_BYTES,
                           4 * tc::B_PART_BYTES);
          cp_async_commit();
        }
        tmem_ld16(tD_h + tlane + n0, val);
        const float hscale = sm.hdr.inv_o * sm.hdr.inv_w1;
#pragma unroll
        for (int gq = 0; gq < 4; ++gq) {
          const int n = n0 + 4 * gq;
          float h[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) h[e] = fmaxf(fmaf(val[4 * gq + e], hscale, sm.b1[c * FD + n + e]), 0.f) * sm.hdr.s_h;
          tc::store_a4(sm.ZH, er, n, h);
        }
        cp_async_wait_all();
        tc::fence_async_smem();
        tc::fence_before_sync();
        __syncthreads();
        tc::fence_after_sync();
        if (tid == 0) {
          tc::mma_block_bf16x3(tD_r, tc::smem_u32(sm.ZH), tc::smem_u32(wb + 2 * tc::B_PART_BYTES), tc::B_PART_BYTES, idesc, c > 0);
          if (c + 1 == FF / FD) tc::commit(&sm.mbar);
        }
      }
      tc::mbar_wait(&sm.mbar, mphase);
      mphase ^= 1;
      tc::fence_after_sync();
      {  // ---- r = LN2(ffn + b2 + o) ; R row store ----
        tmem_ld16(tD_r + tlane + n0, val);
        const float rscale = sm.hdr.inv_h * sm.hdr.inv_w2;
        float s1 = 0.f;
#pragma unroll
        for (int gq = 0; gq < 4; ++gq) {
          const int n = n0 + 4 * gq;
          const float4 o4 = *reinterpret_cast<const float4*>(&sm.Os[er * FLD + n]);
          const float o[4] = {o4.x, o4.y, o4.z, o4.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            val[4 * gq + e] = fmaf(val[4 * gq + e], rscale, sm.b2[n + e]) + o[e];
            s1 += val[4 * gq + e];
          }
        }
        sm.psum[0][ecq][er] = s1;
        __syncthreads();
        const float mean = (((sm.psum[0][0][er] + sm.psum[0][1][er]) + sm.psum[0][2][er]) + sm.psum[0][3][er]) * (1.0f / FD);
        float s2 = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) { const float cdev = val[j] - mean; s2 = fmaf(cdev, cdev, s2); }
        sm.psum[1][ecq][er] = s2;
        __syncthreads();
        const float var = (((sm.psum[1][0][er] + sm.psum[1][1][er]) + sm.psum[1][2][er]) + sm.psum[1][3][er]) * (1.0f / FD);
        const float rstd = 1.0f / sqrtf(var + a.ln_eps);
        float* rrow = Rb + (uint32_t)(t * P + (eact ? q0 + er : 0)) * FD;
#pragma unroll
        for (int gq = 0; gq < 4; ++gq) {
          const int n = n0 + 4 * gq;
          float r4[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float inv = rstd * sm.ln2g[n + e];
            r4[e] = val[4 * gq + e] * inv + (sm.ln2b[n + e] - mean * inv);
          }
          *reinterpret_cast<float4*>(&sm.Os[er * FLD + n]) = make_float4(r4[0], r4[1], r4[2], r4[3]);
          if (eact) __stcs(reinterpret_cast<float4*>(rrow + n), make_float4(r4[0], r4[1], r4[2], r4[3]));   // read again only by the final pass
        }
      }
      tc::fence_before_sync();
      __syncthreads();
      // ---- output layer + Gaussian log-weight: 4 lanes per row ----
      {
        const bool act = pl < np;
        const int L = sm.c.Lq[act ? (q0 + pl) : (P - 1)];
        float rv[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 x4 = *reinterpret_cast<const float4*>(&sm.Os[pl * FLD + 16 * l4 + 4 * i]);
          rv[4 * i] = x4.x; rv[4 * i + 1] = x4.y; rv[4 * i + 2] = x4.z; rv[4 * i + 3] = x4.w;
        }
        float sq = 0.f;
        for (int f = 0; f < Fy; ++f) {
          float part = 0.f;
#pragma unroll
          for (int j = 0; j < 16; ++j) part = fmaf(rv[j], sm.Wo[(16 * l4 + j) * Fy + f], part);
          part += __shfl_xor_sync(SMCT_FULL_MASK, part, 1);
          part += __shfl_xor_sync(SMCT_FULL_MASK, part, 2);
          if (act && l4 == 0 && a.pred) a.pred[((bP + L) * S + t) * Fy + f] = part;
          const float diff = a.y[((long long)b * S + t) * Fy + f] - part;
          sq = fmaf(diff, diff, sq);
        }
        if (act && l4 == 0) sm.c.lw[L] = (-0.5f * sq) / a.sigma_obs;
      }
      __syncthreads();
    }  // tiles

    if (a.loss_sums) {
      const float sq = warp_sum(loss_q), sz = warp_sum(loss_z);
      if (lane == 0) { sm.c.red[wid] = sq; sm.c.red[16 + wid] = sz; }
      __syncthreads();
      if (tid == 0) {
        for (int i = 0; i < 16; ++i) { dloss_q += (double)sm.c.red[i]; dloss_z += (double)sm.c.red[16 + i]; }
      }
    }
    fused_resample_update(sm.c, a, b, t, Acur, Anxt, cur_is_A0);
  }  // timesteps

  for (int p = tid; p < P; p += FNT) a.pos_out[bP + p] = sm.c.pos[p];
  if (tid == 0) {
    a.aflag[b] = cur_is_A0;
    if (a.loss_sums) { atomicAdd(a.loss_sums + 1, dloss_q); atomicAdd(a.loss_sums + 3, dloss_z); }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (wid == 0) tc::tmem_dealloc(tmem, 256);
}

}  // namespace smct

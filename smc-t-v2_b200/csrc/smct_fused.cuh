// smct_fused.cuh — fused persistent forward (SMCT_ALGO_FUSED): ONE launch runs all S particle-filter
// timesteps of a chunk of sequences; one CTA owns one sequence for the whole trajectory, so resampling
// (which couples the particles of a sequence) needs no grid-wide synchronisation.
//
// B200-first design decisions:
//  * path storage instead of a physical ancestor gather: state rows are written once, slot-major
//    (Kp/Vp/Rp[b][s][row][D]); each particle carries an ancestry row A[q][s] (uint16) naming the physical
//    row it inherits at slot s.  Resampling copies 2·t bytes per particle instead of 3·t·D·4.
//  * particles are kept in genealogy (tree) order: after every draw they are sorted by the position of
//    their ancestor, so neighbouring lanes read the same history rows — one fetch per warp instead of
//    one per particle, and shared rows stay L1/L2 resident.
//  * 4 lanes per particle, 128-bit loads, 2-stage shuffle reductions, online softmax over blocks of slots.
//  * the dense part (z projection, LN1, FFN, LN2, output layer, log-weight) is a 128-particle tile GEMM
//    pipeline in shared memory with register-prefetched weight blocks.
//  * the categorical draw keeps TF's CPU semantics: fp64 running sum in logical particle order.
// Results are identical to the staged path / the reference semantics; only the data movement differs.
#pragma once
#include "smct_common.cuh"
#include "../../include/smct.h"

namespace smct {

constexpr int FD = 64;        // d_model handled by the fused kernel
constexpr int FF = 256;       // largest dff (the kernels serve dff = 64, 128, 192, 256: chunks of 64 hidden units)
constexpr int FNT = 512;      // threads per CTA
constexpr int FTILE = 128;    // particles per tile (4 lanes each)
constexpr int FLD = 68;       // row stride (floats) of the activation tiles in shared memory
constexpr int FPMAX = 1024;   // max particles
constexpr int FSB = 2;        // slots per online-softmax block

struct FusedArgs {
  int B, P, S, Sa, Fy, dff;
  int attn_window, len_resampling, noise_z_mode, logvar_dim;
  float ln_eps, sigma_obs;
  long long b_global0;
  unsigned long long seed;
  const float *xln, *q_, *k_, *v_;   // (B,S,64) particle-independent rows (rows_kernel)
  const float* y;                    // (B,S,Fy)
  const float* eps;                  // optional (4,S,B,P,64)
  const double* u;                   // optional (S,B,P)
  const int* i_forced;               // optional (S,B,P)
  smct_params p;
  const unsigned char* wtc;          // bf16 (h,m) weight blocks in the tcgen05 operand layout (tensor-core variant)
  float *Kp, *Vp, *Rp;               // (B,S,P,64) slot-major physical rows
  unsigned char *Ks, *Vs;            // (B,S,P,256 B) fp16 (h,m) split copies of the K/V rows (tensor-core variant)
  unsigned short *A0, *A1;           // (B,P,Sa) ancestry rows, ping-pong
  unsigned short* pos_out;           // (B,P)  final position of each logical particle
  int* aflag;                        // (B)    which ancestry buffer is final
  float *pred, *w, *logw, *noise_q, *noise_z;
  int* i_out;
  double* loss_sums;
};

struct FusedResampleScratch {  // used only between the tiles of two timesteps (draw + genealogy update)
  double cdf[FPMAX];
  unsigned int keys[FPMAX];
  unsigned short src[FPMAX], anc[FPMAX];
};

struct FusedCommon {           // per-sequence bookkeeping shared by both fused kernels
  float lw[FPMAX];
  unsigned short Lq[FPMAX], pos[FPMAX];
  __align__(16) float rowbuf[4][FD];   // xln, q_, k_, v_ of the current step (16-byte aligned: read as float4)
  __align__(16) float stdtab[4][FD];   // exp(0.5 logvar) for k,q,v,z
  __align__(16) float ivar[4][FD];     // exp(-logvar)
  float red[64];
  double dred[2];
};

struct FusedSmem {
  FusedCommon c;
  FusedResampleScratch rs;
  float Zs[FTILE * FLD];
  float Os[FTILE * FLD];
  float Hs[FTILE * FLD];
  float WB[2][FD * FD];
};

__device__ __forceinline__ float dot4(const float4& a, const float4& b, float acc) {
  acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
  return acc;
}

// acc[4][4] (rows 4ty..4ty+3, cols 4tx..4tx+3) += As[128][64] (stride FLD) * Bs[64][64]
__device__ __forceinline__ void gemm_tile(const float* __restrict__ As, const float* __restrict__ Bs, float acc[4][4],
                                          int tx, int ty) {
#pragma unroll 2
  for (int k0 = 0; k0 < FD; k0 += 4) {
    float4 av[4], bv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) av[i] = *reinterpret_cast<const float4*>(As + (4 * ty + i) * FLD + k0);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) bv[kk] = *reinterpret_cast<const float4*>(Bs + (k0 + kk) * FD + 4 * tx);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      acc[i][0] = fmaf(av[i].x, bv[0].x, acc[i][0]); acc[i][1] = fmaf(av[i].x, bv[0].y, acc[i][1]);
      acc[i][2] = fmaf(av[i].x, bv[0].z, acc[i][2]); acc[i][3] = fmaf(av[i].x, bv[0].w, acc[i][3]);
      acc[i][0] = fmaf(av[i].y, bv[1].x, acc[i][0]); acc[i][1] = fmaf(av[i].y, bv[1].y, acc[i][1]);
      acc[i][2] = fmaf(av[i].y, bv[1].z, acc[i][2]); acc[i][3] = fmaf(av[i].y, bv[1].w, acc[i][3]);
      acc[i][0] = fmaf(av[i].z, bv[2].x, acc[i][0]); acc[i][1] = fmaf(av[i].z, bv[2].y, acc[i][1]);
      acc[i][2] = fmaf(av[i].z, bv[2].z, acc[i][2]); acc[i][3] = fmaf(av[i].z, bv[2].w, acc[i][3]);
      acc[i][0] = fmaf(av[i].w, bv[3].x, acc[i][0]); acc[i][1] = fmaf(av[i].w, bv[3].y, acc[i][1]);
      acc[i][2] = fmaf(av[i].w, bv[3].z, acc[i][2]); acc[i][3] = fmaf(av[i].w, bv[3].w, acc[i][3]);
    }
  }
}

// weight block `blk` of the per-tile sequence: 0 = Wz, 1+2c = W1[:, 64c:64c+64], 2+2c = W2[64c:64c+64, :]
__device__ __forceinline__ void wblock_issue(const smct_params& p, int dff, int blk, int tid, float4 wr[2]) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int e = tid + h * FNT;          // float4 index within the 64x64 block (1024 float4)
    const int r = e >> 4, c4 = e & 15;    // row, float4 column
    const float* src;
    if (blk == 0) src = p.Wz + r * FD + 4 * c4;
    else if (blk & 1) src = p.W1 + r * dff + ((blk - 1) >> 1) * FD + 4 * c4;
    else src = p.W2 + (((blk - 2) >> 1) * FD + r) * FD + 4 * c4;
    wr[h] = __ldg(reinterpret_cast<const float4*>(src));
  }
}
__device__ __forceinline__ void wblock_commit(float* WBbuf, int tid, const float4 wr[2]) {
#pragma unroll
  for (int h = 0; h < 2; ++h) reinterpret_cast<float4*>(WBbuf)[tid + h * FNT] = wr[h];
}

// logsumexp normalise, TF-CPU categorical draw (fp64 sequential CDF in logical particle order), then the
// genealogy update: sort the new particles by the position of their ancestor (tree order) and rewrite the
// ancestry rows.  Called by all threads of the CTA once per timestep.
__device__ __forceinline__ void fused_resample_update(FusedCommon& c, FusedResampleScratch& rx, const FusedArgs& a, int b, int t,
                                                      unsigned short*& Acur, unsigned short*& Anxt, int& cur_is_A0) {
  const int tid = threadIdx.x;
  const int P = a.P, S = a.S, Sa = a.Sa;
  const long long bP = (long long)b * P;
  // =========================== weights, categorical draw (TF CPU semantics) ===========================
  float m1 = -INFINITY;
  for (int p = tid; p < P; p += FNT) { const float v = c.lw[p]; if (isfinite(v)) m1 = fmaxf(m1, v); }
  m1 = block_max(m1, c.red);
  float s1 = 0.f;
  for (int p = tid; p < P; p += FNT) s1 += expf(c.lw[p] - m1);
  s1 = block_sum(s1, c.red);
  const float lse = m1 + logf(s1);
  float m2 = -INFINITY;
  for (int p = tid; p < P; p += FNT) { const float v = c.lw[p] - lse; if (isfinite(v)) m2 = fmaxf(m2, v); }
  m2 = block_max(m2, c.red);
  for (int p = tid; p < P; p += FNT) {
    const float raw = c.lw[p];
    const float lg = raw - lse;
    if (a.w) a.w[(bP + p) * S + t] = expf(lg);
    if (a.logw) a.logw[((long long)t * a.B + b) * P + p] = raw;
    rx.cdf[p] = isfinite(lg) ? exp((double)lg - (double)m2) : 0.0;
  }
  __syncthreads();
  if (tid == 0) {
    double run = 0.0;
    for (int p = 0; p < P; ++p) { run += rx.cdf[p]; rx.cdf[p] = run; }
    c.dred[0] = run;
  }
  __syncthreads();
  const double total = c.dred[0];
  for (int p = tid; p < P; p += FNT) {
    const long long ti = ((long long)t * a.B + b) * P + p;
    int idx;
    if (a.i_forced) {
      idx = clamp_index(a.i_forced[ti], P);
    } else {
      const double uu = a.u ? a.u[ti]
                            : philox_uniform(a.seed, (uint32_t)t,
                                             (unsigned long long)(a.b_global0 + b) * (unsigned long long)P + p);
      const double target = uu * total;
      int lo = 0, hi = P;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (rx.cdf[mid] <= target) lo = mid + 1; else hi = mid;
      }
      idx = min(lo, P - 1);
    }
    if (a.i_out) a.i_out[ti] = idx;
    rx.anc[p] = (unsigned short)idx;
  }
  __syncthreads();

  // =========================== genealogy update ===========================
  const bool resample_now = (a.len_resampling < 0) || (t < a.len_resampling);
  if (resample_now) {
    // sort the new particles by the position of their ancestor (ties by logical id): tree order is preserved
    int np2 = 1;
    while (np2 < P) np2 <<= 1;
    for (int p = tid; p < np2; p += FNT)
      rx.keys[p] = (p < P) ? (((unsigned int)c.pos[rx.anc[p]] << 16) | (unsigned int)p) : 0xFFFFFFFFu;
    __syncthreads();
    for (int k = 2; k <= np2; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = tid; i < np2; i += FNT) {
          const int ixj = i ^ j;
          if (ixj > i) {
            const unsigned int kx = rx.keys[i], yv = rx.keys[ixj];
            const bool up = ((i & k) == 0);
            if ((kx > yv) == up) { rx.keys[i] = yv; rx.keys[ixj] = kx; }
          }
        }
        __syncthreads();
      }
    }
    for (int qn = tid; qn < P; qn += FNT) {
      const unsigned int key = rx.keys[qn];
      rx.src[qn] = (unsigned short)(key >> 16);
      c.Lq[qn] = (unsigned short)(key & 0xFFFFu);
    }
    __syncthreads();
    for (int qn = tid; qn < P; qn += FNT) c.pos[c.Lq[qn]] = (unsigned short)qn;
    // ancestry rows: Anxt[q'][0..t) = Acur[src[q']][0..t) ; Anxt[q'][t] = src[q']   (8 slots = one uint4)
    const int nch = (t >> 3) + 1;
    for (int i = tid; i < P * nch; i += FNT) {
      const int qn = i / nch, ch = i - qn * nch;
      const int sq = rx.src[qn];
      uint4 v = *reinterpret_cast<const uint4*>(Acur + (long long)sq * Sa + 8 * ch);
      if ((t >> 3) == ch) {  // patch slot t (little-endian: element e is half (e&1) of word (e>>1))
        const int e = t & 7;
        unsigned int wv = (e >> 1) == 0 ? v.x : ((e >> 1) == 1 ? v.y : ((e >> 1) == 2 ? v.z : v.w));
        wv = (e & 1) ? ((wv & 0x0000FFFFu) | ((unsigned int)sq << 16)) : ((wv & 0xFFFF0000u) | (unsigned int)sq);
        if ((e >> 1) == 0) v.x = wv; else if ((e >> 1) == 1) v.y = wv; else if ((e >> 1) == 2) v.z = wv; else v.w = wv;
      }
      *reinterpret_cast<uint4*>(Anxt + (long long)qn * Sa + 8 * ch) = v;
    }
    unsigned short* tmp = Acur; Acur = Anxt; Anxt = tmp;
    cur_is_A0 ^= 1;
  } else {
    for (int qn = tid; qn < P; qn += FNT) Acur[(long long)qn * Sa + t] = (unsigned short)qn;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(FNT, 1) fused_forward_kernel(const FusedArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FusedSmem& sm = *reinterpret_cast<FusedSmem*>(smem_raw);
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
  const int tx = tid & 15, ty = tid >> 4;       // GEMM thread tile
  const int pl = tid >> 2, l4 = tid & 3;        // attention: particle-in-tile, lane-in-particle
  double dloss_q = 0.0, dloss_z = 0.0;          // thread 0 only

  for (int i = tid; i < P; i += FNT) { sm.c.Lq[i] = (unsigned short)i; sm.c.pos[i] = (unsigned short)i; }
  if (tid < 4 * FD) {
    const int which = tid >> 6, d = tid & 63;
    const float* lv = which == 0 ? a.p.logvar_k : (which == 1 ? a.p.logvar_q : (which == 2 ? a.p.logvar_v : a.p.logvar_z));
    const float v = lv[a.logvar_dim > 1 ? d : 0];
    sm.c.stdtab[which][d] = expf(0.5f * v);
    sm.c.ivar[which][d] = expf(-v);
  }
  __syncthreads();

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
      float4 wr[2];
      wblock_issue(a.p, a.dff, 0, tid, wr);   // Wz block travels while the tile does attention

      // =========================== phase A: noise, slot-t rows, attention ===========================
      {
        const int q = q0 + pl;
        const bool act = pl < np;
        const int qc = act ? q : (P - 1);
        const int L = sm.c.Lq[qc];
        const unsigned long long gbp = (unsigned long long)(a.b_global0 + b) * (unsigned long long)P + L;
        const long long lrow = ((bP + L) * S + t) * FD;              // logical (b,L,t) row of (B,P,S,64) outputs
        const long long prow = ((long long)t * P + qc) * FD;         // physical slot-t row of this particle
        float4 qv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c = l4 + 4 * i, d0 = 4 * c;
          const float4 q_0 = *reinterpret_cast<const float4*>(&sm.c.rowbuf[1][d0]);
          const float4 k_0 = *reinterpret_cast<const float4*>(&sm.c.rowbuf[2][d0]);
          const float4 v_0 = *reinterpret_cast<const float4*>(&sm.c.rowbuf[3][d0]);
          float ek[4], eq[4], ev[4], ez[4];
          if (a.eps) {
            const long long BPD = (long long)a.B * P * FD;
            const long long e0 = ((long long)t * a.B + b) * P * FD + (long long)L * FD + d0;
            const float4 t0 = *reinterpret_cast<const float4*>(a.eps + (0LL * S) * BPD + e0);
            const float4 t1 = *reinterpret_cast<const float4*>(a.eps + (1LL * S) * BPD + e0);
            const float4 t2 = *reinterpret_cast<const float4*>(a.eps + (2LL * S) * BPD + e0);
            const float4 t3 = *reinterpret_cast<const float4*>(a.eps + (3LL * S) * BPD + e0);
            ek[0] = t0.x; ek[1] = t0.y; ek[2] = t0.z; ek[3] = t0.w;
            eq[0] = t1.x; eq[1] = t1.y; eq[2] = t1.z; eq[3] = t1.w;
            ev[0] = t2.x; ev[1] = t2.y; ev[2] = t2.z; ev[3] = t2.w;
            ez[0] = t3.x; ez[1] = t3.y; ez[2] = t3.z; ez[3] = t3.w;
          } else {
            philox_normal4(a.seed, 0u, (uint32_t)t, gbp * 16ull + c, ek);
            philox_normal4(a.seed, 1u, (uint32_t)t, gbp * 16ull + c, eq);
            philox_normal4(a.seed, 2u, (uint32_t)t, gbp * 16ull + c, ev);
            philox_normal4(a.seed, 3u, (uint32_t)t, gbp * 16ull + c, ez);
          }
          const float kin[4] = {k_0.x, k_0.y, k_0.z, k_0.w};
          const float qin[4] = {q_0.x, q_0.y, q_0.z, q_0.w};
          const float vin[4] = {v_0.x, v_0.y, v_0.z, v_0.w};
          float ko[4], qo[4], vo[4], nq[4], nz0[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            ko[e] = __fadd_rn(kin[e], __fmul_rn(sm.c.stdtab[0][d0 + e], ek[e]));
            qo[e] = __fadd_rn(qin[e], __fmul_rn(sm.c.stdtab[1][d0 + e], eq[e]));
            vo[e] = __fadd_rn(vin[e], __fmul_rn(sm.c.stdtab[2][d0 + e], ev[e]));
            nq[e] = qo[e] - qin[e];
            nz0[e] = __fmul_rn(sm.c.stdtab[3][d0 + e], ez[e]);
            loss_q = act ? fmaf(nq[e] * nq[e], sm.c.ivar[1][d0 + e], loss_q) : loss_q;
          }
          qv[i] = make_float4(qo[0], qo[1], qo[2], qo[3]);
          *reinterpret_cast<float4*>(&sm.Os[pl * FLD + d0]) = make_float4(nz0[0], nz0[1], nz0[2], nz0[3]);
          if (act) {
            *reinterpret_cast<float4*>(Kb + prow + d0) = make_float4(ko[0], ko[1], ko[2], ko[3]);
            *reinterpret_cast<float4*>(Vb + prow + d0) = make_float4(vo[0], vo[1], vo[2], vo[3]);
            if (a.noise_q) *reinterpret_cast<float4*>(a.noise_q + lrow + d0) = make_float4(nq[0], nq[1], nq[2], nq[3]);
          }
        }
        // causal attention over the particle's genealogy: slot s lives in physical row A[q][s] (s < t) or q (s == t)
        const unsigned short* Arow = Acur + (long long)qc * Sa;
        float m = -INFINITY, l = 0.f;
        float4 zacc[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) zacc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s = s_lo; s <= t; s += FSB) {
          float sc[FSB];
          float4 vr[FSB][4];
#pragma unroll
          for (int j = 0; j < FSB; ++j) {
            const int sj = s + j;
            const bool val = sj <= t;
            const int row = (sj < t) ? (int)Arow[sj] : qc;
            const long long off = ((long long)(val ? sj : t) * P + row) * FD;
            float part = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int d0 = 4 * (l4 + 4 * i);
              float4 kr = make_float4(0.f, 0.f, 0.f, 0.f);
              vr[j][i] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (val && act) {
                kr = *reinterpret_cast<const float4*>(Kb + off + d0);
                vr[j][i] = *reinterpret_cast<const float4*>(Vb + off + d0);
              }
              part = dot4(qv[i], kr, part);
            }
            part += __shfl_xor_sync(SMCT_FULL_MASK, part, 1);
            part += __shfl_xor_sync(SMCT_FULL_MASK, part, 2);
            sc[j] = val ? part / 8.0f : -INFINITY;   // / sqrt(d_k), d_k = 64
          }
          float mn = m;
#pragma unroll
          for (int j = 0; j < FSB; ++j) mn = fmaxf(mn, sc[j]);
          const float corr = __expf(m - mn);
          float pj[FSB], psum = 0.f;
#pragma unroll
          for (int j = 0; j < FSB; ++j) { pj[j] = __expf(sc[j] - mn); psum += pj[j]; }
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
        const float inv_l = 1.0f / l;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int d0 = 4 * (l4 + 4 * i);
          const float4 zz = act ? make_float4(zacc[i].x * inv_l, zacc[i].y * inv_l, zacc[i].z * inv_l, zacc[i].w * inv_l)
                                : make_float4(0.f, 0.f, 0.f, 0.f);
          *reinterpret_cast<float4*>(&sm.Zs[pl * FLD + d0]) = zz;
        }
      }
      wblock_commit(sm.WB[0], tid, wr);
      __syncthreads();

      // =========================== phase B: dense tile pipeline ===========================
      float acc[4][4];
      // ---- z = dense(z_) + noise ; o_pre = z + x ----
      wblock_issue(a.p, a.dff, 1, tid, wr);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
      gemm_tile(sm.Zs, sm.WB[0], acc, tx, ty);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = 4 * ty + i;
        const bool act = r < np;
        float zo[4], nzv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int n = 4 * tx + j;
          const float zp = acc[i][j] + __ldg(a.p.bz + n);
          const float z = __fadd_rn(zp, sm.Os[r * FLD + n]);
          nzv[j] = (a.noise_z_mode == SMCT_NOISE_Z_CHECKOUT) ? (z - sm.Zs[r * FLD + n]) : (z - zp);
          loss_z = act ? fmaf(nzv[j] * nzv[j], sm.c.ivar[3][n], loss_z) : loss_z;
          zo[j] = z + sm.c.rowbuf[0][n];
        }
        *reinterpret_cast<float4*>(&sm.Os[r * FLD + 4 * tx]) = make_float4(zo[0], zo[1], zo[2], zo[3]);
        if (act && a.noise_z) {
          const int L = sm.c.Lq[q0 + r];
          *reinterpret_cast<float4*>(a.noise_z + ((bP + L) * S + t) * FD + 4 * tx) = make_float4(nzv[0], nzv[1], nzv[2], nzv[3]);
        }
      }
      wblock_commit(sm.WB[1], tid, wr);
      __syncthreads();
      // ---- o = LN1(o_pre): warp per row ----
      for (int r = wid; r < FTILE; r += FNT / 32) {
        const float v0 = sm.Os[r * FLD + lane], v1 = sm.Os[r * FLD + lane + 32];
        const float mean = warp_sum(v0 + v1) * (1.0f / FD);
        const float c0 = v0 - mean, c1 = v1 - mean;
        const float rstd = 1.0f / sqrtf(warp_sum(fmaf(c0, c0, c1 * c1)) * (1.0f / FD) + a.ln_eps);
        const float i0 = rstd * __ldg(a.p.ln1_g + lane), i1 = rstd * __ldg(a.p.ln1_g + lane + 32);
        sm.Os[r * FLD + lane] = v0 * i0 + (__ldg(a.p.ln1_b + lane) - mean * i0);
        sm.Os[r * FLD + lane + 32] = v1 * i1 + (__ldg(a.p.ln1_b + lane + 32) - mean * i1);
      }
      __syncthreads();
      // ---- FFN: r_pre = relu(o W1 + b1) W2 + b2 + o, dff in 4 chunks of 64 ----
      float acc2[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc2[i][j] = 0.f;
      for (int c = 0; c < a.dff / FD; ++c) {
        // GEMM1 chunk c : weight block 1+2c lives in WB[1]; prefetch block 2+2c
        wblock_issue(a.p, a.dff, 2 + 2 * c, tid, wr);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
        gemm_tile(sm.Os, sm.WB[1], acc, tx, ty);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float h[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) h[j] = fmaxf(acc[i][j] + __ldg(a.p.b1 + c * FD + 4 * tx + j), 0.f);
          *reinterpret_cast<float4*>(&sm.Hs[(4 * ty + i) * FLD + 4 * tx]) = make_float4(h[0], h[1], h[2], h[3]);
        }
        wblock_commit(sm.WB[0], tid, wr);
        __syncthreads();
        // GEMM2 chunk c : weight block 2+2c lives in WB[0]; prefetch block 1+2(c+1) (or nothing)
        const bool more = (c + 1) < a.dff / FD;
        if (more) wblock_issue(a.p, a.dff, 1 + 2 * (c + 1), tid, wr);
        gemm_tile(sm.Hs, sm.WB[0], acc2, tx, ty);
        if (more) wblock_commit(sm.WB[1], tid, wr);
        __syncthreads();
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = 4 * ty + i;
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int n = 4 * tx + j;
          o[j] = (acc2[i][j] + __ldg(a.p.b2 + n)) + sm.Os[r * FLD + n];
        }
        *reinterpret_cast<float4*>(&sm.Zs[r * FLD + 4 * tx]) = make_float4(o[0], o[1], o[2], o[3]);
      }
      __syncthreads();
      // ---- r = LN2(r_pre); R row store; output layer; log-weight: warp per row ----
      for (int r = wid; r < np; r += FNT / 32) {
        const int q = q0 + r;
        const int L = sm.c.Lq[q];
        const float v0 = sm.Zs[r * FLD + lane], v1 = sm.Zs[r * FLD + lane + 32];
        const float mean = warp_sum(v0 + v1) * (1.0f / FD);
        const float c0 = v0 - mean, c1 = v1 - mean;
        const float rstd = 1.0f / sqrtf(warp_sum(fmaf(c0, c0, c1 * c1)) * (1.0f / FD) + a.ln_eps);
        const float i0 = rstd * __ldg(a.p.ln2_g + lane), i1 = rstd * __ldg(a.p.ln2_g + lane + 32);
        const float r0 = v0 * i0 + (__ldg(a.p.ln2_b + lane) - mean * i0);
        const float r1 = v1 * i1 + (__ldg(a.p.ln2_b + lane + 32) - mean * i1);
        float* rrow = Rb + ((long long)t * P + q) * FD;
        rrow[lane] = r0;
        rrow[lane + 32] = r1;
        float sq = 0.f;
        for (int f = 0; f < Fy; ++f) {
          const float part = fmaf(r0, __ldg(a.p.Wo + lane * Fy + f), r1 * __ldg(a.p.Wo + (lane + 32) * Fy + f));
          const float pr = warp_sum(part);
          if (lane == 0 && a.pred) a.pred[((bP + L) * S + t) * Fy + f] = pr;
          const float diff = a.y[((long long)b * S + t) * Fy + f] - pr;
          sq = fmaf(diff, diff, sq);
        }
        if (lane == 0) sm.c.lw[L] = (-0.5f * sq) / a.sigma_obs;
      }
      __syncthreads();
    }  // tiles

    // =========================== loss partials of this step ===========================
    if (a.loss_sums) {
      const float sq = warp_sum(loss_q), sz = warp_sum(loss_z);
      if (lane == 0) { sm.c.red[wid] = sq; sm.c.red[16 + wid] = sz; }
      __syncthreads();
      if (tid == 0) {
        for (int i = 0; i < 16; ++i) { dloss_q += (double)sm.c.red[i]; dloss_z += (double)sm.c.red[16 + i]; }
      }
    }

    fused_resample_update(sm.c, sm.rs, a, b, t, Acur, Anxt, cur_is_A0);
  }  // timesteps

  for (int p = tid; p < P; p += FNT) a.pos_out[bP + p] = sm.c.pos[p];
  if (tid == 0) {
    a.aflag[b] = cur_is_A0;
    if (a.loss_sums) { atomicAdd(a.loss_sums + 1, dloss_q); atomicAdd(a.loss_sums + 3, dloss_z); }
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
inline size_t fused_align(size_t x) { return (x + 255) / 256 * 256; }

inline bool fused_supported(const smct_shape& s, const smct_io& io) {
  return s.D == FD && s.dff >= FD && s.dff <= FF && s.dff % FD == 0 && s.H == 1 && s.full_model == 1 && s.noise_on == 1 &&
         s.weights_mode == SMCT_WEIGHTS_GAUSSIAN && s.P <= FPMAX && s.P >= 1 && s.S <= 4096 && s.F_y <= 16 &&
         io.attn == nullptr;
}

inline size_t fused_forward_ws(const smct_shape& s) {
  if (s.D != FD) return 0;
  const size_t Sa = (size_t)(s.S + 7) / 8 * 8;
  size_t n = 0;
  n += 6 * fused_align((size_t)s.B * s.S * FD * sizeof(float));
  n += 5 * fused_align((size_t)s.B * s.S * s.P * FD * sizeof(float));   // Kp, Vp, Rp + split copies of K, V
  n += 2 * fused_align((size_t)s.B * s.P * Sa * sizeof(unsigned short));
  n += fused_align((size_t)s.B * s.P * sizeof(unsigned short));
  n += fused_align((size_t)s.B * sizeof(int));
  n += fused_align((size_t)(1 + 2 * (FF / FD)) * 16384 + 256 + 2048);   // tensor-core weight blocks + header + noise-scale table
  return n + 1024;
}

}  // namespace smct

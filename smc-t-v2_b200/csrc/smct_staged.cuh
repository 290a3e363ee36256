// smct_staged.cuh — "staged" kernels: one launch per stage of a particle-filter timestep.
//   rows_kernel     : encode + LN1 + q_/k_/v_ projections of particle-independent rows
//   step_kernel     : noise, slot-t store, causal attention over each particle's K/V history,
//                     z projection, LN1/FFN/LN2, output layer, log-weight           (SMC_TransformerCell.py:129-162)
//   resample_kernel : logsumexp normalise + TF-CPU categorical (fp64 sequential CDF + upper_bound) (:164-166)
//   gather_kernel   : ancestor gather of state rows                                 (transformer_utils.py:39-47)
//   final_kernel    : pred_resampl, K - wk(X), V - wv(X), loss sums                 (SMC_Transformer.py:231-236)
// These are the per-stage entry points of the C-ABI and the data-movement-faithful forward
// (SMCT_ALGO_STAGED); the fused forward lives in smct_fused.cuh.
#pragma once
#include "smct_common.cuh"
#include "../../include/smct.h"

namespace smct {

constexpr int PT = 32;          // particles per CTA tile in step_kernel
constexpr int NT = 256;         // threads per CTA of the tile kernels (backward dense, step_kernel for d_model > 64)
constexpr int NW = NT / 32;
constexpr int NT_WIDE = 512;    // step_kernel / seq_forward_kernel for d_model <= 64: twice the warps for the per-particle
                                // phases (they are bound by the latency of one warp walking a particle's history)
constexpr int NW_MAX = NT_WIDE / 32;
constexpr int RPT = 4;          // rows per thread in the tile GEMM
constexpr int RG = PT / RPT;

// ------------------------------------------------------------------------------------------
// rows_kernel
// ------------------------------------------------------------------------------------------
struct RowsArgs {
  long long rows;
  int D, Fx, input_mode, do_ln;
  float sqrtD, ln_eps;
  const void* x_in;
  const float *W_in, *b_in, *ln_g, *ln_b, *Wq, *bq, *Wk, *bk, *Wv, *bv;
  float *X, *xln, *q_, *k_, *v_, *kx, *vx;  // rows x D each, optional
};

__global__ void __launch_bounds__(128) rows_kernel(const RowsArgs a) {
  extern __shared__ float sm[];
  const int D = a.D;
  float* xs = sm;          // raw encoded row
  float* xl = sm + D;      // layer-normed row
  float* scratch = sm + 2 * D;
  for (long long row = blockIdx.x; row < a.rows; row += gridDim.x) {
    __syncthreads();
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
      float x;
      if (a.input_mode == SMCT_INPUT_ENCODED) {
        x = ((const float*)a.x_in)[row * D + d];
      } else if (a.input_mode == SMCT_INPUT_EMBEDDING) {
        const int tok = clamp_index(((const int*)a.x_in)[row], a.Fx);
        x = a.W_in[(long long)tok * D + d] * a.sqrtD;
      } else {
        const float* xr = (const float*)a.x_in + row * a.Fx;
        float acc = 0.f;
        for (int f = 0; f < a.Fx; ++f) acc = fmaf(xr[f], a.W_in[f * D + d], acc);
        x = (acc + a.b_in[d]) * a.sqrtD;
      }
      xs[d] = x;
      if (a.X) a.X[row * D + d] = x;
    }
    __syncthreads();
    if (a.do_ln) {
      float s = 0.f;
      for (int d = threadIdx.x; d < D; d += blockDim.x) s += xs[d];
      const float mean = block_sum(s, scratch) / (float)D;
      float v = 0.f;
      for (int d = threadIdx.x; d < D; d += blockDim.x) { const float c = xs[d] - mean; v = fmaf(c, c, v); }
      const float var = block_sum(v, scratch) / (float)D;
      const float rstd = 1.0f / sqrtf(var + a.ln_eps);
      for (int d = threadIdx.x; d < D; d += blockDim.x) {
        const float inv = rstd * a.ln_g[d];
        xl[d] = xs[d] * inv + (a.ln_b[d] - mean * inv);
      }
    } else {
      for (int d = threadIdx.x; d < D; d += blockDim.x) xl[d] = xs[d];
    }
    __syncthreads();
    const bool proj = a.q_ || a.k_ || a.v_ || a.kx || a.vx;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
      if (a.xln) a.xln[row * D + d] = xl[d];
      if (!proj) continue;
      float q = a.bq[d], k = a.bk[d], v = a.bv[d], kx = a.bk[d], vx = a.bv[d];
      for (int j = 0; j < D; ++j) {
        const float xj = xl[j], rj = xs[j];
        const float wq = a.Wq[j * D + d], wk = a.Wk[j * D + d], wv = a.Wv[j * D + d];
        q = fmaf(xj, wq, q); k = fmaf(xj, wk, k); v = fmaf(xj, wv, v);
        kx = fmaf(rj, wk, kx); vx = fmaf(rj, wv, vx);
      }
      if (a.q_) a.q_[row * D + d] = q;
      if (a.k_) a.k_[row * D + d] = k;
      if (a.v_) a.v_[row * D + d] = v;
      if (a.kx) a.kx[row * D + d] = kx;
      if (a.vx) a.vx[row * D + d] = vx;
    }
  }
}

// ------------------------------------------------------------------------------------------
// step_kernel
// ------------------------------------------------------------------------------------------
struct StepArgs {
  int B, P, S, D, H, dff, Fy;
  int t, s_lo;
  int full_model, weights_mode, noise_z_mode, noise_on, logvar_dim, attn_only;
  float ln_eps, sigma_obs;
  const float *xln, *q_, *k_, *v_;   // addr = base + b*row_sb + p*row_sp
  long long row_sb, row_sp;
  const void* y; long long y_sb;     // y row of sequence b: base + b*y_sb (elements)
  const float *eps_k, *eps_q, *eps_v, *eps_z;  // (B,P,D) of this step, or null -> Philox
  unsigned long long seed; long long b_global0; int t_rng;
  float *K, *V, *R;                  // (B,P,S,D)
  smct_params p;
  float* pred; long long pred_st;    // out addr = base + (b*P+p)*stride
  float* r_out; long long r_st;
  float* z_out; long long z_st;
  float *noise_k, *noise_q, *noise_v, *noise_z; long long nz_st;
  float* attn; long long attn_st_bp, attn_st_h;
  float* logw;                       // (B,P)
  double* loss_sums;                 // [1] += sum noise_q^2/var_q ; [3] += sum noise_z^2/var_z
};

// C[row][c] = bias[c] + sum_k A[row][k] * W[k][c]  for the PT-row tile; epilogue(row, c, value).
// Every output accumulates bias + the products in the order k = 0..Kdim-1 (fmaf), whichever blocking is used, so the
// two paths below give identical bits.  Blocked path (N % 4 == 0, 16-byte aligned W): a thread owns 2 rows x 4 columns,
// one 128-bit weight load per k feeds 8 FMAs (the scalar path issues 5 loads per 4 FMAs and is LSU-bound).
template <typename Epi>
__device__ __forceinline__ void tile_gemm(const float* __restrict__ A, int lda, const float* __restrict__ W,
                                          const float* __restrict__ bias, int Kdim, int N, Epi epi) {
  if ((N & 3) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0 && (Kdim & 3) == 0 && (lda & 3) == 0 &&
      (reinterpret_cast<uintptr_t>(A) & 15) == 0) {
    // k-vectorised form of the blocked path (activation rows padded to a multiple of 4 floats: the backward dense tile):
    // two 128-bit shared loads + four 128-bit weight loads feed 32 FMAs (the form below: 12 loads).  Every output still
    // sums its products in k order: identical bits.
    constexpr int RH = PT / 2;
    const int N4 = N >> 2, K4 = Kdim >> 2;
    if (N4 * (PT / 4) >= (int)blockDim.x) {
      // wide outputs (every thread still has work with 4 rows x 4 columns): 8 loads per 64 FMAs
      constexpr int RQ = PT / 4;   // rows rg, rg + RQ, rg + 2 RQ, rg + 3 RQ
      for (int j = threadIdx.x; j < N4 * RQ; j += (int)blockDim.x) {
        const int c = 4 * (j % N4), rg = j / N4;
        float acc[4][4];
#pragma unroll
        for (int e = 0; e < 4; ++e) { const float b0 = bias ? bias[c + e] : 0.f; acc[0][e] = acc[1][e] = acc[2][e] = acc[3][e] = b0; }
        const float* Wc = W + c;
#pragma unroll 2
        for (int k4 = 0; k4 < K4; ++k4) {
          float av[4][4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 t = *reinterpret_cast<const float4*>(A + (rg + i * RQ) * lda + 4 * k4);
            av[i][0] = t.x; av[i][1] = t.y; av[i][2] = t.z; av[i][3] = t.w;
          }
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const float4 w = __ldg(reinterpret_cast<const float4*>(Wc + (long long)(4 * k4 + kk) * N));
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              acc[i][0] = fmaf(av[i][kk], w.x, acc[i][0]); acc[i][1] = fmaf(av[i][kk], w.y, acc[i][1]);
              acc[i][2] = fmaf(av[i][kk], w.z, acc[i][2]); acc[i][3] = fmaf(av[i][kk], w.w, acc[i][3]);
            }
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int e = 0; e < 4; ++e) epi(rg + i * RQ, c + e, acc[i][e]);
      }
      return;
    }
    for (int j = threadIdx.x; j < N4 * RH; j += (int)blockDim.x) {
      const int c = 4 * (j % N4), rg = j / N4;
      float acc[2][4];
#pragma unroll
      for (int e = 0; e < 4; ++e) { const float b0 = bias ? bias[c + e] : 0.f; acc[0][e] = b0; acc[1][e] = b0; }
      const float4* A0 = reinterpret_cast<const float4*>(A + rg * lda);
      const float4* A1 = reinterpret_cast<const float4*>(A + (rg + RH) * lda);
      const float* Wc = W + c;
#pragma unroll 2
      for (int k4 = 0; k4 < K4; ++k4) {
        const float4 a0 = A0[k4], a1 = A1[k4];
        const float av0[4] = {a0.x, a0.y, a0.z, a0.w}, av1[4] = {a1.x, a1.y, a1.z, a1.w};
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const float4 w = __ldg(reinterpret_cast<const float4*>(Wc + (long long)(4 * k4 + kk) * N));
          acc[0][0] = fmaf(av0[kk], w.x, acc[0][0]); acc[0][1] = fmaf(av0[kk], w.y, acc[0][1]);
          acc[0][2] = fmaf(av0[kk], w.z, acc[0][2]); acc[0][3] = fmaf(av0[kk], w.w, acc[0][3]);
          acc[1][0] = fmaf(av1[kk], w.x, acc[1][0]); acc[1][1] = fmaf(av1[kk], w.y, acc[1][1]);
          acc[1][2] = fmaf(av1[kk], w.z, acc[1][2]); acc[1][3] = fmaf(av1[kk], w.w, acc[1][3]);
        }
      }
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int e = 0; e < 4; ++e) epi(rg + i * RH, c + e, acc[i][e]);
    }
    return;
  }
  if ((N & 3) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0) {
    constexpr int RH = PT / 2;   // rows rg and rg + RH
    const int N4 = N >> 2;
    for (int j = threadIdx.x; j < N4 * RH; j += (int)blockDim.x) {
      const int c = 4 * (j % N4), rg = j / N4;
      float acc[2][4];
#pragma unroll
      for (int e = 0; e < 4; ++e) { const float b0 = bias ? bias[c + e] : 0.f; acc[0][e] = b0; acc[1][e] = b0; }
      const float* Ar = A + rg * lda;
#pragma unroll 4
      for (int k = 0; k < Kdim; ++k) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(W + (long long)k * N + c));
        const float a0 = Ar[k], a1 = Ar[RH * lda + k];
        acc[0][0] = fmaf(a0, w.x, acc[0][0]); acc[0][1] = fmaf(a0, w.y, acc[0][1]);
        acc[0][2] = fmaf(a0, w.z, acc[0][2]); acc[0][3] = fmaf(a0, w.w, acc[0][3]);
        acc[1][0] = fmaf(a1, w.x, acc[1][0]); acc[1][1] = fmaf(a1, w.y, acc[1][1]);
        acc[1][2] = fmaf(a1, w.z, acc[1][2]); acc[1][3] = fmaf(a1, w.w, acc[1][3]);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int e = 0; e < 4; ++e) epi(rg + i * RH, c + e, acc[i][e]);
    }
    return;
  }
  for (int j = threadIdx.x; j < N * RG; j += (int)blockDim.x) {
    const int c = j % N, rg = j / N;
    float acc[RPT];
    const float b0 = bias ? bias[c] : 0.f;
#pragma unroll
    for (int i = 0; i < RPT; ++i) acc[i] = b0;
    const float* Ar = A + rg * lda;
#pragma unroll 4
    for (int k = 0; k < Kdim; ++k) {
      const float w = __ldg(W + (long long)k * N + c);
#pragma unroll
      for (int i = 0; i < RPT; ++i) acc[i] = fmaf(Ar[i * RG * lda + k], w, acc[i]);
    }
#pragma unroll
    for (int i = 0; i < RPT; ++i) epi(rg + i * RG, c, acc[i]);
  }
}

// one timestep of one 32-particle tile (particles p0..p0+31 of sequence b); sm: step_smem_bytes() of shared memory
template <int KD>
__device__ __forceinline__ void step_tile(const StepArgs& a, const int b, const int p0, float* sm) {
  const int D = a.D, Dp = D + 1, dff = a.dff, Fp = dff + 1, S = a.S, t = a.t;
  const int nt = (int)blockDim.x, nw = nt >> 5;   // 256 or 512 threads
  float* Zs = sm;
  float* Os = Zs + PT * Dp;
  float* Hs = Os + PT * Dp;
  float* scr = Hs + (a.full_model ? PT * Fp : 0);
  float* red = scr + nw * 3 * D;  // 2*nw floats for loss partials
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int np = min(PT, a.P - p0);
  const int Dh = D / a.H;
  const int GD = (D + 3) >> 2;
  const float sqrt_dh = sqrtf((float)Dh);
  const bool multi = a.logvar_dim > 1;
  float loss_q = 0.f, loss_z = 0.f;

  // zero the tile so that rows >= np are defined for the GEMM phases
  for (int i = tid; i < 2 * PT * Dp; i += nt) sm[i] = 0.f;
  __syncthreads();

  // ---------------- phase A: noise, slot-t store, attention (one warp per particle) ----------------
  for (int pl = wid; pl < np; pl += nw) {
    const int p = p0 + pl;
    const long long bp = (long long)b * a.P + p;
    const long long row = (long long)b * a.row_sb + (long long)p * a.row_sp;
    const unsigned long long gbp = (unsigned long long)(a.b_global0 + b) * (unsigned long long)a.P + p;
    float* qs = scr + wid * 3 * D;
    float* ks = qs + D;
    float* vs = ks + D;
#pragma unroll
    for (int k = 0; k < KD; ++k) {
      const int d = lane + 32 * k;
      if (d < D) {
        const float q0 = a.q_[row + d], k0 = a.k_[row + d], v0 = a.v_[row + d];
        float q = q0, kk = k0, v = v0;
        if (a.noise_on) {
          const int li = multi ? d : 0;
          const float ek = a.eps_k ? a.eps_k[bp * D + d] : philox_normal1(a.seed, 0u, a.t_rng, gbp, GD, d);
          const float eq = a.eps_q ? a.eps_q[bp * D + d] : philox_normal1(a.seed, 1u, a.t_rng, gbp, GD, d);
          const float ev = a.eps_v ? a.eps_v[bp * D + d] : philox_normal1(a.seed, 2u, a.t_rng, gbp, GD, d);
          kk = __fadd_rn(k0, __fmul_rn(expf(0.5f * a.p.logvar_k[li]), ek));
          q = __fadd_rn(q0, __fmul_rn(expf(0.5f * a.p.logvar_q[li]), eq));
          v = __fadd_rn(v0, __fmul_rn(expf(0.5f * a.p.logvar_v[li]), ev));
          const float nq = q - q0;
          if (a.noise_k) a.noise_k[bp * a.nz_st + d] = kk - k0;
          if (a.noise_q) a.noise_q[bp * a.nz_st + d] = nq;
          if (a.noise_v) a.noise_v[bp * a.nz_st + d] = v - v0;
          loss_q = fmaf(nq * nq, expf(-a.p.logvar_q[li]), loss_q);
        }
        a.K[(bp * S + t) * D + d] = kk;
        a.V[(bp * S + t) * D + d] = v;
        qs[d] = q; ks[d] = kk; vs[d] = v;
      }
    }
    __syncwarp();
    for (int h = 0; h < a.H; ++h) {
      float qh[KD], acc[KD];
#pragma unroll
      for (int k = 0; k < KD; ++k) {
        const int dd = lane + 32 * k;
        qh[k] = (dd < Dh) ? qs[h * Dh + dd] : 0.f;
        acc[k] = 0.f;
      }
      float m = -INFINITY, l = 0.f;
      const float* Kp = a.K + (bp * S) * D + h * Dh;
      const float* Vp = a.V + (bp * S) * D + h * Dh;
      int s = a.s_lo;
      for (; s + 3 < t; s += 4) {
        float kr[4][KD], vr[4][KD], sc[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int k = 0; k < KD; ++k) {
            const int dd = lane + 32 * k;
            const bool ok = dd < Dh;
            kr[j][k] = ok ? Kp[(long long)(s + j) * D + dd] : 0.f;
            vr[j][k] = ok ? Vp[(long long)(s + j) * D + dd] : 0.f;
          }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float part = 0.f;
#pragma unroll
          for (int k = 0; k < KD; ++k) part = fmaf(qh[k], kr[j][k], part);
          sc[j] = part;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
          for (int j = 0; j < 4; ++j) sc[j] += __shfl_xor_sync(SMCT_FULL_MASK, sc[j], o);
#pragma unroll
        for (int j = 0; j < 4; ++j) sc[j] = sc[j] / sqrt_dh;
        const float mn = fmaxf(fmaxf(m, fmaxf(sc[0], sc[1])), fmaxf(sc[2], sc[3]));
        const float corr = expf(m - mn);
        float pj[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) pj[j] = expf(sc[j] - mn);
        l = l * corr + ((pj[0] + pj[1]) + (pj[2] + pj[3]));
#pragma unroll
        for (int k = 0; k < KD; ++k) {
          float v = acc[k] * corr;
#pragma unroll
          for (int j = 0; j < 4; ++j) v = fmaf(pj[j], vr[j][k], v);
          acc[k] = v;
        }
        m = mn;
      }
      for (; s <= t; ++s) {  // tail rows; the current slot t comes from registers/smem
        float part = 0.f, vr[KD];
#pragma unroll
        for (int k = 0; k < KD; ++k) {
          const int dd = lane + 32 * k;
          const bool ok = dd < Dh;
          float kv, vv;
          if (s == t) { kv = ok ? ks[h * Dh + dd] : 0.f; vv = ok ? vs[h * Dh + dd] : 0.f; }
          else { kv = ok ? Kp[(long long)s * D + dd] : 0.f; vv = ok ? Vp[(long long)s * D + dd] : 0.f; }
          part = fmaf(qh[k], kv, part);
          vr[k] = vv;
        }
        const float sc = warp_sum(part) / sqrt_dh;
        const float mn = fmaxf(m, sc);
        const float corr = expf(m - mn), pj = expf(sc - mn);
        l = l * corr + pj;
#pragma unroll
        for (int k = 0; k < KD; ++k) acc[k] = fmaf(pj, vr[k], acc[k] * corr);
        m = mn;
      }
#pragma unroll
      for (int k = 0; k < KD; ++k) {
        const int dd = lane + 32 * k;
        if (dd < Dh) Zs[pl * Dp + h * Dh + dd] = acc[k] / l;
      }
      if (a.attn) {  // second pass: normalised attention weights of the allowed slots
        float* ap = a.attn + bp * a.attn_st_bp + (long long)h * a.attn_st_h;
        for (int s2 = a.s_lo; s2 <= t; ++s2) {
          float part = 0.f;
#pragma unroll
          for (int k = 0; k < KD; ++k) {
            const int dd = lane + 32 * k;
            const bool ok = dd < Dh;
            const float kv = (s2 == t) ? (ok ? ks[h * Dh + dd] : 0.f) : (ok ? Kp[(long long)s2 * D + dd] : 0.f);
            part = fmaf(qh[k], kv, part);
          }
          const float sc = warp_sum(part) / sqrt_dh;
          if (lane == 0) ap[s2] = expf(sc - m) / l;
        }
      }
    }
    __syncwarp();
  }
  __syncthreads();

  // ---------------- phase B: z = dense(z_) + noise ----------------
  {
    auto epi = [&](int pl, int n, float zp) {
      if (pl >= np) return;
      const long long bp = (long long)b * a.P + p0 + pl;
      float z = zp;
      if (a.noise_on) {
        const int li = multi ? n : 0;
        const unsigned long long gbp = (unsigned long long)(a.b_global0 + b) * (unsigned long long)a.P + p0 + pl;
        const float ez = a.eps_z ? a.eps_z[bp * D + n] : philox_normal1(a.seed, 3u, a.t_rng, gbp, GD, n);
        z = __fadd_rn(zp, __fmul_rn(expf(0.5f * a.p.logvar_z[li]), ez));
        const float nz = (a.noise_z_mode == SMCT_NOISE_Z_CHECKOUT) ? (z - Zs[pl * Dp + n]) : (z - zp);
        if (a.noise_z) a.noise_z[bp * a.nz_st + n] = nz;
        loss_z = fmaf(nz * nz, expf(-a.p.logvar_z[li]), loss_z);
      }
      Os[pl * Dp + n] = z;
      if (a.z_out) a.z_out[bp * a.z_st + n] = z;
    };
    tile_gemm(Zs, Dp, a.p.Wz, a.p.bz, D, D, epi);
  }
  if (a.loss_sums && a.noise_on) {
    const float sq = warp_sum(loss_q), sz = warp_sum(loss_z);
    if (lane == 0) { red[wid] = sq; red[nw + wid] = sz; }
  }
  __syncthreads();
  if (a.loss_sums && a.noise_on && tid == 0) {
    double sq = 0.0, sz = 0.0;
    for (int i = 0; i < nw; ++i) { sq += (double)red[i]; sz += (double)red[nw + i]; }
    atomicAdd(a.loss_sums + 1, sq);
    atomicAdd(a.loss_sums + 3, sz);
  }
  if (a.attn_only) return;

  if (a.full_model) {
    // ---------------- phase C: o = LN1(z + x) (warp per row) ----------------
    for (int pl = wid; pl < np; pl += nw) {
      const long long row = (long long)b * a.row_sb + (long long)(p0 + pl) * a.row_sp;
      float v[KD], s1 = 0.f;
#pragma unroll
      for (int k = 0; k < KD; ++k) {
        const int d = lane + 32 * k;
        v[k] = (d < D) ? (Os[pl * Dp + d] + a.xln[row + d]) : 0.f;
        s1 += v[k];
      }
      const float mean = warp_sum(s1) / (float)D;
      float s2 = 0.f;
#pragma unroll
      for (int k = 0; k < KD; ++k) {
        const int d = lane + 32 * k;
        const float c = (d < D) ? (v[k] - mean) : 0.f;
        s2 = fmaf(c, c, s2);
      }
      const float rstd = 1.0f / sqrtf(warp_sum(s2) / (float)D + a.ln_eps);
#pragma unroll
      for (int k = 0; k < KD; ++k) {
        const int d = lane + 32 * k;
        if (d < D) {
          const float inv = rstd * a.p.ln1_g[d];
          Os[pl * Dp + d] = v[k] * inv + (a.p.ln1_b[d] - mean * inv);
        }
      }
    }
    __syncthreads();
    // ---------------- phase D: h = relu(o W1 + b1); r_pre = h W2 + b2 + o ----------------
    tile_gemm(Os, Dp, a.p.W1, a.p.b1, D, dff, [&](int pl, int j, float v) { Hs[pl * Fp + j] = fmaxf(v, 0.f); });
    __syncthreads();
    tile_gemm(Hs, Fp, a.p.W2, a.p.b2, dff, D, [&](int pl, int n, float v) { Zs[pl * Dp + n] = v + Os[pl * Dp + n]; });
    __syncthreads();
  }
  // ---------------- phase E: r = LN2(.), R store, output layer, log-weight (warp per row) ----------------
  for (int pl = wid; pl < np; pl += nw) {
    const int p = p0 + pl;
    const long long bp = (long long)b * a.P + p;
    float r[KD];
    if (a.full_model) {
      float s1 = 0.f;
#pragma unroll
      for (int k = 0; k < KD; ++k) {
        const int d = lane + 32 * k;
        r[k] = (d < D) ? Zs[pl * Dp + d] : 0.f;
        s1 += r[k];
      }
      const float mean = warp_sum(s1) / (float)D;
      float s2 = 0.f;
#pragma unroll
      for (int k = 0; k < KD; ++k) {
        const int d = lane + 32 * k;
        const float c = (d < D) ? (r[k] - mean) : 0.f;
        s2 = fmaf(c, c, s2);
      }
      const float rstd = 1.0f / sqrtf(warp_sum(s2) / (float)D + a.ln_eps);
#pragma unroll
      for (int k = 0; k < KD; ++k) {
        const int d = lane + 32 * k;
        if (d < D) {
          const float inv = rstd * a.p.ln2_g[d];
          r[k] = r[k] * inv + (a.p.ln2_b[d] - mean * inv);
        }
      }
    } else {
#pragma unroll
      for (int k = 0; k < KD; ++k) {
        const int d = lane + 32 * k;
        r[k] = (d < D) ? Os[pl * Dp + d] : 0.f;
      }
    }
#pragma unroll
    for (int k = 0; k < KD; ++k) {
      const int d = lane + 32 * k;
      if (d < D) {
        a.R[(bp * S + t) * D + d] = r[k];
        if (a.r_out) a.r_out[bp * a.r_st + d] = r[k];
      }
    }
    // output layer (no bias) + weight
    float sq = 0.f, mx = -INFINITY, se = 0.f, py = 0.f;
    int ytok = 0;
    const bool do_w = a.logw && a.y;
    if (do_w && a.weights_mode == SMCT_WEIGHTS_CLASSIFICATION) ytok = ((const int*)a.y)[(long long)b * a.y_sb];
    for (int f = 0; f < a.Fy; ++f) {
      float part = 0.f;
#pragma unroll
      for (int k = 0; k < KD; ++k) {
        const int d = lane + 32 * k;
        if (d < D) part = fmaf(r[k], __ldg(a.p.Wo + (long long)d * a.Fy + f), part);
      }
      const float pr = warp_sum(part);
      if (lane == 0 && a.pred) a.pred[bp * a.pred_st + f] = pr;
      if (!do_w) continue;
      if (a.weights_mode == SMCT_WEIGHTS_GAUSSIAN) {
        const float diff = ((const float*)a.y)[(long long)b * a.y_sb + f] - pr;
        sq = fmaf(diff, diff, sq);
      } else {
        const float mn = fmaxf(mx, pr);
        se = se * expf(mx - mn) + expf(pr - mn);
        mx = mn;
        if (f == ytok) py = pr;
      }
    }
    if (lane == 0 && do_w) {
      a.logw[bp] = (a.weights_mode == SMCT_WEIGHTS_GAUSSIAN) ? ((-0.5f * sq) / a.sigma_obs) : (expf(py - mx) / se);
    }
  }
}

template <int KD>
__global__ void __launch_bounds__(KD <= 2 ? NT_WIDE : NT) step_kernel(const StepArgs a) {
  extern __shared__ float sm[];
  step_tile<KD>(a, blockIdx.y, blockIdx.x * PT, sm);
}

// ------------------------------------------------------------------------------------------
// resample_kernel: one CTA per sequence
// ------------------------------------------------------------------------------------------
struct ResampleArgs {
  int B, P, normalise;
  const float* logw;        // (B,P)
  const double* u;          // (B,P) or null -> Philox
  const int* i_forced;      // (B,P) or null
  unsigned long long seed; long long b_global0; int t_rng;
  float* w_out; long long w_st;   // w_out[(b*P+p)*w_st]
  float* logw_copy;         // (B,P) optional copy of the input
  float* logits_out;        // (B,P) optional
  int* i_out;               // (B,P) optional (workspace index buffer)
  int* i_out2;              // (B,P) optional second copy (user trace)
};

// draw of sequence b: cdf = P doubles of shared memory, scratch = 32 floats, total_s = 1 double (all shared)
__device__ __forceinline__ void resample_seq(const ResampleArgs& a, const int b, double* cdf, float* scratch, double* total_sp) {
  const int P = a.P, tid = threadIdx.x;
  const float* lw = a.logw + (long long)b * P;
  // 1. max over finite logits
  float m = -INFINITY;
  for (int p = tid; p < P; p += blockDim.x) { const float v = lw[p]; if (isfinite(v)) m = fmaxf(m, v); }
  m = block_max(m, scratch);
  float shift = 0.f, m2 = m;
  if (a.normalise) {
    float s = 0.f;
    for (int p = tid; p < P; p += blockDim.x) s += expf(lw[p] - m);
    s = block_sum(s, scratch);
    shift = m + logf(s);  // logsumexp
    float mm = -INFINITY;
    for (int p = tid; p < P; p += blockDim.x) { const float v = lw[p] - shift; if (isfinite(v)) mm = fmaxf(mm, v); }
    m2 = block_max(mm, scratch);
  }
  // 2. logits, weights, fp64 exponentials
  for (int p = tid; p < P; p += blockDim.x) {
    const float raw = lw[p];
    const float lg = a.normalise ? (raw - shift) : raw;
    const long long bp = (long long)b * P + p;
    if (a.w_out) a.w_out[bp * a.w_st] = a.normalise ? expf(lg) : raw;
    if (a.logw_copy) a.logw_copy[bp] = raw;
    if (a.logits_out) a.logits_out[bp] = lg;
    cdf[p] = isfinite(lg) ? exp((double)lg - (double)m2) : 0.0;
  }
  __syncthreads();
  // 3. running total, strictly sequential in fp64 (the order of TF's CPU kernel)
  if (tid == 0) {
    double run = 0.0;
    for (int p = 0; p < P; ++p) { run += cdf[p]; cdf[p] = run; }
    *total_sp = run;
  }
  __syncthreads();
  const double total = *total_sp;
  // 4. one draw per particle slot: upper_bound(cdf, u * total)
  for (int p = tid; p < P; p += blockDim.x) {
    const long long bp = (long long)b * P + p;
    int idx;
    if (a.i_forced) {
      idx = clamp_index(a.i_forced[bp], P);
    } else {
      const double uu = a.u ? a.u[bp]
                            : philox_uniform(a.seed, (uint32_t)a.t_rng,
                                             (unsigned long long)(a.b_global0 + b) * (unsigned long long)P + p);
      const double target = uu * total;
      int lo = 0, hi = P;  // first index with cdf[idx] > target
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (cdf[mid] <= target) lo = mid + 1; else hi = mid;
      }
      idx = min(lo, P - 1);
    }
    if (a.i_out) a.i_out[bp] = idx;
    if (a.i_out2) a.i_out2[bp] = idx;
  }
}

__global__ void __launch_bounds__(256) resample_kernel(const ResampleArgs a) {
  extern __shared__ double cdf[];  // P doubles
  __shared__ float scratch[32];
  __shared__ double total_s;
  resample_seq(a, blockIdx.x, cdf, scratch, &total_s);
}

// ------------------------------------------------------------------------------------------
// gather_kernel: dst[b,m,:n] = src[b,idx[b,m],:n] for up to 3 tensors (blockIdx.z)
// ------------------------------------------------------------------------------------------
struct GatherArgs {
  int B, P;
  int ppc;                  // particle rows per CTA (short rows: several rows share a CTA instead of one tiny CTA each)
  long long row_elems, n_elems;
  const float* src[3];
  float* dst[3];
  const int* idx;
};

__global__ void __launch_bounds__(256) gather_kernel(const GatherArgs a) {
  const int z = blockIdx.z;
  const float* sbase = z == 0 ? a.src[0] : (z == 1 ? a.src[1] : a.src[2]);
  float* dbase = z == 0 ? a.dst[0] : (z == 1 ? a.dst[1] : a.dst[2]);
  const long long n = a.n_elems;
  const bool vec = ((a.row_elems & 3) == 0) && ((n & 3) == 0) && ((((uintptr_t)sbase) | ((uintptr_t)dbase)) & 15) == 0;
  const long long nbp = (long long)a.B * a.P;
  const long long bp0 = (long long)blockIdx.x * a.ppc;
  const int rows = (int)min((long long)a.ppc, nbp - bp0);
  const long long per = vec ? (n >> 2) : n;                       // copy units per row
  if (a.ppc == 1) {   // long rows: one row per CTA (and grid.y slices of it), no per-unit index arithmetic
    const int b = (int)(bp0 / a.P);
    const float* __restrict__ src = sbase + ((long long)b * a.P + clamp_index(a.idx[bp0], a.P)) * a.row_elems;
    float* __restrict__ dst = dbase + bp0 * a.row_elems;
    if (vec) {
      for (long long i = (long long)blockIdx.y * blockDim.x + threadIdx.x; i < per; i += (long long)gridDim.y * blockDim.x)
        reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(src)[i];
    } else {
      for (long long i = (long long)blockIdx.y * blockDim.x + threadIdx.x; i < per; i += (long long)gridDim.y * blockDim.x)
        dst[i] = src[i];
    }
    return;
  }
  const long long total = per * rows;
  for (long long i = (long long)blockIdx.y * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.y * blockDim.x) {
    const long long bp = bp0 + i / per, e = i % per;
    const int b = (int)(bp / a.P);
    const long long so = ((long long)b * a.P + clamp_index(a.idx[bp], a.P)) * a.row_elems, dofs = bp * a.row_elems;
    if (vec) reinterpret_cast<float4*>(dbase + dofs)[e] = reinterpret_cast<const float4*>(sbase + so)[e];
    else dbase[dofs + e] = sbase[so + e];
  }
}

// ------------------------------------------------------------------------------------------
// seq_forward_kernel (SMCT_ALGO_STAGED_SEQ): the S x (step, draw, gather) loop of the staged path in ONE launch,
// one CTA per sequence (sequences never interact; the particles of a sequence do).  Same device code as the staged
// kernels, hence bit-identical results; made for the reference's own small shapes (P <= 128), where the staged path
// is bound by ~5 launches per timestep.
// ------------------------------------------------------------------------------------------
struct SeqArgs {
  StepArgs step;                 // step-independent fields (fill_step_common) + loss_sums
  int attn_window, len_resampling, normalise, n_gather;
  int cl;                        // CTAs per sequence (thread-block cluster size; 1 = no cluster)
  const float *xln, *q_, *k_, *v_;   // (B,S,D)
  const void* y;                 // (B,S,Fy) f32 | (B,S) i32
  const float* eps;              // optional (4,S,B,P,D)
  const double* u;               // optional (S,B,P)
  const int* i_forced;           // optional (S,B,P)
  unsigned long long seed;
  float *K0, *V0, *R0, *K2, *V2, *R2;   // io state and its ping-pong twin
  float* logw_ws; int* idx_ws;   // (B,P) scratch
  float *pred, *noise_q, *noise_z, *attn, *w, *logw_out;
  int* i_out;
  // SMCT_ALGO_SMALL only (its epilogue replaces final_kernel): (B,S,D) rows wk(X), wv(X); output heads
  const float *kx, *vx;
  float *pred_resampl, *noise_K, *noise_V;
};

// A sequence is served by a thread-block CLUSTER of CL CTAs (CL = 1: the plain one-CTA-per-sequence form): CTA `rank`
// owns the particle tiles rank, rank + CL, ... of every timestep and the matching share of the ancestor gather; the draw
// (which needs all P log-weights) runs on rank 0.  The cluster barrier (release / acquire at cluster scope) orders the
// global-memory traffic between the phases, so the S x (step, draw, gather) loop of C2 / C3 (P = 100 / 60: 4 / 2 tiles)
// is ONE launch instead of 74 without serialising the tiles of a sequence on one SM.
__device__ __forceinline__ void seq_cluster_sync(int cl) {
  if (cl > 1) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  } else {
    __syncthreads();
  }
}

template <int KD>
__global__ void __launch_bounds__(KD <= 2 ? NT_WIDE : NT) seq_forward_kernel(const SeqArgs g) {
  extern __shared__ float sm[];
  __shared__ float scratch[32];
  __shared__ double total_s;
  const int CL = g.cl;
  const int b = blockIdx.x / CL, rank = blockIdx.x % CL, tid = threadIdx.x;
  const int B = g.step.B, P = g.step.P, S = g.step.S, D = g.step.D, H = g.step.H, Fy = g.step.Fy;
  const size_t BPD = (size_t)B * P * D, BP = (size_t)B * P;
  float *Kc, *Vc, *Rc, *Ka, *Va, *Ra;
  if (g.n_gather % 2 == 0) { Kc = g.K0; Vc = g.V0; Rc = g.R0; Ka = g.K2; Va = g.V2; Ra = g.R2; }
  else { Kc = g.K2; Vc = g.V2; Rc = g.R2; Ka = g.K0; Va = g.V0; Ra = g.R0; }
  for (int t = 0; t < S; ++t) {
    StepArgs a = g.step;
    a.t = t;
    a.s_lo = (g.attn_window >= 0 && t > g.attn_window) ? t - g.attn_window : 0;
    a.xln = g.xln + (size_t)t * D; a.q_ = g.q_ + (size_t)t * D; a.k_ = g.k_ + (size_t)t * D; a.v_ = g.v_ + (size_t)t * D;
    a.row_sb = (long long)S * D; a.row_sp = 0;
    if (a.weights_mode == SMCT_WEIGHTS_GAUSSIAN) { a.y = (const float*)g.y + (size_t)t * Fy; a.y_sb = (long long)S * Fy; }
    else { a.y = (const int*)g.y + t; a.y_sb = S; }
    if (g.eps) {
      a.eps_k = g.eps + ((size_t)0 * S + t) * BPD; a.eps_q = g.eps + ((size_t)1 * S + t) * BPD;
      a.eps_v = g.eps + ((size_t)2 * S + t) * BPD; a.eps_z = g.eps + ((size_t)3 * S + t) * BPD;
    }
    a.seed = g.seed; a.t_rng = t;
    a.K = Kc; a.V = Vc; a.R = Rc;
    a.pred = g.pred ? g.pred + (size_t)t * Fy : nullptr; a.pred_st = (long long)S * Fy;
    a.noise_q = g.noise_q ? g.noise_q + (size_t)t * D : nullptr;
    a.noise_z = g.noise_z ? g.noise_z + (size_t)t * D : nullptr;
    a.nz_st = (long long)S * D;
    a.attn = g.attn ? g.attn + (size_t)t * S : nullptr;
    a.attn_st_bp = (long long)H * S * S; a.attn_st_h = (long long)S * S;
    a.logw = g.logw_ws;
    for (int p0 = rank * PT; p0 < P; p0 += CL * PT) {
      step_tile<KD>(a, b, p0, sm);
      __syncthreads();
    }
    if (a.noise_on) {
      seq_cluster_sync(CL);   // the log-weights of all tiles are in global memory
      ResampleArgs r;
      r.B = B; r.P = P; r.normalise = g.normalise;
      r.logw = g.logw_ws;
      r.u = g.u ? g.u + (size_t)t * BP : nullptr;
      r.i_forced = g.i_forced ? g.i_forced + (size_t)t * BP : nullptr;
      r.seed = g.seed; r.b_global0 = a.b_global0; r.t_rng = t;
      r.w_out = g.w ? g.w + t : nullptr; r.w_st = S;
      r.logw_copy = g.logw_out ? g.logw_out + (size_t)t * BP : nullptr;
      r.logits_out = nullptr;
      r.i_out = g.idx_ws;
      r.i_out2 = g.i_out ? g.i_out + (size_t)t * BP : nullptr;
      if (rank == 0) resample_seq(r, b, reinterpret_cast<double*>(sm), scratch, &total_s);
      seq_cluster_sync(CL);   // the ancestor indices are in global memory
      if (g.len_resampling < 0 || t < g.len_resampling) {
        // ancestor gather of this sequence: rows 0..t of K, V, R (transformer_utils.py:39-47)
        const long long rowlen = (long long)(t + 1) * D, stride = (long long)S * D;
        const int m_per = (P + CL - 1) / CL, m_lo = rank * m_per, m_hi = min(P, m_lo + m_per);   // this CTA's destination particles
        const long long n = (long long)max(0, m_hi - m_lo) * rowlen;
        const int* idx = g.idx_ws + (size_t)b * P;
        for (int z = 0; z < 3; ++z) {
          const float* src = (z == 0 ? Kc : (z == 1 ? Vc : Rc)) + (size_t)b * P * stride;
          float* dst = (z == 0 ? Ka : (z == 1 ? Va : Ra)) + (size_t)b * P * stride;
          for (long long i = tid; i < n; i += (long long)blockDim.x) {
            const int m = m_lo + (int)(i / rowlen);
            const long long off = i - (long long)(m - m_lo) * rowlen;
            dst[(long long)m * stride + off] = src[(long long)idx[m] * stride + off];
          }
        }
        float* tp;
        tp = Kc; Kc = Ka; Ka = tp; tp = Vc; Vc = Va; Va = tp; tp = Rc; Rc = Ra; Ra = tp;
        seq_cluster_sync(CL);   // the gathered rows are visible to the CTAs that own them in the next step
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// final_kernel: per (b,p,s) row: pred_resampl = R Wo ; noise_K = K - kx ; noise_V = V - vx ; loss sums
// one warp per row
// ------------------------------------------------------------------------------------------
struct FinalArgs {
  int B, P, S, D, Fy, weights_mode, logvar_dim;
  const float *K, *V, *R, *kx, *vx, *Wo, *logvar_k, *logvar_v;
  const void* y;  // (B,S,Fy) f32 for the gaussian classic loss (optional)
  float *pred_resampl, *noise_K, *noise_V;
  double* loss_sums;  // [0] K, [2] V, [4] sum (y - pred_resampl)^2
};

__global__ void __launch_bounds__(256) final_kernel(const FinalArgs a) {
  __shared__ double red[3][8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const long long nrows = (long long)a.B * a.P * a.S;
  const int D = a.D;
  const bool multi = a.logvar_dim > 1;
  float lk = 0.f, lv = 0.f, ly = 0.f;
  for (long long row = (long long)blockIdx.x * 8 + wid; row < nrows; row += (long long)gridDim.x * 8) {
    const int s = (int)(row % a.S);
    const int b = (int)(row / ((long long)a.P * a.S));
    const long long xs = ((long long)b * a.S + s) * D;
    for (int f = 0; f < a.Fy; ++f) {
      float part = 0.f;
      for (int d = lane; d < D; d += 32) part = fmaf(a.R[row * D + d], __ldg(a.Wo + (long long)d * a.Fy + f), part);
      const float pr = warp_sum(part);
      if (lane == 0) {
        if (a.pred_resampl) a.pred_resampl[row * a.Fy + f] = pr;
        if (a.y && a.weights_mode == SMCT_WEIGHTS_GAUSSIAN) {
          const float diff = ((const float*)a.y)[((long long)b * a.S + s) * a.Fy + f] - pr;
          ly = fmaf(diff, diff, ly);
        }
      }
    }
    if (a.kx) {
      for (int d = lane; d < D; d += 32) {
        const float nk = a.K[row * D + d] - a.kx[xs + d];
        const float nv = a.V[row * D + d] - a.vx[xs + d];
        if (a.noise_K) a.noise_K[row * D + d] = nk;
        if (a.noise_V) a.noise_V[row * D + d] = nv;
        lk = fmaf(nk * nk, expf(-a.logvar_k[multi ? d : 0]), lk);
        lv = fmaf(nv * nv, expf(-a.logvar_v[multi ? d : 0]), lv);
      }
    }
  }
  if (a.loss_sums) {
    lk = warp_sum(lk); lv = warp_sum(lv); ly = warp_sum(ly);
    if (lane == 0) { red[0][wid] = lk; red[1][wid] = lv; red[2][wid] = ly; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double s0 = 0, s1 = 0, s2 = 0;
      for (int i = 0; i < 8; ++i) { s0 += red[0][i]; s1 += red[1][i]; s2 += red[2][i]; }
      atomicAdd(a.loss_sums + 0, s0);
      atomicAdd(a.loss_sums + 2, s1);
      atomicAdd(a.loss_sums + 4, s2);
    }
  }
}

// ------------------------------------------------------------------------------------------
// metric_kernel: the training metric of train_step_SMC_T on the UNRESAMPLED predictions (B,P,S,Fy), one warp per (b,s):
//   classification (train/utils.py:3-20): -log( mean_p softmax(pred[b,p,s,:])[y[b,s]] ), floored at 1e-7 like Keras'
//   SparseCategoricalCrossentropy on probabilities;  regression (pyc): sum_f (mean_p pred[b,p,s,f] - y[b,s,f])^2.
// out[0] += sum_{b,s} m[b,s] * metric, out[1] += sum_{b,s} m[b,s]  (mask (B,S) optional).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) metric_kernel(int B, int P, int S, int Fy, int mode, const float* pred, const void* y,
                                                     const float* mask, double* out) {
  __shared__ double red[2][8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  double acc = 0.0, cnt = 0.0;
  for (long long bs = (long long)blockIdx.x * 8 + wid; bs < (long long)B * S; bs += (long long)gridDim.x * 8) {
    const int b = (int)(bs / S), s = (int)(bs % S);
    const float mk = mask ? mask[bs] : 1.f;
    float val = 0.f;
    if (mode == SMCT_WEIGHTS_CLASSIFICATION) {
      const int tgt = clamp_index(((const int*)y)[bs], Fy);
      float psum = 0.f;                                   // sum over particles of softmax(...)[tgt]
      for (int p = 0; p < P; ++p) {
        const float* pr = pred + (((long long)b * P + p) * S + s) * Fy;
        float mx = -INFINITY;
        for (int f = lane; f < Fy; f += 32) mx = fmaxf(mx, pr[f]);
        mx = warp_max(mx);
        float se = 0.f;
        for (int f = lane; f < Fy; f += 32) se += expf(pr[f] - mx);
        se = warp_sum(se);
        psum += expf(pr[tgt] - mx) / se;
      }
      val = -logf(fmaxf(psum / (float)P, 1e-7f));
    } else {
      for (int f = 0; f < Fy; ++f) {
        float part = 0.f;
        for (int p = lane; p < P; p += 32) part += pred[(((long long)b * P + p) * S + s) * Fy + f];
        const float d = warp_sum(part) / (float)P - ((const float*)y)[bs * Fy + f];
        val = fmaf(d, d, val);
      }
    }
    acc += (double)val * (double)mk;
    cnt += (double)mk;
  }
  if (lane == 0) { red[0][wid] = acc; red[1][wid] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a0 = 0, a1 = 0;
    for (int i = 0; i < 8; ++i) { a0 += red[0][i]; a1 += red[1][i]; }
    atomicAdd(out, a0);
    atomicAdd(out + 1, a1);
  }
}

// ------------------------------------------------------------------------------------------
// EM variance update of src/train/utils.py:22-51 on the device.  em_err_kernel: err[w][j] = mean over (p,s,d) of
// noise_w[j]^2 for the four recorded noises w = k,q,v,z (order of the reference's updates: v,k,q,z — the four
// recurrences are independent) and every batch element j; em_apply_kernel: for j = 0..B-1 in order
//   logvar <- log((1 - it^-a) exp(logvar) + it^-a err[j])        (EM_update, :48-51)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) em_err_kernel(int B, long long per_b, const float* nk, const float* nq, const float* nv,
                                                     const float* nz, double* err) {
  __shared__ double red[8];
  const int w = blockIdx.x / B, j = blockIdx.x % B;
  const float* n = (w == 0 ? nk : (w == 1 ? nq : (w == 2 ? nv : nz))) + (long long)j * per_b;
  double acc = 0.0;
  for (long long i = threadIdx.x; i < per_b; i += blockDim.x) { const float v = n[i]; acc += (double)v * (double)v; }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(SMCT_FULL_MASK, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int i = 0; i < 8; ++i) t += red[i];
    err[(long long)w * B + j] = t / (double)per_b;
  }
}
__global__ void em_apply_kernel(int B, const double* err, float* lk, float* lq, float* lv, float* lz, float it, float em_param) {
  const int w = threadIdx.x;
  if (w >= 4) return;
  float* lp = w == 0 ? lk : (w == 1 ? lq : (w == 2 ? lv : lz));
  const float g = powf(it, -em_param);
  float cur = lp[0];
  for (int j = 0; j < B; ++j) cur = logf((1.0f - g) * expf(cur) + g * (float)err[(long long)w * B + j]);   // fp32 like the reference
  lp[0] = cur;
}

// ------------------------------------------------------------------------------------------
// small utility kernels
// ------------------------------------------------------------------------------------------
__global__ void fill_noise_kernel(unsigned long long seed, int which, int t, long long b0, int B, int P, int D,
                                  float* out) {
  const int GD = (D + 3) >> 2;
  const long long n = (long long)B * P * GD;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long bp = i / GD;
    const int g = (int)(i % GD);
    const unsigned long long gbp = (unsigned long long)(b0 * P) + (unsigned long long)bp;
    float nn[4];
    philox_normal4(seed, (uint32_t)which, (uint32_t)t, gbp * (unsigned long long)GD + g, nn);
    for (int e = 0; e < 4; ++e) {
      const int d = g * 4 + e;
      if (d < D) out[bp * D + d] = nn[e];
    }
  }
}

__global__ void fill_uniform_kernel(unsigned long long seed, int t, long long b0, int B, int P, double* out) {
  const long long n = (long long)B * P;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = philox_uniform(seed, (uint32_t)t, (unsigned long long)(b0 * P) + (unsigned long long)i);
}

__global__ void logweights_kernel(int B, int P, int Fy, int mode, float sigma_obs, const float* pred, const void* y,
                                  float* logw) {
  const long long n = (long long)B * P;
  for (long long bp = (long long)blockIdx.x * blockDim.x + threadIdx.x; bp < n; bp += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(bp / P);
    const float* pr = pred + bp * Fy;
    if (mode == SMCT_WEIGHTS_GAUSSIAN) {
      float sq = 0.f;
      for (int f = 0; f < Fy; ++f) { const float d = ((const float*)y)[(long long)b * Fy + f] - pr[f]; sq = fmaf(d, d, sq); }
      logw[bp] = (-0.5f * sq) / sigma_obs;
    } else {
      float mx = -INFINITY;
      for (int f = 0; f < Fy; ++f) mx = fmaxf(mx, pr[f]);
      float se = 0.f;
      for (int f = 0; f < Fy; ++f) se += expf(pr[f] - mx);
      logw[bp] = expf(pr[clamp_index(((const int*)y)[b], Fy)] - mx) / se;
    }
  }
}

// sum over (b,p,s) rows of ||y[b,s]-pred[b,p,s]||^2 (gaussian) or of the sparse cross-entropy (classification)
__global__ void __launch_bounds__(256) classic_loss_kernel(int B, int P, int S, int Fy, int mode, const float* pred,
                                                           const void* y, const float* mask, double* out) {
  __shared__ double red[8];
  const long long rows = (long long)B * P * S;
  double acc = 0.0;
  for (long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x; row < rows; row += (long long)gridDim.x * blockDim.x) {
    const int s = (int)(row % S);
    const int b = (int)(row / ((long long)P * S));
    const float* pr = pred + row * Fy;
    if (mode == SMCT_WEIGHTS_GAUSSIAN) {
      const float* yr = (const float*)y + ((long long)b * S + s) * Fy;
      float sq = 0.f;
      for (int f = 0; f < Fy; ++f) { const float d = yr[f] - pr[f]; sq = fmaf(d, d, sq); }
      acc += (double)sq * (mask ? (double)mask[(long long)b * S + s] : 1.0);
    } else {
      float mx = -INFINITY;
      for (int f = 0; f < Fy; ++f) mx = fmaxf(mx, pr[f]);
      float se = 0.f;
      for (int f = 0; f < Fy; ++f) se += expf(pr[f] - mx);
      acc += (double)(mx + logf(se) - pr[clamp_index(((const int*)y)[(long long)b * S + s], Fy)]) *
             (mask ? (double)mask[(long long)b * S + s] : 1.0);
    }
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(SMCT_FULL_MASK, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int i = 0; i < 8; ++i) t += red[i];
    atomicAdd(out, t);
  }
}

__global__ void __launch_bounds__(256) sumsq_scaled_kernel(long long rows, int D, const float* n, const float* logvar,
                                                           int logvar_dim, double* out) {
  __shared__ double red[8];
  const long long total = rows * D;
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const float v = n[i];
    acc += (double)(v * v * expf(-logvar[logvar_dim > 1 ? (int)(i % D) : 0]));
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(SMCT_FULL_MASK, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0;
    for (int i = 0; i < 8; ++i) s += red[i];
    atomicAdd(out, s);
  }
}

}  // namespace smct

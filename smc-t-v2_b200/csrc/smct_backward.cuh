// smct_backward.cuh — backward pass of the SMC-Transformer forward.
//
// tape.gradient(loss, trainable_variables) of train_step_SMC_T (src/train/train_step_functions.py:43-63) for
// the regression loss: loss = mean_{b,p,s} [ 1/2 sum_n ||noise_n||^2/sigma_n^2 + 1/2 ||y - pred_resampl||^2/Sigma_obs ].
//
// Path-wise formulation.  Weights and ancestor indices are stop_gradient (SMC_TransformerCell.py:167), the
// noise scale is stop_gradient (self_attention_SMC.py:98-104), and the loss only sees the FINAL resampled
// states.  The final state of particle p, K[b,p,0..S-1], V[..], R[..], is exactly the lineage of p, so the
// loss is a sum over B*P independent "paths": each path is one causal-attention + LN/FFN/LN + linear-head
// sequence of length S whose inputs (x rows, noise constants) are known.  The backward is therefore driven
// entirely by the forward's OUTPUTS (K, V final; ancestor indices; the RNG seed or injected eps) and needs no
// genealogy bookkeeping.
// Node de-duplication.  Paths that share the ancestor at slot s share the whole prefix 0..s, hence the row
// (q_s, z_s, r_s, pred_s) and every derivative of it: the loss is a sum over the NODES (s, ancestor) of the final
// genealogy tree, each weighted by the number m of final particles that descend from it.  node_*_kernel build
// the compact node list (representative final particle, logical ancestor id, multiplicity); all row-wise
// kernels below run once per node with the loss gradient scaled by m.  With resampling at every step the tree
// has a few thousand nodes per sequence instead of P*S rows; without resampling every row is its own node.
//   noise constants:  k_s - k_(x_s) and v_s - v_(x_s) are recovered from the stored rows;
//                     eps_q / eps_z of row (path p, slot s) belong to the logical particle anc[b,p,s] that p
//                     descends from at time s (traced from i_out) and are regenerated (Philox) or read (eps).
// Kernels (correctness-first, generic d_model <= 128, 1 head, full model):
//   anc_trace_kernel        lineage of every final particle in logical ids, (B,S,P)
//   node_count/scan/fill    unique (slot, ancestor) nodes per (b,s): representative particle, multiplicity
//   bwd_attn_fwd_kernel     recompute q, attention output z_ and the softmax statistics per row
//   bwd_dense_kernel        per 32-particle tile at one (b,s): recompute dense forward, back-propagate to dz_,
//                           weight gradients into CTA-private replicas (no float atomics on weights)
//   bwd_attn_bwd_kernel     per group of paths: dq, dk, dv (+ the direct K/V noise terms), reduced over particles
//   bwd_rows_kernel         shared rows (b,s): Q/K/V projections, LN1(x), input encoding
//   reduce_replicas_kernel  sum of the replicas into the gradient buffers
#pragma once
#include "smct_common.cuh"
#include "smct_staged.cuh"
#include "../../include/smct.h"

namespace smct {

// offsets (in floats) of every parameter inside one gradient replica
struct GradLayout {
  int Wq, bq, Wk, bk, Wv, bv, Wz, bz, ln1_g, ln1_b, ln2_g, ln2_b, W1, b1, W2, b2, Wo, W_in, b_in, total;
};

struct BwdArgs {
  int B, P, S, D, dff, Fy, Fx, attn_window, len_resampling, input_mode, weights_mode;
  int logvar_dim;                         // 1: scalar log-variances; D: per dimension (noise_dim = "multi")
  float ln_eps, sigma_obs, inv_n, sqrtD;
  long long b_global0;
  unsigned long long seed;
  const void* x_in;
  const float *X, *xln, *q_, *kx, *vx;   // (B,S,D) rows (rows_kernel)
  const void* y;                          // (B,S,Fy) f32 (gaussian) | (B,S) i32 class ids (classification)
  const float* eps;                       // optional (4,S,B,P,D)
  const float* classic_w;                 // optional (B,S): weight of the classic-loss term of row (b,s) (masked losses)
  smct_params p;
  const float *WzT, *W1T, *W2T;           // transposed weights
  const float *WqT, *WkT, *WvT;           // (out,in): thread d of bwd_rows walks column d — coalesced
  const float *K, *V;                     // (B,P,S,D) final states
  const int* i_out;                       // (S,B,P)
  unsigned short* anc;                    // (B,S,P) logical ancestor of final particle p at slot s
  // compact node list (nodes of (b,s) occupy [node_off[b*S+s], node_off[b*S+s+1]), ordered by ancestor id)
  const int* node_off;                    // (B*S+1)
  const int* node_bs;                     // (N) b*S+s
  const unsigned short *node_p, *node_L;  // (N) representative final particle, logical ancestor id
  const float* node_m;                    // (N) multiplicity
  unsigned short* node_of;                // (B,P,S) index inside its (b,s) segment of the node whose representative is p (0xFFFF: none)
  float *Zm, *dZm, *stat;                 // (N,D), (N,D), (N,2) per node
  float* Qm;                              // (N,D) q row of the node (k_, noise regenerated once by bwd_attn_fwd)
  float *dq_sum, *dk_sum, *dv_sum, *dxln_sum, *dkx_sum, *dvx_sum;   // (B,S,D)
  float* rep;                             // nrep x layout.total gradient replicas
  int nrep;
  GradLayout lay;
};

__global__ void anc_trace_kernel(int B, int P, int S, int len_resampling, const int* i_out, unsigned short* anc) {
  const long long n = (long long)B * P;
  for (long long bp = (long long)blockIdx.x * blockDim.x + threadIdx.x; bp < n; bp += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(bp / P);
    int l = (int)(bp % P);
    for (int t = S - 1; t >= 0; --t) {
      if (len_resampling < 0 || t < len_resampling) l = i_out[((long long)t * B + b) * P + l];
      anc[((long long)b * S + t) * P + (bp % P)] = (unsigned short)l;
    }
  }
}


// Histogram of the ancestors at slot s over the final particles of sequence b (one CTA per (b,s)).
// cnt[L] = multiplicity, rep[L] = smallest final particle descending from L.  Returns the number of nodes.
__device__ __forceinline__ int node_histogram(int P, const unsigned short* anc_bs, int* cnt, int* rep, int* scratch) {
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int i = tid; i < P; i += nt) { cnt[i] = 0; rep[i] = 0x7fffffff; }
  __syncthreads();
  for (int p = tid; p < P; p += nt) {
    const int L = anc_bs[p];
    atomicAdd(&cnt[L], 1);
    atomicMin(&rep[L], p);
  }
  __syncthreads();
  int local = 0;
  for (int i = tid; i < P; i += nt) local += (cnt[i] > 0);
  local = __reduce_add_sync(SMCT_FULL_MASK, local);
  if ((tid & 31) == 0) scratch[tid >> 5] = local;
  __syncthreads();
  int total = 0;
  for (int w = 0; w < (nt >> 5); ++w) total += scratch[w];
  return total;
}

__global__ void __launch_bounds__(256) node_count_kernel(int P, const unsigned short* anc, int* node_cnt) {
  extern __shared__ int nsm[];
  __shared__ int scratch[8];
  const int total = node_histogram(P, anc + (long long)blockIdx.x * P, nsm, nsm + P, scratch);
  if (threadIdx.x == 0) node_cnt[blockIdx.x] = total;
}

// exclusive scan of n counts (one CTA of 1024 threads); off[n] = total
__global__ void __launch_bounds__(1024) node_scan_kernel(int n, const int* cnt, int* off) {
  __shared__ int part[1024];
  const int tid = threadIdx.x;
  const int per = (n + 1023) / 1024;
  const int i0 = min(n, tid * per), i1 = min(n, i0 + per);
  int sum = 0;
  for (int i = i0; i < i1; ++i) sum += cnt[i];
  part[tid] = sum;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    const int v = (tid >= o) ? part[tid - o] : 0;
    __syncthreads();
    part[tid] += v;
    __syncthreads();
  }
  int run = part[tid] - sum;
  for (int i = i0; i < i1; ++i) { off[i] = run; run += cnt[i]; }
  if (tid == 1023) off[n] = part[1023];
}

// node_of (optional, pre-filled with 0xFFFF): node_of[b][rep][s] = index of the node inside its (b,s) segment.  The
// representative (smallest final particle below the node) can only decrease towards the root, so the nodes represented
// by one particle are a chain of consecutive slots [s_top, S-1] on that particle's lineage.
__global__ void __launch_bounds__(256) node_fill_kernel(int P, int S, const unsigned short* anc, const int* node_off,
                                                        int* node_bs, unsigned short* node_p, unsigned short* node_L,
                                                        float* node_m, unsigned short* node_of) {
  extern __shared__ int nsm[];
  __shared__ int scratch[8];
  __shared__ int wbase[9];
  int* cnt = nsm;
  int* rep = nsm + P;
  node_histogram(P, anc + (long long)blockIdx.x * P, cnt, rep, scratch);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int seg0 = node_off[blockIdx.x];
  const int b_ = blockIdx.x / S, s_ = blockIdx.x % S;
  int base = seg0;
  for (int i0 = 0; i0 < P; i0 += 256) {      // ordered compaction, 256 ancestors per pass
    const int i = i0 + tid;
    const bool live = (i < P) && (cnt[i] > 0);
    const unsigned bal = __ballot_sync(SMCT_FULL_MASK, live);
    __syncthreads();
    if (lane == 0) wbase[wid + 1] = __popc(bal);
    __syncthreads();
    if (tid == 0) { wbase[0] = 0; for (int w = 0; w < 8; ++w) wbase[w + 1] += wbase[w]; }
    __syncthreads();
    if (live) {
      const int u = base + wbase[wid] + __popc(bal & ((1u << lane) - 1u));
      node_bs[u] = blockIdx.x;
      node_p[u] = (unsigned short)rep[i];
      node_L[u] = (unsigned short)i;
      node_m[u] = (float)cnt[i];
      if (node_of) node_of[((long long)b_ * P + rep[i]) * S + s_] = (unsigned short)(u - seg0);
    }
    base += wbase[8];
  }
}

__global__ void transpose_kernel(const float* in, float* out, int rows, int cols) {   // out[c][r] = in[r][c]
  const long long n = (long long)rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i % cols);
    out[(long long)c * rows + r] = in[i];
  }
}

// 2^x on the special-function unit (the half-warp attention kernels keep scores and softmax maxima in the log2 domain,
// like the fused forward: one MUFU instead of the ~8-instruction expf)
__device__ __forceinline__ float bwd_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ float half_sum(float v) {   // sum over the 16 lanes of a half-warp (all 32 lanes participate)
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(SMCT_FULL_MASK, v, o);
  return v;
}

__device__ __forceinline__ float bwd_eps(const BwdArgs& a, int which, int b, int s, int L, int d) {
  if (a.eps) return a.eps[((((long long)which * a.S + s) * a.B + b) * a.P + L) * a.D + d];
  const unsigned long long gbp = (unsigned long long)(a.b_global0 + b) * (unsigned long long)a.P + L;
  return philox_normal1(a.seed, (uint32_t)which, (uint32_t)s, gbp, (a.D + 3) >> 2, d);
}

// ------------------------------------------------------------------------------------------
// attention forward recompute: one warp per node (b, s, ancestor)
// ------------------------------------------------------------------------------------------
template <int KD>
__global__ void __launch_bounds__(256) bwd_attn_fwd_kernel(const BwdArgs a) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int D = a.D, S = a.S;
  const long long nnodes = a.node_off[a.B * S];
  const bool multi = a.logvar_dim > 1;
  float std_q[KD];
#pragma unroll
  for (int k = 0; k < KD; ++k) { const int d = lane + 32 * k; std_q[k] = expf(0.5f * a.p.logvar_q[(multi && d < D) ? d : 0]); }
  const float sqrt_d = sqrtf((float)D);
  for (long long u = (long long)blockIdx.x * 8 + wid; u < nnodes; u += (long long)gridDim.x * 8) {
    const int bs = a.node_bs[u];
    const int s = bs % S, b = bs / S;
    const long long bp = (long long)b * a.P + a.node_p[u];
    const int L = a.node_L[u];
    const int s_lo = (a.attn_window >= 0 && s > a.attn_window) ? s - a.attn_window : 0;
    float q[KD], acc[KD];
#pragma unroll
    for (int k = 0; k < KD; ++k) {
      const int d = lane + 32 * k;
      q[k] = (d < D) ? __fadd_rn(a.q_[(long long)bs * D + d], __fmul_rn(std_q[k], bwd_eps(a, 1, b, s, L, d))) : 0.f;
      acc[k] = 0.f;
    }
    const float* Kp = a.K + bp * S * D;
    const float* Vp = a.V + bp * S * D;
    float m = -INFINITY, l = 0.f;
    for (int t = s_lo; t <= s; ++t) {
      float part = 0.f, vr[KD];
#pragma unroll
      for (int k = 0; k < KD; ++k) {
        const int d = lane + 32 * k;
        const float kv = (d < D) ? Kp[(long long)t * D + d] : 0.f;
        vr[k] = (d < D) ? Vp[(long long)t * D + d] : 0.f;
        part = fmaf(q[k], kv, part);
      }
      const float sc = warp_sum(part) / sqrt_d;
      const float mn = fmaxf(m, sc);
      const float corr = expf(m - mn), pj = expf(sc - mn);
      l = l * corr + pj;
#pragma unroll
      for (int k = 0; k < KD; ++k) acc[k] = fmaf(pj, vr[k], acc[k] * corr);
      m = mn;
    }
#pragma unroll
    for (int k = 0; k < KD; ++k) {
      const int d = lane + 32 * k;
      if (d < D) { a.Zm[u * D + d] = acc[k] / l; a.Qm[u * D + d] = q[k]; }
    }
    if (lane == 0) { a.stat[u * 2] = m; a.stat[u * 2 + 1] = l; }
  }
}

// ------------------------------------------------------------------------------------------
// attention forward recompute, half-warp form (d_model a multiple of 4, <= 64): one warp per node; a row is one float4
// per lane of a half-warp, the two halves take alternate slots with their own running (max, sum, z) and merge at the end
// (half the dependent online-softmax steps per node, 4-step reductions, 128-bit loads).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bwd_attn_fwd_half_kernel(const BwdArgs a) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int l16 = lane & 15, hsel = lane >> 4;
  const int D = a.D, S = a.S;
  const long long nnodes = a.node_off[a.B * S];
  const bool multi = a.logvar_dim > 1;
  const int d0 = 4 * l16;
  const bool okd = d0 < D;
  float std_q[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) std_q[e] = expf(0.5f * a.p.logvar_q[(multi && okd) ? d0 + e : 0]);
  const float qk_scale = 1.4426950408889634f / sqrtf((float)D);   // scores in the log2 domain (stat[] holds the log2-domain maximum)
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long u = (long long)blockIdx.x * 8 + wid; u < nnodes; u += (long long)gridDim.x * 8) {
    const int bs = a.node_bs[u];
    const int s = bs % S, b = bs / S;
    const long long bp = (long long)b * a.P + a.node_p[u];
    const int L = a.node_L[u];
    const int s_lo = (a.attn_window >= 0 && s > a.attn_window) ? s - a.attn_window : 0;
    float4 q = zero4;
    if (okd) {
      const float4 qb = *reinterpret_cast<const float4*>(a.q_ + (long long)bs * D + d0);
      q.x = __fadd_rn(qb.x, __fmul_rn(std_q[0], bwd_eps(a, 1, b, s, L, d0)));
      q.y = __fadd_rn(qb.y, __fmul_rn(std_q[1], bwd_eps(a, 1, b, s, L, d0 + 1)));
      q.z = __fadd_rn(qb.z, __fmul_rn(std_q[2], bwd_eps(a, 1, b, s, L, d0 + 2)));
      q.w = __fadd_rn(qb.w, __fmul_rn(std_q[3], bwd_eps(a, 1, b, s, L, d0 + 3)));
    }
    const float* Kp = a.K + bp * S * D;
    const float* Vp = a.V + bp * S * D;
    float m = -INFINITY, l = 0.f;
    float4 acc = zero4;
#pragma unroll 4   // four rows' loads in flight per half-warp (the loop is bound by the latency of its K / V row loads)
    for (int tb = s_lo; tb <= s; tb += 2) {
      const int t = tb + hsel;
      const bool valid = t <= s;                      // uniform per half-warp; all lanes execute the shuffles
      const float4 kv = (valid && okd) ? *reinterpret_cast<const float4*>(Kp + (long long)t * D + d0) : zero4;
      const float4 vr = (valid && okd) ? *reinterpret_cast<const float4*>(Vp + (long long)t * D + d0) : zero4;
      const float sc = half_sum(fmaf(q.x, kv.x, fmaf(q.y, kv.y, fmaf(q.z, kv.z, q.w * kv.w)))) * qk_scale;
      if (valid) {
        const float mn = fmaxf(m, sc);
        const float corr = bwd_ex2(m - mn), pj = bwd_ex2(sc - mn);
        l = l * corr + pj;
        acc.x = fmaf(pj, vr.x, acc.x * corr); acc.y = fmaf(pj, vr.y, acc.y * corr);
        acc.z = fmaf(pj, vr.z, acc.z * corr); acc.w = fmaf(pj, vr.w, acc.w * corr);
        m = mn;
      }
    }
    // merge the two halves (the upper one may be empty: m = -inf, l = 0)
    const float mo = __shfl_xor_sync(SMCT_FULL_MASK, m, 16), lo = __shfl_xor_sync(SMCT_FULL_MASK, l, 16);
    const float4 ao = make_float4(__shfl_xor_sync(SMCT_FULL_MASK, acc.x, 16), __shfl_xor_sync(SMCT_FULL_MASK, acc.y, 16),
                                  __shfl_xor_sync(SMCT_FULL_MASK, acc.z, 16), __shfl_xor_sync(SMCT_FULL_MASK, acc.w, 16));
    const float mt = fmaxf(m, mo);
    const float ca = bwd_ex2(m - mt), cb = (mo == -INFINITY) ? 0.f : bwd_ex2(mo - mt);
    // both halves form the same totals in the same order (lower half's terms first)
    const float l_lo = hsel ? lo : l, l_hi = hsel ? l : lo, c_lo = hsel ? cb : ca, c_hi = hsel ? ca : cb;
    const float lt = l_lo * c_lo + l_hi * c_hi;
    const float4 a_lo = hsel ? ao : acc, a_hi = hsel ? acc : ao;
    if (hsel == 0 && okd) {
      const float inv_l = 1.0f / lt;
      *reinterpret_cast<float4*>(a.Zm + u * D + d0) = make_float4((a_lo.x * c_lo + a_hi.x * c_hi) * inv_l, (a_lo.y * c_lo + a_hi.y * c_hi) * inv_l,
                                                                  (a_lo.z * c_lo + a_hi.z * c_hi) * inv_l, (a_lo.w * c_lo + a_hi.w * c_hi) * inv_l);
      *reinterpret_cast<float4*>(a.Qm + u * D + d0) = q;
    }
    if (lane == 0) { a.stat[u * 2] = mt; a.stat[u * 2 + 1] = lt; }
  }
}

// ------------------------------------------------------------------------------------------
// dense forward recompute + backward: persistent CTAs, one tile = 32 consecutive nodes of the flat node list
// (a tile may span several (b,s); every row carries its own b*S+s, ancestor id and multiplicity)
// ------------------------------------------------------------------------------------------
// rep[(k,n)] += sum_rows A[row][k] * Bm[row][n]   (CTA-private replica: plain read-modify-write).  Every output sums
// its PT products in row order; the blocked path (2 k x 4 n per thread: 6 shared loads per 8 FMAs instead of 2 per 1)
// gives the same bits as the scalar one.
__device__ __forceinline__ void tile_outer_accum(const float* A, int lda, const float* Bm, int ldb, int Kdim, int N, float* rep) {
  if ((Kdim & 1) == 0 && (N & 3) == 0 && (lda & 1) == 0 && (ldb & 3) == 0 && (reinterpret_cast<uintptr_t>(A) & 7) == 0 &&
      (reinterpret_cast<uintptr_t>(Bm) & 15) == 0) {
    if ((Kdim & 3) == 0 && (lda & 3) == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0) {
      // 4 k x 4 n per thread: two 128-bit shared loads per 16 FMAs; same sums (row order)
      const int N4q = N >> 2, K4q = Kdim >> 2;
      for (int idx = threadIdx.x; idx < K4q * N4q; idx += NT) {
        const int k = 4 * (idx / N4q), n = 4 * (idx % N4q);
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
#pragma unroll 4
        for (int r = 0; r < PT; ++r) {
          const float4 a4 = *reinterpret_cast<const float4*>(A + r * lda + k);
          const float4 b = *reinterpret_cast<const float4*>(Bm + r * ldb + n);
          const float av[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            acc[i][0] = fmaf(av[i], b.x, acc[i][0]); acc[i][1] = fmaf(av[i], b.y, acc[i][1]);
            acc[i][2] = fmaf(av[i], b.z, acc[i][2]); acc[i][3] = fmaf(av[i], b.w, acc[i][3]);
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int e = 0; e < 4; ++e) rep[(k + i) * N + n + e] += acc[i][e];
      }
      return;
    }
    // vector loads (rows padded to a multiple of 4 floats): one 64-bit + one 128-bit shared load per 8 FMAs; same sums
    const int N4 = N >> 2, K2 = Kdim >> 1;
    for (int idx = threadIdx.x; idx < K2 * N4; idx += NT) {
      const int k = 2 * (idx / N4), n = 4 * (idx % N4);
      float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll 4
      for (int r = 0; r < PT; ++r) {
        const float2 a01 = *reinterpret_cast<const float2*>(A + r * lda + k);
        const float4 b = *reinterpret_cast<const float4*>(Bm + r * ldb + n);
        acc[0][0] = fmaf(a01.x, b.x, acc[0][0]); acc[0][1] = fmaf(a01.x, b.y, acc[0][1]);
        acc[0][2] = fmaf(a01.x, b.z, acc[0][2]); acc[0][3] = fmaf(a01.x, b.w, acc[0][3]);
        acc[1][0] = fmaf(a01.y, b.x, acc[1][0]); acc[1][1] = fmaf(a01.y, b.y, acc[1][1]);
        acc[1][2] = fmaf(a01.y, b.z, acc[1][2]); acc[1][3] = fmaf(a01.y, b.w, acc[1][3]);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int e = 0; e < 4; ++e) rep[(k + i) * N + n + e] += acc[i][e];
    }
    return;
  }
  if ((Kdim & 1) == 0 && (N & 3) == 0) {
    const int N4 = N >> 2, K2 = Kdim >> 1;
    for (int idx = threadIdx.x; idx < K2 * N4; idx += NT) {
      const int k = 2 * (idx / N4), n = 4 * (idx % N4);
      float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll 4
      for (int r = 0; r < PT; ++r) {
        const float a0 = A[r * lda + k], a1 = A[r * lda + k + 1];
        const float b0 = Bm[r * ldb + n], b1 = Bm[r * ldb + n + 1], b2 = Bm[r * ldb + n + 2], b3 = Bm[r * ldb + n + 3];
        acc[0][0] = fmaf(a0, b0, acc[0][0]); acc[0][1] = fmaf(a0, b1, acc[0][1]);
        acc[0][2] = fmaf(a0, b2, acc[0][2]); acc[0][3] = fmaf(a0, b3, acc[0][3]);
        acc[1][0] = fmaf(a1, b0, acc[1][0]); acc[1][1] = fmaf(a1, b1, acc[1][1]);
        acc[1][2] = fmaf(a1, b2, acc[1][2]); acc[1][3] = fmaf(a1, b3, acc[1][3]);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int e = 0; e < 4; ++e) rep[(k + i) * N + n + e] += acc[i][e];
    }
    return;
  }
  for (int idx = threadIdx.x; idx < Kdim * N; idx += NT) {
    const int k = idx / N, n = idx - k * N;
    float acc = 0.f;
#pragma unroll 8
    for (int r = 0; r < PT; ++r) acc = fmaf(A[r * lda + k], Bm[r * ldb + n], acc);
    rep[idx] += acc;
  }
}

template <int KD>
__global__ void __launch_bounds__(NT) bwd_dense_kernel(const BwdArgs a) {
  extern __shared__ float sm[];
  // row strides: a multiple of 4 floats when the width is (vector loads in the tile GEMMs), else width + 1
  const int D = a.D, Dp = (a.D & 3) ? a.D + 1 : a.D + 4, dff = a.dff, Fp = (a.dff & 3) ? a.dff + 1 : a.dff + 4, S = a.S, Fy = a.Fy;
  float* Zt = sm;                 // z_ tile
  float* Ut = Zt + PT * Dp;       // normalised LN1 input (u hat)
  float* Ot = Ut + PT * Dp;       // o = LN1(z + x)
  float* Gt = Ot + PT * Dp;       // g = ffn + o, then dg
  float* dOt = Gt + PT * Dp;      // do, then du
  float* Ht = dOt + PT * Dp;      // h = relu(o W1 + b1)
  float* dHt = Ht + PT * Fp;      // d hpre
  float* rstd1 = dHt + PT * Fp;   // PT
  float* g_small = rstd1 + PT;    // bz, ln1_g, ln1_b, ln2_g, ln2_b, b2 (6*D), b1 (dff), Wo (D*Fy), dxs (D)
  float* g_bz = g_small; float* g_l1g = g_bz + D; float* g_l1b = g_l1g + D; float* g_l2g = g_l1b + D;
  float* g_l2b = g_l2g + D; float* g_b2 = g_l2b + D; float* g_b1 = g_b2 + D; float* g_Wo = g_b1 + dff;
  float* dxs = g_Wo + D * Fy;
  const int n_small = 6 * D + dff + D * Fy;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  float* rep = a.rep + (long long)blockIdx.x * a.lay.total;
  const float std_z0 = expf(0.5f * a.p.logvar_z[0]);
  const bool multi_z = a.logvar_dim > 1;
  const float c_obs = a.inv_n / a.sigma_obs;
  __shared__ int rbs[PT], rL[PT];
  __shared__ float rm[PT];
  for (int i = tid; i < n_small; i += NT) g_small[i] = 0.f;
  const long long nnodes = a.node_off[a.B * S];
  const long long ntiles = (nnodes + PT - 1) / PT;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long u0 = tile * PT;
    const int np = (int)min((long long)PT, nnodes - u0);
    __syncthreads();
    if (tid < PT) {
      const bool ok = tid < np;
      rbs[tid] = ok ? a.node_bs[u0 + tid] : 0;
      rL[tid] = ok ? (int)a.node_L[u0 + tid] : 0;
      rm[tid] = ok ? a.node_m[u0 + tid] : 0.f;
    }
    // S1: z_ tile
    for (int i = tid; i < PT * D; i += NT) {
      const int pl = i / D, d = i - pl * D;
      Zt[pl * Dp + d] = (pl < np) ? a.Zm[(u0 + pl) * D + d] : 0.f;
    }
    for (int i = tid; i < D; i += NT) dxs[i] = 0.f;
    __syncthreads();
    // S2: u = dense(z_) + noise_z + x
    tile_gemm(Zt, Dp, a.p.Wz, a.p.bz, D, D, [&](int pl, int n, float v) {
      float u = 0.f;
      if (pl < np) {
        const int bs = rbs[pl];
        const float std_z = multi_z ? expf(0.5f * a.p.logvar_z[n]) : std_z0;
        const float z = __fadd_rn(v, __fmul_rn(std_z, bwd_eps(a, 3, bs / S, bs % S, rL[pl], n)));
        u = z + a.xln[(long long)bs * D + n];
      }
      Ut[pl * Dp + n] = u;
    });
    __syncthreads();
    // S3: LN1
    for (int pl = wid; pl < PT; pl += NW) {
      float v[KD], s1 = 0.f;
#pragma unroll
      for (int k = 0; k < KD; ++k) { const int d = lane + 32 * k; v[k] = (d < D) ? Ut[pl * Dp + d] : 0.f; s1 += v[k]; }
      const float mean = warp_sum(s1) / (float)D;
      float s2 = 0.f;
#pragma unroll
      for (int k = 0; k < KD; ++k) { const int d = lane + 32 * k; const float c = (d < D) ? v[k] - mean : 0.f; s2 = fmaf(c, c, s2); }
      const float rstd = 1.0f / sqrtf(warp_sum(s2) / (float)D + a.ln_eps);
      if (lane == 0) rstd1[pl] = rstd;
#pragma unroll
      for (int k = 0; k < KD; ++k) {
        const int d = lane + 32 * k;
        if (d < D) {
          const float uh = (v[k] - mean) * rstd;
          Ut[pl * Dp + d] = uh;
          Ot[pl * Dp + d] = uh * a.p.ln1_g[d] + a.p.ln1_b[d];
        }
      }
    }
    __syncthreads();
    // S4, S5: FFN forward
    tile_gemm(Ot, Dp, a.p.W1, a.p.b1, D, dff, [&](int pl, int j, float v) { Ht[pl * Fp + j] = fmaxf(v, 0.f); });
    __syncthreads();
    tile_gemm(Ht, Fp, a.p.W2, a.p.b2, dff, D, [&](int pl, int n, float v) { Gt[pl * Dp + n] = v + Ot[pl * Dp + n]; });
    __syncthreads();
    // S6: LN2, output head, loss gradient, LN2 backward
    for (int pl = wid; pl < PT; pl += NW) {
      float gh[KD], r[KD], s1 = 0.f;
#pragma unroll
      for (int k = 0; k < KD; ++k) { const int d = lane + 32 * k; gh[k] = (d < D) ? Gt[pl * Dp + d] : 0.f; s1 += gh[k]; }
      const float mean = warp_sum(s1) / (float)D;
      float s2 = 0.f;
#pragma unroll
      for (int k = 0; k < KD; ++k) { const int d = lane + 32 * k; const float c = (d < D) ? gh[k] - mean : 0.f; s2 = fmaf(c, c, s2); }
      const float rstd = 1.0f / sqrtf(warp_sum(s2) / (float)D + a.ln_eps);
      float dr[KD];
#pragma unroll
      for (int k = 0; k < KD; ++k) {
        const int d = lane + 32 * k;
        gh[k] = (d < D) ? (gh[k] - mean) * rstd : 0.f;
        r[k] = (d < D) ? gh[k] * a.p.ln2_g[d] + a.p.ln2_b[d] : 0.f;
        dr[k] = 0.f;
      }
      if (pl < np) {
        // d loss / d pred: gaussian 1/2||y - pred||^2/Sigma_obs (regression era) | sparse cross-entropy of the logits
        // (compute_classic_loss, SMC_Transformer.py:148-164); both averaged over (b,p,s) and scaled by the multiplicity
        const bool ce = a.weights_mode == SMCT_WEIGHTS_CLASSIFICATION;
        float mx = -INFINITY, se = 0.f;
        int ytok = 0;
        if (ce) {
          ytok = ((const int*)a.y)[rbs[pl]];
          for (int f = 0; f < Fy; ++f) {
            float part = 0.f;
#pragma unroll
            for (int k = 0; k < KD; ++k) { const int d = lane + 32 * k; if (d < D) part = fmaf(r[k], __ldg(a.p.Wo + (long long)d * Fy + f), part); }
            const float pr = warp_sum(part);
            const float mn = fmaxf(mx, pr);
            se = se * expf(mx - mn) + expf(pr - mn);
            mx = mn;
          }
        }
        for (int f = 0; f < Fy; ++f) {
          float part = 0.f;
#pragma unroll
          for (int k = 0; k < KD; ++k) { const int d = lane + 32 * k; if (d < D) part = fmaf(r[k], __ldg(a.p.Wo + (long long)d * Fy + f), part); }
          const float pr = warp_sum(part);
          const float cw = a.classic_w ? a.classic_w[rbs[pl]] : 1.f;   // attention_mask weighting of compute_classic_loss
          const float dpred = ce ? (expf(pr - mx) / se - (f == ytok ? 1.f : 0.f)) * (a.inv_n * rm[pl] * cw)
                                 : (pr - ((const float*)a.y)[(long long)rbs[pl] * Fy + f]) * (c_obs * rm[pl] * cw);   // m final particles share this node
#pragma unroll
          for (int k = 0; k < KD; ++k) {
            const int d = lane + 32 * k;
            if (d < D) { dr[k] = fmaf(dpred, __ldg(a.p.Wo + (long long)d * Fy + f), dr[k]); atomicAdd(&g_Wo[d * Fy + f], r[k] * dpred); }
          }
        }
      }
      float t1 = 0.f, t2 = 0.f, dgh[KD];
#pragma unroll
      for (int k = 0; k < KD; ++k) {
        const int d = lane + 32 * k;
        dgh[k] = (d < D) ? dr[k] * a.p.ln2_g[d] : 0.f;
        t1 += dgh[k]; t2 = fmaf(dgh[k], gh[k], t2);
      }
      const float m1 = warp_sum(t1) / (float)D, m2 = warp_sum(t2) / (float)D;
#pragma unroll
      for (int k = 0; k < KD; ++k) {
        const int d = lane + 32 * k;
        if (d < D) {
          const float dg = rstd * (dgh[k] - m1 - gh[k] * m2);
          Gt[pl * Dp + d] = dg;
          if (pl < np) { atomicAdd(&g_l2g[d], dr[k] * gh[k]); atomicAdd(&g_l2b[d], dr[k]); atomicAdd(&g_b2[d], dg); }
        }
      }
    }
    __syncthreads();
    // S7: d hpre = (dg W2^T) * relu'
    tile_gemm(Gt, Dp, a.W2T, nullptr, D, dff, [&](int pl, int j, float v) { dHt[pl * Fp + j] = (Ht[pl * Fp + j] > 0.f) ? v : 0.f; });
    // S8: dW2 += h^T dg      (reads Ht, Gt only: no hazard with S7's writes to dHt)
    tile_outer_accum(Ht, Fp, Gt, Dp, dff, D, rep + a.lay.W2);
    __syncthreads();
    // S9: do = d hpre W1^T + dg ; db1
    tile_gemm(dHt, Fp, a.W1T, nullptr, dff, D, [&](int pl, int n, float v) { dOt[pl * Dp + n] = v + Gt[pl * Dp + n]; });
    for (int j = tid; j < dff; j += NT) {
      float acc = 0.f;
      for (int r = 0; r < PT; ++r) acc += dHt[r * Fp + j];
      g_b1[j] += acc;
    }
    // S10: dW1 += o^T d hpre
    tile_outer_accum(Ot, Dp, dHt, Fp, D, dff, rep + a.lay.W1);
    __syncthreads();
    // S11: LN1 backward -> du (in dOt), dgamma1/dbeta1, sum over particles of du
    for (int pl = wid; pl < PT; pl += NW) {
      float dov[KD], uh[KD], duh[KD], t1 = 0.f, t2 = 0.f;
#pragma unroll
      for (int k = 0; k < KD; ++k) {
        const int d = lane + 32 * k;
        dov[k] = (d < D) ? dOt[pl * Dp + d] : 0.f;
        uh[k] = (d < D) ? Ut[pl * Dp + d] : 0.f;
        duh[k] = (d < D) ? dov[k] * a.p.ln1_g[d] : 0.f;
        t1 += duh[k]; t2 = fmaf(duh[k], uh[k], t2);
      }
      const float m1 = warp_sum(t1) / (float)D, m2 = warp_sum(t2) / (float)D;
      const float rstd = rstd1[pl];
#pragma unroll
      for (int k = 0; k < KD; ++k) {
        const int d = lane + 32 * k;
        if (d < D) {
          const float du = (pl < np) ? rstd * (duh[k] - m1 - uh[k] * m2) : 0.f;
          dOt[pl * Dp + d] = du;
          if (pl < np) { atomicAdd(&g_l1g[d], dov[k] * uh[k]); atomicAdd(&g_l1b[d], dov[k]); }
        }
      }
    }
    __syncthreads();
    // S12: dbz, d xln (segmented sum over the tile's rows: rows of one (b,s) are contiguous), dWz ; S13: dz_ = du Wz^T
    for (int d = tid; d < D; d += NT) {
      float tot = 0.f, run = 0.f;
      int cur = rbs[0];
      for (int pl = 0; pl < np; ++pl) {
        if (rbs[pl] != cur) { atomicAdd(a.dxln_sum + (long long)cur * D + d, run); run = 0.f; cur = rbs[pl]; }
        const float du = dOt[pl * Dp + d];
        run += du; tot += du;
      }
      atomicAdd(a.dxln_sum + (long long)cur * D + d, run);
      g_bz[d] += tot;
    }
    tile_outer_accum(Zt, Dp, dOt, Dp, D, D, rep + a.lay.Wz);
    tile_gemm(dOt, Dp, a.WzT, nullptr, D, D, [&](int pl, int n, float v) {
      if (pl < np) a.dZm[(u0 + pl) * D + n] = v;
    });
  }
  __syncthreads();
  for (int i = tid; i < D; i += NT) {
    rep[a.lay.bz + i] += g_bz[i]; rep[a.lay.ln1_g + i] += g_l1g[i]; rep[a.lay.ln1_b + i] += g_l1b[i];
    rep[a.lay.ln2_g + i] += g_l2g[i]; rep[a.lay.ln2_b + i] += g_l2b[i]; rep[a.lay.b2 + i] += g_b2[i];
  }
  for (int i = tid; i < dff; i += NT) rep[a.lay.b1 + i] += g_b1[i];
  for (int i = tid; i < D * Fy; i += NT) rep[a.lay.Wo + i] += g_Wo[i];
}

// ------------------------------------------------------------------------------------------
// attention backward: BWD_G CTAs per sequence share its nodes (strided, so that long and short rows mix).
// Every warp of a CTA owns the target slots t = w, w+16, w+32, ... (TJ of them per pass of 16*TJ slots) and walks over
// ALL nodes of the CTA: the dk/dv sums of its slots stay in registers (no shared-memory atomics — those are CAS loops
// for fp32), dq of a node is reduced over the warps with one native global RED per element.
// ------------------------------------------------------------------------------------------
constexpr int BWD_G = 8;        // CTAs per sequence
constexpr int BWD_ANT = 512;    // threads per CTA
constexpr int BWD_TJ = 8;       // target slots per warp and pass

template <int KD>
__global__ void __launch_bounds__(BWD_ANT) bwd_attn_bwd_kernel(const BwdArgs a) {
  const int D = a.D, S = a.S;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  constexpr int NWA = BWD_ANT / 32;
  const int b = blockIdx.x / BWD_G, g = blockIdx.x % BWD_G;
  const float sqrt_d = sqrtf((float)D);
  float c_k[KD], c_v[KD];   // 1/sigma_k^2, 1/sigma_v^2 (per dimension when logvar_dim = D) times the loss normalisation
#pragma unroll
  for (int k = 0; k < KD; ++k) {
    const int d = lane + 32 * k;
    const int li = (a.logvar_dim > 1 && d < D) ? d : 0;
    c_k[k] = expf(-a.p.logvar_k[li]) * a.inv_n; c_v[k] = expf(-a.p.logvar_v[li]) * a.inv_n;
  }
  const long long u_begin = a.node_off[(long long)b * S], u_end = a.node_off[(long long)(b + 1) * S];
  for (int tbase = 0; tbase < S; tbase += NWA * BWD_TJ) {
    float dk[BWD_TJ][KD], dv[BWD_TJ][KD];
#pragma unroll
    for (int j = 0; j < BWD_TJ; ++j)
#pragma unroll
      for (int k = 0; k < KD; ++k) { dk[j][k] = 0.f; dv[j][k] = 0.f; }
    for (long long u = u_begin + g; u < u_end; u += BWD_G) {
      const int bs = a.node_bs[u];
      const int s = bs % S;
      if (s < tbase + wid) continue;                       // no target slot of this warp is visible to the node
      const int s_lo = (a.attn_window >= 0 && s > a.attn_window) ? s - a.attn_window : 0;
      const long long bp = (long long)b * a.P + a.node_p[u];
      const float* Kp = a.K + bp * S * D;
      const float* Vp = a.V + bp * S * D;
      float q[KD], dz[KD], dq[KD], part0 = 0.f;
#pragma unroll
      for (int k = 0; k < KD; ++k) {
        const int d = lane + 32 * k;
        const bool ok = d < D;
        q[k] = ok ? a.Qm[u * D + d] : 0.f;
        dz[k] = ok ? a.dZm[u * D + d] : 0.f;
        const float zz = ok ? a.Zm[u * D + d] : 0.f;
        part0 = fmaf(dz[k], zz, part0);
        dq[k] = 0.f;
      }
      const float delta = warp_sum(part0);     // sum_tau a_tau (dz . v_tau) = dz . z_
      const float m = a.stat[u * 2], l = a.stat[u * 2 + 1];
      bool any = false;
#pragma unroll
      for (int j = 0; j < BWD_TJ; ++j) {
        const int t = tbase + wid + NWA * j;
        if (t > s || t < s_lo) continue;                   // warp-uniform
        any = true;
        float pk = 0.f, pv = 0.f, kr[KD], vr[KD];
#pragma unroll
        for (int k = 0; k < KD; ++k) {
          const int d = lane + 32 * k;
          kr[k] = (d < D) ? Kp[(long long)t * D + d] : 0.f;
          vr[k] = (d < D) ? Vp[(long long)t * D + d] : 0.f;
          pk = fmaf(q[k], kr[k], pk);
          pv = fmaf(dz[k], vr[k], pv);
        }
        const float sc = warp_sum(pk) / sqrt_d;
        const float dA = warp_sum(pv);
        const float at = expf(sc - m) / l;
        const float dS = at * (dA - delta) / sqrt_d;
#pragma unroll
        for (int k = 0; k < KD; ++k) {
          dq[k] = fmaf(dS, kr[k], dq[k]);
          dk[j][k] = fmaf(dS, q[k], dk[j][k]);
          dv[j][k] = fmaf(at, dz[k], dv[j][k]);
        }
        if (t == s) {
          // direct terms of the loss on the resampled K/V noises of the node's own row: 1/2 ||K - kx||^2 / sigma_k^2
          // (same for V), once per final particle that descends from the node
          const float mult = a.node_m[u];
#pragma unroll
          for (int k = 0; k < KD; ++k) {
            const int d = lane + 32 * k;
            if (d < D) {
              const float tk = mult * c_k[k] * (kr[k] - a.kx[(long long)bs * D + d]);
              const float tv = mult * c_v[k] * (vr[k] - a.vx[(long long)bs * D + d]);
              dk[j][k] += tk; dv[j][k] += tv;
              atomicAdd(a.dkx_sum + (long long)bs * D + d, -tk);
              atomicAdd(a.dvx_sum + (long long)bs * D + d, -tv);
            }
          }
        }
      }
      if (any) {
#pragma unroll
        for (int k = 0; k < KD; ++k) {
          const int d = lane + 32 * k;
          if (d < D) atomicAdd(a.dq_sum + (long long)bs * D + d, dq[k]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < BWD_TJ; ++j) {
      const int t = tbase + wid + NWA * j;
      if (t < S) {
#pragma unroll
        for (int k = 0; k < KD; ++k) {
          const int d = lane + 32 * k;
          if (d < D) {
            atomicAdd(a.dk_sum + ((long long)b * S + t) * D + d, dk[j][k]);
            atomicAdd(a.dv_sum + ((long long)b * S + t) * D + d, dv[j][k]);
          }
        }
      }
    }
  }
}

constexpr int BWD_TJ2 = 4;      // target slots per half-warp and pass

// ------------------------------------------------------------------------------------------
// attention backward over CHAINS (d_model a multiple of 4, <= 64).  A row of d_model floats is held by 16 lanes (128-bit
// loads), so one warp works on TWO target slots at a time and every reduction is 4 shuffle steps for two (node, slot)
// pairs; BWD_CG CTAs per sequence share its nodes; the half-warp (w, h) owns the target slots 2w + h, + 2 NWA, ... of a pass
// and keeps their dk / dv sums in registers.  The nodes are walked by representative particle.  All nodes represented by particle p lie on p's
// lineage, so they attend over the SAME rows K[b,p,0..s], V[b,p,0..s]: a half-warp loads the rows of its target slots
// once per chain and keeps them in registers for every node of the chain.  The previous form (nodes in list order) reloaded 512 B per
// (node, slot) pair (27 GB of DRAM reads per 296 C4 sequences, 59 % of its stall samples on those loads, 17.7 ms);
// here the K / V tensors are read once.
// ------------------------------------------------------------------------------------------
constexpr int BWD_CNT = 256;     // threads per CTA of the chain kernel (two or more CTAs per SM walk different chains)
constexpr int BWD_CG = 16;       // CTAs per sequence

__global__ void __launch_bounds__(BWD_CNT, 2) bwd_attn_bwd_chain_kernel(const BwdArgs a) {
  const int D = a.D, S = a.S, P = a.P;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int l16 = lane & 15, hsel = lane >> 4;
  constexpr int NWA = BWD_CNT / 32;
  const int b = blockIdx.x / BWD_CG, g = blockIdx.x % BWD_CG;
  const float inv_sqrt_d = 1.0f / sqrtf((float)D);
  const float qk_scale = 1.4426950408889634f * inv_sqrt_d;   // log2-domain scores, like bwd_attn_fwd_half_kernel's stat[]
  const int d0 = 4 * l16;
  const bool okd = d0 < D;
  float ck[4], cv[4];   // 1/sigma^2 (per dimension when logvar_dim = D) times the loss normalisation
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int li = (a.logvar_dim > 1 && okd) ? d0 + e : 0;
    ck[e] = expf(-a.p.logvar_k[li]) * a.inv_n; cv[e] = expf(-a.p.logvar_v[li]) * a.inv_n;
  }
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  auto ld4g = [&](const float* p) { return okd ? *reinterpret_cast<const float4*>(p + d0) : zero4; };
  auto dot4 = [](const float4& x, const float4& y) { return fmaf(x.x, y.x, fmaf(x.y, y.y, fmaf(x.z, y.z, x.w * y.w))); };
  auto axpy4 = [](float s_, const float4& x, float4& y) { y.x = fmaf(s_, x.x, y.x); y.y = fmaf(s_, x.y, y.y); y.z = fmaf(s_, x.z, y.z); y.w = fmaf(s_, x.w, y.w); };
  for (int tbase = 0; tbase < S; tbase += 2 * NWA * BWD_TJ2) {
    float4 dk[BWD_TJ2], dv[BWD_TJ2];
#pragma unroll
    for (int j = 0; j < BWD_TJ2; ++j) { dk[j] = zero4; dv[j] = zero4; }
    const int t0 = tbase + 2 * wid + hsel;               // this half-warp's first target slot of the pass
    const int s_min = tbase + 2 * wid;                   // nodes below it see none of this warp's target slots
    for (int p = g; p < P; p += BWD_CG) {                // representatives of this CTA (strided: long and short chains mix)
      const unsigned short* chain = a.node_of + ((long long)b * P + p) * S;
      unsigned idx = chain[S - 1];
      if (idx == 0xFFFFu) continue;                      // p represents no node (CTA-uniform)
      const long long bp = (long long)b * P + p;
      const float* Kp = a.K + bp * S * D;
      const float* Vp = a.V + bp * S * D;
      float4 kr[BWD_TJ2], vr[BWD_TJ2];                   // the lineage's rows at this half-warp's target slots: once per chain
#pragma unroll
      for (int j = 0; j < BWD_TJ2; ++j) {
        const int t = t0 + 2 * NWA * j;
        kr[j] = t < S ? ld4g(Kp + (long long)t * D) : zero4;
        vr[j] = t < S ? ld4g(Vp + (long long)t * D) : zero4;
      }
      // the chain is the suffix [s_top, S-1]; the rows of the NEXT node are requested before the current one is worked on
      int s = S - 1;
      long long u = (long long)a.node_off[b * S + s] + idx;
      float4 q = ld4g(a.Qm + u * D), dz = ld4g(a.dZm + u * D), zr = ld4g(a.Zm + u * D);
      float2 ml = *reinterpret_cast<const float2*>(a.stat + u * 2);
      while (true) {
        const int bs = b * S + s;
        // ---- request the next node of the chain ----
        const bool more = (s > s_min) && (s > 0);
        unsigned idx_n = 0xFFFFu;
        if (more) idx_n = chain[s - 1];
        const bool have_n = idx_n != 0xFFFFu;
        long long u_n = 0;
        float4 q_n = zero4, dz_n = zero4, zr_n = zero4;
        float2 ml_n = make_float2(0.f, 1.f);
        if (have_n) {
          u_n = (long long)a.node_off[bs - 1] + idx_n;
          q_n = ld4g(a.Qm + u_n * D); dz_n = ld4g(a.dZm + u_n * D); zr_n = ld4g(a.Zm + u_n * D);
          ml_n = *reinterpret_cast<const float2*>(a.stat + u_n * 2);
        }
        // ---- the current node ----
        if (s >= s_min) {
          const int s_lo = (a.attn_window >= 0 && s > a.attn_window) ? s - a.attn_window : 0;
          float4 dq = zero4;
          const float delta = half_sum(dot4(dz, zr));      // sum_tau a_tau (dz . v_tau) = dz . z_
          const float m = ml.x, inv_l = 1.0f / ml.y;
#pragma unroll
          for (int j = 0; j < BWD_TJ2; ++j) {
            const int t = t0 + 2 * NWA * j;
            const bool valid = (t <= s) && (t >= s_lo);    // uniform per half-warp; the shuffles are executed by all lanes
            const float sc = half_sum(dot4(q, kr[j])) * qk_scale;
            const float dA = half_sum(dot4(dz, vr[j]));
            if (valid) {
              const float at = bwd_ex2(sc - m) * inv_l;
              const float dS = at * (dA - delta) * inv_sqrt_d;
              axpy4(dS, kr[j], dq);
              axpy4(dS, q, dk[j]);
              axpy4(at, dz, dv[j]);
              if (t == s && okd) {
                // direct terms of the loss on the resampled K/V noises of the node's own row: 1/2 ||K - kx||^2 / sigma_k^2
                // (same for V), once per final particle that descends from the node
                const float mult = a.node_m[u];
                const float4 kx = *reinterpret_cast<const float4*>(a.kx + (long long)bs * D + d0);
                const float4 vx = *reinterpret_cast<const float4*>(a.vx + (long long)bs * D + d0);
                const float tk[4] = {mult * ck[0] * (kr[j].x - kx.x), mult * ck[1] * (kr[j].y - kx.y),
                                     mult * ck[2] * (kr[j].z - kx.z), mult * ck[3] * (kr[j].w - kx.w)};
                const float tv[4] = {mult * cv[0] * (vr[j].x - vx.x), mult * cv[1] * (vr[j].y - vx.y),
                                     mult * cv[2] * (vr[j].z - vx.z), mult * cv[3] * (vr[j].w - vx.w)};
                dk[j].x += tk[0]; dk[j].y += tk[1]; dk[j].z += tk[2]; dk[j].w += tk[3];
                dv[j].x += tv[0]; dv[j].y += tv[1]; dv[j].z += tv[2]; dv[j].w += tv[3];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  atomicAdd(a.dkx_sum + (long long)bs * D + d0 + e, -tk[e]);
                  atomicAdd(a.dvx_sum + (long long)bs * D + d0 + e, -tv[e]);
                }
              }
            }
          }
          // dq of the node: the two halves worked on different slots; combine, then one native global RED per element
          dq.x += __shfl_xor_sync(SMCT_FULL_MASK, dq.x, 16); dq.y += __shfl_xor_sync(SMCT_FULL_MASK, dq.y, 16);
          dq.z += __shfl_xor_sync(SMCT_FULL_MASK, dq.z, 16); dq.w += __shfl_xor_sync(SMCT_FULL_MASK, dq.w, 16);
          if (hsel == 0 && okd) {
            float* dst = a.dq_sum + (long long)bs * D + d0;
            atomicAdd(dst, dq.x); atomicAdd(dst + 1, dq.y); atomicAdd(dst + 2, dq.z); atomicAdd(dst + 3, dq.w);
          }
        }
        if (!have_n) break;                              // warp-uniform (s_min depends on the warp only)
        --s; u = u_n; q = q_n; dz = dz_n; zr = zr_n; ml = ml_n;
      }
    }
#pragma unroll
    for (int j = 0; j < BWD_TJ2; ++j) {
      const int t = t0 + 2 * NWA * j;
      if (t < S && okd) {
        float* dkp = a.dk_sum + ((long long)b * S + t) * D + d0;
        float* dvp = a.dv_sum + ((long long)b * S + t) * D + d0;
        atomicAdd(dkp, dk[j].x); atomicAdd(dkp + 1, dk[j].y); atomicAdd(dkp + 2, dk[j].z); atomicAdd(dkp + 3, dk[j].w);
        atomicAdd(dvp, dv[j].x); atomicAdd(dvp + 1, dv[j].y); atomicAdd(dvp + 2, dv[j].z); atomicAdd(dvp + 3, dv[j].w);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// shared rows (b,s): q_/k_/v_ projections of LN1(X), kx/vx projections of X, LN1 backward, input encoding
// persistent CTAs of 128 threads, CTA-private replica
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) bwd_rows_kernel(const BwdArgs a) {
  extern __shared__ float sm[];
  const int D = a.D, S = a.S;
  float* Xs = sm;          // raw encoded row
  float* xh = Xs + D;      // normalised
  float* xl = xh + D;      // LN1(X)
  float* dq = xl + D; float* dk = dq + D; float* dv = dk + D; float* dkx = dv + D; float* dvx = dkx + D;
  float* t1 = dvx + D;     // d xln total
  float* dX = t1 + D;
  float* scratch = dX + D; // 32
  const int tid = threadIdx.x;
  float* rep = a.rep + (long long)blockIdx.x * a.lay.total;
  const long long nrows = (long long)a.B * S;
  // The D x D weight gradients of the CTA's rows are accumulated in REGISTERS when the element -> thread map is the same
  // for every row (128 % D == 0, at most 32 elements per thread and matrix): element (k, n) with n = tid % D,
  // k = tid / D + i (128 / D).  (The first version added every row's outer products to the replica in global memory:
  // 12 288 read-modify-writes per row, 32 us per row, 4.2 ms per 296 C4 sequences at 9 % issue activity.)
  const bool regacc = (128 % D == 0) && (D * D <= 128 * 32);
  const int rn = tid % D, rk0 = tid / D, rkstep = regacc ? 128 / D : 0;
  float aq[32], ak[32], av[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) aq[i] = ak[i] = av[i] = 0.f;
  float g_ln1g = 0.f, g_ln1b = 0.f, g_bq = 0.f, g_bk = 0.f, g_bv = 0.f;   // element d = tid of the vectors (D <= 128)
  for (long long row = blockIdx.x; row < nrows; row += gridDim.x) {
    __syncthreads();
    float s1 = 0.f;
    for (int d = tid; d < D; d += 128) {
      Xs[d] = a.X[row * D + d];
      dq[d] = a.dq_sum[row * D + d]; dk[d] = a.dk_sum[row * D + d]; dv[d] = a.dv_sum[row * D + d];
      dkx[d] = a.dkx_sum[row * D + d]; dvx[d] = a.dvx_sum[row * D + d];
      s1 += Xs[d];
    }
    const float mean = block_sum(s1, scratch) / (float)D;
    float s2 = 0.f;
    for (int d = tid; d < D; d += 128) { const float c = Xs[d] - mean; s2 = fmaf(c, c, s2); }
    const float rstd = 1.0f / sqrtf(block_sum(s2, scratch) / (float)D + a.ln_eps);
    for (int d = tid; d < D; d += 128) {
      xh[d] = (Xs[d] - mean) * rstd;
      xl[d] = xh[d] * a.p.ln1_g[d] + a.p.ln1_b[d];
    }
    __syncthreads();
    // d xln = (from the residual z + x) + dq Wq^T + dk Wk^T + dv Wv^T
    float u1 = 0.f, u2 = 0.f;
    for (int d = tid; d < D; d += 128) {
      float acc = a.dxln_sum[row * D + d];
      for (int n = 0; n < D; ++n)   // (transposed copies: consecutive threads read consecutive words; W[d][n] would be 32 lines per load)
        acc = fmaf(dq[n], a.WqT[n * D + d], fmaf(dk[n], a.WkT[n * D + d], fmaf(dv[n], a.WvT[n * D + d], acc)));
      t1[d] = acc;
      const float dxh = acc * a.p.ln1_g[d];
      u1 += dxh; u2 = fmaf(dxh, xh[d], u2);
    }
    const float m1 = block_sum(u1, scratch) / (float)D;
    const float m2 = block_sum(u2, scratch) / (float)D;
    for (int d = tid; d < D; d += 128) {
      float acc = rstd * (t1[d] * a.p.ln1_g[d] - m1 - xh[d] * m2);
      for (int n = 0; n < D; ++n) acc = fmaf(dkx[n], a.WkT[n * D + d], fmaf(dvx[n], a.WvT[n * D + d], acc));
      dX[d] = acc;
      g_ln1g = fmaf(t1[d], xh[d], g_ln1g);
      g_ln1b += t1[d];
      g_bq += dq[d];
      g_bk += dk[d] + dkx[d];
      g_bv += dv[d] + dvx[d];
    }
    __syncthreads();
    if (regacc) {
      const float dqn = dq[rn], dkn = dk[rn], dvn = dv[rn], dkxn = dkx[rn], dvxn = dvx[rn];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int k = rk0 + i * rkstep;
        if (k < D) {
          const float xlk = xl[k], xsk = Xs[k];
          aq[i] = fmaf(xlk, dqn, aq[i]);
          ak[i] = fmaf(xlk, dkn, fmaf(xsk, dkxn, ak[i]));
          av[i] = fmaf(xlk, dvn, fmaf(xsk, dvxn, av[i]));
        }
      }
    } else {
      for (int idx = tid; idx < D * D; idx += 128) {
        const int k = idx / D, n = idx - k * D;
        rep[a.lay.Wq + idx] += xl[k] * dq[n];
        rep[a.lay.Wk + idx] += fmaf(xl[k], dk[n], Xs[k] * dkx[n]);
        rep[a.lay.Wv + idx] += fmaf(xl[k], dv[n], Xs[k] * dvx[n]);
      }
    }
    // input encoding: X = (x W_in + b_in) * sqrt(D)  |  X = emb[tok] * sqrt(D)
    if (a.input_mode == SMCT_INPUT_LINEAR) {
      const float* xr = (const float*)a.x_in + row * a.Fx;
      for (int idx = tid; idx < a.Fx * D; idx += 128) {
        const int f = idx / D, d = idx - f * D;
        rep[a.lay.W_in + idx] += xr[f] * dX[d] * a.sqrtD;
      }
      for (int d = tid; d < D; d += 128) rep[a.lay.b_in + d] += dX[d] * a.sqrtD;
    } else if (a.input_mode == SMCT_INPUT_EMBEDDING) {
      const int tok = clamp_index(((const int*)a.x_in)[row], a.Fx);
      for (int d = tid; d < D; d += 128) rep[a.lay.W_in + (long long)tok * D + d] += dX[d] * a.sqrtD;
    }
  }
  if (tid < D) {
    rep[a.lay.ln1_g + tid] += g_ln1g; rep[a.lay.ln1_b + tid] += g_ln1b;
    rep[a.lay.bq + tid] += g_bq; rep[a.lay.bk + tid] += g_bk; rep[a.lay.bv + tid] += g_bv;
  }
  if (regacc) {
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int k = rk0 + i * rkstep;
      if (k < D) {
        const int idx = k * D + rn;
        rep[a.lay.Wq + idx] += aq[i]; rep[a.lay.Wk + idx] += ak[i]; rep[a.lay.Wv + idx] += av[i];
      }
    }
  }
}

__global__ void reduce_replicas_kernel(const float* rep, int nrep, int total, float* out) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int r = 0; r < nrep; ++r) acc += rep[(long long)r * total + i];
    out[i] += acc;
  }
}

// scatter the flat reduced gradient into the caller's per-parameter buffers (accumulating) + log-variance grads
struct ScatterArgs {
  GradLayout lay;
  int D, dff, Fy, win_elems;
  const float* flat;
  const double* loss_sums;   // forward's sums: [0] K, [1] q, [2] V, [3] z
  int scalar_logvar;         // 1: scalar log-variances (their gradients come from loss_sums here); 0: per dimension
  float inv_n, logdet_term;  // logdet_term = 0.5 * D * (rows of this call) * inv_n for the checkout loss variant, else 0
  smct_grads g;
};

__global__ void scatter_grads_kernel(const ScatterArgs a) {
  const int D = a.D, dff = a.dff;
  const int i0 = blockIdx.x * blockDim.x + threadIdx.x, st = gridDim.x * blockDim.x;
  auto put = [&](float* dst, int off, int n) {
    if (!dst) return;
    for (int i = i0; i < n; i += st) dst[i] += a.flat[off + i];
  };
  put(a.g.Wq, a.lay.Wq, D * D); put(a.g.bq, a.lay.bq, D); put(a.g.Wk, a.lay.Wk, D * D); put(a.g.bk, a.lay.bk, D);
  put(a.g.Wv, a.lay.Wv, D * D); put(a.g.bv, a.lay.bv, D); put(a.g.Wz, a.lay.Wz, D * D); put(a.g.bz, a.lay.bz, D);
  put(a.g.ln1_g, a.lay.ln1_g, D); put(a.g.ln1_b, a.lay.ln1_b, D); put(a.g.ln2_g, a.lay.ln2_g, D); put(a.g.ln2_b, a.lay.ln2_b, D);
  put(a.g.W1, a.lay.W1, D * dff); put(a.g.b1, a.lay.b1, dff); put(a.g.W2, a.lay.W2, dff * D); put(a.g.b2, a.lay.b2, D);
  put(a.g.Wo, a.lay.Wo, D * a.Fy); put(a.g.W_in, a.lay.W_in, a.win_elems); put(a.g.b_in, a.lay.b_in, D);
  if (i0 == 0 && a.loss_sums && a.scalar_logvar) {
    if (a.g.logvar_k) a.g.logvar_k[0] += (float)(-0.5 * a.loss_sums[0] * a.inv_n) + a.logdet_term;
    if (a.g.logvar_q) a.g.logvar_q[0] += (float)(-0.5 * a.loss_sums[1] * a.inv_n) + a.logdet_term;
    if (a.g.logvar_v) a.g.logvar_v[0] += (float)(-0.5 * a.loss_sums[2] * a.inv_n) + a.logdet_term;
    if (a.g.logvar_z) a.g.logvar_z[0] += (float)(-0.5 * a.loss_sums[3] * a.inv_n) + a.logdet_term;
  }
}

// Per-dimension log-variance gradients (noise_dim = "multi", self_attention_SMC.py:57-61,102-107): the loss term of a
// recorded noise tensor n (rows, D) is  inv_n * sum_rows 1/2 sum_d n_d^2 exp(-logvar_d)  (+ 1/2 sum_d logvar_d per row in
// the checkout variant), hence  d/dlogvar_d = -1/2 inv_n exp(-logvar_d) sum_rows n_d^2 + logdet.  One pass over the
// tensor; fp64 column sums (acc: D doubles, zeroed by the caller).
__global__ void __launch_bounds__(256) colsumsq_kernel(long long rows, int D, const float* __restrict__ n, double* acc) {
  const int tid = threadIdx.x;
  const int per = 256 / D > 0 ? 256 / D : 1;          // rows handled side by side by one block (D <= 256)
  const int d = tid % D, sub = tid / D;
  if (sub >= per) return;
  double sum = 0.0;
  for (long long r = (long long)blockIdx.x * per + sub; r < rows; r += (long long)gridDim.x * per) {
    const float v = n[r * D + d];
    sum += (double)v * (double)v;
  }
  atomicAdd(acc + d, sum);
}
__global__ void logvar_grad_multi_kernel(int D, const double* acc, const float* logvar, float inv_n, float logdet_per_dim, float* grad) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d < D) grad[d] += (float)(-0.5 * acc[d] * (double)expf(-logvar[d]) * (double)inv_n) + logdet_per_dim;
}

// Adam (tf.keras.optimizers.Adam: m, v moments, bias-corrected step size, epsilon outside the sqrt)
__global__ void adam_kernel(long long n, float* p, const float* g, float* m, float* v, float lr_t, float b1, float b2, float eps) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i];
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    p[i] -= lr_t * mi / (sqrtf(vi) + eps);
  }
}

}  // namespace smct

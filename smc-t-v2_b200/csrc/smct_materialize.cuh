// smct_materialize.cuh — the last pass of the fused forwards: K, V, R (B,P,S,64) <- slot-major physical rows through
// the final ancestry table (resample(params, i_t), transformer_utils.py:39-47, applied S times by the reference),
// fused with the output heads of SMC_Transformer.call (SMC_Transformer.py:231-236): pred_resampl = R·Wo,
// noise_K = K - wk(X), noise_V = V - wv(X) and the loss sums.
//
// The pass is a pure write stream: 3 x (B,P,S,64) fp32 = 29.8 GB per 296 C4 sequences, against 2.4 GB of distinct
// source rows (neighbours in genealogy order share all but ~7 % of their ancestors).  A plain fill of that size takes
// 4.0 ms on this part (7.5 TB/s).  Two earlier versions are in the history: one row per loop trip behind a
// pos -> ancestry -> row dependency chain (7.7 ms), and register-cached rows with per-thread stores (6.3 ms, LSU queue
// throttling).  This one keeps an IMAGE of 32 rows of one particle's K, V, R (3 x 8 KB) in shared memory, walks
// consecutive genealogy positions, rewrites only the rows whose ancestor changed, and ships the whole image with three
// bulk copies (cp.async.bulk shared -> global, the TMA engine): the SM issues a few dozen instructions per 24 KB
// written.  Two images alternate per CTA and four CTAs share an SM, so updates run under the stores in flight.
#pragma once
#include "smct_common.cuh"
#include "smct_fused.cuh"

namespace smct {

struct MatArgs {
  int B, P, S, Sa, Fy, logvar_dim;
  const float *Kp, *Vp, *Rp;
  const unsigned short *A0, *A1, *pos;
  const int* aflag;
  const float *kx, *vx, *Wo, *logvar_k, *logvar_v;
  const float* y;            // (B,S,Fy) or null
  float *K, *V, *R;
  float *pred_resampl, *noise_K, *noise_V;
  double* loss_sums;         // [0] K, [2] V, [4] sum (y - pred_resampl)^2
};

constexpr int MAT_ROWS = 32;      // rows of one image (one s-tile): 3 x 8 KB per image, ~55 KB per CTA -> 4 CTAs per SM
constexpr int MAT_RPH = MAT_ROWS / 16;   // rows per half-warp
constexpr int MAT_PB = 64;        // consecutive genealogy positions per CTA
constexpr int MAT_NT = 256;
constexpr int MAT_FYMAX = 16;
constexpr int MAT_INFO = 2 + MAT_FYMAX;   // per row: sum (K-kx)^2/var_k, sum (V-vx)^2/var_v, pred[Fy]

struct MatSmem {
  float img[2][3][MAT_ROWS][FD];          // [buffer][K,V,R][row][dim]
  float info[2][MAT_ROWS][MAT_INFO];
  float ytile[MAT_ROWS][MAT_FYMAX];
  unsigned short have[2][MAT_ROWS];       // physical row currently in the image (0xFFFF: none)
  int lp[MAT_PB];                         // logical particle at genealogy position q0 + i
  double red[3][MAT_NT / 32];
};

__device__ __forceinline__ uint32_t mat_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(mat_smem_u32(ssrc)), "r"(bytes)
               : "memory");
}

// grid = B x ceil(P / MAT_PB) x ceil(S / MAT_ROWS): one CTA ships one s-tile of MAT_PB consecutive genealogy positions
__global__ void __launch_bounds__(MAT_NT, 4) materialize_final_kernel(const MatArgs a) {
  extern __shared__ unsigned char mat_smem_raw[];
  MatSmem& sm = *reinterpret_cast<MatSmem*>(mat_smem_raw + ((128u - (mat_smem_u32(mat_smem_raw) & 127u)) & 127u));
  const int Fy = a.Fy, S = a.S, P = a.P;
  const int pblocks = (P + MAT_PB - 1) / MAT_PB, stiles = (S + MAT_ROWS - 1) / MAT_ROWS;
  const int b = (int)(blockIdx.x / (unsigned)(pblocks * stiles));
  const int rem = (int)(blockIdx.x % (unsigned)(pblocks * stiles));
  const int q0 = (rem / stiles) * MAT_PB;      // the CTA owns MAT_PB consecutive POSITIONS of the genealogy order
  const int s_base = (rem % stiles) * MAT_ROWS;
  const int pn = min(MAT_PB, P - q0), nrow = min(MAT_ROWS, S - s_base);
  const int tid = threadIdx.x, c4 = tid & 15, g = tid >> 4, lane = tid & 31, wid = tid >> 5;
  const unsigned hmask = (lane & 16) ? 0xFFFF0000u : 0x0000FFFFu;   // shuffles stay inside the half-warp
  const unsigned short* const Ab = (a.aflag[b] ? a.A0 : a.A1) + (long long)b * P * a.Sa;
  for (int p = tid; p < P; p += MAT_NT) {   // invert pos (logical -> position) for the CTA's positions
    const int q = (int)a.pos[(long long)b * P + p] - q0;
    if (q >= 0 && q < pn) sm.lp[q] = p;
  }
  if (tid < MAT_ROWS) { sm.have[0][tid] = 0xFFFFu; sm.have[1][tid] = 0xFFFFu; }
  if (a.y)
    for (int i = tid; i < nrow * Fy; i += MAT_NT) sm.ytile[i / Fy][i % Fy] = a.y[((long long)b * S + s_base) * Fy + i];
  float ivk[4], ivv[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int d = a.logvar_dim > 1 ? (4 * c4 + e) : 0;
    ivk[e] = a.kx ? expf(-a.logvar_k[d]) : 0.f;
    ivv[e] = a.kx ? expf(-a.logvar_v[d]) : 0.f;
  }
  const bool want_pred = a.pred_resampl != nullptr || a.y != nullptr;
  float lk = 0.f, lv = 0.f, ly = 0.f;
  const int r0 = MAT_RPH * g;                   // half-warp g owns rows r0, r0 + 1 of the image
  // the half-warp's ancestry entries of position q0 + pi as one 4-byte word (Sa and s_base + r0 are even), one particle ahead
  auto load_anc = [&](int pi) -> uint32_t {
    if (pi >= pn || r0 >= nrow) return 0u;
    return *reinterpret_cast<const uint32_t*>(Ab + (long long)(q0 + pi) * a.Sa + s_base + r0);
  };
  uint32_t av_next = load_anc(0);
  for (int pi = 0; pi < pn; ++pi) {
    const int buf = pi & 1;
    const uint32_t av = av_next;
    av_next = load_anc(pi + 1);
    // the bulk copies that read this buffer two images ago must have finished reading shared memory
    if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    __syncthreads();
    // ---- update the image: only rows whose ancestor differs from what the buffer holds ----
#pragma unroll
    for (int j = 0; j < MAT_RPH; ++j) {
      const int r = r0 + j;
      const uint32_t row = (j & 1) ? (av >> 16) : (av & 0xFFFFu);
      if (r < nrow && row != (uint32_t)sm.have[buf][r]) {      // uniform over the half-warp
        const long long src = (((long long)b * S + s_base + r) * P + row) * FD + 4 * c4;
        const float4 k4 = *reinterpret_cast<const float4*>(a.Kp + src);
        const float4 v4 = *reinterpret_cast<const float4*>(a.Vp + src);
        const float4 r4 = *reinterpret_cast<const float4*>(a.Rp + src);
        *reinterpret_cast<float4*>(&sm.img[buf][0][r][4 * c4]) = k4;
        *reinterpret_cast<float4*>(&sm.img[buf][1][r][4 * c4]) = v4;
        *reinterpret_cast<float4*>(&sm.img[buf][2][r][4 * c4]) = r4;
        if (a.kx) {
          const long long xs = ((long long)b * S + s_base + r) * FD + 4 * c4;
          const float4 kx4 = *reinterpret_cast<const float4*>(a.kx + xs);
          const float4 vx4 = *reinterpret_cast<const float4*>(a.vx + xs);
          const float nk[4] = {k4.x - kx4.x, k4.y - kx4.y, k4.z - kx4.z, k4.w - kx4.w};
          const float nv[4] = {v4.x - vx4.x, v4.y - vx4.y, v4.z - vx4.z, v4.w - vx4.w};
          float dk = 0.f, dv = 0.f;
#pragma unroll
          for (int e = 0; e < 4; ++e) { dk = fmaf(nk[e] * nk[e], ivk[e], dk); dv = fmaf(nv[e] * nv[e], ivv[e], dv); }
#pragma unroll
          for (int o = 8; o >= 1; o >>= 1) { dk += __shfl_xor_sync(hmask, dk, o); dv += __shfl_xor_sync(hmask, dv, o); }
          if (c4 == 0) { sm.info[buf][r][0] = dk; sm.info[buf][r][1] = dv; }
        }
        if (want_pred) {
          for (int f = 0; f < Fy; ++f) {
            const float* w = a.Wo + (4 * c4) * Fy + f;
            float part = fmaf(r4.x, __ldg(w), fmaf(r4.y, __ldg(w + Fy), fmaf(r4.z, __ldg(w + 2 * Fy), r4.w * __ldg(w + 3 * Fy))));
#pragma unroll
            for (int o = 8; o >= 1; o >>= 1) part += __shfl_xor_sync(hmask, part, o);
            if (c4 == 0) sm.info[buf][r][2 + f] = part;
          }
        }
        if (c4 == 0) sm.have[buf][r] = (unsigned short)row;
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the image writes become visible to the bulk-copy engine
    __syncthreads();
    // ---- ship the image: three bulk copies (one per state tensor) ----
    const long long bp = (long long)b * P + sm.lp[pi];
    const long long dst = (bp * S + s_base) * FD;
    if (tid == 0) {
      const uint32_t bytes = (uint32_t)nrow * FD * sizeof(float);
      bulk_store(a.K + dst, &sm.img[buf][0][0][0], bytes);
      bulk_store(a.V + dst, &sm.img[buf][1][0][0], bytes);
      bulk_store(a.R + dst, &sm.img[buf][2][0][0], bytes);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    // ---- per-particle heads from the per-row summaries (read-only on this buffer until its next update) ----
    if (tid < nrow) {
      const float* inf = sm.info[buf][tid];
      if (a.kx) { lk += inf[0]; lv += inf[1]; }
      if (want_pred) {
        for (int f = 0; f < Fy; ++f) {
          const float pr = inf[2 + f];
          if (a.pred_resampl) a.pred_resampl[(bp * S + s_base + tid) * Fy + f] = pr;
          if (a.y) { const float diff = sm.ytile[tid][f] - pr; ly = fmaf(diff, diff, ly); }
        }
      }
    }
    if (a.noise_K || a.noise_V) {   // optional full noise tensors (EM / diagnostics): plain stores from the image
      for (int j = 0; j < MAT_RPH; ++j) {
        const int r = r0 + j;
        if (r < nrow) {
          const long long xs = ((long long)b * S + s_base + r) * FD + 4 * c4;
          const long long o = dst + (long long)r * FD + 4 * c4;
          if (a.noise_K) {
            const float4 k4 = *reinterpret_cast<const float4*>(&sm.img[buf][0][r][4 * c4]);
            const float4 x4 = *reinterpret_cast<const float4*>(a.kx + xs);
            *reinterpret_cast<float4*>(a.noise_K + o) = make_float4(k4.x - x4.x, k4.y - x4.y, k4.z - x4.z, k4.w - x4.w);
          }
          if (a.noise_V) {
            const float4 v4 = *reinterpret_cast<const float4*>(&sm.img[buf][1][r][4 * c4]);
            const float4 x4 = *reinterpret_cast<const float4*>(a.vx + xs);
            *reinterpret_cast<float4*>(a.noise_V + o) = make_float4(v4.x - x4.x, v4.y - x4.y, v4.z - x4.z, v4.w - x4.w);
          }
        }
      }
    }
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // shared memory must outlive the copies
  if (a.loss_sums) {
    lk = warp_sum(lk); lv = warp_sum(lv); ly = warp_sum(ly);
    if (lane == 0) { sm.red[0][wid] = lk; sm.red[1][wid] = lv; sm.red[2][wid] = ly; }
    __syncthreads();
    if (tid == 0) {
      double s0 = 0, s1 = 0, s2 = 0;
      for (int i = 0; i < MAT_NT / 32; ++i) { s0 += sm.red[0][i]; s1 += sm.red[1][i]; s2 += sm.red[2][i]; }
      atomicAdd(a.loss_sums + 0, s0);
      atomicAdd(a.loss_sums + 2, s1);
      atomicAdd(a.loss_sums + 4, s2);
    }
  }
}

}  // namespace smct

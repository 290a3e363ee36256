// smct_small.cuh — SMCT_ALGO_SMALL: the whole forward of one sequence in one CTA for the reference's own small shapes
// (BASELINE.json configs 0-2: d_model 8 / 8 / 32, dff 8 / 8 / 32, 10 / 100 / 60 particles, 24 steps).
//
// The generic step tile (smct_staged.cuh) gives every particle a whole warp and walks through ~10 barrier-separated
// phases per timestep; at d_model = 8 three quarters of its lanes idle and a timestep costs 18-36 us.  Here a particle is
// served by LG = d_model / 8 lanes (1, 2 or 4) that keep its q, k, v, z, o, r rows in registers (8 dims per lane) from the
// noise generation to the log-weight, with shuffles inside the lane group for the dot products, LayerNorm statistics and
// the small dense layers (weights in shared memory).  One barrier separates the particle phase from the draw
// (resample_seq, shared with the staged path: TF-CPU categorical semantics) and one the draw from the genealogy update.
// PATH STORAGE like the fused forward: K / V / R rows are written once, slot-major ([s][particle][D]; K and V in shared
// memory when 2*S*P*D floats fit, else in the workspace), and every particle carries an ancestry row A[m][s] (uint16 in
// shared memory, ping-pong): resampling copies t+1 indices per particle instead of 3*(t+1)*D floats.  The epilogue writes
// K, V, R (B,P,S,D) through the final ancestry table together with the output heads of SMC_Transformer.call
// (pred_resampl = R Wo, K - wk(X), V - wv(X)) and the loss sums — no separate final pass.
// Served: gaussian weights, 1 head, full model, noise on, d_model in {8,16,32}, dff a multiple of LG and <= 128,
// P * LG <= 1024, F_y <= 8, no attention-weights output.  Same RNG streams as every other path (results agree with the
// staged path to rounding: the sums are formed in a different order).
#pragma once
#include "smct_staged.cuh"

namespace smct {

constexpr int SMALL_FY = 8;

struct SmallSmemLayout {   // offsets in floats into the dynamic shared memory
  int Wz, W1, W2, Wo, bz, b1, b2, l1g, l1b, l2g, l2b, sig, lw, idx, anc, hist, hist_smem, cdf_bytes_off, total_bytes;
};

inline SmallSmemLayout small_layout(int D, int dff, int Fy, int P, int S) {
  SmallSmemLayout l;
  int o = 0;
  auto take = [&](int n) { int r = o; o += (n + 3) / 4 * 4; return r; };
  l.Wz = take(D * D); l.W1 = take(D * dff); l.W2 = take(dff * D); l.Wo = take(D * Fy);
  l.bz = take(D); l.b1 = take(dff); l.b2 = take(D); l.l1g = take(D); l.l1b = take(D); l.l2g = take(D); l.l2b = take(D);
  l.sig = take(8 * D);   // std k,q,v,z ; 1/var q, z, k, v
  l.lw = take(P); l.idx = take(P);   // log-weights and drawn ancestors of the current step
  l.anc = take(P * S);               // two ancestry tables of P*S uint16 (= P*S floats)
  const long long hist_floats = 2LL * S * P * D;
  l.hist_smem = ((long long)(o + hist_floats) * 4 + (long long)P * 8 + 64 <= 200LL * 1024) ? 1 : 0;   // K, V rows in shared memory
  l.hist = l.hist_smem ? take((int)hist_floats) : 0;
  l.cdf_bytes_off = (o * 4 + 7) / 8 * 8;
  l.total_bytes = l.cdf_bytes_off + P * 8;
  return l;
}

inline bool small_supported(const smct_shape& s, const smct_io& io) {
  if (!(s.D == 8 || s.D == 16 || s.D == 32)) return false;
  const int LG = s.D / 8;
  return s.H == 1 && s.full_model == 1 && s.noise_on == 1 && s.weights_mode == SMCT_WEIGHTS_GAUSSIAN && io.attn == nullptr &&
         s.dff >= LG && s.dff % LG == 0 && s.dff <= 128 && s.F_y <= SMALL_FY && s.P >= 1 &&
         (long long)s.P * LG <= (LG == 1 ? 1024 : 512);   // 8 * LG-float rows live in registers: <= 512 threads for LG > 1
}

template <int LG>
__device__ __forceinline__ float group_allsum(float v) {   // sum over the LG lanes of a particle
#pragma unroll
  for (int o = LG / 2; o > 0; o >>= 1) v += __shfl_xor_sync(SMCT_FULL_MASK, v, o);
  return v;
}

// every lane of the group gets the full row (8 * LG values) from the 8 values each lane holds
template <int LG>
__device__ __forceinline__ void group_allgather(const float v[8], float full[8 * LG], int lane) {
  const int base = lane & ~(LG - 1);
#pragma unroll
  for (int j = 0; j < LG; ++j)
#pragma unroll
    for (int e = 0; e < 8; ++e) full[8 * j + e] = __shfl_sync(SMCT_FULL_MASK, v[e], base + j);
}

// po[8*LG] partial outputs per lane -> out[8]: lane gl of the group ends with the group sum of outputs 8 gl .. 8 gl + 7
template <int LG>
__device__ __forceinline__ void group_reduce_scatter(float po[8 * LG], float out[8], int gl) {
  if (LG == 4) {
    const bool hi2 = (gl & 2) != 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float send = hi2 ? po[i] : po[16 + i];
      const float keep = hi2 ? po[16 + i] : po[i];
      po[i] = keep + __shfl_xor_sync(SMCT_FULL_MASK, send, 2);
    }
  }
  if (LG >= 2) {
    const bool hi1 = (gl & 1) != 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float send = hi1 ? po[i] : po[8 + i];
      const float keep = hi1 ? po[8 + i] : po[i];
      out[i] = keep + __shfl_xor_sync(SMCT_FULL_MASK, send, 1);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) out[i] = po[i];
  }
}

template <int LG, int NTHR>   // NTHR: upper bound of the block size (register budget: 64 K / NTHR per thread)
__global__ void __launch_bounds__(NTHR) small_seq_forward_kernel(const SeqArgs g, const SmallSmemLayout L) {
  constexpr int D = 8 * LG;
  extern __shared__ __align__(16) float smf[];
  __shared__ float scratch[32];
  __shared__ double total_s;
  __shared__ float redl[64];
  const StepArgs& a0 = g.step;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = (int)blockDim.x >> 5;
  const int B = a0.B, P = a0.P, S = a0.S, dff = a0.dff, Fy = a0.Fy;
  const int p = tid / LG, gl = tid % LG;
  const bool act = p < P;
  const int pc = act ? p : P - 1;
  const int d0 = 8 * gl;                                    // this lane's dims [d0, d0 + 8)
  const bool multi = a0.logvar_dim > 1;
  const size_t BPD = (size_t)B * P * D, BP = (size_t)B * P;
  float* sWz = smf + L.Wz; float* sW1 = smf + L.W1; float* sW2 = smf + L.W2; float* sWo = smf + L.Wo;
  float* sbz = smf + L.bz; float* sb1 = smf + L.b1; float* sb2 = smf + L.b2;
  float* sl1g = smf + L.l1g; float* sl1b = smf + L.l1b; float* sl2g = smf + L.l2g; float* sl2b = smf + L.l2b;
  float* ssig = smf + L.sig;
  float* slw = smf + L.lw;                                  // resample_seq indexes its arrays by b * P + p
  int* sidx = reinterpret_cast<int*>(smf + L.idx);
  double* cdf = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(smf) + L.cdf_bytes_off);
  for (int i = tid; i < D * D; i += blockDim.x) sWz[i] = a0.p.Wz[i];
  for (int i = tid; i < D * dff; i += blockDim.x) { sW1[i] = a0.p.W1[i]; sW2[i] = a0.p.W2[i]; }
  for (int i = tid; i < D * Fy; i += blockDim.x) sWo[i] = a0.p.Wo[i];
  for (int i = tid; i < dff; i += blockDim.x) sb1[i] = a0.p.b1[i];
  for (int i = tid; i < D; i += blockDim.x) {
    sbz[i] = a0.p.bz[i]; sb2[i] = a0.p.b2[i];
    sl1g[i] = a0.p.ln1_g[i]; sl1b[i] = a0.p.ln1_b[i]; sl2g[i] = a0.p.ln2_g[i]; sl2b[i] = a0.p.ln2_b[i];
    const int li = multi ? i : 0;
    ssig[i] = expf(0.5f * a0.p.logvar_k[li]); ssig[D + i] = expf(0.5f * a0.p.logvar_q[li]);
    ssig[2 * D + i] = expf(0.5f * a0.p.logvar_v[li]); ssig[3 * D + i] = expf(0.5f * a0.p.logvar_z[li]);
    ssig[4 * D + i] = expf(-a0.p.logvar_q[li]); ssig[5 * D + i] = expf(-a0.p.logvar_z[li]);
    ssig[6 * D + i] = expf(-a0.p.logvar_k[li]); ssig[7 * D + i] = expf(-a0.p.logvar_v[li]);
  }
  __syncthreads();

  // physical rows, slot-major [s][particle][D]: K, V in shared memory when they fit, else in the workspace; R in the workspace
  const size_t seq_off = (size_t)b * S * P * D;
  float* const Kph = L.hist_smem ? (smf + L.hist) : (g.K2 + seq_off);
  float* const Vph = L.hist_smem ? (smf + L.hist + (size_t)S * P * D) : (g.V2 + seq_off);
  float* const Rph = g.R2 + seq_off;
  unsigned short* Acur = reinterpret_cast<unsigned short*>(smf + L.anc);     // A[m][s]: physical row of particle m at slot s
  unsigned short* Anxt = Acur + (size_t)P * S;
  const long long bp = (long long)b * P + pc;
  const unsigned long long gbp = (unsigned long long)(a0.b_global0 + b) * (unsigned long long)P + pc;
  const float inv_sqrt_d = 1.0f / sqrtf((float)D);
  double dloss_q = 0.0, dloss_z = 0.0;   // thread 0 only

  for (int t = 0; t < S; ++t) {
    const int s_lo = (g.attn_window >= 0 && t > g.attn_window) ? t - g.attn_window : 0;
    const long long rowbs = ((long long)b * S + t) * D + d0;
    float loss_q = 0.f, loss_z = 0.f;
    // ---------------- noise, slot-t rows ----------------
    float q[8], k[8], v[8], xl[8];
    {
      float q0[8], k0[8], v0[8], ek[8], eq[8], ev[8];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float4 t0 = *reinterpret_cast<const float4*>(g.q_ + rowbs + 4 * h);
        const float4 t1 = *reinterpret_cast<const float4*>(g.k_ + rowbs + 4 * h);
        const float4 t2 = *reinterpret_cast<const float4*>(g.v_ + rowbs + 4 * h);
        const float4 t3 = *reinterpret_cast<const float4*>(g.xln + rowbs + 4 * h);
        q0[4 * h] = t0.x; q0[4 * h + 1] = t0.y; q0[4 * h + 2] = t0.z; q0[4 * h + 3] = t0.w;
        k0[4 * h] = t1.x; k0[4 * h + 1] = t1.y; k0[4 * h + 2] = t1.z; k0[4 * h + 3] = t1.w;
        v0[4 * h] = t2.x; v0[4 * h + 1] = t2.y; v0[4 * h + 2] = t2.z; v0[4 * h + 3] = t2.w;
        xl[4 * h] = t3.x; xl[4 * h + 1] = t3.y; xl[4 * h + 2] = t3.z; xl[4 * h + 3] = t3.w;
        if (g.eps) {
          const size_t e0 = (size_t)bp * D + d0 + 4 * h;
          const float4 n0 = *reinterpret_cast<const float4*>(g.eps + ((size_t)0 * S + t) * BPD + e0);
          const float4 n1 = *reinterpret_cast<const float4*>(g.eps + ((size_t)1 * S + t) * BPD + e0);
          const float4 n2 = *reinterpret_cast<const float4*>(g.eps + ((size_t)2 * S + t) * BPD + e0);
          ek[4 * h] = n0.x; ek[4 * h + 1] = n0.y; ek[4 * h + 2] = n0.z; ek[4 * h + 3] = n0.w;
          eq[4 * h] = n1.x; eq[4 * h + 1] = n1.y; eq[4 * h + 2] = n1.z; eq[4 * h + 3] = n1.w;
          ev[4 * h] = n2.x; ev[4 * h + 1] = n2.y; ev[4 * h + 2] = n2.z; ev[4 * h + 3] = n2.w;
        } else {   // group index of philox_normal1: particle * ceil(D/4) + d/4
          const unsigned long long grp = gbp * (unsigned long long)(2 * LG) + (unsigned long long)(2 * gl + h);
          philox_normal4(g.seed, 0u, (uint32_t)t, grp, &ek[4 * h]);
          philox_normal4(g.seed, 1u, (uint32_t)t, grp, &eq[4 * h]);
          philox_normal4(g.seed, 2u, (uint32_t)t, grp, &ev[4 * h]);
        }
      }
      float nq[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        k[e] = __fadd_rn(k0[e], __fmul_rn(ssig[d0 + e], ek[e]));
        q[e] = __fadd_rn(q0[e], __fmul_rn(ssig[D + d0 + e], eq[e]));
        v[e] = __fadd_rn(v0[e], __fmul_rn(ssig[2 * D + d0 + e], ev[e]));
        nq[e] = q[e] - q0[e];
        loss_q = act ? fmaf(nq[e] * nq[e], ssig[4 * D + d0 + e], loss_q) : loss_q;
      }
      if (act) {
        float* kd = Kph + ((size_t)t * P + pc) * D + d0;
        float* vd = Vph + ((size_t)t * P + pc) * D + d0;
        *reinterpret_cast<float4*>(kd) = make_float4(k[0], k[1], k[2], k[3]);
        *reinterpret_cast<float4*>(kd + 4) = make_float4(k[4], k[5], k[6], k[7]);
        *reinterpret_cast<float4*>(vd) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(vd + 4) = make_float4(v[4], v[5], v[6], v[7]);
        if (g.noise_q) {
          float* nd = g.noise_q + (bp * S + t) * D + d0;
          *reinterpret_cast<float4*>(nd) = make_float4(nq[0], nq[1], nq[2], nq[3]);
          *reinterpret_cast<float4*>(nd + 4) = make_float4(nq[4], nq[5], nq[6], nq[7]);
        }
      }
    }
    // ---------------- attention over the particle's own history (rows s_lo..t-1 in global memory, row t in registers) ----------------
    float z_[8];
    {
      float m = -INFINITY, l = 0.f, acc[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = 0.f;
      const unsigned short* Arow = Acur + (size_t)pc * S;
      // (the slot-t row of this step was stored above by this very lane: reading it back keeps the loop body uniform, so
      // that the unrolled iterations' loads are issued together)
#pragma unroll 4
      for (int s = s_lo; s <= t; ++s) {
        float kr[8], vr[8];
        {
          const size_t ro = ((size_t)s * P + (s == t ? pc : (int)Arow[s])) * D + d0;
          const float4 k0 = *reinterpret_cast<const float4*>(Kph + ro), k1 = *reinterpret_cast<const float4*>(Kph + ro + 4);
          const float4 v0 = *reinterpret_cast<const float4*>(Vph + ro), v1 = *reinterpret_cast<const float4*>(Vph + ro + 4);
          kr[0] = k0.x; kr[1] = k0.y; kr[2] = k0.z; kr[3] = k0.w; kr[4] = k1.x; kr[5] = k1.y; kr[6] = k1.z; kr[7] = k1.w;
          vr[0] = v0.x; vr[1] = v0.y; vr[2] = v0.z; vr[3] = v0.w; vr[4] = v1.x; vr[5] = v1.y; vr[6] = v1.z; vr[7] = v1.w;
        }
        float part = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) part = fmaf(q[e], kr[e], part);
        const float sc = group_allsum<LG>(part) * inv_sqrt_d;
        const float mn = fmaxf(m, sc);
        const float corr = expf(m - mn), pj = expf(sc - mn);
        l = l * corr + pj;
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = fmaf(pj, vr[e], acc[e] * corr);
        m = mn;
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) z_[e] = acc[e] / l;
    }
    // ---------------- z = dense(z_) + noise ; o = LN1(z + x) ----------------
    float o[8];
    {
      float zf[D];
      group_allgather<LG>(z_, zf, lane);
      float zp[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) zp[e] = sbz[d0 + e];
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const float4 w0 = *reinterpret_cast<const float4*>(sWz + d * D + d0), w1 = *reinterpret_cast<const float4*>(sWz + d * D + d0 + 4);
        zp[0] = fmaf(zf[d], w0.x, zp[0]); zp[1] = fmaf(zf[d], w0.y, zp[1]); zp[2] = fmaf(zf[d], w0.z, zp[2]); zp[3] = fmaf(zf[d], w0.w, zp[3]);
        zp[4] = fmaf(zf[d], w1.x, zp[4]); zp[5] = fmaf(zf[d], w1.y, zp[5]); zp[6] = fmaf(zf[d], w1.z, zp[6]); zp[7] = fmaf(zf[d], w1.w, zp[7]);
      }
      float ez[8];
      if (g.eps) {
        const size_t e0 = ((size_t)3 * S + t) * BPD + (size_t)bp * D + d0;
        const float4 n0 = *reinterpret_cast<const float4*>(g.eps + e0), n1 = *reinterpret_cast<const float4*>(g.eps + e0 + 4);
        ez[0] = n0.x; ez[1] = n0.y; ez[2] = n0.z; ez[3] = n0.w; ez[4] = n1.x; ez[5] = n1.y; ez[6] = n1.z; ez[7] = n1.w;
      } else {
        const unsigned long long grp = gbp * (unsigned long long)(2 * LG) + (unsigned long long)(2 * gl);
        philox_normal4(g.seed, 3u, (uint32_t)t, grp, &ez[0]);
        philox_normal4(g.seed, 3u, (uint32_t)t, grp + 1, &ez[4]);
      }
      float u[8], nz[8], s1 = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float z = __fadd_rn(zp[e], __fmul_rn(ssig[3 * D + d0 + e], ez[e]));
        nz[e] = (a0.noise_z_mode == SMCT_NOISE_Z_CHECKOUT) ? (z - z_[e]) : (z - zp[e]);
        loss_z = act ? fmaf(nz[e] * nz[e], ssig[5 * D + d0 + e], loss_z) : loss_z;
        u[e] = z + xl[e];
        s1 += u[e];
      }
      if (act && g.noise_z) {
        float* nd = g.noise_z + (bp * S + t) * D + d0;
        *reinterpret_cast<float4*>(nd) = make_float4(nz[0], nz[1], nz[2], nz[3]);
        *reinterpret_cast<float4*>(nd + 4) = make_float4(nz[4], nz[5], nz[6], nz[7]);
      }
      const float mean = group_allsum<LG>(s1) / (float)D;
      float s2 = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) { const float c = u[e] - mean; s2 = fmaf(c, c, s2); }
      const float rstd = 1.0f / sqrtf(group_allsum<LG>(s2) / (float)D + a0.ln_eps);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float inv = rstd * sl1g[d0 + e];
        o[e] = u[e] * inv + (sl1b[d0 + e] - mean * inv);
      }
    }
    // ---------------- FFN: r_pre = relu(o W1 + b1) W2 + b2 + o ; r = LN2(r_pre) ----------------
    float r[8];
    {
      float of[D], po[D];
      group_allgather<LG>(o, of, lane);
#pragma unroll
      for (int n = 0; n < D; ++n) po[n] = 0.f;
      for (int jj = 0; jj < dff / LG; ++jj) {     // this lane's hidden units j = jj * LG + gl
        const int j = jj * LG + gl;
        float h = sb1[j];
#pragma unroll
        for (int d = 0; d < D; ++d) h = fmaf(of[d], sW1[d * dff + j], h);
        h = fmaxf(h, 0.f);
#pragma unroll
        for (int n4 = 0; n4 < D / 4; ++n4) {
          const float4 w = *reinterpret_cast<const float4*>(sW2 + j * D + 4 * n4);
          po[4 * n4] = fmaf(h, w.x, po[4 * n4]); po[4 * n4 + 1] = fmaf(h, w.y, po[4 * n4 + 1]);
          po[4 * n4 + 2] = fmaf(h, w.z, po[4 * n4 + 2]); po[4 * n4 + 3] = fmaf(h, w.w, po[4 * n4 + 3]);
        }
      }
      float ffn[8];
      group_reduce_scatter<LG>(po, ffn, gl);
      float s1 = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) { r[e] = (ffn[e] + sb2[d0 + e]) + o[e]; s1 += r[e]; }
      const float mean = group_allsum<LG>(s1) / (float)D;
      float s2 = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) { const float c = r[e] - mean; s2 = fmaf(c, c, s2); }
      const float rstd = 1.0f / sqrtf(group_allsum<LG>(s2) / (float)D + a0.ln_eps);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float inv = rstd * sl2g[d0 + e];
        r[e] = r[e] * inv + (sl2b[d0 + e] - mean * inv);
      }
      if (act) {
        float* rd = Rph + ((size_t)t * P + pc) * D + d0;
        *reinterpret_cast<float4*>(rd) = make_float4(r[0], r[1], r[2], r[3]);
        *reinterpret_cast<float4*>(rd + 4) = make_float4(r[4], r[5], r[6], r[7]);
      }
    }
    // ---------------- output layer, Gaussian log-weight ----------------
    {
      float sq = 0.f;
      for (int f = 0; f < Fy; ++f) {
        float part = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) part = fmaf(r[e], sWo[(d0 + e) * Fy + f], part);
        const float pr = group_allsum<LG>(part);
        if (act && gl == 0 && g.pred) g.pred[(bp * S + t) * Fy + f] = pr;
        const float diff = ((const float*)g.y)[((long long)b * S + t) * Fy + f] - pr;
        sq = fmaf(diff, diff, sq);
      }
      if (act && gl == 0) slw[pc] = (-0.5f * sq) / a0.sigma_obs;
    }
    // ---------------- loss partials of this step ----------------
    if (a0.loss_sums) {
      const float sq = warp_sum(loss_q), sz = warp_sum(loss_z);
      if (lane == 0) { redl[wid] = sq; redl[32 + wid] = sz; }
    }
    __syncthreads();   // slot-t rows, log-weights and loss partials of the sequence are complete
    if (a0.loss_sums && tid == 0) {
      for (int i = 0; i < nw; ++i) { dloss_q += (double)redl[i]; dloss_z += (double)redl[32 + i]; }
    }
    // ---------------- draw (TF-CPU categorical semantics) and ancestor gather ----------------
    ResampleArgs rs;
    rs.B = B; rs.P = P; rs.normalise = g.normalise;
    rs.logw = slw - (size_t)b * P;                         // shared memory: no global round trip between the phases
    rs.u = g.u ? g.u + (size_t)t * BP : nullptr;
    rs.i_forced = g.i_forced ? g.i_forced + (size_t)t * BP : nullptr;
    rs.seed = g.seed; rs.b_global0 = a0.b_global0; rs.t_rng = t;
    rs.w_out = g.w ? g.w + t : nullptr; rs.w_st = S;
    rs.logw_copy = g.logw_out ? g.logw_out + (size_t)t * BP : nullptr;
    rs.logits_out = nullptr;
    rs.i_out = sidx - (size_t)b * P;
    rs.i_out2 = g.i_out ? g.i_out + (size_t)t * BP : nullptr;
    resample_seq(rs, b, cdf, scratch, &total_s);
    __syncthreads();
    // ---------------- genealogy update: A_new[m][0..t) = A[idx[m]][0..t), A_new[m][t] = idx[m] ----------------
    if (g.len_resampling < 0 || t < g.len_resampling) {
      const int n = P * (t + 1);
      for (int i = tid; i < n; i += (int)blockDim.x) {
        const int mm = i / (t + 1), sl = i - mm * (t + 1);
        const int src = sidx[mm];
        Anxt[(size_t)mm * S + sl] = (sl == t) ? (unsigned short)src : Acur[(size_t)src * S + sl];
      }
      unsigned short* tp = Acur; Acur = Anxt; Anxt = tp;
    } else {
      for (int mm = tid; mm < P; mm += (int)blockDim.x) Acur[(size_t)mm * S + t] = (unsigned short)mm;
    }
    __syncthreads();
  }

  // ---------------- epilogue: K, V, R (B,P,S,D) through the final ancestry table + the output heads + loss sums ----------------
  {
    float lk = 0.f, lv = 0.f, ly = 0.f;
    const unsigned short* Arow = Acur + (size_t)pc * S;
#pragma unroll 2
    for (int s2 = 0; s2 < S; ++s2) {
      const size_t ro = ((size_t)s2 * P + Arow[s2]) * D + d0;
      const float4 k0 = *reinterpret_cast<const float4*>(Kph + ro), k1 = *reinterpret_cast<const float4*>(Kph + ro + 4);
      const float4 v0 = *reinterpret_cast<const float4*>(Vph + ro), v1 = *reinterpret_cast<const float4*>(Vph + ro + 4);
      const float4 r0 = *reinterpret_cast<const float4*>(Rph + ro), r1 = *reinterpret_cast<const float4*>(Rph + ro + 4);
      const size_t dst = ((size_t)bp * S + s2) * D + d0;
      if (act) {
        *reinterpret_cast<float4*>(g.K0 + dst) = k0; *reinterpret_cast<float4*>(g.K0 + dst + 4) = k1;
        *reinterpret_cast<float4*>(g.V0 + dst) = v0; *reinterpret_cast<float4*>(g.V0 + dst + 4) = v1;
        *reinterpret_cast<float4*>(g.R0 + dst) = r0; *reinterpret_cast<float4*>(g.R0 + dst + 4) = r1;
      }
      if (g.kx) {   // noise_K = K - wk(X), noise_V = V - wv(X) with the un-normalised X (SMC_Transformer.py:235-236)
        const size_t xs = ((size_t)b * S + s2) * D + d0;
        const float4 a0k = *reinterpret_cast<const float4*>(g.kx + xs), a1k = *reinterpret_cast<const float4*>(g.kx + xs + 4);
        const float4 a0v = *reinterpret_cast<const float4*>(g.vx + xs), a1v = *reinterpret_cast<const float4*>(g.vx + xs + 4);
        const float nk[8] = {k0.x - a0k.x, k0.y - a0k.y, k0.z - a0k.z, k0.w - a0k.w, k1.x - a1k.x, k1.y - a1k.y, k1.z - a1k.z, k1.w - a1k.w};
        const float nv[8] = {v0.x - a0v.x, v0.y - a0v.y, v0.z - a0v.z, v0.w - a0v.w, v1.x - a1v.x, v1.y - a1v.y, v1.z - a1v.z, v1.w - a1v.w};
        if (act && g.noise_K) {
          *reinterpret_cast<float4*>(g.noise_K + dst) = make_float4(nk[0], nk[1], nk[2], nk[3]);
          *reinterpret_cast<float4*>(g.noise_K + dst + 4) = make_float4(nk[4], nk[5], nk[6], nk[7]);
        }
        if (act && g.noise_V) {
          *reinterpret_cast<float4*>(g.noise_V + dst) = make_float4(nv[0], nv[1], nv[2], nv[3]);
          *reinterpret_cast<float4*>(g.noise_V + dst + 4) = make_float4(nv[4], nv[5], nv[6], nv[7]);
        }
        if (act) {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            lk = fmaf(nk[e] * nk[e], ssig[6 * D + d0 + e], lk);
            lv = fmaf(nv[e] * nv[e], ssig[7 * D + d0 + e], lv);
          }
        }
      }
      const float rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
      for (int f = 0; f < Fy; ++f) {
        float part = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) part = fmaf(rr[e], sWo[(d0 + e) * Fy + f], part);
        const float pr = group_allsum<LG>(part);
        if (act && gl == 0) {
          if (g.pred_resampl) g.pred_resampl[((size_t)bp * S + s2) * Fy + f] = pr;
          const float diff = ((const float*)g.y)[((size_t)b * S + s2) * Fy + f] - pr;
          ly = fmaf(diff, diff, ly);
        }
      }
    }
    if (a0.loss_sums) {
      lk = warp_sum(lk); lv = warp_sum(lv); ly = warp_sum(ly);
      __syncthreads();
      if (lane == 0) { redl[wid] = lk; redl[32 + wid] = lv; scratch[wid] = ly; }
      __syncthreads();
      if (tid == 0) {
        double s0 = 0, s1 = 0, s2 = 0;
        for (int i = 0; i < nw; ++i) { s0 += (double)redl[i]; s1 += (double)redl[32 + i]; s2 += (double)scratch[i]; }
        atomicAdd(a0.loss_sums + 0, s0); atomicAdd(a0.loss_sums + 2, s1); atomicAdd(a0.loss_sums + 4, s2);
      }
    }
  }
  if (tid == 0 && a0.loss_sums) { atomicAdd(a0.loss_sums + 1, dloss_q); atomicAdd(a0.loss_sums + 3, dloss_z); }
}

}  // namespace smct

// smct_fused_tc.cuh — fused persistent forward, tensor-core variant (default SMCT_ALGO_FUSED).
// Same structure as smct_fused.cuh (one CTA per sequence, path storage, genealogy order), but the dense
// part of every 128-particle tile runs on the 5th-gen tensor cores:
//     z  = z_·Wz            -> tcgen05.mma, fp32 accumulator in TMEM (cols   0.. 63)
//     h  = o·W1[:, chunk]   -> tcgen05.mma, one TMEM range per chunk  (cols  64..319), up to 4 chunks of 64 hidden units
//     r' = h·W2[chunk, :]   -> tcgen05.mma accumulating over chunks   (cols 320..383)
// Operands are fp16 (h,m) splits of power-of-two-scaled fp32 values (three MMA terms, see smct_tc.cuh) so that
// the result keeps ~22 mantissa bits and the 1e-4 parity bar holds with margin.  Epilogues (bias, noise, residual, LayerNorm, relu,
// re-split to fp16, output layer, log-weight) are done by all 16 warps straight from TMEM (tcgen05.ld):
// warp w owns TMEM lanes 32*(w%4).. (rows) and columns 16*(w/4).. of the 64-wide accumulator.
// Weight blocks (fp16 h,m, pre-arranged in the operand layout by tc_weight_prep_kernel): W1 stays resident in shared
// memory, Wz and the W2 chunks stream from L2 with cp.async through two buffers.  The four first FFN GEMMs are issued
// at once into four TMEM ranges, so the relu/split epilogue of a chunk overlaps the tensor-core work of the others.
// The attention over the genealogy runs on warp-level MMAs (smct_fused_attn.cuh).
#pragma once
#include "smct_fused.cuh"
#include "smct_tc.cuh"
#include "smct_fused_attn.cuh"

namespace smct {

struct FusedTcSmem {
  FusedCommon c;
  __align__(16) float bz[FD], b1[FF], b2[FD], ln1g[FD], ln1b[FD], ln2g[FD], ln2b[FD], Wo[FD * 16];   // read as float4
  float psum[2][4][FTILE];
  float rs[FTILE];                                    // per-row inverse power-of-two scale of the z_ operand
  TcHeader hdr;                                       // weight / activation scales (tc_weight_prep_kernel)
  float Os[FTILE * FLD];                              // fp32 tile: z_ -> LN1 input -> o -> LN2 input -> r
  __align__(128) unsigned char ZH[2 * tc::A_PART_BYTES];   // A operand: z_ (h,m), later the FFN hidden chunk
  __align__(128) unsigned char Ob[2 * tc::A_PART_BYTES];   // A operand: o = LN1(z + x) (h,m)
  __align__(128) unsigned char W1all[FF / FD][2 * tc::B_PART_BYTES];   // B operands: the four W1 chunks (h,m), resident for the whole kernel
  // stream buffers: Ws[0] <- Wz, W2 chunk 1, W2 chunk 3 ; Ws[1] <- W2 chunk 0, W2 chunk 2 (per tile).  Between the tiles
  // of two timesteps the same bytes hold the scratch of the draw / genealogy update (FusedResampleScratch, 16 KB).
  __align__(128) unsigned char Ws[2][2 * tc::B_PART_BYTES];
  __align__(8) uint2 cdesc[FNT / 32][16];             // per-warp column descriptors of the attention groups
  __align__(8) uint64_t mbar_z, mbar_h[FF / FD], mbar_r;
  uint32_t tmem_slot;
};
static_assert(sizeof(FusedResampleScratch) <= 2 * 2 * tc::B_PART_BYTES, "resample scratch must fit the weight stream buffers");
static_assert(sizeof(FusedTcSmem) <= 227 * 1024, "fused tc kernel shared memory exceeds 227 KB");

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(tc::smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
__device__ __forceinline__ void copy_block_async(unsigned char* dst, const unsigned char* src, int bytes) {
  for (int i = threadIdx.x; i < bytes / 16; i += FNT) cp_async16(dst + 16 * i, src + 16 * i);
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
  unsigned char* const Ksb = a.Ks + ((size_t)b * S * P << 8);
  unsigned char* const Vsb = a.Vs + ((size_t)b * S * P << 8);
  const int pl = tid >> 2, l4 = tid & 3;          // noise / output stages: particle-in-tile, lane-in-particle
  const int er = (wid & 3) * 32 + lane;           // epilogue: row of the tile (= TMEM lane)
  const int ecq = wid >> 2;                       // epilogue: column quarter, cols [16*ecq, 16*ecq+16)
  const uint32_t tlane = ((uint32_t)((wid & 3) * 32)) << 16;
  const uint32_t idesc = tc::make_idesc(128, 64);
  double dloss_q = 0.0, dloss_z = 0.0;            // thread 0 only
  uint32_t ph_z = 0, ph_h = 0, ph_r = 0;   // phase parities of the MMA completion barriers

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
  const int nch = a.dff / FD;                     // FFN chunks of 64 hidden units (1..4)
  for (int i = tid; i < a.dff; i += FNT) sm.b1[i] = a.p.b1[i];
  for (int i = tid; i < FD * Fy; i += FNT) sm.Wo[i] = a.p.Wo[i];
  if (tid == 0) sm.hdr = *reinterpret_cast<const TcHeader*>(a.wtc + (size_t)(1 + 2 * nch) * 2 * tc::B_PART_BYTES);
  if (wid == 0) tc::tmem_alloc(&sm.tmem_slot, 512);
  if (tid == 0) {
    tc::mbar_init(&sm.mbar_z, 1); tc::mbar_init(&sm.mbar_r, 1);
    for (int c = 0; c < FF / FD; ++c) tc::mbar_init(&sm.mbar_h[c], 1);
    tc::fence_mbar_init();
  }
  for (int c = 0; c < nch; ++c)   // W1 stays in shared memory for all tiles and timesteps
    copy_block_async(sm.W1all[c], a.wtc + (size_t)(1 + 2 * c) * 2 * tc::B_PART_BYTES, 2 * tc::B_PART_BYTES);
  cp_async_commit();
  cp_async_wait_all();
  tc::fence_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = sm.tmem_slot;
  const uint32_t tD_z = tmem, tD_h = tmem + 64, tD_r = tmem + 64 + FF;   // z | four hidden chunks | r
  FusedResampleScratch& rsx = *reinterpret_cast<FusedResampleScratch*>(&sm.Ws[0][0]);

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
      copy_block_async(sm.Ws[0], a.wtc, 2 * tc::B_PART_BYTES);                                      // Wz
      copy_block_async(sm.Ws[1], a.wtc + (size_t)2 * 2 * tc::B_PART_BYTES, 2 * tc::B_PART_BYTES);   // W2 chunk 0
      cp_async_commit();

      // =========================== phase A1: noise, slot-t rows (fp32 + fp16 split), q tile ===========================
      {
        const int q = q0 + pl;
        const bool act = pl < np;
        const int qc = act ? q : (P - 1);
        const int L = sm.c.Lq[qc];
        const unsigned long long gbp = (unsigned long long)(a.b_global0 + b) * (unsigned long long)P + L;
        const long long lrow = ((bP + L) * S + t) * FD;
        const uint32_t prow = (uint32_t)(t * P + qc) * FD;
        float ko[16], vo[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c = 4 * l4 + i, d0 = 4 * c;          // lane l4 owns dims [16 l4, 16 l4 + 16)
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
          float qo[4], nq[4], sdk[4], sdq[4], sdv[4], ivq[4];
          ld4(&sm.c.stdtab[0][d0], sdk); ld4(&sm.c.stdtab[1][d0], sdq); ld4(&sm.c.stdtab[2][d0], sdv); ld4(&sm.c.ivar[1][d0], ivq);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            ko[4 * i + e] = __fadd_rn(kin[e], __fmul_rn(sdk[e], ek[e]));
            qo[e] = __fadd_rn(qin[e], __fmul_rn(sdq[e], eq[e]));
            vo[4 * i + e] = __fadd_rn(vin[e], __fmul_rn(sdv[e], ev[e]));
            nq[e] = qo[e] - qin[e];
            loss_q = act ? fmaf(nq[e] * nq[e], ivq[e], loss_q) : loss_q;
          }
          // q tile for the MMA fragments: scaled by log2(e)/sqrt(d_k) (scores in the log2 domain) and the fp16 scale
          const float qs = sm.hdr.s_q;
          *reinterpret_cast<float4*>(&sm.Os[pl * FLD + d0]) = make_float4(qo[0] * qs, qo[1] * qs, qo[2] * qs, qo[3] * qs);
          if (act) {
            // fp32 rows: read again only by the final pass
            __stcs(reinterpret_cast<float4*>(Kb + prow + d0), make_float4(ko[4 * i], ko[4 * i + 1], ko[4 * i + 2], ko[4 * i + 3]));
            __stcs(reinterpret_cast<float4*>(Vb + prow + d0), make_float4(vo[4 * i], vo[4 * i + 1], vo[4 * i + 2], vo[4 * i + 3]));
            if (a.noise_q) __stcs(reinterpret_cast<float4*>(a.noise_q + lrow + d0), make_float4(nq[0], nq[1], nq[2], nq[3]));
          }
        }
        if (act) {   // fp16 (h,m) split copies of the new rows in the MMA operand order (smct_fused_attn.cuh)
          unsigned char* ksr = Ksb + ((size_t)(t * P + qc) << 8);
          unsigned char* vsr = Vsb + ((size_t)(t * P + qc) << 8);
          const float sk = sm.hdr.s_k, sv = sm.hdr.s_v;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            float x[8];
            uint4 h, m;
#pragma unroll
            for (int e = 0; e < 8; ++e) x[e] = ko[8 * half + e] * sk;
            attn::split8(x, h, m);
            *reinterpret_cast<uint4*>(ksr + 16 * (half * 4 + l4)) = h;
            *reinterpret_cast<uint4*>(ksr + 16 * (8 + half * 4 + l4)) = m;
#pragma unroll
            for (int e = 0; e < 8; ++e) x[e] = vo[8 * half + e] * sv;
            attn::split8(x, h, m);
            *reinterpret_cast<uint4*>(vsr + 32 * l4 + 16 * half) = h;
            *reinterpret_cast<uint4*>(vsr + 128 + 32 * l4 + 16 * half) = m;
          }
        }
      }
      __syncthreads();   // q tile complete; the slot-t rows of this tile are visible to the whole CTA

      // =========================== phase A2: attention over the genealogy (warp-level MMAs) ===========================
      {
        const int grp = wid & 7, hsel = wid >> 3;    // 16 particles per group; the two warps of a group split the slot blocks
        const int r = lane >> 2, c = lane & 3;
        attn::State st;
        uint4* Qf = reinterpret_cast<uint4*>(sm.Ob) + grp * 256;   // this group's Q fragments (Ob is free until the LN1 epilogue)
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {   // the two warps of a group split the four k-steps
          const int j = 2 * hsel + jj;
          const float4 x0 = *reinterpret_cast<const float4*>(&sm.Os[(16 * grp + r) * FLD + 16 * c + 4 * j]);
          const float4 x1 = *reinterpret_cast<const float4*>(&sm.Os[(16 * grp + r + 8) * FLD + 16 * c + 4 * j]);
          uint4 qh, qm;
          attn::split2(attn::sat16(x0.x), attn::sat16(x0.y), qh.x, qm.x);
          attn::split2(attn::sat16(x1.x), attn::sat16(x1.y), qh.y, qm.y);
          attn::split2(attn::sat16(x0.z), attn::sat16(x0.w), qh.z, qm.z);
          attn::split2(attn::sat16(x1.z), attn::sat16(x1.w), qh.w, qm.w);
          Qf[j * 32 + lane] = qh;
          Qf[(4 + j) * 32 + lane] = qm;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) st.Z[i][0] = st.Z[i][1] = st.Z[i][2] = st.Z[i][3] = 0.f;
        st.m0 = st.m1 = -INFINITY; st.l0 = st.l1 = 0.f;
        __syncthreads();   // Q fragments complete; Os may now be overwritten by the partial results
        const int l15 = lane & 15, hsl = lane >> 4;
        const int myq = min(q0 + 16 * grp + l15, P - 1);
        const unsigned short* Arow = Acur + (long long)myq * Sa;
        uint2* cd = sm.cdesc[wid];
        const float inv_qk = sm.hdr.inv_qk;
        const uint32_t ltmask = (1u << lane) - 1u;
        int ncols = 0;
        // column builder.  Units of 4 slots alternate between the two warps of a group.  A particle opens a column when
        // its ancestor differs from its left neighbour's (the ancestry is monotone in the position).  Old slots are
        // usually shared by all 16 particles: a unit whose ancestry words are identical across the group adds its 4
        // columns at once; otherwise the two half-warps look at two consecutive slots per pass.
        const int u_lo = s_lo >> 2;
        const int u_first = u_lo + (((u_lo & 1) != hsel) ? 1 : 0);
        uint2 av_next = make_uint2(0u, 0u);
        if (4 * u_first < t) av_next = *reinterpret_cast<const uint2*>(Arow + 4 * u_first);
        for (int u = u_first; u <= (t >> 2); u += 2) {
          const uint2 av = av_next;
          av_next = make_uint2(0u, 0u);
          if (4 * (u + 2) < t) av_next = *reinterpret_cast<const uint2*>(Arow + 4 * (u + 2));   // requested one unit ahead
          if (4 * u >= s_lo && 4 * u + 3 < t) {
            const uint32_t x0 = __shfl_sync(SMCT_FULL_MASK, av.x, 0), y0 = __shfl_sync(SMCT_FULL_MASK, av.y, 0);
            if (__all_sync(SMCT_FULL_MASK, av.x == x0 && av.y == y0)) {
              if (ncols + 4 > 16) {
                attn::process_group(st, cd, ncols, Qf, Ksb, Vsb, P, inv_qk, lane);
                ncols = 0;
              }
              if (lane < 4) {
                const uint32_t w = lane < 2 ? x0 : y0;
                const uint32_t row = (lane & 1) ? (w >> 16) : (w & 0xFFFFu);
                cd[ncols + lane] = make_uint2(((uint32_t)(4 * u + lane) << 16) | row, 16u << 8);
                attn::prefetch_column(Ksb, Vsb, 4 * u + lane, (int)row, P);
              }
              __syncwarp();
              ncols += 4;
              continue;
            }
          }
#pragma unroll 1
          for (int pass = 0; pass < 2; ++pass) {
            const int s = 4 * u + 2 * pass + hsl;
            const uint32_t w = pass == 0 ? av.x : av.y;
            int aval = hsl ? (int)(w >> 16) : (int)(w & 0xFFFFu);
            if (s == t) aval = myq;                       // the current slot: the particle's own row
            const int prev = __shfl_up_sync(SMCT_FULL_MASK, aval, 1);
            const bool head = (s >= s_lo) && (s <= t) && (l15 == 0 || aval != prev);
            const uint32_t hb = __ballot_sync(SMCT_FULL_MASK, head);
            uint32_t pend = hb;
            while (pend) {
              uint32_t take = pend;
              if (ncols + __popc(take) > 16) {
                const uint32_t low = pend & 0xFFFFu;
                if (low != 0u && ncols + __popc(low) <= 16) {
                  take = low;
                } else if (ncols > 0) {
                  attn::process_group(st, cd, ncols, Qf, Ksb, Vsb, P, inv_qk, lane);
                  ncols = 0;
                  continue;
                } else {
                  take = low ? low : pend;                // an empty group always has room for one slot
                }
              }
              if (head && ((take >> lane) & 1u)) {
                const int col = ncols + __popc(take & ltmask);
                const uint32_t above = (hsl ? (hb >> 16) : (hb & 0xFFFFu)) >> (l15 + 1);
                const int hi = above ? (l15 + __ffs((int)above)) : 16;
                cd[col] = make_uint2(((uint32_t)s << 16) | (uint32_t)aval, ((uint32_t)hi << 8) | (uint32_t)l15);
                attn::prefetch_column(Ksb, Vsb, s, aval, P);
              }
              __syncwarp();
              ncols += __popc(take);
              pend &= ~take;
            }
          }
        }
        if (ncols > 0) attn::process_group(st, cd, ncols, Qf, Ksb, Vsb, P, inv_qk, lane);
        // ---- merge the two halves of every group through shared memory ----
        const int R0 = 16 * grp + r, R1 = R0 + 8;
        if (hsel == 1) {
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {   // dims 16c + 4 k4 .. +3 : Z[i][0] = dim 16c+i, Z[i][1] = dim 16c+8+i
            const int i0 = (k4 & 1) * 4, e = k4 >> 1;
            *reinterpret_cast<float4*>(&sm.Os[R0 * FLD + 16 * c + 4 * k4]) = make_float4(st.Z[i0][e], st.Z[i0 + 1][e], st.Z[i0 + 2][e], st.Z[i0 + 3][e]);
            *reinterpret_cast<float4*>(&sm.Os[R1 * FLD + 16 * c + 4 * k4]) = make_float4(st.Z[i0][2 + e], st.Z[i0 + 1][2 + e], st.Z[i0 + 2][2 + e], st.Z[i0 + 3][2 + e]);
          }
          if (c == 0) { sm.psum[0][0][R0] = st.m0; sm.psum[0][1][R0] = st.l0; sm.psum[0][0][R1] = st.m1; sm.psum[0][1][R1] = st.l1; }
        }
        __syncthreads();
        if (hsel == 0) {
          const float mB0 = sm.psum[0][0][R0], lB0 = sm.psum[0][1][R0], mB1 = sm.psum[0][0][R1], lB1 = sm.psum[0][1][R1];
          const float mt0 = fmaxf(st.m0, mB0), mt1 = fmaxf(st.m1, mB1);
          const float cA0 = attn::ex2f(st.m0 - mt0), cB0 = attn::ex2f(mB0 - mt0);
          const float cA1 = attn::ex2f(st.m1 - mt1), cB1 = attn::ex2f(mB1 - mt1);
          const float il0 = sm.hdr.inv_v / fmaf(st.l0, cA0, lB0 * cB0), il1 = sm.hdr.inv_v / fmaf(st.l1, cA1, lB1 * cB1);
          float z0[16], z1[16];
          float rmax0 = 0.f, rmax1 = 0.f;
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const int i0 = (k4 & 1) * 4, e = k4 >> 1;
            const float4 b0 = *reinterpret_cast<const float4*>(&sm.Os[R0 * FLD + 16 * c + 4 * k4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&sm.Os[R1 * FLD + 16 * c + 4 * k4]);
            const float pb0[4] = {b0.x, b0.y, b0.z, b0.w}, pb1[4] = {b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              z0[4 * k4 + k] = fmaf(st.Z[i0 + k][e], cA0, pb0[k] * cB0) * il0;
              z1[4 * k4 + k] = fmaf(st.Z[i0 + k][2 + e], cA1, pb1[k] * cB1) * il1;
              rmax0 = fmaxf(rmax0, fabsf(z0[4 * k4 + k])); rmax1 = fmaxf(rmax1, fabsf(z1[4 * k4 + k]));
            }
          }
          rmax0 = fmaxf(rmax0, __shfl_xor_sync(SMCT_FULL_MASK, rmax0, 1)); rmax0 = fmaxf(rmax0, __shfl_xor_sync(SMCT_FULL_MASK, rmax0, 2));
          rmax1 = fmaxf(rmax1, __shfl_xor_sync(SMCT_FULL_MASK, rmax1, 1)); rmax1 = fmaxf(rmax1, __shfl_xor_sync(SMCT_FULL_MASK, rmax1, 2));
          const int ez0 = tc::scale_exponent(rmax0), ez1 = tc::scale_exponent(rmax1);   // exact power-of-two scales of the rows' fp16 operand
          const float zs0 = ldexpf(1.f, ez0), zs1 = ldexpf(1.f, ez1);
          if (c == 0) { sm.rs[R0] = ldexpf(1.f, -ez0); sm.rs[R1] = ldexpf(1.f, -ez1); }
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const int d0 = 16 * c + 4 * k4;
            *reinterpret_cast<float4*>(&sm.Os[R0 * FLD + d0]) = make_float4(z0[4 * k4], z0[4 * k4 + 1], z0[4 * k4 + 2], z0[4 * k4 + 3]);
            *reinterpret_cast<float4*>(&sm.Os[R1 * FLD + d0]) = make_float4(z1[4 * k4], z1[4 * k4 + 1], z1[4 * k4 + 2], z1[4 * k4 + 3]);
          }
#pragma unroll
          for (int k8 = 0; k8 < 2; ++k8) {
            float zz0[8], zz1[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) { zz0[e] = z0[8 * k8 + e] * zs0; zz1[e] = z1[8 * k8 + e] * zs1; }
            tc::store_a8(sm.ZH, R0, 16 * c + 8 * k8, zz0);
            tc::store_a8(sm.ZH, R1, 16 * c + 8 * k8, zz1);
          }
        }
      }
      cp_async_wait_all();
      tc::fence_async_smem();
      tc::fence_before_sync();
      __syncthreads();
      tc::fence_after_sync();

      // =========================== phase B: dense tile on the tensor cores ===========================
      if (wid == 0 && tc::elect_one()) {
        tc::mma_block_f16x3(tD_z, tc::smem_u32(sm.ZH), tc::smem_u32(sm.Ws[0]), tc::B_PART_BYTES, idesc, false);
        tc::commit(&sm.mbar_z);
      }
      const bool eact = er < np;
      const int eL = sm.c.Lq[eact ? (q0 + er) : (P - 1)];
      const int n0 = 16 * ecq;
      float val[16];
      float ezv[16];   // z noise of this thread's 16 columns: generated while the tensor core works
      {
        const unsigned long long gbp = (unsigned long long)(a.b_global0 + b) * (unsigned long long)P + eL;
#pragma unroll
        for (int gq = 0; gq < 4; ++gq) {
          const int n = n0 + 4 * gq;
          if (a.eps) {
            const float4 t3 = *reinterpret_cast<const float4*>(a.eps + (3LL * S) * ((long long)a.B * P * FD) +
                                                               ((long long)t * a.B + b) * P * FD + (long long)eL * FD + n);
            ezv[4 * gq] = t3.x; ezv[4 * gq + 1] = t3.y; ezv[4 * gq + 2] = t3.z; ezv[4 * gq + 3] = t3.w;
          } else {
            philox_normal4(a.seed, 3u, (uint32_t)t, gbp * 16ull + (n >> 2), &ezv[4 * gq]);
          }
        }
      }
      tc::mbar_wait(&sm.mbar_z, ph_z);
      ph_z ^= 1;
      tc::fence_after_sync();
      // Wz has been consumed: W2 chunk 1 travels into its buffer during the LN1 epilogue
      if (nch > 1) {
        copy_block_async(sm.Ws[0], a.wtc + (size_t)4 * 2 * tc::B_PART_BYTES, 2 * tc::B_PART_BYTES);
        cp_async_commit();
      }
      {  // ---- z = dense(z_) + noise ; LN1(z + x) ----
        tmem_ld16(tD_z + tlane + n0, val);
        const float zscale = sm.rs[er] * sm.hdr.inv_wz;
        float s1 = 0.f;
#pragma unroll
        for (int gq = 0; gq < 4; ++gq) {
          const int n = n0 + 4 * gq;
          const float4 zold = *reinterpret_cast<const float4*>(&sm.Os[er * FLD + n]);
          const float z_[4] = {zold.x, zold.y, zold.z, zold.w};
          float nzv[4], cbz[4], csd[4], civ[4], cx[4];
          ld4(&sm.bz[n], cbz); ld4(&sm.c.stdtab[3][n], csd); ld4(&sm.c.ivar[3][n], civ); ld4(&sm.c.rowbuf[0][n], cx);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float zp = fmaf(val[4 * gq + e], zscale, cbz[e]);   // scale is a power of two: exact
            const float z = __fadd_rn(zp, __fmul_rn(csd[e], ezv[4 * gq + e]));
            nzv[e] = (a.noise_z_mode == SMCT_NOISE_Z_CHECKOUT) ? (z - z_[e]) : (z - zp);
            loss_z = eact ? fmaf(nzv[e] * nzv[e], civ[e], loss_z) : loss_z;
            val[4 * gq + e] = z + cx[e];
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
          float o[4], cg[4], cb[4];
          ld4(&sm.ln1g[n], cg); ld4(&sm.ln1b[n], cb);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float inv = rstd * cg[e];
            o[e] = val[4 * gq + e] * inv + (cb[e] - mean * inv);
          }
          *reinterpret_cast<float4*>(&sm.Os[er * FLD + n]) = make_float4(o[0], o[1], o[2], o[3]);
#pragma unroll
          for (int e = 0; e < 4; ++e) val[4 * gq + e] = o[e] * sm.hdr.s_o;
        }
        tc::store_a8(sm.Ob, er, n0, &val[0]);
        tc::store_a8(sm.Ob, er, n0 + 8, &val[8]);
      }
      tc::fence_async_smem();
      tc::fence_before_sync();
      __syncthreads();
      tc::fence_after_sync();
      // ---- FFN: the four first GEMMs (o · W1 chunk, W1 resident) are issued at once into four TMEM ranges; the relu /
      //      split epilogue of chunk c then overlaps the tensor-core work of the later chunks and of the second GEMMs ----
      if (wid == 0 && tc::elect_one()) {
        for (int c = 0; c < nch; ++c) {
          tc::mma_block_f16x3(tD_h + FD * c, tc::smem_u32(sm.Ob), tc::smem_u32(sm.W1all[c]), tc::B_PART_BYTES, idesc, false);
          tc::commit(&sm.mbar_h[c]);
        }
      }
      for (int c = 0; c < nch; ++c) {
        tc::mbar_wait(&sm.mbar_h[c], ph_h);
        tc::fence_after_sync();
        tmem_ld16(tD_h + FD * c + tlane + n0, val);
        const float hscale = sm.hdr.inv_o * sm.hdr.inv_w1;
        float hv[16];
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          float cb1[4];
          ld4(&sm.b1[c * FD + n0 + 4 * j4], cb1);
#pragma unroll
          for (int e = 0; e < 4; ++e) hv[4 * j4 + e] = fmaxf(fmaf(val[4 * j4 + e], hscale, cb1[e]), 0.f) * sm.hdr.s_h;
        }
        if (c > 0) {   // the previous second GEMM has read ZH and its W2 buffer: both may be refilled
          tc::mbar_wait(&sm.mbar_r, ph_r);
          ph_r ^= 1;
          tc::fence_after_sync();
          if (c + 1 < nch) {
            copy_block_async(sm.Ws[(c + 1 + 1) & 1], a.wtc + (size_t)(2 + 2 * (c + 1)) * 2 * tc::B_PART_BYTES, 2 * tc::B_PART_BYTES);
            cp_async_commit();
          }
        }
#pragma unroll
        for (int g8 = 0; g8 < 2; ++g8) tc::store_a8(sm.ZH, er, n0 + 8 * g8, &hv[8 * g8]);
        if (c == 0 || c + 1 == nch) cp_async_wait_all(); else cp_async_wait_1();   // W2 chunk c has landed (chunk c+1 may still travel)
        tc::fence_async_smem();
        tc::fence_before_sync();
        __syncthreads();
        tc::fence_after_sync();
        if (wid == 0 && tc::elect_one()) {
          tc::mma_block_f16x3(tD_r, tc::smem_u32(sm.ZH), tc::smem_u32(sm.Ws[(c + 1) & 1]), tc::B_PART_BYTES, idesc, c > 0);
          tc::commit(&sm.mbar_r);
        }
      }
      ph_h ^= 1;
      tc::mbar_wait(&sm.mbar_r, ph_r);
      ph_r ^= 1;
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
          float cb2[4];
          ld4(&sm.b2[n], cb2);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            val[4 * gq + e] = fmaf(val[4 * gq + e], rscale, cb2[e]) + o[e];
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
          float r4[4], cg[4], cb[4];
          ld4(&sm.ln2g[n], cg); ld4(&sm.ln2b[n], cb);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float inv = rstd * cg[e];
            r4[e] = val[4 * gq + e] * inv + (cb[e] - mean * inv);
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
    fused_resample_update(sm.c, rsx, a, b, t, Acur, Anxt, cur_is_A0);
  }  // timesteps

  for (int p = tid; p < P; p += FNT) a.pos_out[bP + p] = sm.c.pos[p];
  if (tid == 0) {
    a.aflag[b] = cur_is_A0;
    if (a.loss_sums) { atomicAdd(a.loss_sums + 1, dloss_q); atomicAdd(a.loss_sums + 3, dloss_z); }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (wid == 0) tc::tmem_dealloc(tmem, 512);
}

}  // namespace smct

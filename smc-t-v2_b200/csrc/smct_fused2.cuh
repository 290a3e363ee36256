// smct_fused2.cuh — fused persistent forward, second generation (default SMCT_ALGO_FUSED).
//
// Same algorithm and storage as smct_fused_tc.cuh (one CTA per sequence for the whole trajectory, path storage with
// ancestry rows, genealogy order, attention over distinct (slot,row) columns on warp-level MMAs, dense part on tcgen05
// with fp16 (h,m) split operands and fp32 TMEM accumulators) but restructured around what the round-1 profile showed:
// the first-generation kernel (512 threads, 1 CTA/SM) was a chain of __syncthreads-separated, latency-bound phases
// (IPC 0.37 per scheduler; 12 % of the time at the barrier behind the attention, ~40 % of the dense part waiting for
// barriers / MMA round trips, one thread busy during the draw).
//
//  * TWO CTAs PER SM, 256 threads each (<= 113 KB shared memory, 256 TMEM columns, 128 registers): while one sequence
//    sits in a synchronisation bubble, a memory-latency chain or the serial part of the draw, the other one issues.
//    The fp32 activation tile of generation one is gone: z_ and o are kept only as fp16 (h,m) operand tiles (22 bits)
//    and read back from there where the epilogues need them; W1 is streamed like the other weight blocks.
//  * NO CTA-WIDE BARRIER INSIDE A TILE.  A warp owns 16 particles from the noise generation through the end of the
//    attention (Q fragments, slot-t rows and column descriptors are warp-private: __syncwarp only).  The dense part is
//    sequenced with mbarriers: the eight warps ARRIVE on `ready` when their operand rows are written, one elected lane
//    of warp 0 waits and issues the tcgen05.mma's, everybody waits on the tcgen05.commit barriers.  The two warps that
//    share a TMEM lane quadrant (column halves of the same 32 rows) exchange LayerNorm partials through a 64-thread
//    named barrier (Chan's pairwise mean / M2 combination: one exchange per LayerNorm).
//  * WEIGHTS ARRIVE BY BULK COPY (cp.async.bulk, the TMA engine's linear mode, SASS UBLKCP): the elected lane posts
//    the 16 KB operand blocks (pre-arranged by tc_weight_prep_kernel) with mbarrier expect_tx; no thread spends issue
//    slots or registers on the weight stream.  Two buffers: A <- W1 chunks, B <- Wz, W2 chunks; a block is requested as
//    soon as the commit barrier of the MMA that read its buffer has completed, one epilogue ahead of its use.
//    Between timesteps the buffers hold the scratch of the draw / genealogy update.
//
// Served shapes: d_model = 64, dff in {64,128,192,256}, 1 head, full model, gaussian weights, P <= 1024, F_y <= 4.
#pragma once
#include "smct_fused.cuh"
#include "smct_tc.cuh"
#include "smct_fused_attn.cuh"

namespace smct {

constexpr int F2NT = 256;     // threads per CTA (8 warps); 2 CTAs per SM
constexpr int F2NW = F2NT / 32;
constexpr int F2FY = 4;       // largest F_y served by this kernel
constexpr int F2TMEM = 256;   // TMEM columns per CTA: z | h even | h odd | r

struct Fused2Smem {
  __align__(128) unsigned char ZH[2 * tc::A_PART_BYTES];   // A operand: z_ (h,m), later the FFN hidden chunk
  __align__(128) unsigned char Ob[2 * tc::A_PART_BYTES];   // Q fragments during the attention; A operand o = LN1(z + x) (h,m)
  __align__(128) unsigned char W[2][2 * tc::B_PART_BYTES]; // weight stream: [0] <- W1 chunks, [1] <- Wz, W2 chunks; resample scratch
  float lw[FPMAX];                                         // unnormalised log-weights by logical particle
  unsigned short Lq[FPMAX];                                // logical particle at genealogy position q
  __align__(16) float rowbuf[4][FD];                       // xln, q_, k_, v_ of the current step
  __align__(16) float bz[FD], b2[FD], ln1g[FD], ln1b[FD], ln2g[FD], ln2b[FD];
  __align__(16) float b1[FF];
  __align__(16) float Wo[FD * F2FY];
  __align__(16) float4 xch[2][FTILE];                      // pair exchange: LayerNorm partials / output-layer partials;
                                                           // scratch of the block reductions during the draw
  float rs[FTILE];                                         // per-row inverse power-of-two scale of the z_ operand
  __align__(8) uint2 cdesc[F2NW][16];                      // per-warp column descriptors of the attention groups
  TcHeader hdr;
  __align__(16) float sig[8];                              // scalar noise scales: std k,q,v,z ; 1/var q, z
  __align__(8) uint64_t mbar_ready, mbar_full[2], mbar_z, mbar_h[2], mbar_r;
  uint32_t tmem_slot;
};
struct Fused2Scratch {   // scratch of the draw / genealogy update; lives in the weight stream buffers between timesteps
  double cdf[FPMAX];                    // fp64 CDF of the draw; afterwards the per-ancestor child counts and their scan (uint32)
  unsigned short cnt[F2NW][FPMAX];      // children of ancestor position a among the logical ids of segment w
  unsigned short src[FPMAX], anc[FPMAX], rank[FPMAX];
};
static_assert(sizeof(Fused2Scratch) <= 2 * 2 * tc::B_PART_BYTES, "resample scratch must fit the weight stream buffers");
static_assert(sizeof(Fused2Smem) + 128 <= 113 * 1024, "two CTAs of the fused kernel must fit one SM");

namespace f2 {

__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* mbar) {
  const uint32_t mb = tc::smem_u32(mbar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(tc::smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(mb)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* mbar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(mbar)) : "memory");
}
// the two warps that share TMEM lane quadrant q (warps q and q + 4)
__device__ __forceinline__ void pair_sync(int quadrant) {   // literal barrier ids: the kernel then reserves 5 barriers, not 16
  switch (quadrant) {
    case 0: asm volatile("bar.sync 1, 64;" ::: "memory"); break;
    case 1: asm volatile("bar.sync 2, 64;" ::: "memory"); break;
    case 2: asm volatile("bar.sync 3, 64;" ::: "memory"); break;
    default: asm volatile("bar.sync 4, 64;" ::: "memory"); break;
  }
}
// operand rows written by this warp are complete: make them visible to the tensor-core (async) proxy and tell warp 0
__device__ __forceinline__ void warp_rows_ready(uint64_t* mbar, int lane) {
  tc::fence_async_smem();
  tc::fence_before_sync();
  __syncwarp();
  if (lane == 0) mbar_arrive(mbar);
}
// 8 consecutive k-values of row r of an A operand tile, reconstructed from the (h,m) split (scaled values)
__device__ __forceinline__ void load_a8(const unsigned char* base, int r, int k0, float x[8]) {
  const uint32_t off = tc::operand_offset(r, k0);
  const uint4 h = *reinterpret_cast<const uint4*>(base + off);
  const uint4 m = *reinterpret_cast<const uint4*>(base + tc::A_PART_BYTES + off);
  const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, mw[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hw[i]));
    const float2 mf = __half22float2(*reinterpret_cast<const __half2*>(&mw[i]));
    x[2 * i] = hf.x + mf.x;
    x[2 * i + 1] = hf.y + mf.y;
  }
}

// logsumexp normalise, TF-CPU categorical draw (fp64 sequential CDF in logical particle order), genealogy update.
// Same semantics as fused_resample_update (smct_fused.cuh); the position table lives in global memory (pos_out).
__device__ __forceinline__ void resample_update(Fused2Smem& sm, Fused2Scratch& rx, const FusedArgs& a, int b, int t,
                                                unsigned short*& Acur, unsigned short*& Anxt, int& cur_is_A0) {
  const int tid = threadIdx.x;
  const int P = a.P, S = a.S, Sa = a.Sa;
  const long long bP = (long long)b * P;
  float* red = reinterpret_cast<float*>(&sm.xch[0][0]);
  double* dred = reinterpret_cast<double*>(&sm.xch[1][0]);
  unsigned short* pos = a.pos_out + bP;
  float m1 = -INFINITY;
  for (int p = tid; p < P; p += F2NT) { const float v = sm.lw[p]; if (isfinite(v)) m1 = fmaxf(m1, v); }
  m1 = block_max(m1, red);
  float s1 = 0.f;
  for (int p = tid; p < P; p += F2NT) s1 += expf(sm.lw[p] - m1);
  s1 = block_sum(s1, red);
  const float lse = m1 + logf(s1);
  float m2 = -INFINITY;
  for (int p = tid; p < P; p += F2NT) { const float v = sm.lw[p] - lse; if (isfinite(v)) m2 = fmaxf(m2, v); }
  m2 = block_max(m2, red);
  for (int p = tid; p < P; p += F2NT) {
    const float raw = sm.lw[p];
    const float lg = raw - lse;
    if (a.w) a.w[(bP + p) * S + t] = expf(lg);
    if (a.logw) a.logw[((long long)t * a.B + b) * P + p] = raw;
    rx.cdf[p] = isfinite(lg) ? exp((double)lg - (double)m2) : 0.0;
  }
  __syncthreads();
  if (tid == 0) {
    double run = 0.0;
    for (int p = 0; p < P; ++p) { run += rx.cdf[p]; rx.cdf[p] = run; }
    dred[0] = run;
  }
  __syncthreads();
  const double total = dred[0];
  for (int p = tid; p < P; p += F2NT) {
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

  const bool resample_now = (a.len_resampling < 0) || (t < a.len_resampling);
  if (resample_now) {
    // New genealogy order = the new particles sorted by (position of the ancestor, logical id): a stable counting sort.
    // The 8 warps own 8 contiguous segments of logical ids; cnt[w][a] counts the children of ancestor position a in
    // segment w, so  position(p) = (children of smaller a) + (children of a in earlier segments) + (rank of p among the
    // children of a in its own segment).  Deterministic (no atomics): inside a warp, __match_any groups equal keys.
    unsigned short* cnt = &rx.cnt[0][0];                                       // [F2NW][P]
    unsigned short* rank = rx.rank;                                            // rank of p among the children of a in its segment
    const int lane = tid & 31, wid = tid >> 5;
    for (int i = tid; i < F2NW * P; i += F2NT) cnt[i] = 0;
    __syncthreads();
    const int seg = (P + F2NW - 1) / F2NW;
    const int p_lo = wid * seg, p_hi = min(P, p_lo + seg);
    for (int p0 = p_lo; p0 < p_hi; p0 += 32) {
      const int p = p0 + lane;
      const bool on = p < p_hi;
      const int akey = on ? (int)pos[rx.anc[p]] : (0x10000 + lane);           // inactive lanes: unique keys
      const unsigned m = __match_any_sync(SMCT_FULL_MASK, akey);
      if (on) {
        const int before = cnt[wid * P + akey];
        rank[p] = (unsigned short)(before + __popc(m & ((1u << lane) - 1u)));
        rx.anc[p] = (unsigned short)akey;                                      // from here on: the ancestor's POSITION
      }
      __syncwarp();
      if (on && (m & ((1u << lane) - 1u)) == 0u) cnt[wid * P + akey] = (unsigned short)(cnt[wid * P + akey] + __popc(m));
      __syncwarp();
    }
    __syncthreads();
    // per ancestor position: exclusive prefix over the segments and the total (the CDF is dead: its bytes hold the totals)
    unsigned int* tot = reinterpret_cast<unsigned int*>(rx.cdf);                // [P] totals, then their exclusive scan
    for (int aa = tid; aa < P; aa += F2NT) {
      unsigned int run = 0;
#pragma unroll
      for (int w = 0; w < F2NW; ++w) { const unsigned int c = cnt[w * P + aa]; cnt[w * P + aa] = (unsigned short)run; run += c; }
      tot[aa] = run;
    }
    __syncthreads();
    {  // exclusive scan of tot[0..P): each thread owns 4 consecutive entries (P <= 1024 = 4 * F2NT)
      unsigned int v[4], sum = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) { const int aa = 4 * tid + k; v[k] = aa < P ? tot[aa] : 0u; sum += v[k]; }
      unsigned int inc = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const unsigned int n = __shfl_up_sync(SMCT_FULL_MASK, inc, o); if (lane >= o) inc += n; }
      unsigned int* wsum = reinterpret_cast<unsigned int*>(red);
      if (lane == 31) wsum[wid] = inc;
      __syncthreads();
      unsigned int base = 0;
      for (int w = 0; w < wid; ++w) base += wsum[w];
      unsigned int run = base + inc - sum;
#pragma unroll
      for (int k = 0; k < 4; ++k) { const int aa = 4 * tid + k; if (aa < P) tot[aa] = run; run += v[k]; }
    }
    __syncthreads();
    for (int p = tid; p < P; p += F2NT) {
      const int akey = rx.anc[p];
      const int qn = (int)tot[akey] + (int)cnt[(p / seg) * P + akey] + (int)rank[p];
      sm.Lq[qn] = (unsigned short)p;
      pos[p] = (unsigned short)qn;
      rx.src[qn] = (unsigned short)akey;   // position of the ancestor of the particle now at qn
    }
    __syncthreads();
    // ancestry rows: Anxt[q'][0..t) = Acur[src[q']][0..t) ; Anxt[q'][t] = src[q']   (8 slots = one uint4; 4 copies in flight)
    const int nch = (t >> 3) + 1;
    const int total = P * nch;
    for (int i0 = tid; i0 < total; i0 += 4 * F2NT) {
      uint4 v[4];
      int sqv[4], qv[4], cv[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int i = i0 + k * F2NT;
        const bool ok = i < total;
        const int qn = ok ? i / nch : 0;
        qv[k] = qn; cv[k] = ok ? i - qn * nch : 0; sqv[k] = rx.src[qn];
        v[k] = *reinterpret_cast<const uint4*>(Acur + (long long)sqv[k] * Sa + 8 * cv[k]);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (i0 + k * F2NT >= total) continue;
        uint4 x = v[k];
        if ((t >> 3) == cv[k]) {   // patch slot t (little-endian: element e is half (e&1) of word (e>>1))
          const int e = t & 7;
          const unsigned int sq = (unsigned int)sqv[k];
          unsigned int wv = (e >> 1) == 0 ? x.x : ((e >> 1) == 1 ? x.y : ((e >> 1) == 2 ? x.z : x.w));
          wv = (e & 1) ? ((wv & 0x0000FFFFu) | (sq << 16)) : ((wv & 0xFFFF0000u) | sq);
          if ((e >> 1) == 0) x.x = wv; else if ((e >> 1) == 1) x.y = wv; else if ((e >> 1) == 2) x.z = wv; else x.w = wv;
        }
        *reinterpret_cast<uint4*>(Anxt + (long long)qv[k] * Sa + 8 * cv[k]) = x;
      }
    }
    unsigned short* tmp = Acur; Acur = Anxt; Anxt = tmp;
    cur_is_A0 ^= 1;
  } else {
    for (int qn = tid; qn < P; qn += F2NT) Acur[(long long)qn * Sa + t] = (unsigned short)qn;
  }
}

}  // namespace f2

// INJECT: eps are injected (tests) instead of drawn; MULTI: per-dimension log-variances; CHECKOUT: noise_z = z - z_
// (self_attention_SMC.py:190).  Compile-time variants keep the instruction footprint of the hot loop small (the kernel
// is sensitive to instruction-cache misses: its warps run different phases at the same time).
template <bool INJECT, bool MULTI, bool CHECKOUT>
__global__ void __launch_bounds__(F2NT, 2) fused2_forward_kernel(const FusedArgs a) {
  extern __shared__ unsigned char smem_raw_f2[];
  Fused2Smem& sm = *reinterpret_cast<Fused2Smem*>(smem_raw_f2 + ((128u - (tc::smem_u32(smem_raw_f2) & 127u)) & 127u));
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
  const int r8 = lane >> 2, c4 = lane & 3;        // attention / noise: row-in-octet, dims 16 c4 .. 16 c4 + 15
  const int equad = wid & 3, ehalf = wid >> 2;    // epilogue: TMEM lane quadrant, column half
  const int er = equad * 32 + lane;               // epilogue: row of the tile (= TMEM lane)
  const int n0 = 32 * ehalf;                      // epilogue: columns [n0, n0 + 32)
  const uint32_t tlane = ((uint32_t)(equad * 32)) << 16;
  const uint32_t idesc = tc::make_idesc(128, 64);
  const int nch = a.dff / FD;                     // FFN chunks of 64 hidden units (1..4)
  constexpr bool multi = MULTI;
  const unsigned char* const hdr_g = a.wtc + (size_t)(1 + 2 * nch) * 2 * tc::B_PART_BYTES;
  const float* const sigtab = reinterpret_cast<const float*>(hdr_g + 256);   // [8][64]: std k,q,v,z ; 1/var k,q,v,z
  double dloss_q = 0.0, dloss_z = 0.0;            // thread 0 only
  uint32_t ph_ready = 0, ph_full0 = 0, ph_full1 = 0;   // phase parities tracked by warp 0
  uint32_t ph_z = 0, ph_h0 = 0, ph_h1 = 0, ph_r = 0;   // phase parities of the MMA completion barriers (all threads)

  for (int i = tid; i < P; i += F2NT) { sm.Lq[i] = (unsigned short)i; a.pos_out[bP + i] = (unsigned short)i; }
  for (int i = tid; i < FD; i += F2NT) {
    sm.bz[i] = a.p.bz[i]; sm.b2[i] = a.p.b2[i];
    sm.ln1g[i] = a.p.ln1_g[i]; sm.ln1b[i] = a.p.ln1_b[i]; sm.ln2g[i] = a.p.ln2_g[i]; sm.ln2b[i] = a.p.ln2_b[i];
  }
  // b1 is kept pre-multiplied by the operand scale s_h (a power of two, so relu(x a + b) s_h == relu(x (a s_h) + b s_h) bit for bit)
  for (int i = tid; i < a.dff; i += F2NT) sm.b1[i] = a.p.b1[i] * reinterpret_cast<const TcHeader*>(hdr_g)->s_h;
  for (int i = tid; i < FD * Fy; i += F2NT) sm.Wo[i] = a.p.Wo[i];
  if (tid < 8) sm.sig[tid] = sigtab[(tid < 4 ? tid : (tid < 6 ? 4 + 2 * (tid - 4) + 1 : 0)) * FD];   // std k,q,v,z ; ivar q, ivar z
  if (tid == 0) sm.hdr = *reinterpret_cast<const TcHeader*>(hdr_g);
  if (wid == 0) tc::tmem_alloc(&sm.tmem_slot, F2TMEM);
  if (tid == 0) {
    tc::mbar_init(&sm.mbar_ready, F2NW);
    tc::mbar_init(&sm.mbar_full[0], 1); tc::mbar_init(&sm.mbar_full[1], 1);
    tc::mbar_init(&sm.mbar_z, 1); tc::mbar_init(&sm.mbar_h[0], 1); tc::mbar_init(&sm.mbar_h[1], 1); tc::mbar_init(&sm.mbar_r, 1);
    tc::fence_mbar_init();
  }
  if (tid < 4 * FD) {
    const int which = tid >> 6, d = tid & 63;
    const float* src = which == 0 ? a.xln : (which == 1 ? a.q_ : (which == 2 ? a.k_ : a.v_));
    sm.rowbuf[which][d] = src[((long long)b * S) * FD + d];
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = sm.tmem_slot;
  const uint32_t tD_z = tmem, tD_h0 = tmem + 64, tD_r = tmem + 192;   // z | h even chunk | h odd chunk | r
  Fused2Scratch& rsx = *reinterpret_cast<Fused2Scratch*>(&sm.W[0][0]);
  const unsigned char* const wblk = a.wtc;
  constexpr uint32_t WB = 2 * tc::B_PART_BYTES;   // bytes of one weight block (h and m parts)
  const float s_q = sm.hdr.s_q, s_k = sm.hdr.s_k, s_v = sm.hdr.s_v;

  for (int t = 0; t < S; ++t) {
    const int s_lo = (a.attn_window >= 0 && t > a.attn_window) ? t - a.attn_window : 0;
    float loss_q = 0.f, loss_z = 0.f;

    for (int g = 0; g < ntiles; ++g) {
      const int q0 = g * FTILE;
      const int np = min(FTILE, P - q0);
      // ---- weight blocks for the start of this tile: Wz -> W[1], W1 chunk 0 -> W[0] (both buffers are free: every MMA of the
      //      previous tile has completed; after the draw their bytes were touched through the generic proxy) ----
      if (wid == 0) {
        if (tc::elect_one()) {
          tc::fence_async_smem();
          f2::bulk_load(sm.W[1], wblk, WB, &sm.mbar_full[1]);
          f2::bulk_load(sm.W[0], wblk + (size_t)1 * WB, WB, &sm.mbar_full[0]);
        }
        __syncwarp();
      }

      // =========================== phase A (warp-private): noise, slot-t rows, Q fragments, attention ===========================
      const bool warp_act = 16 * wid < np;
      if (warp_act) {
        uint4* Qf = reinterpret_cast<uint4*>(sm.Ob) + wid * 256;   // this warp's Q fragments
        // rolled loops (two particles x four groups of 4 dims): one copy of the generator in the instruction stream
#pragma unroll 1
        for (int pp = 0; pp < 2; ++pp) {
          const int pl = 16 * wid + r8 + 8 * pp;       // particle of the tile
          const bool act = pl < np;
          const int qc = act ? (q0 + pl) : (P - 1);
          const int L = sm.Lq[qc];
          const unsigned long long gbp = (unsigned long long)(a.b_global0 + b) * (unsigned long long)P + L;
          const long long lrow = ((bP + L) * S + t) * FD;
          const uint32_t prow = (uint32_t)(t * P + qc) * FD;
          unsigned char* const ksr = Ksb + ((size_t)(t * P + qc) << 8);   // fp16 (h,m) split copies of the new rows in the
          unsigned char* const vsr = Vsb + ((size_t)(t * P + qc) << 8);   // MMA operand order (smct_fused_attn.cuh)
#pragma unroll 1
          for (int i = 0; i < 4; ++i) {
            const int cg = 4 * c4 + i, d0 = 4 * cg;      // this lane owns dims [16 c4, 16 c4 + 16)
            const float4 q_0 = *reinterpret_cast<const float4*>(&sm.rowbuf[1][d0]);
            const float4 k_0 = *reinterpret_cast<const float4*>(&sm.rowbuf[2][d0]);
            const float4 v_0 = *reinterpret_cast<const float4*>(&sm.rowbuf[3][d0]);
            float ek[4], eq[4], ev[4];
            if (INJECT) {
              const long long BPD = (long long)a.B * P * FD;
              const long long e0 = ((long long)t * a.B + b) * P * FD + (long long)L * FD + d0;
              const float4 t0 = *reinterpret_cast<const float4*>(a.eps + (0LL * S) * BPD + e0);
              const float4 t1 = *reinterpret_cast<const float4*>(a.eps + (1LL * S) * BPD + e0);
              const float4 t2 = *reinterpret_cast<const float4*>(a.eps + (2LL * S) * BPD + e0);
              ek[0] = t0.x; ek[1] = t0.y; ek[2] = t0.z; ek[3] = t0.w;
              eq[0] = t1.x; eq[1] = t1.y; eq[2] = t1.z; eq[3] = t1.w;
              ev[0] = t2.x; ev[1] = t2.y; ev[2] = t2.z; ev[3] = t2.w;
            } else {
              philox_normal4(a.seed, 0u, (uint32_t)t, gbp * 16ull + cg, ek);
              philox_normal4(a.seed, 1u, (uint32_t)t, gbp * 16ull + cg, eq);
              philox_normal4(a.seed, 2u, (uint32_t)t, gbp * 16ull + cg, ev);
            }
            const float kin[4] = {k_0.x, k_0.y, k_0.z, k_0.w};
            const float qin[4] = {q_0.x, q_0.y, q_0.z, q_0.w};
            const float vin[4] = {v_0.x, v_0.y, v_0.z, v_0.w};
            float sdk[4], sdq[4], sdv[4], ivq[4];
            if (multi) {
              ld4(sigtab + 0 * FD + d0, sdk); ld4(sigtab + 1 * FD + d0, sdq); ld4(sigtab + 2 * FD + d0, sdv); ld4(sigtab + 5 * FD + d0, ivq);
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e) { sdk[e] = sm.sig[0]; sdq[e] = sm.sig[1]; sdv[e] = sm.sig[2]; ivq[e] = sm.sig[4]; }
            }
            float ko[4], vo[4], qs[4], nq[4], xk[4], xv[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              ko[e] = __fadd_rn(kin[e], __fmul_rn(sdk[e], ek[e]));
              const float qo = __fadd_rn(qin[e], __fmul_rn(sdq[e], eq[e]));
              vo[e] = __fadd_rn(vin[e], __fmul_rn(sdv[e], ev[e]));
              nq[e] = qo - qin[e];
              loss_q = act ? fmaf(nq[e] * nq[e], ivq[e], loss_q) : loss_q;
              qs[e] = attn::sat16(qo * s_q);   // scaled by log2(e)/sqrt(d_k) (scores in the log2 domain) and the fp16 scale
              xk[e] = attn::sat16(ko[e] * s_k);
              xv[e] = attn::sat16(vo[e] * s_v);
            }
            if (act) {
              // fp32 rows: read again only by the final pass
              __stcs(reinterpret_cast<float4*>(Kb + prow + d0), make_float4(ko[0], ko[1], ko[2], ko[3]));
              __stcs(reinterpret_cast<float4*>(Vb + prow + d0), make_float4(vo[0], vo[1], vo[2], vo[3]));
              if (a.noise_q) __stcs(reinterpret_cast<float4*>(a.noise_q + lrow + d0), make_float4(nq[0], nq[1], nq[2], nq[3]));
              // split rows: these 4 dims are one half (8 bytes) of the 16-byte unit of dims 16 c4 + 8 (i>>1) .. + 7
              uint2 h, m;
              attn::split2(xk[0], xk[1], h.x, m.x);
              attn::split2(xk[2], xk[3], h.y, m.y);
              *reinterpret_cast<uint2*>(ksr + 16 * ((i >> 1) * 4 + c4) + 8 * (i & 1)) = h;
              *reinterpret_cast<uint2*>(ksr + 16 * (8 + (i >> 1) * 4 + c4) + 8 * (i & 1)) = m;
              attn::split2(xv[0], xv[1], h.x, m.x);
              attn::split2(xv[2], xv[3], h.y, m.y);
              *reinterpret_cast<uint2*>(vsr + 32 * c4 + 8 * i) = h;
              *reinterpret_cast<uint2*>(vsr + 128 + 32 * c4 + 8 * i) = m;
            }
            // Q fragments of the warp's 16 x 64 tile (rows r8 and r8 + 8): k-step i covers dims 16 c4 + 4 i .. + 3 of this
            // lane.  The first particle parks its four values in the m slot; the second one builds both units.
            float4* const park = reinterpret_cast<float4*>(&Qf[(4 + i) * 32 + lane]);
            if (pp == 0) {
              *park = make_float4(qs[0], qs[1], qs[2], qs[3]);
            } else {
              const float4 qa = *park;
              uint4 qh, qm;
              attn::split2(qa.x, qa.y, qh.x, qm.x);
              attn::split2(qs[0], qs[1], qh.y, qm.y);
              attn::split2(qa.z, qa.w, qh.z, qm.z);
              attn::split2(qs[2], qs[3], qh.w, qm.w);
              Qf[i * 32 + lane] = qh;
              Qf[(4 + i) * 32 + lane] = qm;
            }
          }
        }
        __syncwarp();   // Q fragments and the slot-t rows of the warp's particles are visible to the whole warp

        // ---- attention over the genealogy: columns = distinct (slot, row) pairs of the warp's 16 particles ----
        attn::State st;
#pragma unroll
        for (int i = 0; i < 8; ++i) st.Z[i][0] = st.Z[i][1] = st.Z[i][2] = st.Z[i][3] = 0.f;
        st.m0 = st.m1 = -INFINITY; st.l0 = st.l1 = 0.f;
        const int l15 = lane & 15, hsl = lane >> 4;
        const int myq = min(q0 + 16 * wid + l15, P - 1);
        const unsigned short* Arow = Acur + (long long)myq * Sa;
        uint2* cd = sm.cdesc[wid];
        const float inv_qk = sm.hdr.inv_qk;
        const uint32_t ltmask = (1u << lane) - 1u;
        int ncols = 0;
        // A particle opens a column when its ancestor differs from its left neighbour's (the ancestry is monotone in the
        // position).  Old slots are usually shared by all 16 particles: a unit of 4 slots whose ancestry words are identical
        // across the warp adds its 4 columns at once; otherwise the two half-warps look at two consecutive slots per pass.
        const int u_first = s_lo >> 2;
        uint2 av_next = make_uint2(0u, 0u);
        if (4 * u_first < t) av_next = *reinterpret_cast<const uint2*>(Arow + 4 * u_first);
        for (int u = u_first; u <= (t >> 2); ++u) {
          const uint2 av = av_next;
          av_next = make_uint2(0u, 0u);
          if (4 * (u + 1) < t) av_next = *reinterpret_cast<const uint2*>(Arow + 4 * (u + 1));   // requested one unit ahead
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

        // ---- z_ = Z / l : rows R0, R1 of the tile, dims 16 c4 .. + 15 -> scaled fp16 (h,m) operand rows ----
        const int R0 = 16 * wid + r8, R1 = R0 + 8;
        const float il0 = sm.hdr.inv_v / st.l0, il1 = sm.hdr.inv_v / st.l1;
        float z0[16], z1[16];
        float rmax0 = 0.f, rmax1 = 0.f;
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {   // dims 16 c4 + 4 k4 .. + 3 : Z[i][0] = dim 16 c4 + i, Z[i][1] = dim 16 c4 + 8 + i
          const int i0 = (k4 & 1) * 4, e = k4 >> 1;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            z0[4 * k4 + k] = st.Z[i0 + k][e] * il0;
            z1[4 * k4 + k] = st.Z[i0 + k][2 + e] * il1;
            rmax0 = fmaxf(rmax0, fabsf(z0[4 * k4 + k])); rmax1 = fmaxf(rmax1, fabsf(z1[4 * k4 + k]));
          }
        }
        rmax0 = fmaxf(rmax0, __shfl_xor_sync(SMCT_FULL_MASK, rmax0, 1)); rmax0 = fmaxf(rmax0, __shfl_xor_sync(SMCT_FULL_MASK, rmax0, 2));
        rmax1 = fmaxf(rmax1, __shfl_xor_sync(SMCT_FULL_MASK, rmax1, 1)); rmax1 = fmaxf(rmax1, __shfl_xor_sync(SMCT_FULL_MASK, rmax1, 2));
        const int ez0 = tc::scale_exponent(rmax0), ez1 = tc::scale_exponent(rmax1);   // exact power-of-two scales of the rows' fp16 operand
        const float zs0 = ldexpf(1.f, ez0), zs1 = ldexpf(1.f, ez1);
        if (c4 == 0) { sm.rs[R0] = ldexpf(1.f, -ez0); sm.rs[R1] = ldexpf(1.f, -ez1); }
#pragma unroll
        for (int k8 = 0; k8 < 2; ++k8) {
          float zz0[8], zz1[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) { zz0[e] = z0[8 * k8 + e] * zs0; zz1[e] = z1[8 * k8 + e] * zs1; }
          tc::store_a8(sm.ZH, R0, 16 * c4 + 8 * k8, zz0);
          tc::store_a8(sm.ZH, R1, 16 * c4 + 8 * k8, zz1);
        }
      } else {   // no particle of the tile in this warp's 16 rows: zero operand rows
        const int R0 = 16 * wid + r8, R1 = R0 + 8;
        const float zz[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (c4 == 0) { sm.rs[R0] = 1.f; sm.rs[R1] = 1.f; }
#pragma unroll
        for (int k8 = 0; k8 < 2; ++k8) {
          tc::store_a8(sm.ZH, R0, 16 * c4 + 8 * k8, zz);
          tc::store_a8(sm.ZH, R1, 16 * c4 + 8 * k8, zz);
        }
      }
      f2::warp_rows_ready(&sm.mbar_ready, lane);   // ready #1: the z_ operand rows (and rs) of this warp

      // =========================== phase B: dense tile on the tensor cores ===========================
      const bool eact = er < np;
      const int eL = sm.Lq[eact ? (q0 + er) : (P - 1)];
      float ezv[32];   // z noise of this thread's 32 columns: generated while the other warps finish / the tensor core works
      {
        const unsigned long long gbp = (unsigned long long)(a.b_global0 + b) * (unsigned long long)P + eL;
#pragma unroll
        for (int gq = 0; gq < 8; ++gq) {
          const int n = n0 + 4 * gq;
          if (INJECT) {
            const float4 t3 = *reinterpret_cast<const float4*>(a.eps + (3LL * S) * ((long long)a.B * P * FD) +
                                                               ((long long)t * a.B + b) * P * FD + (long long)eL * FD + n);
            ezv[4 * gq] = t3.x; ezv[4 * gq + 1] = t3.y; ezv[4 * gq + 2] = t3.z; ezv[4 * gq + 3] = t3.w;
          } else {
            const unsigned long long grp = gbp * 16ull + (n >> 2);
            const float4 nz = philox_normal4_call((uint32_t)a.seed, (uint32_t)(a.seed >> 32), 3u, (uint32_t)t, (uint32_t)grp, (uint32_t)(grp >> 32));
            ezv[4 * gq] = nz.x; ezv[4 * gq + 1] = nz.y; ezv[4 * gq + 2] = nz.z; ezv[4 * gq + 3] = nz.w;
          }
        }
      }
      if (wid == 0) {
        if (tc::elect_one()) {
          tc::mbar_wait(&sm.mbar_ready, ph_ready);
          tc::fence_after_sync();
          tc::mbar_wait(&sm.mbar_full[1], ph_full1);   // Wz
          tc::mma_block_f16x3(tD_z, tc::smem_u32(sm.ZH), tc::smem_u32(sm.W[1]), tc::B_PART_BYTES, idesc, false);
          tc::commit(&sm.mbar_z);
        }
        __syncwarp();
        ph_ready ^= 1; ph_full1 ^= 1;
      }
      tc::mbar_wait(&sm.mbar_z, ph_z);
      ph_z ^= 1;
      tc::fence_after_sync();
      if (wid == 0) {   // Wz has been consumed: W2 chunk 0 travels into its buffer during the LN1 epilogue
        if (tc::elect_one()) f2::bulk_load(sm.W[1], wblk + (size_t)2 * WB, WB, &sm.mbar_full[1]);
        __syncwarp();
      }
      float val[32];
      {  // ---- z = dense(z_) + noise ; o = LN1(z + x) -> A operand ----
        tc::tmem_ld32(tD_z + tlane + n0, val);
        const float rsr = sm.rs[er];
        const float zscale = rsr * sm.hdr.inv_wz;
        float s1 = 0.f;
#pragma unroll
        for (int gq = 0; gq < 8; ++gq) {
          const int n = n0 + 4 * gq;
          float z_[4] = {0.f, 0.f, 0.f, 0.f};
          if (CHECKOUT) {   // z_ back from its operand rows (22 bits)
            const uint32_t off = tc::operand_offset(er, n);
            const uint2 h = *reinterpret_cast<const uint2*>(sm.ZH + off), m = *reinterpret_cast<const uint2*>(sm.ZH + tc::A_PART_BYTES + off);
            const float2 h0 = __half22float2(*reinterpret_cast<const __half2*>(&h.x)), h1 = __half22float2(*reinterpret_cast<const __half2*>(&h.y));
            const float2 m0 = __half22float2(*reinterpret_cast<const __half2*>(&m.x)), m1 = __half22float2(*reinterpret_cast<const __half2*>(&m.y));
            z_[0] = (h0.x + m0.x) * rsr; z_[1] = (h0.y + m0.y) * rsr; z_[2] = (h1.x + m1.x) * rsr; z_[3] = (h1.y + m1.y) * rsr;
          }
          float nzv[4], cbz[4], csd[4], civ[4], cx[4];
          ld4(&sm.bz[n], cbz); ld4(&sm.rowbuf[0][n], cx);
          if (multi) { ld4(sigtab + 3 * FD + n, csd); ld4(sigtab + 7 * FD + n, civ); }
          else {
#pragma unroll
            for (int e = 0; e < 4; ++e) { csd[e] = sm.sig[3]; civ[e] = sm.sig[5]; }
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float zp = fmaf(val[4 * gq + e], zscale, cbz[e]);   // scale is a power of two: exact
            const float z = __fadd_rn(zp, __fmul_rn(csd[e], ezv[4 * gq + e]));
            nzv[e] = CHECKOUT ? (z - z_[e]) : (z - zp);
            loss_z = eact ? fmaf(nzv[e] * nzv[e], civ[e], loss_z) : loss_z;
            val[4 * gq + e] = z + cx[e];
            s1 += val[4 * gq + e];
          }
          if (eact && a.noise_z)
            __stcs(reinterpret_cast<float4*>(a.noise_z + ((bP + eL) * S + t) * FD + n), make_float4(nzv[0], nzv[1], nzv[2], nzv[3]));
        }
        // LayerNorm statistics of the 64-wide row from the two 32-wide halves (Chan et al.: mean and M2 combine exactly)
        const float mh = s1 * (1.0f / 32.0f);
        float m2h = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) { const float cdev = val[j] - mh; m2h = fmaf(cdev, cdev, m2h); }
        sm.xch[ehalf][er] = make_float4(mh, m2h, 0.f, 0.f);
        f2::pair_sync(equad);
        const float4 po = sm.xch[ehalf ^ 1][er];   // the other half's partials (own ones stay in registers)
        const float mA = ehalf ? po.x : mh, mB = ehalf ? mh : po.x, qA = ehalf ? po.y : m2h, qB = ehalf ? m2h : po.y;
        const float mean = 0.5f * (mA + mB);
        const float dm = mB - mA;
        const float var = ((qA + qB) + 16.0f * dm * dm) * (1.0f / FD);
        const float rstd = 1.0f / sqrtf(var + a.ln_eps);
#pragma unroll
        for (int gq = 0; gq < 8; ++gq) {
          const int n = n0 + 4 * gq;
          float cg[4], cb[4];
          ld4(&sm.ln1g[n], cg); ld4(&sm.ln1b[n], cb);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float inv = rstd * cg[e];
            val[4 * gq + e] = (val[4 * gq + e] * inv + (cb[e] - mean * inv)) * sm.hdr.s_o;
          }
        }
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) tc::store_a8(sm.Ob, er, n0 + 8 * g8, &val[8 * g8]);
      }
      f2::warp_rows_ready(&sm.mbar_ready, lane);   // ready #2: o operand rows
      if (wid == 0) {
        if (tc::elect_one()) {
          tc::mbar_wait(&sm.mbar_ready, ph_ready);
          tc::fence_after_sync();
          tc::mbar_wait(&sm.mbar_full[0], ph_full0);   // W1 chunk 0
          tc::mma_block_f16x3(tD_h0, tc::smem_u32(sm.Ob), tc::smem_u32(sm.W[0]), tc::B_PART_BYTES, idesc, false);
          tc::commit(&sm.mbar_h[0]);
        }
        __syncwarp();
        ph_ready ^= 1; ph_full0 ^= 1;
      }
      // ---- FFN: h_c = relu(o W1[:,c] + b1) -> A operand -> r += h_c W2[c,:] ; the first GEMM of chunk c+1 is issued right
      //      behind the second GEMM of chunk c (TMEM ranges h even / h odd), so it overlaps the epilogue of chunk c+1's wait ----
      for (int c = 0; c < nch; ++c) {
        if (c & 1) { tc::mbar_wait(&sm.mbar_h[1], ph_h1); ph_h1 ^= 1; }
        else { tc::mbar_wait(&sm.mbar_h[0], ph_h0); ph_h0 ^= 1; }
        tc::fence_after_sync();
        // h_c complete => every earlier MMA complete: W[0] (W1 chunk c) and, for c > 0, W[1] (W2 chunk c-1) and ZH are free
        if (wid == 0) {
          if (tc::elect_one()) {
            if (c + 1 < nch) f2::bulk_load(sm.W[0], wblk + (size_t)(1 + 2 * (c + 1)) * WB, WB, &sm.mbar_full[0]);
            if (c > 0) f2::bulk_load(sm.W[1], wblk + (size_t)(2 + 2 * c) * WB, WB, &sm.mbar_full[1]);
          }
          __syncwarp();
        }
        tc::tmem_ld32(tD_h0 + 64 * (c & 1) + tlane + n0, val);
        const float hscale = sm.hdr.inv_o * sm.hdr.inv_w1 * sm.hdr.s_h;   // (sm.b1 holds b1 s_h)
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          float cb1[4];
          ld4(&sm.b1[c * FD + n0 + 4 * j4], cb1);
#pragma unroll
          for (int e = 0; e < 4; ++e) val[4 * j4 + e] = fmaxf(fmaf(val[4 * j4 + e], hscale, cb1[e]), 0.f);
        }
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) tc::store_a8(sm.ZH, er, n0 + 8 * g8, &val[8 * g8]);
        f2::warp_rows_ready(&sm.mbar_ready, lane);   // ready #3+c: hidden operand rows of chunk c
        if (wid == 0) {
          if (tc::elect_one()) {
            tc::mbar_wait(&sm.mbar_ready, ph_ready);
            tc::fence_after_sync();
            tc::mbar_wait(&sm.mbar_full[1], ph_full1);   // W2 chunk c
            tc::mma_block_f16x3(tD_r, tc::smem_u32(sm.ZH), tc::smem_u32(sm.W[1]), tc::B_PART_BYTES, idesc, c > 0);
            if (c + 1 == nch) {
              tc::commit(&sm.mbar_r);
            } else {
              tc::mbar_wait(&sm.mbar_full[0], ph_full0);   // W1 chunk c+1
              tc::mma_block_f16x3(tD_h0 + 64 * ((c + 1) & 1), tc::smem_u32(sm.Ob), tc::smem_u32(sm.W[0]), tc::B_PART_BYTES, idesc, false);
              tc::commit(&sm.mbar_h[(c + 1) & 1]);
            }
          }
          __syncwarp();
          ph_ready ^= 1; ph_full1 ^= 1;
          if (c + 1 < nch) ph_full0 ^= 1;
        }
      }
      tc::mbar_wait(&sm.mbar_r, ph_r);
      ph_r ^= 1;
      tc::fence_after_sync();
      {  // ---- r = LN2(ffn + b2 + o) ; R row store ; output layer ; Gaussian log-weight ----
        tc::tmem_ld32(tD_r + tlane + n0, val);
        const float rscale = sm.hdr.inv_h * sm.hdr.inv_w2;
        const float inv_o = sm.hdr.inv_o;
        float s1 = 0.f;
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) {
          float o8[8];
          f2::load_a8(sm.Ob, er, n0 + 8 * g8, o8);   // o back from its operand rows (22 bits)
          float cb2[8];
          ld4(&sm.b2[n0 + 8 * g8], cb2); ld4(&sm.b2[n0 + 8 * g8 + 4], cb2 + 4);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            val[8 * g8 + e] = fmaf(val[8 * g8 + e], rscale, cb2[e]) + o8[e] * inv_o;
            s1 += val[8 * g8 + e];
          }
        }
        const float mh = s1 * (1.0f / 32.0f);
        float m2h = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) { const float cdev = val[j] - mh; m2h = fmaf(cdev, cdev, m2h); }
        sm.xch[ehalf][er] = make_float4(mh, m2h, 0.f, 0.f);
        f2::pair_sync(equad);
        const float4 po = sm.xch[ehalf ^ 1][er];   // the other half's partials (own ones stay in registers)
        const float mA = ehalf ? po.x : mh, mB = ehalf ? mh : po.x, qA = ehalf ? po.y : m2h, qB = ehalf ? m2h : po.y;
        const float mean = 0.5f * (mA + mB);
        const float dm = mB - mA;
        const float var = ((qA + qB) + 16.0f * dm * dm) * (1.0f / FD);
        const float rstd = 1.0f / sqrtf(var + a.ln_eps);
        float* rrow = Rb + (uint32_t)(t * P + (eact ? q0 + er : 0)) * FD;
#pragma unroll
        for (int gq = 0; gq < 8; ++gq) {
          const int n = n0 + 4 * gq;
          float cg[4], cb[4];
          ld4(&sm.ln2g[n], cg); ld4(&sm.ln2b[n], cb);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float inv = rstd * cg[e];
            val[4 * gq + e] = val[4 * gq + e] * inv + (cb[e] - mean * inv);
          }
          if (eact)   // read again only by the final pass
            __stcs(reinterpret_cast<float4*>(rrow + n), make_float4(val[4 * gq], val[4 * gq + 1], val[4 * gq + 2], val[4 * gq + 3]));
        }
        // output layer: each column half forms its partial dot products and parks them in the exchange slot it has just
        // read (xch[other half][er]: nobody else reads that slot any more); the lower half then adds the two in a fixed order
        float* const park = reinterpret_cast<float*>(&sm.xch[ehalf ^ 1][er]);
#pragma unroll 1
        for (int f = 0; f < Fy; ++f) {
          float acc = 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) acc = fmaf(val[j], sm.Wo[(n0 + j) * Fy + f], acc);
          park[f] = acc;
        }
        f2::pair_sync(equad);
        if (ehalf == 0) {
          const float* const lo = reinterpret_cast<const float*>(&sm.xch[1][er]);   // own partials (columns 0..31)
          const float* const hi = reinterpret_cast<const float*>(&sm.xch[0][er]);   // the upper half's (columns 32..63)
          float sq = 0.f;
#pragma unroll 1
          for (int f = 0; f < Fy; ++f) {
            const float pr = lo[f] + hi[f];
            if (eact && a.pred) a.pred[((bP + eL) * S + t) * Fy + f] = pr;
            const float diff = a.y[((long long)b * S + t) * Fy + f] - pr;
            sq = fmaf(diff, diff, sq);
          }
          if (eact) sm.lw[eL] = (-0.5f * sq) / a.sigma_obs;
        }
      }
      tc::fence_before_sync();   // the TMEM loads of this tile are ordered before the next tile's MMAs (via the ready arrivals)
    }  // tiles

    // =========================== end of the timestep: loss partials, draw, genealogy update ===========================
    __syncthreads();   // all tiles done: lw complete, operand / weight buffers idle
    Fused2Scratch& rx = rsx;
    if (a.loss_sums) {
      float* red = reinterpret_cast<float*>(&sm.xch[0][0]);
      const float sq = warp_sum(loss_q), sz = warp_sum(loss_z);
      if (lane == 0) { red[64 + wid] = sq; red[64 + F2NW + wid] = sz; }
      __syncthreads();
      if (tid == 0) {
        for (int i = 0; i < F2NW; ++i) { dloss_q += (double)red[64 + i]; dloss_z += (double)red[64 + F2NW + i]; }
      }
    }
    f2::resample_update(sm, rx, a, b, t, Acur, Anxt, cur_is_A0);
    if (t + 1 < S && tid < 4 * FD) {
      const int which = tid >> 6, d = tid & 63;
      const float* src = which == 0 ? a.xln : (which == 1 ? a.q_ : (which == 2 ? a.k_ : a.v_));
      sm.rowbuf[which][d] = src[((long long)b * S + t + 1) * FD + d];
    }
    __syncthreads();
  }  // timesteps

  if (tid == 0) {
    a.aflag[b] = cur_is_A0;
    if (a.loss_sums) { atomicAdd(a.loss_sums + 1, dloss_q); atomicAdd(a.loss_sums + 3, dloss_z); }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (wid == 0) tc::tmem_dealloc(tmem, F2TMEM);
}

inline bool fused2_supported(const smct_shape& s, const smct_io& io) {
  return fused_supported(s, io) && s.F_y <= F2FY;
}

}  // namespace smct

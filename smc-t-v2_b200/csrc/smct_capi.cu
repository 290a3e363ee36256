// smct_capi.cu — extern "C" entry points of libsmct_b200.so (declared in include/smct.h).
// Host-side argument checking, workspace carving and kernel launches.  No CPU fallback.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <algorithm>
#include <atomic>

#include "../../include/smct.h"
#include "smct_common.cuh"
#include "smct_staged.cuh"
#include "smct_small.cuh"
#include "smct_fused.cuh"
#include "smct_materialize.cuh"
#include "smct_tc.cuh"
#include "smct_fused_tc.cuh"
#include "smct_fused2.cuh"
#include "smct_backward.cuh"

#ifndef SMCT_SOURCE_HASH   // set by smc-t-v2_b200/build.py: sha256 over csrc/ and include/smct.h
#define SMCT_SOURCE_HASH "unhashed-build------------------"
#endif

namespace {

thread_local std::string g_last_error;
std::atomic<long long> g_launches{0};  // kernels launched by this library (bench.py's gpu_launches)

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

#define SMCT_CUDA(expr)                                                                         \
  do {                                                                                          \
    cudaError_t e__ = (expr);                                                                   \
    if (e__ != cudaSuccess)                                                                     \
      return fail(SMCT_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

#define SMCT_LAUNCH_CHECK(name)                                                                  \
  do {                                                                                          \
    g_launches.fetch_add(1, std::memory_order_relaxed);                                         \
    cudaError_t e__ = cudaGetLastError();                                                       \
    if (e__ != cudaSuccess)                                                                     \
      return fail(SMCT_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(e__));     \
  } while (0)

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

struct Carver {
  char* base; size_t off, cap;
  template <typename T> T* take(size_t n) {
    T* p = (T*)(base + off);
    off = align_up(off + n * sizeof(T));
    return p;
  }
};

// cudaFuncSetAttribute applies to the CURRENT device: the "already done" flag of a launch site is kept per device so
// that a process driving several GPUs configures each of them (one flag word per site, bit = device ordinal).
struct PerDeviceOnce {
  std::atomic<unsigned long long> done{0};
  bool need() {
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    return (done.fetch_or(bit) & bit) == 0;
  }
};

int check_shape(const smct_shape* s) {
  if (!s) return fail(SMCT_ERR_INVALID, "shape is NULL");
  if (s->struct_bytes != (int32_t)sizeof(smct_shape))
    return fail(SMCT_ERR_INVALID, "smct_shape.struct_bytes=%d, library expects %d (ABI %d)", s->struct_bytes,
                (int)sizeof(smct_shape), SMCT_ABI_VERSION);
  if (s->B < 1 || s->P < 1 || s->S < 1 || s->D < 1 || s->H < 1 || s->F_y < 1)
    return fail(SMCT_ERR_INVALID, "B,P,S,D,H,F_y must be >= 1 (got %d,%d,%d,%d,%d,%d)", s->B, s->P, s->S, s->D, s->H, s->F_y);
  if (s->D % s->H) return fail(SMCT_ERR_INVALID, "d_model %d not divisible by num_heads %d", s->D, s->H);
  if (s->D > 256) return fail(SMCT_ERR_UNSUPPORTED, "d_model %d > 256 unsupported", s->D);
  if (s->P > 16384) return fail(SMCT_ERR_UNSUPPORTED, "particles %d > 16384 unsupported", s->P);
  if (s->full_model && s->dff < 1) return fail(SMCT_ERR_INVALID, "full_model needs dff >= 1");
  if (s->logvar_dim != 1 && s->logvar_dim != s->D) return fail(SMCT_ERR_INVALID, "logvar_dim must be 1 or D");
  if (s->weights_mode != SMCT_WEIGHTS_GAUSSIAN && s->weights_mode != SMCT_WEIGHTS_CLASSIFICATION)
    return fail(SMCT_ERR_INVALID, "bad weights_mode %d", s->weights_mode);
  if (s->input_mode < 0 || s->input_mode > 2) return fail(SMCT_ERR_INVALID, "bad input_mode %d", s->input_mode);
  if (s->weights_mode == SMCT_WEIGHTS_GAUSSIAN && !(s->sigma_obs > 0.f))
    return fail(SMCT_ERR_INVALID, "sigma_obs must be > 0");
  return SMCT_OK;
}

size_t step_smem_bytes(int D, int dff, int full_model) {
  const size_t Dp = D + 1, Fp = dff + 1;
  return (2 * smct::PT * Dp + (full_model ? smct::PT * Fp : 0) + smct::NW_MAX * 3 * (size_t)D + 2 * smct::NW_MAX) * sizeof(float);
}

template <int KD>
int launch_step_t(const smct::StepArgs& a, cudaStream_t st) {
  static PerDeviceOnce attr_once;
  const size_t smem = step_smem_bytes(a.D, a.dff, a.full_model);
  if (smem > 200 * 1024) return fail(SMCT_ERR_UNSUPPORTED, "step tile needs %zu B shared memory (D=%d dff=%d)", smem, a.D, a.dff);
  if (attr_once.need()) {
    SMCT_CUDA(cudaFuncSetAttribute(smct::step_kernel<KD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  dim3 grid((a.P + smct::PT - 1) / smct::PT, a.B);
  smct::step_kernel<KD><<<grid, KD <= 2 ? smct::NT_WIDE : smct::NT, smem, st>>>(a);
  SMCT_LAUNCH_CHECK("step_kernel");
  return SMCT_OK;
}

int launch_step(const smct::StepArgs& a, cudaStream_t st) {
  if (a.D <= 32) return launch_step_t<1>(a, st);
  if (a.D <= 64) return launch_step_t<2>(a, st);
  if (a.D <= 128) return launch_step_t<4>(a, st);
  return launch_step_t<8>(a, st);
}

int launch_rows(const smct::RowsArgs& a, cudaStream_t st) {
  const size_t smem = (2 * (size_t)a.D + 32) * sizeof(float);
  const int grid = (int)std::min<long long>(a.rows, 148LL * 16);
  smct::rows_kernel<<<grid, 128, smem, st>>>(a);
  SMCT_LAUNCH_CHECK("rows_kernel");
  return SMCT_OK;
}

int launch_resample(const smct::ResampleArgs& a, cudaStream_t st) {
  static PerDeviceOnce attr_once;
  const size_t smem = (size_t)a.P * sizeof(double);
  if (smem > 48 * 1024 && attr_once.need()) {
    SMCT_CUDA(cudaFuncSetAttribute(smct::resample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  }
  smct::resample_kernel<<<a.B, 256, smem, st>>>(a);
  SMCT_LAUNCH_CHECK("resample_kernel");
  return SMCT_OK;
}

int launch_gather(const smct::GatherArgs& a_in, int ntensors, cudaStream_t st) {
  smct::GatherArgs a = a_in;
  const long long per = ((a.row_elems & 3) == 0 && (a.n_elems & 3) == 0) ? a.n_elems / 4 : a.n_elems;
  // short rows (the reference's small shapes): several particle rows per CTA, so that a CTA has ~1024 copy units
  a.ppc = (int)std::max<long long>(1, std::min<long long>(64, 1024 / std::max<long long>(per, 1)));
  const long long nbp = (long long)a.B * a.P;
  int gy = (int)std::max<long long>(1, std::min<long long>(64, (per * a.ppc + 256 * 8 - 1) / (256 * 8)));
  dim3 grid((unsigned)((nbp + a.ppc - 1) / a.ppc), gy, ntensors);
  smct::gather_kernel<<<grid, 256, 0, st>>>(a);
  SMCT_LAUNCH_CHECK("gather_kernel");
  return SMCT_OK;
}

__global__ void fill_f32_kernel(float* p, long long n, float v) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}

void fill_step_common(smct::StepArgs& a, const smct_shape& s, const smct_params& pr) {
  memset(&a, 0, sizeof(a));
  a.B = s.B; a.P = s.P; a.S = s.S; a.D = s.D; a.H = s.H; a.dff = s.dff; a.Fy = s.F_y;
  a.full_model = s.full_model; a.weights_mode = s.weights_mode; a.noise_z_mode = s.noise_z_mode;
  a.noise_on = s.noise_on; a.logvar_dim = s.logvar_dim;
  a.ln_eps = s.ln_eps; a.sigma_obs = s.sigma_obs;
  a.b_global0 = s.b_global0;
  a.p = pr;
}

// ------------------------------------------------------------------------------------------
// staged forward: S x (step, resample, gather) with ping-pong state buffers
// ------------------------------------------------------------------------------------------
size_t staged_forward_ws(const smct_shape& s) {
  const size_t rows = (size_t)s.B * std::max(s.S, s.P);
  size_t n = 0;
  n += 6 * align_up(rows * s.D * sizeof(float));
  n += align_up((size_t)s.B * s.P * sizeof(float));
  n += align_up((size_t)s.B * s.P * sizeof(int));
  n += 3 * align_up((size_t)s.B * s.P * s.S * s.D * sizeof(float));
  return n + 1024;
}

template <int KD>
int launch_seq_t(const smct::SeqArgs& g, cudaStream_t st) {
  static PerDeviceOnce attr_once;
  const size_t smem = std::max(step_smem_bytes(g.step.D, g.step.dff, g.step.full_model), (size_t)g.step.P * sizeof(double));
  if (smem > 200 * 1024) return fail(SMCT_ERR_UNSUPPORTED, "per-sequence forward needs %zu B shared memory", smem);
  if (attr_once.need()) {
    SMCT_CUDA(cudaFuncSetAttribute(smct::seq_forward_kernel<KD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  const int nthreads = KD <= 2 ? smct::NT_WIDE : smct::NT;
  if (g.cl <= 1) {
    smct::seq_forward_kernel<KD><<<g.step.B, nthreads, smem, st>>>(g);
  } else {   // one thread-block cluster per sequence
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(g.step.B * g.cl), 1, 1);
    cfg.blockDim = dim3((unsigned)nthreads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)g.cl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    SMCT_CUDA(cudaLaunchKernelEx(&cfg, smct::seq_forward_kernel<KD>, g));
  }
  SMCT_LAUNCH_CHECK("seq_forward_kernel");
  return SMCT_OK;
}

template <int LG, int NTHR>
int launch_small_t(const smct::SeqArgs& g, const smct::SmallSmemLayout& L, int nthreads, cudaStream_t st) {
  static PerDeviceOnce attr_once;
  if (attr_once.need()) {
    SMCT_CUDA(cudaFuncSetAttribute(smct::small_seq_forward_kernel<LG, NTHR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024));
  }
  smct::small_seq_forward_kernel<LG, NTHR><<<g.step.B, nthreads, L.total_bytes, st>>>(g, L);
  SMCT_LAUNCH_CHECK("small_seq_forward_kernel");
  return SMCT_OK;
}

int launch_small(const smct::SeqArgs& g, cudaStream_t st) {
  const int D = g.step.D, LG = D / 8, P = g.step.P;
  const smct::SmallSmemLayout L = smct::small_layout(D, g.step.dff, g.step.Fy, P, g.step.S);
  if (L.total_bytes > 208 * 1024) return fail(SMCT_ERR_UNSUPPORTED, "small forward needs %d B shared memory", L.total_bytes);
  const int nthreads = (P * LG + 31) / 32 * 32;
  if (LG == 1) return nthreads <= 256 ? launch_small_t<1, 256>(g, L, nthreads, st) : launch_small_t<1, 1024>(g, L, nthreads, st);
  if (LG == 2) return nthreads <= 256 ? launch_small_t<2, 256>(g, L, nthreads, st) : launch_small_t<2, 512>(g, L, nthreads, st);
  return nthreads <= 256 ? launch_small_t<4, 256>(g, L, nthreads, st) : launch_small_t<4, 512>(g, L, nthreads, st);
}

int forward_staged(const smct_shape& s, const smct_params& pr, const smct_io& io, void* ws, size_t ws_bytes,
                   cudaStream_t st, bool seq, bool small = false) {
  if (ws_bytes < staged_forward_ws(s)) return fail(SMCT_ERR_WORKSPACE, "workspace %zu < required %zu", ws_bytes, staged_forward_ws(s));
  const int B = s.B, P = s.P, S = s.S, D = s.D;
  const size_t rows = (size_t)B * std::max(S, P);
  Carver c{(char*)ws, 0, ws_bytes};
  float* xln = c.take<float>(rows * D);
  float* q_ = c.take<float>(rows * D);
  float* k_ = c.take<float>(rows * D);
  float* v_ = c.take<float>(rows * D);
  float* kx = c.take<float>(rows * D);
  float* vx = c.take<float>(rows * D);
  float* logw = c.take<float>((size_t)B * P);
  int* idx = c.take<int>((size_t)B * P);
  const size_t BPSD = (size_t)B * P * S * D;
  float* K2 = c.take<float>(BPSD);
  float* V2 = c.take<float>(BPSD);
  float* R2 = c.take<float>(BPSD);

  if (io.loss_sums) SMCT_CUDA(cudaMemsetAsync(io.loss_sums, 0, 8 * sizeof(double), st));
  if (io.attn) SMCT_CUDA(cudaMemsetAsync(io.attn, 0, (size_t)B * P * s.H * S * S * sizeof(float), st));

  smct::RowsArgs ra;
  memset(&ra, 0, sizeof(ra));
  ra.rows = (long long)B * S; ra.D = D; ra.Fx = s.F_x; ra.input_mode = s.input_mode; ra.do_ln = 1;
  ra.sqrtD = sqrtf((float)D); ra.ln_eps = s.ln_eps; ra.x_in = io.x_in;
  ra.W_in = pr.W_in; ra.b_in = pr.b_in; ra.ln_g = pr.ln1_g; ra.ln_b = pr.ln1_b;
  ra.Wq = pr.Wq; ra.bq = pr.bq; ra.Wk = pr.Wk; ra.bk = pr.bk; ra.Wv = pr.Wv; ra.bv = pr.bv;
  ra.xln = xln; ra.q_ = q_; ra.k_ = k_; ra.v_ = v_; ra.kx = kx; ra.vx = vx;
  if (int rc = launch_rows(ra, st)) return rc;

  const int n_gather = s.noise_on ? (s.len_resampling < 0 ? S : std::min(S, std::max(0, s.len_resampling))) : 0;
  float *Kc, *Vc, *Rc, *Ka, *Va, *Ra;
  if (n_gather % 2 == 0) { Kc = io.K; Vc = io.V; Rc = io.R; Ka = K2; Va = V2; Ra = R2; }
  else { Kc = K2; Vc = V2; Rc = R2; Ka = io.K; Va = io.V; Ra = io.R; }

  if (!s.noise_on) {  // SMC_Transformer.py:246: weights are ones, nothing is drawn, no noise is recorded
    if (io.w) {
      fill_f32_kernel<<<256, 256, 0, st>>>(io.w, (long long)B * P * S, 1.0f);
      SMCT_LAUNCH_CHECK("fill_f32_kernel");
    }
    const size_t nBPS = (size_t)B * P * S;
    if (io.i_out) SMCT_CUDA(cudaMemsetAsync(io.i_out, 0, nBPS * sizeof(int), st));
    if (io.logw) SMCT_CUDA(cudaMemsetAsync(io.logw, 0, nBPS * sizeof(float), st));
    if (io.noise_q) SMCT_CUDA(cudaMemsetAsync(io.noise_q, 0, nBPS * D * sizeof(float), st));
    if (io.noise_z) SMCT_CUDA(cudaMemsetAsync(io.noise_z, 0, nBPS * D * sizeof(float), st));
  }

  const size_t BPD = (size_t)B * P * D, BP = (size_t)B * P;
  if (seq) {   // one launch: one CTA per sequence runs all S steps (same device code as the loop below)
    smct::SeqArgs g;
    memset(&g, 0, sizeof(g));
    fill_step_common(g.step, s, pr);
    g.step.loss_sums = io.loss_sums;
    g.attn_window = s.attn_window; g.len_resampling = s.len_resampling;
    g.normalise = (s.weights_mode == SMCT_WEIGHTS_GAUSSIAN); g.n_gather = n_gather;
    g.xln = xln; g.q_ = q_; g.k_ = k_; g.v_ = v_; g.y = io.y; g.eps = io.eps; g.u = io.u; g.i_forced = io.i_forced;
    g.seed = io.seed;
    g.K0 = io.K; g.V0 = io.V; g.R0 = io.R; g.K2 = K2; g.V2 = V2; g.R2 = R2;
    g.logw_ws = logw; g.idx_ws = idx;
    {   // CTAs per sequence: one per particle tile, rounded up to a power of two, at most a portable cluster (8)
      const int tiles = (P + smct::PT - 1) / smct::PT;
      g.cl = 1;
      while (g.cl < tiles && g.cl < 8) g.cl *= 2;
    }
    g.pred = io.pred; g.noise_q = io.noise_q; g.noise_z = io.noise_z; g.attn = io.attn; g.w = io.w; g.logw_out = io.logw;
    g.i_out = io.i_out;
    if (small) {   // the small kernel's epilogue writes the final state, the output heads and the loss sums itself
      const bool want_noise_s = io.noise_K || io.noise_V || io.loss_sums;
      g.kx = want_noise_s ? kx : nullptr; g.vx = want_noise_s ? vx : nullptr;
      g.pred_resampl = io.pred_resampl; g.noise_K = io.noise_K; g.noise_V = io.noise_V;
      return launch_small(g, st);
    }
    int rc;
    if (D <= 32) rc = launch_seq_t<1>(g, st);
    else if (D <= 64) rc = launch_seq_t<2>(g, st);
    else if (D <= 128) rc = launch_seq_t<4>(g, st);
    else rc = launch_seq_t<8>(g, st);
    if (rc) return rc;
    Kc = io.K; Vc = io.V; Rc = io.R;   // n_gather swaps end in the caller's buffers by construction
  }
  for (int t = 0; t < (seq ? 0 : S); ++t) {
    smct::StepArgs a;
    fill_step_common(a, s, pr);
    a.t = t;
    a.s_lo = (s.attn_window >= 0 && t > s.attn_window) ? t - s.attn_window : 0;
    a.xln = xln + (size_t)t * D; a.q_ = q_ + (size_t)t * D; a.k_ = k_ + (size_t)t * D; a.v_ = v_ + (size_t)t * D;
    a.row_sb = (long long)S * D; a.row_sp = 0;
    if (s.weights_mode == SMCT_WEIGHTS_GAUSSIAN) { a.y = (const float*)io.y + (size_t)t * s.F_y; a.y_sb = (long long)S * s.F_y; }
    else { a.y = (const int*)io.y + t; a.y_sb = S; }
    if (io.eps) {
      a.eps_k = io.eps + ((size_t)0 * S + t) * BPD; a.eps_q = io.eps + ((size_t)1 * S + t) * BPD;
      a.eps_v = io.eps + ((size_t)2 * S + t) * BPD; a.eps_z = io.eps + ((size_t)3 * S + t) * BPD;
    }
    a.seed = io.seed; a.t_rng = t;
    a.K = Kc; a.V = Vc; a.R = Rc;
    a.pred = io.pred ? io.pred + (size_t)t * s.F_y : nullptr; a.pred_st = (long long)S * s.F_y;
    a.noise_q = io.noise_q ? io.noise_q + (size_t)t * D : nullptr;
    a.noise_z = io.noise_z ? io.noise_z + (size_t)t * D : nullptr;
    a.nz_st = (long long)S * D;
    a.attn = io.attn ? io.attn + (size_t)t * S : nullptr;
    a.attn_st_bp = (long long)s.H * S * S; a.attn_st_h = (long long)S * S;
    a.logw = logw;
    a.loss_sums = io.loss_sums;
    if (int rc = launch_step(a, st)) return rc;

    if (s.noise_on) {
      smct::ResampleArgs r;
      memset(&r, 0, sizeof(r));
      r.B = B; r.P = P; r.normalise = (s.weights_mode == SMCT_WEIGHTS_GAUSSIAN);
      r.logw = logw;
      r.u = io.u ? io.u + (size_t)t * BP : nullptr;
      r.i_forced = io.i_forced ? io.i_forced + (size_t)t * BP : nullptr;
      r.seed = io.seed; r.b_global0 = s.b_global0; r.t_rng = t;
      r.w_out = io.w ? io.w + t : nullptr; r.w_st = S;
      r.logw_copy = io.logw ? io.logw + (size_t)t * BP : nullptr;
      r.i_out = idx;
      r.i_out2 = io.i_out ? io.i_out + (size_t)t * BP : nullptr;
      if (int rc = launch_resample(r, st)) return rc;
      if (s.len_resampling < 0 || t < s.len_resampling) {
        smct::GatherArgs g;
        memset(&g, 0, sizeof(g));
        g.B = B; g.P = P; g.row_elems = (long long)S * D; g.n_elems = (long long)(t + 1) * D;
        g.src[0] = Kc; g.src[1] = Vc; g.src[2] = Rc; g.dst[0] = Ka; g.dst[1] = Va; g.dst[2] = Ra; g.idx = idx;
        if (int rc = launch_gather(g, 3, st)) return rc;
        std::swap(Kc, Ka); std::swap(Vc, Va); std::swap(Rc, Ra);
      }
    }
  }
  if (Kc != io.K) return fail(SMCT_ERR_INVALID, "internal: ping-pong parity error");

  smct::FinalArgs f;
  memset(&f, 0, sizeof(f));
  f.B = B; f.P = P; f.S = S; f.D = D; f.Fy = s.F_y; f.weights_mode = s.weights_mode; f.logvar_dim = s.logvar_dim;
  f.K = io.K; f.V = io.V; f.R = io.R;
  const bool want_noise = io.noise_K || io.noise_V || io.loss_sums;
  f.kx = want_noise ? kx : nullptr; f.vx = want_noise ? vx : nullptr;
  f.Wo = pr.Wo; f.logvar_k = pr.logvar_k; f.logvar_v = pr.logvar_v;
  f.y = (s.weights_mode == SMCT_WEIGHTS_GAUSSIAN) ? io.y : nullptr;
  f.pred_resampl = io.pred_resampl; f.noise_K = io.noise_K; f.noise_V = io.noise_V; f.loss_sums = io.loss_sums;
  const long long nrows = (long long)B * P * S;
  const int grid = (int)std::min<long long>((nrows + 7) / 8, 148LL * 32);
  smct::final_kernel<<<grid, 256, 0, st>>>(f);
  SMCT_LAUNCH_CHECK("final_kernel");
  return SMCT_OK;
}

// ------------------------------------------------------------------------------------------
// fused forward: rows_kernel -> fused_forward_kernel (all S steps, one CTA per sequence)
//                -> materialize_kernel -> final_kernel
// ------------------------------------------------------------------------------------------
int forward_fused(const smct_shape& s, const smct_params& pr, const smct_io& io, void* ws, size_t ws_bytes,
                  cudaStream_t st, int algo) {
  const bool use_tc = algo != SMCT_ALGO_FUSED_FFMA;
  if (ws_bytes < smct::fused_forward_ws(s)) return fail(SMCT_ERR_WORKSPACE, "workspace %zu < required %zu", ws_bytes, smct::fused_forward_ws(s));
  const int B = s.B, P = s.P, S = s.S, D = s.D;
  const int Sa = (S + 7) / 8 * 8;
  Carver c{(char*)ws, 0, ws_bytes};
  const size_t BSD = (size_t)B * S * D;
  float* xln = c.take<float>(BSD);
  float* q_ = c.take<float>(BSD);
  float* k_ = c.take<float>(BSD);
  float* v_ = c.take<float>(BSD);
  float* kx = c.take<float>(BSD);
  float* vx = c.take<float>(BSD);
  const size_t BSPD = (size_t)B * S * P * D;
  float* Kp = c.take<float>(BSPD);
  float* Vp = c.take<float>(BSPD);
  float* Rp = c.take<float>(BSPD);
  unsigned char* Ks = c.take<unsigned char>(BSPD * 4);   // fp16 (h,m) split copies of the K/V rows (attention operands)
  unsigned char* Vs = c.take<unsigned char>(BSPD * 4);
  unsigned short* A0 = c.take<unsigned short>((size_t)B * P * Sa);
  unsigned short* A1 = c.take<unsigned short>((size_t)B * P * Sa);
  unsigned short* pos = c.take<unsigned short>((size_t)B * P);
  int* aflag = c.take<int>((size_t)B);
  unsigned char* wtc = c.take<unsigned char>((size_t)(1 + 2 * (smct::FF / smct::FD)) * 16384 + 256 + 2048);   // blocks, header, noise-scale table

  if (io.loss_sums) SMCT_CUDA(cudaMemsetAsync(io.loss_sums, 0, 8 * sizeof(double), st));
  if (use_tc) {
    smct::TcAttnPrep ap{pr.Wq, pr.bq, pr.Wk, pr.bk, pr.Wv, pr.bv, pr.logvar_q, pr.logvar_k, pr.logvar_v, pr.logvar_z, s.logvar_dim};
    smct::tc_weight_prep_kernel<<<1, 1024, 0, st>>>(pr.Wz, pr.W1, pr.W2, pr.b1, pr.ln1_g, pr.ln1_b, s.dff, wtc, ap);
    SMCT_LAUNCH_CHECK("tc_weight_prep_kernel");
  }

  smct::RowsArgs ra;
  memset(&ra, 0, sizeof(ra));
  ra.rows = (long long)B * S; ra.D = D; ra.Fx = s.F_x; ra.input_mode = s.input_mode; ra.do_ln = 1;
  ra.sqrtD = sqrtf((float)D); ra.ln_eps = s.ln_eps; ra.x_in = io.x_in;
  ra.W_in = pr.W_in; ra.b_in = pr.b_in; ra.ln_g = pr.ln1_g; ra.ln_b = pr.ln1_b;
  ra.Wq = pr.Wq; ra.bq = pr.bq; ra.Wk = pr.Wk; ra.bk = pr.bk; ra.Wv = pr.Wv; ra.bv = pr.bv;
  ra.xln = xln; ra.q_ = q_; ra.k_ = k_; ra.v_ = v_; ra.kx = kx; ra.vx = vx;
  if (int rc = launch_rows(ra, st)) return rc;

  smct::FusedArgs a;
  memset(&a, 0, sizeof(a));
  a.B = B; a.P = P; a.S = S; a.Sa = Sa; a.Fy = s.F_y; a.dff = s.dff;
  a.attn_window = s.attn_window; a.len_resampling = s.len_resampling; a.noise_z_mode = s.noise_z_mode;
  a.logvar_dim = s.logvar_dim; a.ln_eps = s.ln_eps; a.sigma_obs = s.sigma_obs;
  a.b_global0 = s.b_global0; a.seed = io.seed;
  a.xln = xln; a.q_ = q_; a.k_ = k_; a.v_ = v_;
  a.y = (const float*)io.y; a.eps = io.eps; a.u = io.u; a.i_forced = io.i_forced;
  a.p = pr;
  a.Kp = Kp; a.Vp = Vp; a.Rp = Rp; a.Ks = Ks; a.Vs = Vs; a.A0 = A0; a.A1 = A1; a.pos_out = pos; a.aflag = aflag;
  a.pred = io.pred; a.w = io.w; a.logw = io.logw; a.noise_q = io.noise_q; a.noise_z = io.noise_z;
  a.i_out = io.i_out; a.loss_sums = io.loss_sums;
  a.wtc = wtc;
  static PerDeviceOnce attr_once;
  if (attr_once.need()) {
    SMCT_CUDA(cudaFuncSetAttribute(smct::fused_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)sizeof(smct::FusedSmem)));
    SMCT_CUDA(cudaFuncSetAttribute(smct::fused_tc_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)sizeof(smct::FusedTcSmem) + 1024));
  }
  if (algo == SMCT_ALGO_FUSED) {
    // compile-time variants: injected eps (tests) | per-dimension log-variances | the checkout's noise_z
    using KernelT = void (*)(const smct::FusedArgs);
    static const KernelT variants[8] = {
        smct::fused2_forward_kernel<false, false, false>, smct::fused2_forward_kernel<false, false, true>,
        smct::fused2_forward_kernel<false, true, false>,  smct::fused2_forward_kernel<false, true, true>,
        smct::fused2_forward_kernel<true, false, false>,  smct::fused2_forward_kernel<true, false, true>,
        smct::fused2_forward_kernel<true, true, false>,   smct::fused2_forward_kernel<true, true, true>};
    static PerDeviceOnce f2_once;
    if (f2_once.need()) {
      for (KernelT k : variants) {
        SMCT_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(smct::Fused2Smem) + 128));
        SMCT_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
      }
    }
    const int vi = (io.eps ? 4 : 0) | (s.logvar_dim > 1 ? 2 : 0) | (s.noise_z_mode == SMCT_NOISE_Z_CHECKOUT ? 1 : 0);
    variants[vi]<<<B, smct::F2NT, sizeof(smct::Fused2Smem) + 128, st>>>(a);
    SMCT_LAUNCH_CHECK("fused2_forward_kernel");
  } else if (use_tc) {
    smct::fused_tc_forward_kernel<<<B, smct::FNT, sizeof(smct::FusedTcSmem) + 1024, st>>>(a);
    SMCT_LAUNCH_CHECK("fused_tc_forward_kernel");
  } else {
    smct::fused_forward_kernel<<<B, smct::FNT, sizeof(smct::FusedSmem), st>>>(a);
    SMCT_LAUNCH_CHECK("fused_forward_kernel");
  }

  smct::MatArgs m;
  memset(&m, 0, sizeof(m));
  m.B = B; m.P = P; m.S = S; m.Sa = Sa; m.Fy = s.F_y; m.logvar_dim = s.logvar_dim;
  m.Kp = Kp; m.Vp = Vp; m.Rp = Rp; m.A0 = A0; m.A1 = A1; m.pos = pos; m.aflag = aflag;
  const bool want_noise = io.noise_K || io.noise_V || io.loss_sums;
  m.kx = want_noise ? kx : nullptr; m.vx = want_noise ? vx : nullptr;
  m.Wo = pr.Wo; m.logvar_k = pr.logvar_k; m.logvar_v = pr.logvar_v;
  m.y = io.loss_sums ? (const float*)io.y : nullptr;
  m.K = io.K; m.V = io.V; m.R = io.R;
  m.pred_resampl = io.pred_resampl; m.noise_K = io.noise_K; m.noise_V = io.noise_V; m.loss_sums = io.loss_sums;
  {
    static PerDeviceOnce mat_once;
    const size_t mat_smem = sizeof(smct::MatSmem) + 128;
    if (mat_once.need())
      SMCT_CUDA(cudaFuncSetAttribute(smct::materialize_final_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mat_smem));
    const unsigned grid = (unsigned)((long long)B * ((P + smct::MAT_PB - 1) / smct::MAT_PB) * ((S + smct::MAT_ROWS - 1) / smct::MAT_ROWS));
    smct::materialize_final_kernel<<<grid, smct::MAT_NT, mat_smem, st>>>(m);
  }
  SMCT_LAUNCH_CHECK("materialize_final_kernel");
  return SMCT_OK;
}

// ------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------
smct::GradLayout make_layout(const smct_shape& s) {
  smct::GradLayout l;
  const int D = s.D, dff = s.dff;
  int o = 0;
  auto take = [&](int n) { int r = o; o += (n + 3) / 4 * 4; return r; };
  l.Wq = take(D * D); l.bq = take(D); l.Wk = take(D * D); l.bk = take(D); l.Wv = take(D * D); l.bv = take(D);
  l.Wz = take(D * D); l.bz = take(D); l.ln1_g = take(D); l.ln1_b = take(D); l.ln2_g = take(D); l.ln2_b = take(D);
  l.W1 = take(D * dff); l.b1 = take(dff); l.W2 = take(dff * D); l.b2 = take(D); l.Wo = take(D * s.F_y);
  l.W_in = take(s.input_mode == SMCT_INPUT_ENCODED ? 0 : s.F_x * D); l.b_in = take(D);
  l.total = o;
  return l;
}

constexpr int BWD_NREP = 296;

int check_backward_shape(const smct_shape& s) {
  // noise_z_mode checkout (noise_z = z - z_, self_attention_SMC.py:190) makes the noise_z loss term of EVERY per-step
  // particle depend on the parameters, including particles that do not survive to the final genealogy: that term is
  // outside the path-wise formulation (which sees the final lineages only)
  if (s.H != 1 || !s.full_model || !s.noise_on || s.noise_z_mode != SMCT_NOISE_Z_PYC)
    return fail(SMCT_ERR_UNSUPPORTED, "backward supports: 1 head, full model, noise on, noise_z_mode pyc");
  if (s.D > 128) return fail(SMCT_ERR_UNSUPPORTED, "backward: d_model %d > 128", s.D);
  if ((long long)s.S * s.D > 12800) return fail(SMCT_ERR_UNSUPPORTED, "backward: S*d_model %lld > 12800", (long long)s.S * s.D);
  if (s.P > 12288) return fail(SMCT_ERR_UNSUPPORTED, "backward: particles %d > 12288", s.P);
  if ((long long)s.B * s.P * s.S >= (1LL << 31)) return fail(SMCT_ERR_UNSUPPORTED, "backward: B*P*S %lld >= 2^31 (run the batch in chunks)", (long long)s.B * s.P * s.S);
  return SMCT_OK;
}

size_t backward_ws(const smct_shape& s) {
  const size_t BSD = (size_t)s.B * s.S * s.D, BPS = (size_t)s.B * s.P * s.S;
  const smct::GradLayout l = make_layout(s);
  size_t n = 0;
  n += 5 * align_up(BSD * 4);                                   // X, xln, q_, kx, vx
  n += 4 * align_up((size_t)s.D * s.D * 4) + 2 * align_up((size_t)s.D * s.dff * 4);   // transposed weights (Wz, Wq, Wk, Wv; W1, W2)
  n += align_up(BPS * 2);                                       // anc
  n += 2 * align_up(((size_t)s.B * s.S + 1) * 4);               // node_cnt, node_off
  n += 2 * align_up(BPS * 4) + 3 * align_up(BPS * 2);           // node_bs, node_m, node_p, node_L, node_of
  n += 3 * align_up(BPS * s.D * 4) + align_up(BPS * 2 * 4);     // Zm, dZm, Qm, stat
  n += 6 * align_up(BSD * 4);                                   // per-row sums
  n += align_up((size_t)BWD_NREP * l.total * 4) + align_up((size_t)l.total * 4);
  n += align_up((size_t)4 * 256 * sizeof(double));   // per-dimension noise sums (logvar_dim = D)
  return n + 1024;
}

template <int KD>
int launch_backward_t(const smct::BwdArgs& a, cudaStream_t st) {
  const int D = a.D, dff = a.dff;
  const long long nrows = (long long)a.B * a.P * a.S;   // upper bound of the node count (known on the device only)
  if (D % 4 == 0 && D <= 64) smct::bwd_attn_fwd_half_kernel<<<(int)std::min<long long>((nrows + 7) / 8, 148LL * 64), 256, 0, st>>>(a);
  else smct::bwd_attn_fwd_kernel<KD><<<(int)std::min<long long>((nrows + 7) / 8, 148LL * 64), 256, 0, st>>>(a);
  SMCT_LAUNCH_CHECK("bwd_attn_fwd_kernel");
  const size_t smem_d = ((size_t)5 * smct::PT * (D + 4) + 2 * smct::PT * (dff + 4) + smct::PT + 7 * D + dff + D * a.Fy) * sizeof(float);   // (strides: smct_backward.cuh)
  if (smem_d > 200 * 1024) return fail(SMCT_ERR_UNSUPPORTED, "backward dense tile needs %zu B shared memory", smem_d);
  static PerDeviceOnce attr_once;
  if (attr_once.need()) {
    SMCT_CUDA(cudaFuncSetAttribute(smct::bwd_dense_kernel<KD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  const long long ntiles = (nrows + smct::PT - 1) / smct::PT;
  smct::bwd_dense_kernel<KD><<<(int)std::min<long long>(ntiles, a.nrep), smct::NT, smem_d, st>>>(a);
  SMCT_LAUNCH_CHECK("bwd_dense_kernel");
  if (D % 4 == 0 && D <= 64) {   // rows as one float4 per lane of a half-warp: two target slots per warp at a time
    smct::bwd_attn_bwd_chain_kernel<<<a.B * smct::BWD_CG, smct::BWD_CNT, 0, st>>>(a);   // nodes walked by representative particle
    SMCT_LAUNCH_CHECK("bwd_attn_bwd_chain_kernel");
  } else {
    smct::bwd_attn_bwd_kernel<KD><<<a.B * smct::BWD_G, smct::BWD_ANT, 0, st>>>(a);
    SMCT_LAUNCH_CHECK("bwd_attn_bwd_kernel");
  }
  return SMCT_OK;
}

}  // namespace

// ==========================================================================================
// extern "C"
// ==========================================================================================
extern "C" {

int smct_abi_version(void) { return SMCT_ABI_VERSION; }
const char* smct_last_error(void) { return g_last_error.c_str(); }
const char* smct_build_info(void) {
#define SMCT_STR2(x) #x
#define SMCT_STR(x) SMCT_STR2(x)
  return "libsmct_b200 abi " SMCT_STR(SMCT_ABI_VERSION) " | sm_100a | nvcc " SMCT_STR(__CUDACC_VER_MAJOR__) "." SMCT_STR(
      __CUDACC_VER_MINOR__) " | " __DATE__ " " __TIME__ " | src " SMCT_SOURCE_HASH;
}

const char* smct_source_hash(void) {
  static const char tagged[] = "smct-source-hash:" SMCT_SOURCE_HASH;   // the tag lets build.py find the hash in the binary's bytes
  return tagged + 17;
}

long long smct_launch_count(int reset) {
  return reset ? g_launches.exchange(0) : g_launches.load();
}

size_t smct_workspace_bytes(const smct_shape* shape) {
  if (check_shape(shape) != SMCT_OK) return 0;
  return std::max(staged_forward_ws(*shape), smct::fused_forward_ws(*shape));
}

int smct_forward(const smct_shape* shape, const smct_params* params, const smct_io* io, void* workspace,
                 size_t workspace_bytes, void* stream) {
  if (int rc = check_shape(shape)) return rc;
  if (!params || !io) return fail(SMCT_ERR_INVALID, "params/io is NULL");
  if (!io->x_in || !io->K || !io->V || !io->R) return fail(SMCT_ERR_INVALID, "x_in, K, V, R are required");
  if ((((uintptr_t)io->K | (uintptr_t)io->V | (uintptr_t)io->R) & 15u) != 0)   // 128-bit stores and bulk copies
    return fail(SMCT_ERR_INVALID, "K, V, R must be 16-byte aligned");
  if (shape->noise_on && !io->y) return fail(SMCT_ERR_INVALID, "y (targets) required when noise is on");
  if (!workspace) return fail(SMCT_ERR_WORKSPACE, "workspace is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  int algo = shape->algo;
  if (algo == SMCT_ALGO_AUTO)
    algo = smct::fused2_supported(*shape, *io) ? SMCT_ALGO_FUSED : (smct::fused_supported(*shape, *io) ? SMCT_ALGO_FUSED_V1 : SMCT_ALGO_STAGED);
  if (algo == SMCT_ALGO_FUSED || algo == SMCT_ALGO_FUSED_FFMA || algo == SMCT_ALGO_FUSED_V1) {
    const bool ok = algo == SMCT_ALGO_FUSED ? smct::fused2_supported(*shape, *io) : smct::fused_supported(*shape, *io);
    if (!ok) return fail(SMCT_ERR_UNSUPPORTED, "fused forward does not support this shape/io");
    if (!io->y) return fail(SMCT_ERR_INVALID, "y (targets) required");
    return forward_fused(*shape, *params, *io, workspace, workspace_bytes, st, algo);
  }
  // the per-sequence single-launch variant (a thread-block cluster per sequence, one CTA per particle tile) serves the
  // reference's own small shapes; with more than 8 tiles per sequence (P > 256) a CTA would serialise several tiles
  // (measured on B200: C1 0.69 -> 0.44 ms, C2 0.91 -> 0.86 ms; C3 — 256 sequences x 2 tiles, more CTAs than SMs — is
  // faster on the per-stage kernels: 2.75 vs 3.75 ms, so AUTO takes the cluster form only while every CTA gets its own SM)
  if (algo == SMCT_ALGO_SMALL || (shape->algo == SMCT_ALGO_AUTO && smct::small_supported(*shape, *io))) {
    // the reference's own small shapes: the whole forward of a sequence in one CTA, a particle's rows in registers
    if (!smct::small_supported(*shape, *io)) return fail(SMCT_ERR_UNSUPPORTED, "small forward does not support this shape/io");
    return forward_staged(*shape, *params, *io, workspace, workspace_bytes, st, true, true);
  }
  const int seq_tiles = (shape->P + smct::PT - 1) / smct::PT;
  const bool seq = (algo == SMCT_ALGO_STAGED_SEQ) ||
                   (shape->algo == SMCT_ALGO_AUTO && (seq_tiles == 1 || (seq_tiles <= 8 && (long long)shape->B * seq_tiles <= 148)));
  return forward_staged(*shape, *params, *io, workspace, workspace_bytes, st, seq);
}

int smct_cell_step(const smct_shape* shape, const smct_params* params, int32_t t, const float* x,
                   int32_t x_per_particle, const void* y, const float* eps4, const double* u,
                   const int32_t* i_forced, uint64_t seed, float* K, float* V, float* R, float* r, float* pred,
                   float* attn, float* noise_q, float* noise_z, float* w, int32_t* i_out, void* workspace,
                   size_t workspace_bytes, void* stream) {
  if (int rc = check_shape(shape)) return rc;
  const smct_shape& s = *shape;
  if (!params || !x || !K || !V || !R) return fail(SMCT_ERR_INVALID, "params, x, K, V, R are required");
  if (t < 0 || t >= s.S) return fail(SMCT_ERR_INVALID, "timestep %d outside [0,%d)", t, s.S);
  if (s.noise_on && !y) return fail(SMCT_ERR_INVALID, "y required when noise is on");
  if (!workspace || workspace_bytes < staged_forward_ws(s)) return fail(SMCT_ERR_WORKSPACE, "workspace %zu < required %zu", workspace_bytes, staged_forward_ws(s));
  cudaStream_t st = (cudaStream_t)stream;
  const int B = s.B, P = s.P, S = s.S, D = s.D;
  const size_t rows_cap = (size_t)B * std::max(S, P);
  Carver c{(char*)workspace, 0, workspace_bytes};
  float* xln = c.take<float>(rows_cap * D);
  float* q_ = c.take<float>(rows_cap * D);
  float* k_ = c.take<float>(rows_cap * D);
  float* v_ = c.take<float>(rows_cap * D);
  c.take<float>(rows_cap * D); c.take<float>(rows_cap * D);
  float* logw = c.take<float>((size_t)B * P);
  int* idx = c.take<int>((size_t)B * P);
  const size_t BPSD = (size_t)B * P * S * D;
  float* K2 = c.take<float>(BPSD);
  float* V2 = c.take<float>(BPSD);
  float* R2 = c.take<float>(BPSD);

  smct::RowsArgs ra;
  memset(&ra, 0, sizeof(ra));
  ra.rows = x_per_particle ? (long long)B * P : B; ra.D = D; ra.Fx = s.F_x; ra.input_mode = SMCT_INPUT_ENCODED; ra.do_ln = 1;
  ra.sqrtD = 1.f; ra.ln_eps = s.ln_eps; ra.x_in = x;
  ra.ln_g = params->ln1_g; ra.ln_b = params->ln1_b;
  ra.Wq = params->Wq; ra.bq = params->bq; ra.Wk = params->Wk; ra.bk = params->bk; ra.Wv = params->Wv; ra.bv = params->bv;
  ra.xln = xln; ra.q_ = q_; ra.k_ = k_; ra.v_ = v_;
  if (int rc = launch_rows(ra, st)) return rc;
  if (attn) SMCT_CUDA(cudaMemsetAsync(attn, 0, (size_t)B * P * s.H * S * sizeof(float), st));

  smct::StepArgs a;
  fill_step_common(a, s, *params);
  a.t = t; a.s_lo = (s.attn_window >= 0 && t > s.attn_window) ? t - s.attn_window : 0;
  a.xln = xln; a.q_ = q_; a.k_ = k_; a.v_ = v_;
  a.row_sb = x_per_particle ? (long long)P * D : D; a.row_sp = x_per_particle ? D : 0;
  a.y = y; a.y_sb = (s.weights_mode == SMCT_WEIGHTS_GAUSSIAN) ? s.F_y : 1;
  const size_t BPD = (size_t)B * P * D;
  if (eps4) { a.eps_k = eps4; a.eps_q = eps4 + BPD; a.eps_v = eps4 + 2 * BPD; a.eps_z = eps4 + 3 * BPD; }
  a.seed = seed; a.t_rng = t;
  a.K = K; a.V = V; a.R = R;
  a.pred = pred; a.pred_st = s.F_y;
  a.r_out = r; a.r_st = D;
  a.noise_q = noise_q; a.noise_z = noise_z; a.nz_st = D;
  a.attn = attn; a.attn_st_bp = (long long)s.H * S; a.attn_st_h = S;
  a.logw = logw;
  if (int rc = launch_step(a, st)) return rc;
  if (s.noise_on) {
    smct::ResampleArgs rs;
    memset(&rs, 0, sizeof(rs));
    rs.B = B; rs.P = P; rs.normalise = (s.weights_mode == SMCT_WEIGHTS_GAUSSIAN);
    rs.logw = logw; rs.u = u; rs.i_forced = i_forced; rs.seed = seed; rs.b_global0 = s.b_global0; rs.t_rng = t;
    rs.w_out = w; rs.w_st = 1; rs.i_out = idx; rs.i_out2 = i_out;
    if (int rc = launch_resample(rs, st)) return rc;
    if (s.len_resampling < 0 || t < s.len_resampling) {
      // the reference gathers whole rows (all S slots): tf.gather on (B,P,S,3D), SMC_TransformerCell.py:170-172
      smct::GatherArgs g;
      memset(&g, 0, sizeof(g));
      g.B = B; g.P = P; g.row_elems = (long long)S * D; g.n_elems = (long long)S * D;
      g.src[0] = K; g.src[1] = V; g.src[2] = R; g.dst[0] = K2; g.dst[1] = V2; g.dst[2] = R2; g.idx = idx;
      if (int rc = launch_gather(g, 3, st)) return rc;
      SMCT_CUDA(cudaMemcpyAsync(K, K2, BPSD * sizeof(float), cudaMemcpyDeviceToDevice, st));
      SMCT_CUDA(cudaMemcpyAsync(V, V2, BPSD * sizeof(float), cudaMemcpyDeviceToDevice, st));
      SMCT_CUDA(cudaMemcpyAsync(R, R2, BPSD * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
  } else if (w) {
    fill_f32_kernel<<<64, 256, 0, st>>>(w, (long long)B * P, 1.0f);
    SMCT_LAUNCH_CHECK("fill_f32_kernel");
  }
  return SMCT_OK;
}

int smct_attention_step(const smct_shape* shape, const smct_params* params, int32_t t, const float* inputs,
                        const float* eps4, uint64_t seed, float* K, float* V, float* z, float* attn,
                        float* noise_k, float* noise_q, float* noise_v, float* noise_z, void* workspace,
                        size_t workspace_bytes, void* stream) {
  if (int rc = check_shape(shape)) return rc;
  const smct_shape& s = *shape;
  if (!params || !inputs || !K || !V || !z) return fail(SMCT_ERR_INVALID, "params, inputs, K, V, z are required");
  if (t < 0 || t >= s.S) return fail(SMCT_ERR_INVALID, "timestep %d outside [0,%d)", t, s.S);
  if (!workspace || workspace_bytes < staged_forward_ws(s)) return fail(SMCT_ERR_WORKSPACE, "workspace %zu < required %zu", workspace_bytes, staged_forward_ws(s));
  cudaStream_t st = (cudaStream_t)stream;
  const int B = s.B, P = s.P, S = s.S, D = s.D;
  const size_t rows_cap = (size_t)B * std::max(S, P);
  Carver c{(char*)workspace, 0, workspace_bytes};
  c.take<float>(rows_cap * D);
  float* q_ = c.take<float>(rows_cap * D);
  float* k_ = c.take<float>(rows_cap * D);
  float* v_ = c.take<float>(rows_cap * D);

  smct::RowsArgs ra;
  memset(&ra, 0, sizeof(ra));
  ra.rows = (long long)B * P; ra.D = D; ra.Fx = s.F_x; ra.input_mode = SMCT_INPUT_ENCODED; ra.do_ln = 0;
  ra.sqrtD = 1.f; ra.ln_eps = s.ln_eps; ra.x_in = inputs;
  ra.Wq = params->Wq; ra.bq = params->bq; ra.Wk = params->Wk; ra.bk = params->bk; ra.Wv = params->Wv; ra.bv = params->bv;
  ra.q_ = q_; ra.k_ = k_; ra.v_ = v_;
  if (int rc = launch_rows(ra, st)) return rc;
  if (attn) SMCT_CUDA(cudaMemsetAsync(attn, 0, (size_t)B * P * s.H * S * sizeof(float), st));

  smct::StepArgs a;
  fill_step_common(a, s, *params);
  a.full_model = 0; a.attn_only = 1;
  a.t = t; a.s_lo = (s.attn_window >= 0 && t > s.attn_window) ? t - s.attn_window : 0;
  a.xln = inputs; a.q_ = q_; a.k_ = k_; a.v_ = v_;
  a.row_sb = (long long)P * D; a.row_sp = D;
  const size_t BPD = (size_t)B * P * D;
  if (eps4) { a.eps_k = eps4; a.eps_q = eps4 + BPD; a.eps_v = eps4 + 2 * BPD; a.eps_z = eps4 + 3 * BPD; }
  a.seed = seed; a.t_rng = t;
  a.K = K; a.V = V; a.R = nullptr;
  a.z_out = z; a.z_st = D;
  a.noise_k = noise_k; a.noise_q = noise_q; a.noise_v = noise_v; a.noise_z = noise_z; a.nz_st = D;
  a.attn = attn; a.attn_st_bp = (long long)s.H * S; a.attn_st_h = S;
  return launch_step(a, st);
}

int smct_logweights(int32_t B, int32_t P, int32_t F_y, int32_t weights_mode, float sigma_obs, const float* pred,
                    const void* y, float* logw, void* stream) {
  if (B < 1 || P < 1 || F_y < 1 || !pred || !y || !logw) return fail(SMCT_ERR_INVALID, "smct_logweights: bad arguments");
  if (weights_mode == SMCT_WEIGHTS_GAUSSIAN && !(sigma_obs > 0.f)) return fail(SMCT_ERR_INVALID, "sigma_obs must be > 0");
  const long long n = (long long)B * P;
  smct::logweights_kernel<<<(int)std::min<long long>((n + 255) / 256, 148 * 8), 256, 0, (cudaStream_t)stream>>>(
      B, P, F_y, weights_mode, sigma_obs, pred, y, logw);
  SMCT_LAUNCH_CHECK("logweights_kernel");
  return SMCT_OK;
}

int smct_resample(int32_t B, int32_t P, int32_t normalise, const float* logw, const double* u, float* w_out,
                  float* logits_out, int32_t* i_out, void* stream) {
  if (B < 1 || P < 1 || !logw || !u || !i_out) return fail(SMCT_ERR_INVALID, "smct_resample: bad arguments");
  if (P > 16384) return fail(SMCT_ERR_UNSUPPORTED, "particles %d > 16384 unsupported", P);
  smct::ResampleArgs r;
  memset(&r, 0, sizeof(r));
  r.B = B; r.P = P; r.normalise = normalise; r.logw = logw; r.u = u;
  r.w_out = w_out; r.w_st = 1; r.logits_out = logits_out; r.i_out = i_out;
  return launch_resample(r, (cudaStream_t)stream);
}

int smct_gather(int32_t B, int32_t P, int64_t row_elems, int64_t n_elems, const float* src, const int32_t* idx,
                float* dst, void* stream) {
  if (B < 1 || P < 1 || row_elems < 1 || n_elems < 0 || n_elems > row_elems || !src || !idx || !dst)
    return fail(SMCT_ERR_INVALID, "smct_gather: bad arguments");
  if (src == dst) return fail(SMCT_ERR_INVALID, "smct_gather: src and dst must be distinct buffers");
  if (n_elems == 0) return SMCT_OK;
  smct::GatherArgs g;
  memset(&g, 0, sizeof(g));
  g.B = B; g.P = P; g.row_elems = row_elems; g.n_elems = n_elems; g.src[0] = src; g.dst[0] = dst; g.idx = idx;
  return launch_gather(g, 1, (cudaStream_t)stream);
}

int smct_fill_noise(uint64_t seed, int32_t which, int32_t t, int64_t b_global0, int32_t B, int32_t P, int32_t D,
                    float* out, void* stream) {
  if (which < 0 || which > 3 || B < 1 || P < 1 || D < 1 || !out) return fail(SMCT_ERR_INVALID, "smct_fill_noise: bad arguments");
  const long long n = (long long)B * P * ((D + 3) / 4);
  smct::fill_noise_kernel<<<(int)std::min<long long>((n + 255) / 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(
      seed, which, t, b_global0, B, P, D, out);
  SMCT_LAUNCH_CHECK("fill_noise_kernel");
  return SMCT_OK;
}

int smct_fill_uniform(uint64_t seed, int32_t t, int64_t b_global0, int32_t B, int32_t P, double* out, void* stream) {
  if (B < 1 || P < 1 || !out) return fail(SMCT_ERR_INVALID, "smct_fill_uniform: bad arguments");
  const long long n = (long long)B * P;
  smct::fill_uniform_kernel<<<(int)std::min<long long>((n + 255) / 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(
      seed, t, b_global0, B, P, out);
  SMCT_LAUNCH_CHECK("fill_uniform_kernel");
  return SMCT_OK;
}

int smct_sumsq_scaled(int64_t rows, int32_t D, const float* n, const float* logvar, int32_t logvar_dim, double* out,
                      void* stream) {
  if (rows < 1 || D < 1 || !n || !logvar || !out || (logvar_dim != 1 && logvar_dim != D))
    return fail(SMCT_ERR_INVALID, "smct_sumsq_scaled: bad arguments");
  const long long total = rows * D;
  smct::sumsq_scaled_kernel<<<(int)std::min<long long>((total + 255) / 256, 148 * 8), 256, 0, (cudaStream_t)stream>>>(
      rows, D, n, logvar, logvar_dim, out);
  SMCT_LAUNCH_CHECK("sumsq_scaled_kernel");
  return SMCT_OK;
}

int smct_encode(const smct_shape* shape, const smct_params* params, const void* x_in, float* X, float* X_ln,
                void* stream) {
  if (int rc = check_shape(shape)) return rc;
  if (!params || !x_in || (!X && !X_ln)) return fail(SMCT_ERR_INVALID, "smct_encode: bad arguments");
  const smct_shape& s = *shape;
  smct::RowsArgs ra;
  memset(&ra, 0, sizeof(ra));
  ra.rows = (long long)s.B * s.S; ra.D = s.D; ra.Fx = s.F_x; ra.input_mode = s.input_mode; ra.do_ln = X_ln ? 1 : 0;
  ra.sqrtD = sqrtf((float)s.D); ra.ln_eps = s.ln_eps; ra.x_in = x_in;
  ra.W_in = params->W_in; ra.b_in = params->b_in; ra.ln_g = params->ln1_g; ra.ln_b = params->ln1_b;
  ra.X = X; ra.xln = X_ln;
  return launch_rows(ra, (cudaStream_t)stream);
}

size_t smct_backward_workspace_bytes(const smct_shape* shape) {
  if (check_shape(shape) != SMCT_OK) return 0;
  if (check_backward_shape(*shape) != SMCT_OK) return 0;
  return backward_ws(*shape);
}

int smct_backward(const smct_shape* shape, const smct_params* params, const smct_io* io, const smct_grads* grads,
                  float inv_n, int32_t loss_variant, void* workspace, size_t workspace_bytes, void* stream) {
  if (int rc = check_shape(shape)) return rc;
  const smct_shape& s = *shape;
  if (int rc = check_backward_shape(s)) return rc;
  if (!params || !io || !grads) return fail(SMCT_ERR_INVALID, "params/io/grads is NULL");
  if (!io->x_in || !io->y || !io->K || !io->V || !io->i_out)
    return fail(SMCT_ERR_INVALID, "smct_backward needs io->x_in, y, K, V (final states) and i_out from the forward");
  if (!(inv_n > 0.f)) return fail(SMCT_ERR_INVALID, "inv_n must be > 0");
  const bool multi = s.logvar_dim > 1;
  if (multi && (grads->logvar_k || grads->logvar_q || grads->logvar_v || grads->logvar_z) &&
      !(io->noise_K && io->noise_q && io->noise_V && io->noise_z))
    return fail(SMCT_ERR_INVALID, "per-dimension log-variance gradients need the forward's recorded noises: io->noise_K, noise_q, noise_V, noise_z");
  if (!workspace || workspace_bytes < backward_ws(s)) return fail(SMCT_ERR_WORKSPACE, "workspace %zu < required %zu", workspace_bytes, backward_ws(s));
  cudaStream_t st = (cudaStream_t)stream;
  const int B = s.B, P = s.P, S = s.S, D = s.D;
  const size_t BSD = (size_t)B * S * D, BPS = (size_t)B * P * S;
  const smct::GradLayout lay = make_layout(s);
  Carver c{(char*)workspace, 0, workspace_bytes};
  float* X = c.take<float>(BSD); float* xln = c.take<float>(BSD); float* q_ = c.take<float>(BSD);
  float* kx = c.take<float>(BSD); float* vx = c.take<float>(BSD);
  float* WzT = c.take<float>((size_t)D * D); float* W1T = c.take<float>((size_t)D * s.dff); float* W2T = c.take<float>((size_t)D * s.dff);
  float* WqT = c.take<float>((size_t)D * D); float* WkT = c.take<float>((size_t)D * D); float* WvT = c.take<float>((size_t)D * D);
  unsigned short* anc = c.take<unsigned short>(BPS);
  int* node_cnt = c.take<int>((size_t)B * S + 1);
  int* node_off = c.take<int>((size_t)B * S + 1);
  int* node_bs = c.take<int>(BPS);
  float* node_m = c.take<float>(BPS);
  unsigned short* node_p = c.take<unsigned short>(BPS);
  unsigned short* node_L = c.take<unsigned short>(BPS);
  unsigned short* node_of = c.take<unsigned short>(BPS);
  float* Zm = c.take<float>(BPS * D); float* dZm = c.take<float>(BPS * D); float* Qm = c.take<float>(BPS * D); float* stat = c.take<float>(BPS * 2);
  float* sums = c.take<float>(6 * BSD);   // contiguous: one memset
  float* rep = c.take<float>((size_t)BWD_NREP * lay.total);
  float* flat = c.take<float>((size_t)lay.total);
  double* colacc = c.take<double>((size_t)4 * 256);
  // `sums` was carved as one block; re-derive the six (B,S,D) views without alignment gaps
  float* dq_sum = sums; float* dk_sum = sums + BSD; float* dv_sum = sums + 2 * BSD; float* dxln_sum = sums + 3 * BSD;
  float* dkx_sum = sums + 4 * BSD; float* dvx_sum = sums + 5 * BSD;
  SMCT_CUDA(cudaMemsetAsync(sums, 0, 6 * BSD * sizeof(float), st));
  SMCT_CUDA(cudaMemsetAsync(rep, 0, (size_t)BWD_NREP * lay.total * sizeof(float), st));
  SMCT_CUDA(cudaMemsetAsync(flat, 0, (size_t)lay.total * sizeof(float), st));

  smct::RowsArgs ra;
  memset(&ra, 0, sizeof(ra));
  ra.rows = (long long)B * S; ra.D = D; ra.Fx = s.F_x; ra.input_mode = s.input_mode; ra.do_ln = 1;
  ra.sqrtD = sqrtf((float)D); ra.ln_eps = s.ln_eps; ra.x_in = io->x_in;
  ra.W_in = params->W_in; ra.b_in = params->b_in; ra.ln_g = params->ln1_g; ra.ln_b = params->ln1_b;
  ra.Wq = params->Wq; ra.bq = params->bq; ra.Wk = params->Wk; ra.bk = params->bk; ra.Wv = params->Wv; ra.bv = params->bv;
  ra.X = X; ra.xln = xln; ra.q_ = q_; ra.kx = kx; ra.vx = vx;
  if (int rc = launch_rows(ra, st)) return rc;
  smct::transpose_kernel<<<64, 256, 0, st>>>(params->Wz, WzT, D, D);
  SMCT_LAUNCH_CHECK("transpose_kernel");
  smct::transpose_kernel<<<64, 256, 0, st>>>(params->W1, W1T, D, s.dff);
  SMCT_LAUNCH_CHECK("transpose_kernel");
  smct::transpose_kernel<<<64, 256, 0, st>>>(params->W2, W2T, s.dff, D);
  SMCT_LAUNCH_CHECK("transpose_kernel");
  smct::transpose_kernel<<<64, 256, 0, st>>>(params->Wq, WqT, D, D);
  smct::transpose_kernel<<<64, 256, 0, st>>>(params->Wk, WkT, D, D);
  smct::transpose_kernel<<<64, 256, 0, st>>>(params->Wv, WvT, D, D);
  SMCT_LAUNCH_CHECK("transpose_kernel");
  smct::anc_trace_kernel<<<(int)std::min<long long>(((long long)B * P + 255) / 256, 148 * 8), 256, 0, st>>>(
      B, P, S, s.len_resampling, io->i_out, anc);
  SMCT_LAUNCH_CHECK("anc_trace_kernel");
  {  // unique (slot, ancestor) nodes of the final genealogy: count, scan, fill
    const size_t smem_n = (size_t)2 * P * sizeof(int);
    static PerDeviceOnce node_once;
    if (node_once.need()) {
      SMCT_CUDA(cudaFuncSetAttribute(smct::node_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 12288 * (int)sizeof(int)));
      SMCT_CUDA(cudaFuncSetAttribute(smct::node_fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 12288 * (int)sizeof(int)));
    }
    smct::node_count_kernel<<<B * S, 256, smem_n, st>>>(P, anc, node_cnt);
    SMCT_LAUNCH_CHECK("node_count_kernel");
    smct::node_scan_kernel<<<1, 1024, 0, st>>>(B * S, node_cnt, node_off);
    SMCT_LAUNCH_CHECK("node_scan_kernel");
    SMCT_CUDA(cudaMemsetAsync(node_of, 0xFF, BPS * sizeof(unsigned short), st));
    smct::node_fill_kernel<<<B * S, 256, smem_n, st>>>(P, S, anc, node_off, node_bs, node_p, node_L, node_m, node_of);
    SMCT_LAUNCH_CHECK("node_fill_kernel");
  }

  smct::BwdArgs a;
  memset(&a, 0, sizeof(a));
  a.B = B; a.P = P; a.S = S; a.D = D; a.dff = s.dff; a.Fy = s.F_y; a.Fx = s.F_x; a.attn_window = s.attn_window;
  a.len_resampling = s.len_resampling; a.input_mode = s.input_mode; a.weights_mode = s.weights_mode; a.logvar_dim = s.logvar_dim;
  a.ln_eps = s.ln_eps; a.sigma_obs = s.sigma_obs; a.inv_n = inv_n; a.sqrtD = sqrtf((float)D);
  a.b_global0 = s.b_global0; a.seed = io->seed;
  a.x_in = io->x_in; a.X = X; a.xln = xln; a.q_ = q_; a.kx = kx; a.vx = vx;
  a.y = io->y; a.eps = io->eps; a.classic_w = io->classic_w; a.p = *params; a.WzT = WzT; a.W1T = W1T; a.W2T = W2T; a.WqT = WqT; a.WkT = WkT; a.WvT = WvT;
  a.K = io->K; a.V = io->V; a.i_out = io->i_out; a.anc = anc; a.node_off = node_off; a.node_bs = node_bs; a.node_p = node_p; a.node_L = node_L; a.node_m = node_m; a.node_of = node_of; a.Zm = Zm; a.dZm = dZm; a.Qm = Qm; a.stat = stat;
  a.dq_sum = dq_sum; a.dk_sum = dk_sum; a.dv_sum = dv_sum; a.dxln_sum = dxln_sum; a.dkx_sum = dkx_sum; a.dvx_sum = dvx_sum;
  a.rep = rep; a.nrep = BWD_NREP; a.lay = lay;
  int rc;
  if (D <= 32) rc = launch_backward_t<1>(a, st);
  else if (D <= 64) rc = launch_backward_t<2>(a, st);
  else rc = launch_backward_t<4>(a, st);
  if (rc) return rc;
  const size_t smem_r = ((size_t)10 * D + 32) * sizeof(float);
  smct::bwd_rows_kernel<<<(int)std::min<long long>((long long)B * S, BWD_NREP), 128, smem_r, st>>>(a);
  SMCT_LAUNCH_CHECK("bwd_rows_kernel");
  smct::reduce_replicas_kernel<<<(lay.total + 255) / 256, 256, 0, st>>>(rep, BWD_NREP, lay.total, flat);
  SMCT_LAUNCH_CHECK("reduce_replicas_kernel");
  smct::ScatterArgs sa;
  memset(&sa, 0, sizeof(sa));
  sa.lay = lay; sa.D = D; sa.dff = s.dff; sa.Fy = s.F_y; sa.win_elems = (s.input_mode == SMCT_INPUT_ENCODED) ? 0 : s.F_x * D;
  sa.flat = flat; sa.loss_sums = io->loss_sums; sa.inv_n = inv_n;
  sa.logdet_term = (loss_variant == SMCT_LOSS_CHECKOUT) ? 0.5f * D * (float)((double)B * P * S * inv_n) : 0.f;
  sa.g = *grads;
  sa.scalar_logvar = multi ? 0 : 1;
  smct::scatter_grads_kernel<<<64, 256, 0, st>>>(sa);
  SMCT_LAUNCH_CHECK("scatter_grads_kernel");
  if (multi && io->noise_K) {   // per-dimension log-variance gradients from the forward's recorded noises
    SMCT_CUDA(cudaMemsetAsync(colacc, 0, (size_t)4 * 256 * sizeof(double), st));
    const long long rows = (long long)B * P * S;
    const float* nz[4] = {io->noise_K, io->noise_q, io->noise_V, io->noise_z};
    const float* lv[4] = {params->logvar_k, params->logvar_q, params->logvar_v, params->logvar_z};
    float* gv[4] = {grads->logvar_k, grads->logvar_q, grads->logvar_v, grads->logvar_z};
    const float logdet_d = (loss_variant == SMCT_LOSS_CHECKOUT) ? 0.5f * (float)((double)rows * inv_n) : 0.f;
    for (int w = 0; w < 4; ++w) {
      if (!gv[w]) continue;
      const int per = std::max(1, 256 / D);
      smct::colsumsq_kernel<<<(int)std::min<long long>((rows + per - 1) / per, 148LL * 8), 256, 0, st>>>(rows, D, nz[w], colacc + 256 * w);
      SMCT_LAUNCH_CHECK("colsumsq_kernel");
      smct::logvar_grad_multi_kernel<<<(D + 127) / 128, 128, 0, st>>>(D, colacc + 256 * w, lv[w], inv_n, logdet_d, gv[w]);
      SMCT_LAUNCH_CHECK("logvar_grad_multi_kernel");
    }
  }
  return SMCT_OK;
}

int smct_adam_step(int64_t n, float* param, const float* grad, float* m, float* v, float lr_t, float beta1, float beta2,
                   float epsilon, void* stream) {
  if (n < 1 || !param || !grad || !m || !v) return fail(SMCT_ERR_INVALID, "smct_adam_step: bad arguments");
  smct::adam_kernel<<<(int)std::min<long long>((n + 255) / 256, 148 * 8), 256, 0, (cudaStream_t)stream>>>(
      n, param, grad, m, v, lr_t, beta1, beta2, epsilon);
  SMCT_LAUNCH_CHECK("adam_kernel");
  return SMCT_OK;
}

int smct_tc_selftest(const float* A, const float* W, float* D, void* scratch, void* stream) {
  if (!A || !W || !D || !scratch) return fail(SMCT_ERR_INVALID, "smct_tc_selftest: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  smct::tc_weight_prep_kernel<<<1, 1024, 0, st>>>(W, nullptr, nullptr, nullptr, nullptr, nullptr, 0, (unsigned char*)scratch);
  SMCT_LAUNCH_CHECK("tc_weight_prep_kernel");
  const int smem = 2 * smct::tc::A_PART_BYTES + 2 * smct::tc::B_PART_BYTES + 1024;
  static PerDeviceOnce attr_once;
  if (attr_once.need()) {
    SMCT_CUDA(cudaFuncSetAttribute(smct::tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  }
  smct::tc_selftest_kernel<<<1, 128, smem, st>>>(A, (const unsigned char*)scratch, D);
  SMCT_LAUNCH_CHECK("tc_selftest_kernel");
  return SMCT_OK;
}

int smct_classic_loss(int32_t B, int32_t P, int32_t S, int32_t F_y, int32_t weights_mode, const float* pred,
                      const void* y, const float* mask, double* out, void* stream) {
  if (B < 1 || P < 1 || S < 1 || F_y < 1 || !pred || !y || !out) return fail(SMCT_ERR_INVALID, "smct_classic_loss: bad arguments");
  const long long rows = (long long)B * P * S;
  smct::classic_loss_kernel<<<(int)std::min<long long>((rows + 255) / 256, 148 * 8), 256, 0, (cudaStream_t)stream>>>(
      B, P, S, F_y, weights_mode, pred, y, mask, out);
  SMCT_LAUNCH_CHECK("classic_loss_kernel");
  return SMCT_OK;
}

int smct_metric(int32_t B, int32_t P, int32_t S, int32_t F_y, int32_t weights_mode, const float* pred, const void* y,
                const float* mask, double* out, void* stream) {
  if (B < 1 || P < 1 || S < 1 || F_y < 1 || !pred || !y || !out) return fail(SMCT_ERR_INVALID, "smct_metric: bad arguments");
  const long long rows = (long long)B * S;
  smct::metric_kernel<<<(int)std::min<long long>((rows + 7) / 8, 148 * 8), 256, 0, (cudaStream_t)stream>>>(
      B, P, S, F_y, weights_mode, pred, y, mask, out);
  SMCT_LAUNCH_CHECK("metric_kernel");
  return SMCT_OK;
}

int smct_em_update(int32_t B, int64_t per_b, const float* noise_k, const float* noise_q, const float* noise_v,
                   const float* noise_z, float* logvar_k, float* logvar_q, float* logvar_v, float* logvar_z, float it,
                   float em_param, void* scratch, void* stream) {
  if (B < 1 || per_b < 1 || !noise_k || !noise_q || !noise_v || !noise_z || !logvar_k || !logvar_q || !logvar_v || !logvar_z || !scratch)
    return fail(SMCT_ERR_INVALID, "smct_em_update: bad arguments");
  if (!(it > 0.f)) return fail(SMCT_ERR_INVALID, "smct_em_update: it must be > 0");
  cudaStream_t st = (cudaStream_t)stream;
  smct::em_err_kernel<<<4 * B, 256, 0, st>>>(B, per_b, noise_k, noise_q, noise_v, noise_z, (double*)scratch);
  SMCT_LAUNCH_CHECK("em_err_kernel");
  smct::em_apply_kernel<<<1, 32, 0, st>>>(B, (const double*)scratch, logvar_k, logvar_q, logvar_v, logvar_z, it, em_param);
  SMCT_LAUNCH_CHECK("em_apply_kernel");
  return SMCT_OK;
}

}  // extern "C"

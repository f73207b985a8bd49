// Parameter blocks and launcher prototypes shared by the kernel translation units and api.cu.
#pragma once
#include "common.cuh"

namespace volt {

enum { KIND_DENSE = 0, KIND_VOL = 1, KIND_BM = 2 };
enum { MA_EWMA = 0, MA_DEWMA = 1, MA_TEWMA = 2, MA_MEANREVERT = 3, MA_GIVEN = 4 };
constexpr int NSCALARS = 16;
constexpr int NSERIES = 8;  // floats per series handed to the rollout: u.u, u.z1, V[n-1], dx, jitter, 0, 0, 0

// Series-sharded loss without a collective launch (volt_mll_step_sharded).  Every rank owns ring * world 64-bit slots in
// peer-mapped memory; a slot word is {step number : 32 | float bits : 32}, self-contained, so plain system-scope stores and
// loads suffice (no ordering against other data).  The last CTA of the step's kernel stores this rank's partial into slot
// [seq % ring][rank] of EVERY rank's buffer and, when `totals` is given, sums the slots of step seq - lag of its own buffer
// (they arrived while the steps in between ran) into totals[(seq - lag) % ring] -- in rank order, so every rank gets the
// same bits.  lag = 1 ties every rank to the slowest one step by step; lag = 2 lets them drift by a whole step.
struct LossExchange {
  const unsigned long long* peers;   // device array of `world` pointers, entry r = rank r's slots
  const unsigned long long* mine;    // this rank's slots
  float* totals;                     // ring floats, or nullptr (no earlier step to sum yet)
  int world, rank, ring, lag;        // the kernel of step seq sums step seq - lag (lag >= 1)
  unsigned int seq;
};

// value of slot `src` once it carries step `seq`; NaN after ~2 s (a peer that never got there)
__device__ __forceinline__ float exchange_wait_slot(const unsigned long long* src, unsigned int seq) {
  for (long long spin = 0; spin < 10000000LL; ++spin) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(src) : "memory");
    if ((unsigned int)(v >> 32) == seq) return __uint_as_float((unsigned int)v);
    __nanosleep(200);
  }
  return __int_as_float(0x7fc00000);
}

// total of step `seq` from this rank's slots; called by one whole warp, result valid on every lane
__device__ __forceinline__ float exchange_sum_warp(const unsigned long long* mine, int world, int ring, unsigned int seq, int lane) {
  float tot = 0.f;
  for (int base = 0; base < world; base += 32) {
    float val = 0.f;
    if (base + lane < world) val = exchange_wait_slot(mine + (size_t)(seq % (unsigned)ring) * world + base + lane, seq);
    const int n = (world - base < 32) ? world - base : 32;
    for (int i = 0; i < n; ++i) tot += __shfl_sync(0xffffffffu, val, i);
  }
  return tot;
}

// called by one whole warp at the very end of the step (value = this rank's partial, same on every lane)
__device__ __forceinline__ void exchange_partial_warp(const LossExchange& e, float value, int lane) {
  const unsigned long long v = ((unsigned long long)e.seq << 32) | (unsigned long long)__float_as_uint(value);
  for (int r = lane; e.peers && r < e.world; r += 32) {
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(e.peers[r]) + (size_t)(e.seq % (unsigned)e.ring) * e.world + e.rank;
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(dst), "l"(v) : "memory");
  }
  if (e.totals) {
    const unsigned int old = e.seq - (unsigned)e.lag;
#ifdef VOLT_EXCHANGE_DEBUG
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
#endif
    const float tot = exchange_sum_warp(e.mine, e.world, e.ring, old, lane);
    if (lane == 0) e.totals[old % (unsigned)e.ring] = tot;
#ifdef VOLT_EXCHANGE_DEBUG
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (lane == 0) e.totals[e.ring] += (float)(t1 - t0) * 1e-3f;   // microseconds spent waiting for the peers' slots (totals has ring + 1 entries)
#endif
  }
}

struct MllParams {
  int kind, B, T, Tp, nb;
  const float* dense; long long dense_bstride; int ldd;
  const float* V;                       // (B,T) cumtrapz prefix (KIND_VOL)
  const float* x;                       // (T) grid (KIND_BM)
  const float* scale; int scale_stride; // BM vol (per series or shared)
  const float* diag_add; int diag_stride;
  const float* resid;                   // (B,T) or null
  const float* resid2;                  // (B,T) or null: second right-hand side (forward substitution only)
  float* z_out;                         // (B,2,T) or null: z1 = L^-1 resid, z2 = L^-1 resid2
  float* scalars;                       // (B,16): see VOLT_S_* in include/volt_b200.h
  float* alpha;                         // (B,T) or null
  int* info;                            // (B) or null
  float* L_out; long long L_bstride; int ldl;
  float* U_out;                         // optional (B,T,T) contiguous: (L^-1)^T, upper triangular (tensor-core kernel, needs do_inverse)
  int do_inverse;
  float jitter; int max_tries;
  float* scratch;                       // gridDim.x * Tp * Tp
  float* dinv;                          // gridDim.x * nb * 64 * 64
  // tensor-core batched kernel only: prefix built in-kernel from the raw volatility path (V == nullptr), and series
  // b >= ready_from wait until *ready != 0 (their inputs are still in flight on a copy stream; volt_mll_grad_vol_host)
  const float* vol_in; const float* x_in; int x_batched, vol_mode;
  const int* ready; int ready_from;
  int* ready_timeout; long long ready_spins;   // give up after ready_spins polls: *ready_timeout = 1, the CTA stops (the host re-runs)
  // fused training-step epilogue (tensor-core batched kernel; volt_mll_grad_vol_raw): the diagonal term is the GaussianLikelihood
  // noise softplus(raw_noise) + 1e-4 evaluated in-kernel, scalars[VOLT_S_DRAW] = dMLL/draw_noise, scalars[VOLT_S_NOISE] = noise,
  // and the last CTA to finish writes loss_out[0] = -sum_b MLL_b, summed in a fixed order (bitwise reproducible).
  const float* raw_noise; int raw_stride;
  float* loss_out; unsigned int* done_counter;
  // rollout prep (tensor-core batched kernel): per-series hand-over to the rollout kernel, which is launched as a programmatic
  // dependent and starts on a series as soon as its flag is raised -- pack_out (B, NSERIES): u.u, u.z1, V[n-1], dx, jitter;
  // series_flag[b] = 1 (release) once pack_out / info of series b are written
  float* pack_out; const float* pack_x; int* series_flag;
  LossExchange ex;   // series-sharded job: push the partial loss to the peers (ex.peers != nullptr)
  int stage_in;      // host-buffer entry on mapped pinned memory: x_in / vol_in / resid are host pointers -- one bulk read per series
};

// [GPyTorch] GaussianLikelihood / GreaterThan(1e-4): noise = softplus(raw_noise) + 1e-4 (torch's softplus: beta 1, threshold 20)
__device__ __forceinline__ float noise_from_raw_dev(float raw) { return (raw > 20.f ? raw : log1pf(expf(raw))) + 1e-4f; }
__device__ __forceinline__ float sigmoid_dev(float raw) { return 1.f / (1.f + expf(-raw)); }

// CumTrapz of one series by one warp (VolKernel.py:4-10).  torch.cumsum on CPU accumulates float32 data in a double
// accumulator and rounds each prefix to float32; same here (per-lane sequential double sums + a warp scan of the lane
// totals), so the result equals the reference's to the last bit except for double-rounding ties.  `out` may be shared
// or global memory.
__device__ __forceinline__ void cumtrapz_warp(const float* __restrict__ xs, const float* __restrict__ vs, int T, int mode, int half_last,
                                              float* out, int lane) {
  const float dx = xs[1] - xs[0];
  const float w_end = dx * 0.5f;
  const int seg = (T + 31) / 32;
  const int lo = min(lane * seg, T), hi = min(lo + seg, T);
  double s = 0.0;
  for (int i = lo; i < hi; ++i) {
    float v = vs[i];
    if (mode == 2) v = expf(v);
    const float y = mode ? v * v : v;
    float w = (i == 0 || (half_last && i == T - 1)) ? w_end : dx;
    if (T == 1 && half_last) w = dx * 0.25f;
    s += (double)(w * y);
  }
  double incl = s;   // exclusive warp scan of the lane totals
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  double acc = incl - s;
  for (int i = lo; i < hi; ++i) {
    float v = vs[i];
    if (mode == 2) v = expf(v);
    const float y = mode ? v * v : v;
    float w = (i == 0 || (half_last && i == T - 1)) ? w_end : dx;
    if (T == 1 && half_last) w = dx * 0.25f;
    acc += (double)(w * y);
    out[i] = (float)acc;
  }
}

struct RolloutParams {
  int B, n, S, H, k, mean_kind, joint;   // joint = 1: one-shot multi-point draw (no feedback of samples)
  const float* w;          // (k) EWMA weights
  const float* ytrain;     // (B,n) log prices
  const float* e_train;    // (B,n+1) EWMA path of ytrain       (MA kinds)
  const float* ee_train;   // (B,n+1) EWMA(e)[:-1]              (DEWMA/TEWMA)
  const float* mean_test;  // (B,H) parametric test means       (MA_GIVEN)
  const float* series;     // (B,NSERIES)
  const int* series_info;  // (B) potrf info of the shared block
  const float* pred_vol;   // (B,S,H)
  const float* eps;        // (B,S,H) or null
  const float* latent;     // (B) rollout-level mean reversion target (theta) or null
  float theta; int use_theta;
  const float* mr_latent;  // (B) MeanRevertingEMAMean.latent_mean
  float mr_theta;
  float jitter;            // psd_safe_cholesky jitter for pred_cov (1e-4 in rollout_utils, 1e-6 default)
  unsigned long long seed;
  float* samples;          // (B,S,H)
  int* info;               // (B,S) bit0: per-draw pivot failure, bit1: pred_cov needed jitter, bit2: not PSD after jitter
  int Hp;                  // padded tile row length (odd)
  const int* series_flag;  // (B) or null: wait until series_flag[b] != 0 before reading series / series_info (prep still running)
  int b_offset;            // global index of series 0 of this launch (batches > 65535 series are launched in chunks)
};

int launch_mll_batched(MllParams p, cudaStream_t st);      // dispatches on the selected implementation
int launch_mll_batched_simt(MllParams p, cudaStream_t st); // fp32 CUDA-core GEMM micro-kernel (chol_batched.cu)
int launch_mll_batched_tc(MllParams p, cudaStream_t st);   // tcgen05 3xTF32 tensor-core products (chol_tc.cu)
int mll_tc_resident_ctas(int T, int two_rhs);              // series in flight per launch of that kernel
int launch_mll_large(const MllParams& p, int b, cudaStream_t st);  // multi-CTA path for one long series (chol_large.cu)
int launch_rollout(RolloutParams p, cudaStream_t st);
int launch_rollout_normals(unsigned long long seed, int b_offset, int B, int S, int H, int joint, float* out, cudaStream_t st);
int launch_gemm_nt(const float* A, long long lda, long long a_bstride, const float* B, long long ldb, long long b_bstride, float* C,
                   long long ldc, long long c_bstride, int M, int N, int K, int batch, int mode, int tri, int max_ctas, cudaStream_t st);
int launch_gpcv_rows(const float* chol_var, const float* W, const float* var_mean, const float* y, const float* gh_t, const float* gh_w,
                     int nq, int B, int n, float inv_n, float* grad_chol, float* rows, cudaStream_t st);
int launch_adam(float* p, const float* g, float* m, float* v, long long count, float lr, float beta1, float beta2, float eps, int step,
                const float* step_dev, cudaStream_t st);
int launch_rollout_stats(const float* samples, int B, int S, int H, const float* truth, const float* strike, int exp_flag, float* ecdf,
                         float* mean, float* sd, float* nll, float* payoff, cudaStream_t st);
int launch_ma_paths(const float* y, int S, int T, int k, const float* w, int kind, float theta, const float* latent, float* out,
                    float* e_out, float* ee_out, float* resid_out, cudaStream_t st);
int launch_cumtrapz(const float* x, int x_batched, const float* vol, int B, int T, int mode, int half_last, float* V,
                    cudaStream_t st);
int launch_vol_cov(const float* V, const float* add_diag, int add_stride, int B, int T, float* K, cudaStream_t st);
int launch_bm_cov(const float* x1, int n1, const float* x2, int n2, const float* vol, float* K, cudaStream_t st);
int launch_ewma_weights(int k, float* w_dev, cudaStream_t st);
int launch_ewma(const float* y, int S, int T, int k, const float* w_dev, float* out, cudaStream_t st);
int launch_chol_solve(const float* L, long long l_bstride, int ldl, int B, int T, float* rhs, long long r_bstride, int nrhs,
                      int mode, cudaStream_t st);
int launch_posterior(const float* W, int B, int T, int H, const float* Kss, const float* mean_s, float* mean, float* cov,
                     cudaStream_t st);
int launch_bm_posterior_pack(const float* x, int B, int T, const float* xs, int H, const float* y, const float* vol,
                             int vol_stride, float* W0, float* Kss, float* mean_s, float* resid, cudaStream_t st);
int launch_mvn_sample(const float* mean, const float* Lc, const float* eps, int B, int H, int S, int exp_out, float* samples,
                      cudaStream_t st);

}  // namespace volt

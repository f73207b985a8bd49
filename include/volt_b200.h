/* volt_b200 -- C ABI of the B200-native Volt GP hot path (libvolt_b200.so).
 *
 * The reference (g-benton/Volt) has no FFI: its boundary for this path is a set of Python callables that
 * delegate to torch / GPyTorch.  Each entry point below names the reference interface it replaces
 * (paths relative to the reference repository root).  The Python host layer (volt_b200/*.py) binds these
 * with ctypes and re-creates the voltron.* API on top; INTEGRATION.md shows the stub a maintainer of the
 * reference would add.
 *
 * Conventions
 *   - plain C, no torch types.  All `const float*` / `float*` / `int*` arguments are DEVICE pointers unless the
 *     function name ends in `_host`, in which case they are HOST pointers and the call performs the
 *     host<->device copies itself and synchronises before returning.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Device-pointer entry points only
 *     enqueue work.  They synchronise in two cases only: when a library-owned scratch buffer has to grow (first call at a
 *     larger size), and on the multi-CTA path for a few very long series (T >= 1536 and B <= 16), which reads one
 *     4-byte failure flag back per factorisation attempt.
 *   - scratch memory is owned by the library: one arena per (device, stream), grown on demand and reused by every call on
 *     that stream.  Work on one stream is ordered, so calls issued to the same stream never overlap on the scratch; calls
 *     on different streams of one device, from one or several host threads, use different arenas and may run
 *     concurrently.  The `_host` entry points run on library-owned streams with their own arenas, so they may be mixed
 *     freely with device-pointer calls in flight.  volt_release_workspaces() frees every arena of the current device.
 *   - row-major, contiguous, float32 (the reference's precision).  Batched arrays put the series index first.
 *   - return value: 0 on success, <0 on error (VOLT_ERR_*); volt_last_error() returns the message.
 *     Numerical failure (matrix not positive definite) is NOT an error return: it is reported per matrix in `info`
 *     exactly like torch.linalg.cholesky_ex (1-based order of the first non-positive leading minor, 0 = success).
 *   - an empty batch (B == 0 / S == 0 series) is a successful no-op for every batched entry point: no pointer is
 *     dereferenced and nothing is launched (empty shards of a small job); negative sizes are argument errors.
 *   - requires an sm_100 device; there is no CPU fallback.
 */
#ifndef VOLT_B200_H_
#define VOLT_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define VOLT_ABI_VERSION 1

#define VOLT_OK 0
#define VOLT_ERR_ARG (-1)
#define VOLT_ERR_CUDA (-2)
#define VOLT_ERR_ARCH (-3)
#define VOLT_ERR_ALLOC (-4)

/* number of floats per series in the `scalars` output of the MLL entry points */
#define VOLT_NSCALARS 16
/* scalars[b*16 + i]: */
#define VOLT_S_MLL 0       /* -1/2 (r^T A^-1 r + logdet A + T log 2pi) / T   (GPyTorch ExactMarginalLogLikelihood) */
#define VOLT_S_DNOISE 1    /* d MLL / d noise = 1/2 (alpha.alpha - tr A^-1) / T                                     */
#define VOLT_S_LOGDET 2
#define VOLT_S_INVQUAD 3
#define VOLT_S_TRINV 4     /* tr(A^-1) = ||L^-1||_F^2 */
#define VOLT_S_ALAL 5      /* alpha.alpha, alpha = A^-1 r; d MLL / d r = -alpha / T */
#define VOLT_S_ALR 6       /* alpha.r */
#define VOLT_S_JITTER 7    /* jitter the psd_safe_cholesky policy had to add (0 if none) */
#define VOLT_S_Z2Z2 8      /* |L^-1 rhs2|^2   (second right-hand side, used by the rollout) */
#define VOLT_S_Z1Z2 9      /* (L^-1 r).(L^-1 rhs2) */
#define VOLT_S_DRAW 10     /* d MLL / d raw_noise = VOLT_S_DNOISE * sigmoid(raw_noise)   (volt_mll_grad_vol_raw only) */
#define VOLT_S_NOISE 11    /* noise = softplus(raw_noise) + 1e-4                          (volt_mll_grad_vol_raw only) */

/* moving-average mean families (voltron/means/EWMA.py) */
#define VOLT_MA_EWMA 0
#define VOLT_MA_DEWMA 1
#define VOLT_MA_TEWMA 2
#define VOLT_MA_MEANREVERT 3
#define VOLT_MA_GIVEN 4    /* parametric mean evaluated by the caller (constant / linear / log-linear) */

/* input encodings of a volatility path */
#define VOLT_VOL_SIGMA 1      /* sigma      -> integrand sigma^2            */
#define VOLT_VOL_LOGSIGMA 2   /* log sigma  -> integrand exp(log sigma)^2   (model.log_vol_path.exp(), rollout_utils.py:7) */
#define VOLT_VOL_RAW 0        /* integrand given directly (CumTrapz(y, x))  */

const char* volt_last_error(void);
int volt_abi_version(void);
/* 0 if the current CUDA device is sm_100 (B200); VOLT_ERR_ARCH otherwise. */
int volt_device_check(void);
/* number of kernels launched by this library since load (bench.py's gpu_launches counter) */
long long volt_launch_count(void);
/* Select the batched Cholesky/MLL kernel: 1 = tcgen05 3xTF32 tensor-core products (default), 0 = fp32 CUDA-core
 * products (kept for A/B measurement; both are CUDA paths).  Also settable with VOLT_MLL_IMPL=tc|simt.  Returns the
 * previous setting (-1 if never set). */
int volt_set_mll_impl(int impl);
/* Free every cached scratch arena of the current device (synchronises the device first).  The reference's counterpart is
 * torch.cuda.empty_cache() at voltron/rollout_utils.py:49,92; the arenas are re-created on demand by the next call. */
int volt_release_workspaces(void);

/* CumTrapz(y, x)  -- voltron/kernels/VolKernel.py:4-10.
 * x (T) or (B,T) if x_batched; y (B,T) encoded per `vol_mode`; half_last=1 reproduces the reference weights
 * dx*[1/2,1,...,1,1/2]; half_last=0 keeps the last weight at dx (prefix of a longer grid, used by the rollout). */
int volt_cumtrapz(const float* x, int x_batched, const float* y, int B, int T, int vol_mode, int half_last, float* V,
                  void* stream);

/* VolatilityKernel.forward(x, vol_path)  -- voltron/kernels/VolKernel.py:18-41.
 * K[b,i,j] = V[b,min(i,j)] (+ add_diag[b*add_stride] on the diagonal when add_diag != NULL, the likelihood noise of
 * [GPyTorch] GaussianLikelihood.__call__).  K is (B,T,T). */
int volt_vol_cov(const float* x, int x_batched, const float* vol, int vol_mode, int B, int T, const float* add_diag,
                 int add_stride, float* K, void* stream);

/* BMKernel.forward(x1, x2)  -- voltron/kernels/BMKernel.py:38-51.  K[i,j] = vol[0] * min(x1_i, x2_j), K is (n1,n2). */
int volt_bm_cov(const float* x1, int n1, const float* x2, int n2, const float* vol, float* K, void* stream);

/* EWMA(y, k)  -- voltron/means/EWMA.py:20-37.  y (S,T) -> out (S,T+1). */
int volt_ewma(const float* y, int S, int T, int k, float* out, void* stream);

/* EWMAMean / DEWMAMean / TEWMAMean / MeanRevertingEMAMean full-length paths  -- voltron/means/EWMA.py:39-135.
 * y (S,T) -> out (S,T+1).  Optional outputs (NULL to skip): e_out, ee_out (S,T+1) intermediate EWMA paths,
 * resid_out (S,T) = y - out[:, :T] (the training residual y - mean_module(train_x)).
 * latent (S): MeanRevertingEMAMean.latent_mean, only read for VOLT_MA_MEANREVERT. */
int volt_ma_mean(const float* y, int S, int T, int k, int kind, float theta, const float* latent, float* out, float* e_out,
                 float* ee_out, float* resid_out, void* stream);

/* One exact MLL + gradient evaluation per series for the data model (Volatility kernel):
 *   loss = -mll(model(train_x), y); loss.backward()   -- voltron/train_utils.py:247-250 (and :134-137),
 *   [GPyTorch] ExactMarginalLogLikelihood / MultivariateNormal.log_prob / psd_safe_cholesky.
 * A_b = K(x, vol_b) + noise_b I is generated inside the factorisation (never written to HBM).
 * resid (B,T) = y - mean; noise (B) with stride noise_stride (0 = shared scalar).
 * jitter / max_tries: psd_safe_cholesky policy (GPyTorch default 1e-6 / 3); jitter <= 0 disables the retry.
 * Outputs: scalars (B,VOLT_NSCALARS), alpha (B,T) or NULL, info (B) or NULL. */
int volt_mll_grad_vol(const float* x, int x_batched, const float* vol, int vol_mode, const float* resid, const float* noise,
                      int noise_stride, int B, int T, float jitter, int max_tries, float* scalars, float* alpha, int* info,
                      void* stream);

/* The whole training step of the data model in one launch:  voltron/train_utils.py:247-250 with the likelihood's parameter
 * transform ([GPyTorch] GaussianLikelihood: noise = softplus(raw_noise) + 1e-4, the transform `voltron.likelihood.raw_noise`
 * goes through at train_utils.py:222,249) and the reduction to the scalar loss folded into the kernel.
 * raw_noise (B) with stride raw_stride (0 = shared).  Besides the outputs of volt_mll_grad_vol:
 *   scalars[b][VOLT_S_DRAW] = d MLL_b / d raw_noise_b, scalars[b][VOLT_S_NOISE] = noise_b,
 *   loss_out[0] (or NULL) = -sum_b MLL_b, summed in a fixed order by the last CTA to finish (bitwise reproducible): the
 *   rank-local partial of the series-sharded loss, all-reduced by the caller (volt_b200/batched.py).
 * An empty batch writes loss_out[0] = 0. */
int volt_mll_grad_vol_raw(const float* x, int x_batched, const float* vol, int vol_mode, const float* resid, const float* raw_noise,
                          int raw_stride, int B, int T, float jitter, int max_tries, float* scalars, float* alpha, int* info,
                          float* loss_out, void* stream);

/* volt_mll_grad_vol_raw for a series-sharded job (one process per GPU, each rank holds B of the series): the compute step and
 * the exchange of the loss in ONE kernel, no collective launch.  Every rank owns an exchange buffer of ring*world 64-bit
 * slots in peer-mapped memory; after the fixed-order sum, the kernel's last CTA stores {seq, -sum_b MLL_b} (one 64-bit
 * word) into slot [seq % ring][rank] of EVERY rank's buffer over NVLink, and -- when prev_totals is given -- adds up the
 * slots of step seq-lag of its own buffer (they arrived while the steps in between ran), in rank order so every rank gets
 * the same bits: prev_totals[(seq-lag) % ring] = total loss of step seq-lag.  lag = 1 (volt_b200.batched) keeps the ranks
 * in lock step: no step ends before every rank has finished the one before it; lag = 2 lets them drift by a whole step.
 *   peer_slot_ptrs: DEVICE array of `world` pointers, entry r = rank r's buffer mapped into this process
 *                   (torch.distributed._symmetric_memory buffer_ptrs_dev, or cudaIpc / cuMem mappings); slots start zeroed.
 *                   NULL: this call does not publish its partial (see volt_loss_push), it only sums step seq-lag.
 *   local_slots:    this rank's own buffer (needed with prev_totals).   prev_totals: ring floats or NULL (seq <= lag).
 *   seq:            step number, the same on every rank, 1, 2, 3, ...; 1 <= lag < ring, and no rank may run more than
 *                   ring-1 steps ahead of another (prev_totals bounds it to lag).
 * Replaces the scalar the reference builds from the loss on one process (voltron/train_utils.py:249-250); the reference
 * has no multi-process path.  An empty shard (B = 0) pushes 0.
 * volt_loss_gather: total of step seq from this rank's slots (for the newest steps, which no later kernel has summed yet);
 * waits on the device, bounded (~2 s, then NaN), for slots that have not arrived.  out: 1 float. */
int volt_mll_step_sharded(const float* x, int x_batched, const float* vol, int vol_mode, const float* resid, const float* raw_noise,
                          int raw_stride, int B, int T, float jitter, int max_tries, float* scalars, float* alpha, int* info,
                          float* loss_out, const void* peer_slot_ptrs, const void* local_slots, float* prev_totals, int lag, int world,
                          int rank, int ring, unsigned int seq, void* stream);
int volt_loss_gather(const void* local_slots, int world, int ring, unsigned int seq, float* out, void* stream);
/* The push alone, from a value in device memory: stores {seq, *value} into slot [seq % ring][rank] of every rank's buffer
 * (a one-warp kernel; small enough to run next to the resident CTAs of the next step).  With it a caller keeps remote
 * stores out of the step kernel: volt_mll_step_sharded(peer_slot_ptrs = NULL, local_slots, prev_totals, ...) only sums, and
 * this call, on another stream behind an event, publishes loss_out. */
int volt_loss_push(const float* value, const void* peer_slot_ptrs, int world, int rank, int ring, unsigned int seq, void* stream);

/* Same for the vol model BMGP (A = scale_b * min(x_i, x_j) + noise_b I)  -- voltron/train_utils.py:86-90,
 * voltron/models/BMGP.py:20-28.  x (T) shared grid. */
int volt_mll_grad_bm(const float* x, const float* scale, int scale_stride, const float* resid, const float* noise,
                     int noise_stride, int B, int T, float jitter, int max_tries, float* scalars, float* alpha, int* info,
                     void* stream);

/* volt_mll_grad_bm that also returns the inverse factor: linv_t (B,T,T) = (L^-1)^T (upper triangular), L the Cholesky
 * factor of scale*min(x,x') + noise*I.  The GPCV stage needs K^-1 times a T x T matrix every iteration
 * ([GPyTorch] kl_mvn_mvn inside VariationalELBO, voltron/train_utils.py:46-56): K^-1 M = linv_t (linv_t^T M), two GEMMs. */
int volt_mll_grad_bm_inv(const float* x, const float* scale, int scale_stride, const float* resid, const float* noise,
                         int noise_stride, int B, int T, float jitter, int max_tries, float* scalars, float* alpha, int* info,
                         float* linv_t, void* stream);

/* Same for an explicit dense covariance K (B,T,ld) (lower triangle read) + noise_b I  -- the generic
 * MultivariateNormal.log_prob path used when a caller hands over an evaluated kernel matrix. */
int volt_mll_grad_dense(const float* K, long long k_bstride, int ld, const float* resid, const float* noise, int noise_stride,
                        int B, int T, float jitter, int max_tries, float* scalars, float* alpha, int* info, void* stream);

/* Host-buffer variant of volt_mll_grad_vol (all pointers are HOST memory; pinned memory recommended).  Synchronous: the
 * outputs are valid on return.  The copies are inside the call: for a large batch the inputs of the first series are
 * copied, ONE kernel is launched for the whole batch, and the remaining inputs follow on a copy stream underneath it
 * (later series wait in-kernel for an arrival flag written by the copy engine); the CumTrapz prefix is built in-kernel. */
int volt_mll_grad_vol_host(const float* x, const float* vol, const float* resid, const float* noise, int noise_stride, int B,
                           int T, float jitter, int max_tries, float* scalars, float* alpha, int* info);

/* psd_safe_cholesky(A, jitter) / torch.linalg.cholesky_ex  -- voltron/rollout_utils.py:35,46; VoltMagpie.py:87,92.
 * A (B,T,lda) lower triangle read (+ add_diag); L (B,T,ldl) lower factor, strict upper triangle zeroed.
 * jitter_used (B) or NULL. */
int volt_potrf(const float* A, long long a_bstride, int lda, const float* add_diag, int add_stride, int B, int T, float jitter,
               int max_tries, float* L, long long l_bstride, int ldl, float* jitter_used, int* info, void* stream);

/* torch.cholesky_solve(rhs, L)  -- voltron/rollout_utils.py:36,44.  rhs (B,T,nrhs) overwritten with the solution;
 * forward_only=1 stops after L^-1 rhs. */
int volt_potrs(const float* L, long long l_bstride, int ldl, int B, int T, float* rhs, long long r_bstride, int nrhs,
               int forward_only, void* stream);

/* BMGP in eval mode: model.vol_model(test_x)  -- voltron/rollout_utils.py:66, voltron/models/BMGP.py:18-28,
 * [GPyTorch] ExactGP.__call__ / DefaultPredictionStrategy.  x (T), y (B,T) = log vol, xs (H).
 * Outputs mean (B,H), cov (B,H,H), info (B). */
int volt_bmgp_posterior(const float* x, const float* y, int B, int T, const float* xs, int H, const float* vol, int vol_stride,
                        const float* noise, int noise_stride, float* mean, float* cov, int* info, void* stream);

/* MultivariateNormal.sample(): samples[b,s,:] = mean[b] + psd_safe_cholesky(cov[b]) eps[b,:,s]; eps is (B,H,S).
 * exp_out=1 applies exp() (pred_vol = ...sample(...).exp(), rollout_utils.py:66). */
int volt_mvn_sample(const float* mean, const float* cov, const float* eps, int B, int H, int S, float jitter, int exp_out,
                    float* samples, int* info, void* stream);

/* GeneratePrediction + Rollouts  -- voltron/rollout_utils.py:6-93 (joint=0), and the multi-point draw of
 * rollout_utils.py:6-53 / voltron/models/VoltMagpie.py:67-99 (joint=1).
 *   x (n) training grid (uniform; only dx = x[1]-x[0] enters, as in CumTrapz), logy (B,n) log prices on that grid,
 *   vol (B,n) volatility path encoded per vol_mode, pred_vol (B,S,H) sigma draws for the H test points,
 *   eps (B,S,H) base normals or NULL (in-kernel Philox seeded by `seed`),
 *   mean_kind: VOLT_MA_* ; k window; mr_theta / mr_latent (B): MeanRevertingEMAMean parameters;
 *   resid_given (B,n), mean_test (B,H): required for VOLT_MA_GIVEN (y - mean(train_x), mean(test_x));
 *   use_theta / theta / latent (B): the rollout-level mean reversion of rollout_utils.py:41-42;
 *   jitter: psd_safe_cholesky jitter (1e-4 in rollout_utils.py:35,46; 1e-6 in VoltMagpie.py:87,92).
 * Outputs samples (B,S,H) log prices; draw_info (B,S) bit flags (1: non-positive pivot in the per-draw rows,
 * 2: pred_cov needed jitter, 4: pred_cov not PSD after 3 tries); series_info (B) cholesky_ex info of the shared block.
 * A draw with bit 1 set has to be re-run by the caller as its own series, one step at a time (the reference jitters that
 * draw's whole matrix, rollout_utils.py:35); volt_b200.ops.rollout does so and marks repaired draws with bit 8. */
int volt_rollout(const float* x, const float* logy, const float* vol, int vol_mode, const float* pred_vol, const float* eps,
                 int B, int n, int S, int H, int mean_kind, int k, float mr_theta, const float* mr_latent,
                 const float* resid_given, const float* mean_test, int use_theta, float theta, const float* latent, int joint,
                 float jitter, unsigned long long seed, float* samples, int* draw_info, int* series_info, void* stream);

/* The base normals volt_rollout draws itself when eps == NULL (same Philox4x32-10 counters: key = seed, counter =
 * (series, draw, step)), written as an eps tensor (B,S,H): volt_rollout(..., eps = this, ...) reproduces
 * volt_rollout(..., eps = NULL, seed, ...) bit for bit.  Lets a caller re-run single flagged draws (draw_info bit 1, see
 * above) with the numbers the kernel used.  joint: the mode flag of volt_rollout (the two modes index the generator differently). */
int volt_rollout_normals(unsigned long long seed, int B, int S, int H, int joint, float* eps, void* stream);

/* Forecast evaluation reductions over a rollout tensor, one pass over samples (B,S,H):
 *   ecdf[b,h]   = #{s : v < truth[b,h]} / S      voltron/option_utils.py:48-52 (ECDF; the caller passes log prices and the
 *                                                 log of the realised price) and the weather calibration notebook's
 *                                                 ECDF(sample, truth) = sum(sample < truth, 0) / S (calib_plotter cell 2)
 *   mean, sd    = samples.mean(0), samples.std(0) (unbiased), nll = -Normal(mean, sd).log_prob(truth)   (cell 15, GetNLL)
 *   payoff[b,h] = mean_s max(v - strike[b,h], 0)  voltron/option_utils.py:37 (Pricer's Monte-Carlo call valuation)
 * with v = samples[b,s,h], or exp(samples[b,s,h]) when exp_flag != 0 (the notebooks' exp=True).  truth / strike are (B,H)
 * or NULL; every output is (B,H) or NULL (nll needs truth, payoff needs strike). */
int volt_rollout_stats(const float* samples, int B, int S, int H, const float* truth, const float* strike, int exp_flag,
                       float* ecdf, float* mean, float* sd, float* nll, float* payoff, void* stream);

/* GPCV stage (LearnGPCV, voltron/train_utils.py:15-67): per-row terms of the variational ELBO
 * and the gradient of the Cholesky variational factor, for B series of n points.
 *   chol_var (B,n,n): CholeskyVariationalDistribution parameter, lower triangle used (single_task_variational_gp.py:86-88)
 *   W        (B,n,n): K^-1 tril(chol_var), K = BM kernel + 1e-3 I (prior of the UnwhitenedVariationalStrategy)
 *   var_mean, y (B,n): variational mean and the scaled returns (train_utils.py:16-18)
 *   gh_t, gh_w (nq <= 128): Gauss-Hermite nodes / weights (train_utils.py:52 uses 75)
 * grad_chol (B,n,n) = d(-ELBO)/d chol_var with ELBO scaled by inv_n = 1/n (VariationalELBO, train_utils.py:46);
 * rows (B,n,6) = E_q[log p(y_i|f_i)], its derivative in the variational mean, sum_k L[i,k] W[i,k], sum_k W[i,k]^2,
 * log|L[i,i]|, S_ii -- the host layer sums them into the loss and the gradients of the mean / kernel parameters. */
int volt_gpcv_rows(const float* chol_var, const float* W, const float* var_mean, const float* y, const float* gh_t,
                   const float* gh_w, int nq, int B, int n, float inv_n, float* grad_chol, float* rows, void* stream);

/* Batched product C_z (= | -=) A_z B_z^T with fp32-equivalent accuracy on the tensor cores (3xTF32), A_z (M,K), B_z (N,K),
 * C_z (M,N) row-major; ld*: row strides, *_bstride: floats between batch members (all multiples of 4 floats, pointers
 * 16-byte aligned).  Replaces the two matrix products behind K^-1 L_S in the GPCV stage's KL term ([GPyTorch] kl_mvn_mvn
 * inside VariationalELBO, voltron/train_utils.py:46-56, which the reference leaves to torch.matmul) and serves the deferred
 * trailing updates of the long-series factorisation.  subtract != 0: C -= A B^T (the update is applied in the L2 by a TMA
 * reduction; C is never loaded by a thread). */
int volt_gemm_nt(const float* A, long long lda, long long a_bstride, const float* B, long long ldb, long long b_bstride, float* C,
                 long long ldc, long long c_bstride, int M, int N, int K, int batch, int subtract, void* stream);

/* One torch.optim.Adam step (no weight decay / amsgrad) over a flat parameter buffer: the optimiser of every training
 * loop of the reference (train_utils.py:38-41 lr 0.01; :76-78, :123-125, :236-238 lr 0.01).  step counts from 1; when
 * step_dev is not NULL the count is read from that device float instead (the caller increments it before each launch),
 * so that the launch can be replayed from a CUDA graph. */
int volt_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long count, float lr, float beta1,
                   float beta2, float eps, int step, const float* step_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VOLT_B200_H_ */

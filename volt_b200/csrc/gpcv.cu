// GPCV stage: the variational GP on the log-volatility that produces the vol path the hot path
// consumes -- LearnGPCV, voltron/train_utils.py:15-67, with
//   q(u) = N(m, L_S L_S^T)                 CholeskyVariationalDistribution (models/single_task_variational_gp.py:86-88)
//   p(y_i | f_i) = N(0, max(exp f_i, 1e-3)^2)   VolatilityGaussianLikelihood(param="exp") (likelihoods/volatility_likelihood.py:44-52)
//   E_q[log p(y_i | f_i)] by 75-point Gauss-Hermite quadrature (train_utils.py:52, [GPyTorch] GaussHermiteQuadrature1D)
//   KL(q || p) with p(u) = N(c, K + 1e-3 I), K the BM kernel ([GPyTorch] kl_mvn_mvn).
// The T^3 pieces (factorisation of K, tr K^-1, K^-1 (c - m)) come from the batched MLL kernel; this file holds the
// per-row kernel that turns them into the ELBO terms and the gradient of the T x T variational factor, and the Adam
// update over the flat parameter buffer (torch.optim.Adam semantics, train_utils.py:38-41,59).
#include "params.cuh"

namespace volt {

// One CTA per (row i, series b).  Inputs: chol_var (B,n,n; lower triangle used), W = K^-1 tril(chol_var) (B,n,n),
// var_mean, y (B,n), Gauss-Hermite nodes / weights (nq <= 128).
// Outputs: grad_chol[b,i,k] = (-2 gs_i L[i,k] + W[i,k] - [k == i] / L[i,i]) * inv_n for k <= i, 0 above the diagonal
//          (d(-ELBO)/dL_S: likelihood term through S_ii = sum_k L[i,k]^2, KL term 0.5 tr(K^-1 S) - 0.5 logdet S);
//          rows[b,i,0..5] = E_i, dE_i/dm_i, sum_{k<=i} L[i,k] W[i,k], sum_k W[i,k]^2, log|L[i,i]|, S_ii.
__global__ void __launch_bounds__(128) gpcv_rows_kernel(const float* __restrict__ chol_var, const float* __restrict__ W,
                                                        const float* __restrict__ var_mean, const float* __restrict__ y,
                                                        const float* __restrict__ gh_t, const float* __restrict__ gh_w, int nq, int n,
                                                        float inv_n, float* __restrict__ grad_chol, float* __restrict__ rows) {
  __shared__ float red[4][8];
  const int i = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const size_t ro = ((size_t)b * n + i) * n;
  const float* Lr = chol_var + ro;
  const float* Wr = W + ro;
  float s = 0.f, lw = 0.f, ww = 0.f;
  for (int k = tid; k < n; k += 128) {
    const float wv = Wr[k];
    ww = fmaf(wv, wv, ww);
    if (k <= i) {
      const float lv = Lr[k];
      s = fmaf(lv, lv, s);
      lw = fmaf(lv, wv, lw);
    }
  }
  auto block3 = [&](float& a, float& c, float& d) {
    a = warp_sum(a); c = warp_sum(c); d = warp_sum(d);
    if (lane == 0) { red[w][0] = a; red[w][1] = c; red[w][2] = d; }
    __syncthreads();
    a = red[0][0] + red[1][0] + red[2][0] + red[3][0];
    c = red[0][1] + red[1][1] + red[2][1] + red[3][1];
    d = red[0][2] + red[1][2] + red[2][2] + red[3][2];
    __syncthreads();
  };
  block3(s, lw, ww);
  // ---- Gauss-Hermite: f = sqrt(2 S_ii) t + m_i;  l(f) = -log(scale) - 0.5 log 2pi - 0.5 (y / scale)^2, scale = max(e^f, 1e-3)
  const float m = var_mean[(size_t)b * n + i], yi = y[(size_t)b * n + i];
  const float sd2 = sqrtf(2.f * s);
  float e = 0.f, gm = 0.f, gs = 0.f;
  if (tid < nq) {
    const float t = gh_t[tid], wq = gh_w[tid];
    const float f = fmaf(sd2, t, m);
    const float ef = expf(f);
    const bool clamped = !(ef > 1e-3f);
    const float scale = clamped ? 1e-3f : ef;
    const float r = yi / scale;
    const float ll = -logf(scale) - 0.91893853320467274178f - 0.5f * r * r;
    const float dl = clamped ? 0.f : (r * r - 1.f);          // d l / d f
    e = wq * ll;
    gm = wq * dl;
    gs = wq * dl * t;
  }
  block3(e, gm, gs);
  const float isp = 0.56418958354775628695f;                  // 1 / sqrt(pi)
  e *= isp; gm *= isp;
  gs = (sd2 > 0.f) ? gs * isp / sd2 : 0.f;                    // dE/dS_ii = sum w l' t / sqrt(2 S_ii) / sqrt(pi)
  const float lii = Lr[i];
  for (int k = tid; k < n; k += 128) {
    float g = 0.f;
    if (k <= i) {
      g = fmaf(-2.f * gs, Lr[k], Wr[k]);
      if (k == i) g -= 1.f / lii;
      g *= inv_n;
    }
    grad_chol[ro + k] = g;
  }
  if (tid == 0) {
    float* o = rows + ((size_t)b * n + i) * 6;
    o[0] = e; o[1] = gm; o[2] = lw; o[3] = ww; o[4] = logf(fabsf(lii)); o[5] = s;
  }
}

// torch.optim.Adam (no weight decay, no amsgrad): p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps).
// The step count t comes from the host (step_size / inv_sqrt_bc2 precomputed) or, for CUDA-graph replay, from a device
// float the caller increments before every launch (step_dev != NULL).
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            long long count, float step_size, float beta1, float beta2, float inv_sqrt_bc2, float eps, float lr,
                            const float* __restrict__ step_dev) {
  if (step_dev) {
    const float t = *step_dev;
    step_size = lr / (1.f - powf(beta1, t));
    inv_sqrt_bc2 = rsqrtf(1.f - powf(beta2, t));
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i];
    const float mi = fmaf(beta1, m[i], (1.f - beta1) * gi);
    const float vi = fmaf(beta2, v[i], (1.f - beta2) * gi * gi);
    m[i] = mi; v[i] = vi;
    p[i] -= step_size * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
  }
}

int launch_gpcv_rows(const float* chol_var, const float* W, const float* var_mean, const float* y, const float* gh_t, const float* gh_w,
                     int nq, int B, int n, float inv_n, float* grad_chol, float* rows, cudaStream_t st) {
  for (int b0 = 0; b0 < B; b0 += 65535) {
    const int nb = min(65535, B - b0);
    const size_t o2 = (size_t)b0 * n * n, o1 = (size_t)b0 * n;
    dim3 grid(n, nb);
    gpcv_rows_kernel<<<grid, 128, 0, st>>>(chol_var + o2, W + o2, var_mean + o1, y + o1, gh_t, gh_w, nq, n, inv_n, grad_chol + o2,
                                           rows + o1 * 6);
  }
  return check_cuda(cudaGetLastError(), "gpcv_rows_kernel");
}

int launch_adam(float* p, const float* g, float* m, float* v, long long count, float lr, float beta1, float beta2, float eps, int step,
                const float* step_dev, cudaStream_t st) {
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  long long blocks = (count + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  adam_kernel<<<(int)blocks, 256, 0, st>>>(p, g, m, v, count, step_dev ? 0.f : (float)(lr / bc1), beta1, beta2,
                                           step_dev ? 0.f : (float)(1.0 / sqrt(bc2)), eps, lr, step_dev);
  return check_cuda(cudaGetLastError(), "adam_kernel");
}

}  // namespace volt

// Exact-GP prediction helpers on top of a batched Cholesky factor (not on the throughput-critical path:
// called once per series per forecast).
//
//   chol_solve_kernel     torch.cholesky_solve (voltron/rollout_utils.py:36,44) / GPyTorch prediction strategy
//   posterior_kernel      [GPyTorch] ExactGP.__call__ in eval mode (voltron/models/BMGP.py:23-28, rollout_utils.py:66):
//                         mean* = m* + K*^T (K+s2 I)^-1 (y-m),  cov* = K** - K*^T (K+s2 I)^-1 K*
//   mvn_sample_kernel     [GPyTorch] MultivariateNormal.rsample: mean + chol(cov) eps
#include "params.cuh"

namespace volt {

// In-place solve with the lower factor L (B,T,ldl) of rhs (B,T,nrhs) row-major.  One CTA per matrix, one thread per
// right-hand-side column (adjacent threads touch adjacent columns: coalesced rhs traffic, broadcast L loads).
// mode 0: L L^T x = b (potrs);  mode 1: L x = b (forward substitution only).
__global__ void chol_solve_kernel(const float* __restrict__ L, long long l_bstride, int ldl, int T, float* __restrict__ rhs,
                                  long long r_bstride, int nrhs, int mode) {
  const float* Lb = L + (size_t)blockIdx.x * l_bstride;
  float* R = rhs + (size_t)blockIdx.x * r_bstride;
  for (int j = threadIdx.x; j < nrhs; j += blockDim.x) {
    for (int i = 0; i < T; ++i) {
      const float* Li = Lb + (size_t)i * ldl;
      float s = R[(size_t)i * nrhs + j];
      for (int k = 0; k < i; ++k) s = fmaf(-Li[k], R[(size_t)k * nrhs + j], s);
      R[(size_t)i * nrhs + j] = s / Li[i];
    }
    if (mode == 0) {
      for (int i = T - 1; i >= 0; --i) {
        float s = R[(size_t)i * nrhs + j];
        for (int k = i + 1; k < T; ++k) s = fmaf(-Lb[(size_t)k * ldl + i], R[(size_t)k * nrhs + j], s);
        R[(size_t)i * nrhs + j] = s / Lb[(size_t)i * ldl + i];
      }
    }
  }
}

// W = [L^-1 Kxs | L^-1 resid] (B,T,H+1) already forward-substituted.  mean[h] = mean_s[h] + W[:,h].v,
// cov[h,g] = Kss[h,g] - W[:,h].W[:,g].  grid = (B), 256 threads.
__global__ void __launch_bounds__(256) posterior_kernel(const float* __restrict__ W, int T, int H, const float* __restrict__ Kss,
                                                        const float* __restrict__ mean_s, float* __restrict__ mean,
                                                        float* __restrict__ cov) {
  const int b = blockIdx.x;
  const float* Wb = W + (size_t)b * T * (H + 1);
  const int total = H * H + H;
  for (int o = threadIdx.x; o < total; o += blockDim.x) {
    if (o < H * H) {
      const int h = o / H, g = o - h * H;
      float acc = 0.f;
      for (int i = 0; i < T; ++i) acc = fmaf(Wb[(size_t)i * (H + 1) + h], Wb[(size_t)i * (H + 1) + g], acc);
      cov[(size_t)b * H * H + o] = Kss[(size_t)b * H * H + o] - acc;
    } else {
      const int h = o - H * H;
      float acc = 0.f;
      for (int i = 0; i < T; ++i) acc = fmaf(Wb[(size_t)i * (H + 1) + h], Wb[(size_t)i * (H + 1) + H], acc);
      mean[(size_t)b * H + h] = mean_s[(size_t)b * H + h] + acc;
    }
  }
}

// BM-kernel GP posterior inputs (voltron/models/BMGP.py:20-28, kernels/BMKernel.py:40-41):
//   W0[i, h] = vol * min(x_i, xs_h) (h < H), W0[i, H] = y_i - (-1/2 vol^2 x_i);  Kss[h,g] = vol*min(xs_h, xs_g);
//   mean_s[h] = -1/2 vol^2 xs_h.   grid = (B).
__global__ void bm_posterior_pack_kernel(const float* __restrict__ x, int T, const float* __restrict__ xs, int H,
                                         const float* __restrict__ y, const float* __restrict__ vol, int vol_stride,
                                         float* __restrict__ W0, float* __restrict__ Kss, float* __restrict__ mean_s,
                                         float* __restrict__ resid) {
  const int b = blockIdx.x;
  const float v = vol[(size_t)b * vol_stride];
  const float hv2 = -0.5f * (v * v);
  float* Wb = W0 + (size_t)b * T * (H + 1);
  for (int o = threadIdx.x; o < T * (H + 1); o += blockDim.x) {
    const int i = o / (H + 1), h = o - i * (H + 1);
    Wb[o] = (h < H) ? v * fminf(x[i], xs[h]) : y[(size_t)b * T + i] - hv2 * x[i];
  }
  for (int o = threadIdx.x; o < H * H; o += blockDim.x) {
    const int h = o / H, g = o - h * H;
    Kss[(size_t)b * H * H + o] = v * fminf(xs[h], xs[g]);
  }
  for (int h = threadIdx.x; h < H; h += blockDim.x) mean_s[(size_t)b * H + h] = hv2 * xs[h];
  if (resid)
    for (int i = threadIdx.x; i < T; i += blockDim.x) resid[(size_t)b * T + i] = y[(size_t)b * T + i] - hv2 * x[i];
}

// samples[b,s,h] = mean[b,h] + sum_{g<=h} Lc[b,h,g] eps[b,g,s]   (eps laid out (H,S) like GPyTorch's base samples)
__global__ void mvn_sample_kernel(const float* __restrict__ mean, const float* __restrict__ Lc, const float* __restrict__ eps,
                                  int H, int S, int exp_out, float* __restrict__ samples) {
  const int b = blockIdx.y;
  for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < S * H; o += gridDim.x * blockDim.x) {
    const int s = o / H, h = o - s * H;
    float acc = mean[(size_t)b * H + h];
    const float* Lr = Lc + ((size_t)b * H + h) * H;
    for (int g = 0; g <= h; ++g) acc = fmaf(Lr[g], eps[((size_t)b * H + g) * S + s], acc);
    samples[(size_t)b * S * H + o] = exp_out ? expf(acc) : acc;
  }
}

int launch_chol_solve(const float* L, long long l_bstride, int ldl, int B, int T, float* rhs, long long r_bstride, int nrhs,
                      int mode, cudaStream_t st) {
  int threads = (nrhs + 31) / 32 * 32;
  if (threads > 256) threads = 256;
  chol_solve_kernel<<<B, threads, 0, st>>>(L, l_bstride, ldl, T, rhs, r_bstride, nrhs, mode);
  return check_cuda(cudaGetLastError(), "chol_solve_kernel");
}

int launch_posterior(const float* W, int B, int T, int H, const float* Kss, const float* mean_s, float* mean, float* cov,
                     cudaStream_t st) {
  posterior_kernel<<<B, 256, 0, st>>>(W, T, H, Kss, mean_s, mean, cov);
  return check_cuda(cudaGetLastError(), "posterior_kernel");
}

int launch_bm_posterior_pack(const float* x, int B, int T, const float* xs, int H, const float* y, const float* vol,
                             int vol_stride, float* W0, float* Kss, float* mean_s, float* resid, cudaStream_t st) {
  bm_posterior_pack_kernel<<<B, 256, 0, st>>>(x, T, xs, H, y, vol, vol_stride, W0, Kss, mean_s, resid);
  return check_cuda(cudaGetLastError(), "bm_posterior_pack_kernel");
}

int launch_mvn_sample(const float* mean, const float* Lc, const float* eps, int B, int H, int S, int exp_out, float* samples,
                      cudaStream_t st) {
  int bx = (S * H + 255) / 256;
  if (bx > 1024) bx = 1024;
  dim3 grid(bx, B);
  mvn_sample_kernel<<<grid, 256, 0, st>>>(mean, Lc, eps, H, S, exp_out, samples);
  return check_cuda(cudaGetLastError(), "mvn_sample_kernel");
}

}  // namespace volt

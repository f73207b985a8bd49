// Forecast evaluation reductions over the (B, S, H) rollout tensor -- the post-processing the reference does on the CPU
// after every rollout:
//   ECDF / sample percentile   voltron/option_utils.py:48-52 (log prices sorted and counted below the realised price);
//                              experiments/weather calibration notebook, cell 2: sum(samples < truth, 0) / S
//   moment-matched Gaussian NLL  same notebook, cell 15: Normal(preds.mean(0), preds.std(0)).log_prob(truth)
//   Monte-Carlo call valuation   voltron/option_utils.py:37: mean(max(px - strike, 0)) over the draws
// One pass over the samples (HBM bound: 4 B per step-sample read, nothing written but B x H scalars).  Lane = horizon
// step, warps stride over the draws, so a warp reads H consecutive floats per draw; per-lane partial sums are kept in
// double (the moments are differences of large sums) and combined across the warps in shared memory in a fixed order.
#include "params.cuh"

namespace volt {

constexpr int STAT_WARPS = 8;

__global__ void __launch_bounds__(32 * STAT_WARPS) rollout_stats_kernel(const float* __restrict__ samples, int S, int H,
                                                                        const float* __restrict__ truth,
                                                                        const float* __restrict__ strike, int exp_flag,
                                                                        float* __restrict__ ecdf, float* __restrict__ mean,
                                                                        float* __restrict__ sd, float* __restrict__ nll,
                                                                        float* __restrict__ payoff) {
  __shared__ double s_sum[STAT_WARPS][32], s_sq[STAT_WARPS][32], s_pay[STAT_WARPS][32];
  __shared__ unsigned int s_cnt[STAT_WARPS][32];
  const int b = blockIdx.x, h = blockIdx.y * 32 + (threadIdx.x & 31), w = threadIdx.x >> 5;
  const float* sb = samples + (size_t)b * S * H;
  const bool act = h < H;
  const float tr = (act && truth) ? truth[(size_t)b * H + h] : 0.f;
  const float kk = (act && strike) ? strike[(size_t)b * H + h] : 0.f;
  double sum = 0.0, sq = 0.0, pay = 0.0;
  unsigned int cnt = 0;
  if (act) {
    auto take = [&](float v) {
      if (exp_flag) v = expf(v);
      sum += (double)v;
      sq += (double)v * (double)v;
      cnt += (v < tr) ? 1u : 0u;
      pay += (double)fmaxf(v - kk, 0.f);
    };
    int s = w;
    for (; s + 7 * STAT_WARPS < S; s += 8 * STAT_WARPS) {   // eight independent loads in flight per lane
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldcs(sb + (size_t)(s + u * STAT_WARPS) * H + h);
#pragma unroll
      for (int u = 0; u < 8; ++u) take(v[u]);
    }
    for (; s < S; s += STAT_WARPS) take(__ldcs(sb + (size_t)s * H + h));
  }
  const int l = threadIdx.x & 31;
  s_sum[w][l] = sum; s_sq[w][l] = sq; s_pay[w][l] = pay; s_cnt[w][l] = cnt;
  __syncthreads();
  if (w == 0 && act) {
    for (int q = 1; q < STAT_WARPS; ++q) { sum += s_sum[q][l]; sq += s_sq[q][l]; pay += s_pay[q][l]; cnt += s_cnt[q][l]; }
    const double n = (double)S;
    const double mu = sum / n;
    // torch.std: unbiased (S - 1); NaN for a single draw like torch
    const double var = (S > 1) ? fmax(sq - sum * mu, 0.0) / (n - 1.0) : __longlong_as_double(0x7ff8000000000000LL);
    const double sig = sqrt(var);
    const size_t o = (size_t)b * H + h;
    if (ecdf) ecdf[o] = (float)cnt / (float)S;       // torch.sum(smp < x) / S in fp32 (option_utils.py:52)
    if (mean) mean[o] = (float)mu;
    if (sd) sd[o] = (float)sig;
    if (nll && truth) {
      const double z = ((double)tr - mu) / sig;
      nll[o] = (float)(0.5 * z * z + log(sig) + 0.91893853320467274178);   // -Normal(mu, sig).log_prob(truth)
    }
    if (payoff && strike) payoff[o] = (float)(pay / n);
  }
}

int launch_rollout_stats(const float* samples, int B, int S, int H, const float* truth, const float* strike, int exp_flag, float* ecdf,
                         float* mean, float* sd, float* nll, float* payoff, cudaStream_t st) {
  for (int b0 = 0; b0 < B; b0 += 65535 * 32) {  // grid.x limit is 2^31-1; keep the loop for symmetry with the rollout launcher
    const int nb = min(B - b0, 65535 * 32);
    dim3 grid(nb, (H + 31) / 32);
    const size_t off = (size_t)b0 * H;
    rollout_stats_kernel<<<grid, 32 * STAT_WARPS, 0, st>>>(samples + (size_t)b0 * S * H, S, H, truth ? truth + off : nullptr,
                                                          strike ? strike + off : nullptr, exp_flag, ecdf ? ecdf + off : nullptr,
                                                          mean ? mean + off : nullptr, sd ? sd + off : nullptr,
                                                          nll ? nll + off : nullptr, payoff ? payoff + off : nullptr);
  }
  return check_cuda(cudaGetLastError(), "rollout_stats_kernel");
}

}  // namespace volt

// Covariance construction and moving-average kernels (HBM-bound, SIMT + 1-D bulk TMA staging).
//
//   cumtrapz_kernel   voltron/kernels/VolKernel.py:4-10   V = cumsum(w * sigma^2), w = dx*[1/2,1,..,1,1/2]
//   vol_cov_kernel    voltron/kernels/VolKernel.py:30-33  K[b,i,j] = V[b,min(i,j)] (+ add_diag[b] on the diagonal)
//   bm_cov_kernel     voltron/kernels/BMKernel.py:40-41   K[i,j] = vol * min(x1_i, x2_j)
//   ewma_kernel       voltron/means/EWMA.py:20-37         causal k-tap weighted mean, left pad = k copies of y[0]
#include "params.cuh"

namespace volt {

// ------------------------------------------------------------------------------------------ cumtrapz
// One warp per series (cumtrapz_warp, params.cuh).
__global__ void cumtrapz_kernel(const float* __restrict__ x, int x_batched, const float* __restrict__ vol,
                                int B, int T, int mode, int half_last, float* __restrict__ V) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= B) return;
  cumtrapz_warp(x + (x_batched ? (size_t)warp * T : 0), vol + (size_t)warp * T, T, mode, half_last, V + (size_t)warp * T, lane);
}

// ------------------------------------------------------------------------------------------ vol_cov
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// 1-D bulk TMA: global -> shared, completion on an mbarrier (bytes multiple of 16, 16-byte aligned).
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

constexpr int COV_ROWS = 32;      // rows of K per CTA
constexpr int COV_THREADS = 256;  // 8 warps, 4 rows each

// grid = (ceil(T/COV_ROWS), B).  V[b] is staged into shared memory with one bulk-TMA copy; every warp
// then streams whole rows with 128-bit evict-first stores (K is written once and is larger than L2).
__global__ void __launch_bounds__(COV_THREADS) vol_cov_kernel(const float* __restrict__ V, const float* __restrict__ add_diag,
                                                              int add_stride, int T, float* __restrict__ K) {
  extern __shared__ __align__(16) float sV[];
  __shared__ __align__(8) uint64_t bar;
  const int b = blockIdx.y;
  const float* Vb = V + (size_t)b * T;
  const bool tma_ok = ((T & 3) == 0) && ((reinterpret_cast<uintptr_t>(Vb) & 15) == 0);
  if (tma_ok) {
    if (threadIdx.x == 0) {
      mbar_init(&bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      mbar_expect_tx(&bar, (uint32_t)T * 4u);
      tma_bulk_g2s(sV, Vb, (uint32_t)T * 4u, &bar);
    }
    mbar_wait(&bar, 0);
  } else {
    for (int i = threadIdx.x; i < T; i += COV_THREADS) sV[i] = Vb[i];
    __syncthreads();
  }
  const float add = add_diag ? add_diag[(size_t)b * add_stride] : 0.f;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* Kb = K + (size_t)b * T * T;
  const bool vec_ok = ((T & 3) == 0) && ((reinterpret_cast<uintptr_t>(Kb) & 15) == 0);
  for (int rr = warp; rr < COV_ROWS; rr += COV_THREADS / 32) {
    const int i = blockIdx.x * COV_ROWS + rr;
    if (i >= T) break;
    const float vi = sV[i];
    float* row = Kb + (size_t)i * T;
    if (vec_ok) {
      for (int j = lane * 4; j < T; j += 128) {
        float4 o;
        if (j + 3 < i) {
          o = *reinterpret_cast<const float4*>(&sV[j]);
        } else {
          o.x = (j + 0 < i) ? sV[j + 0] : vi;
          o.y = (j + 1 < i) ? sV[j + 1] : vi;
          o.z = (j + 2 < i) ? sV[j + 2] : vi;
          o.w = (j + 3 < i) ? sV[j + 3] : vi;
          if (add != 0.f && i >= j && i < j + 4) (&o.x)[i - j] += add;
        }
        __stcs(reinterpret_cast<float4*>(row + j), o);
      }
    } else {
      for (int j = lane; j < T; j += 32) {
        float o = (j < i) ? sV[j] : vi;
        if (j == i) o += add;
        row[j] = o;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ bm_cov
__global__ void bm_cov_kernel(const float* __restrict__ x1, int n1, const float* __restrict__ x2, int n2,
                              const float* __restrict__ vol, float* __restrict__ K) {
  const float s = vol[0];
  const size_t total = (size_t)n1 * n2;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(idx / n2), j = (int)(idx % n2);
    K[idx] = s * fminf(x1[i], x2[j]);
  }
}

// ------------------------------------------------------------------------------------------ ewma
// grid = (ceil((T+1)/256), S).  Weights (k taps, oldest first) live in shared memory next to the padded
// window of the series this CTA needs: padded[u] = y[0] for u < k, y[u-k] otherwise (EWMA.py:29-30).
__global__ void __launch_bounds__(256) ewma_kernel(const float* __restrict__ y, int T, int k, const float* __restrict__ w,
                                                   float* __restrict__ out) {
  extern __shared__ float sm[];
  float* sw = sm;         // k
  float* sp = sm + k;     // 256 + k - 1 padded samples
  const int s = blockIdx.y;
  const float* ys = y + (size_t)s * T;
  const int j0 = blockIdx.x * 256;
  for (int t = threadIdx.x; t < k; t += 256) sw[t] = w[t];
  const int span = 256 + k - 1;
  for (int u = threadIdx.x; u < span; u += 256) {
    const int p = j0 + u;  // index into the padded series (length T + k)
    float v = 0.f;
    if (p < T + k) v = (p < k) ? ys[0] : ys[p - k];
    sp[u] = v;
  }
  __syncthreads();
  const int j = j0 + threadIdx.x;
  if (j > T) return;
  float acc = 0.f;
  for (int t = 0; t < k; ++t) acc = fmaf(sw[t], sp[threadIdx.x + t], acc);
  out[(size_t)s * (T + 1) + j] = acc;
}

// k-tap weights alpha*(1-alpha)^(k-1..0) / sum, evaluated in float32 like the reference's torch expression
// (EWMA.py:21-24).  One block.
__global__ void ewma_weights_kernel(int k, float* __restrict__ w) {
  __shared__ float red[32];
  const float alpha = (float)(2.0 / (k + 1.0));
  const float base = (float)(1.0 - 2.0 / (k + 1.0));
  float part = 0.f;
  for (int t = threadIdx.x; t < k; t += blockDim.x) {
    const float v = alpha * powf(base, (float)(k - 1 - t));
    w[t] = v;
    part += v;
  }
  const float tot = block_sum(part, red);
  for (int t = threadIdx.x; t < k; t += blockDim.x) w[t] = w[t] / tot;
}

// ------------------------------------------------------------------------------------------ launchers
int launch_cumtrapz(const float* x, int x_batched, const float* vol, int B, int T, int mode, int half_last, float* V,
                    cudaStream_t st) {
  const int threads = 128;
  const int blocks = (B * 32 + threads - 1) / threads;
  cumtrapz_kernel<<<blocks, threads, 0, st>>>(x, x_batched, vol, B, T, mode, half_last, V);
  return check_cuda(cudaGetLastError(), "cumtrapz_kernel");
}

int launch_vol_cov(const float* V, const float* add_diag, int add_stride, int B, int T, float* K, cudaStream_t st) {
  const size_t smem = ((size_t)T * 4 + 15) / 16 * 16;
  if (smem > 200 * 1024) {
    set_error("vol_cov: T=%d too large for the shared-memory staged builder", T);
    return VOLT_ERR_ARG;
  }
  if (smem > 48 * 1024) {
    int s = check_cuda(cudaFuncSetAttribute(vol_cov_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                       "cudaFuncSetAttribute(vol_cov_kernel)");
    if (s) return s;
  }
  for (int b0 = 0; b0 < B; b0 += 65535) {
    const int nb = min(65535, B - b0);
    dim3 grid((T + COV_ROWS - 1) / COV_ROWS, nb);
    vol_cov_kernel<<<grid, COV_THREADS, smem, st>>>(V + (size_t)b0 * T, add_diag ? add_diag + (size_t)b0 * add_stride : nullptr,
                                                    add_stride, T, K + (size_t)b0 * T * T);
  }
  return check_cuda(cudaGetLastError(), "vol_cov_kernel");
}

int launch_bm_cov(const float* x1, int n1, const float* x2, int n2, const float* vol, float* K, cudaStream_t st) {
  const size_t total = (size_t)n1 * n2;
  size_t want = (total + 255) / 256, cap = (size_t)sm_count() * 8;
  int blocks = (int)(want < cap ? want : cap);
  if (blocks < 1) blocks = 1;
  bm_cov_kernel<<<blocks, 256, 0, st>>>(x1, n1, x2, n2, vol, K);
  return check_cuda(cudaGetLastError(), "bm_cov_kernel");
}

int launch_ewma_weights(int k, float* w_dev, cudaStream_t st) {
  ewma_weights_kernel<<<1, 256, 0, st>>>(k, w_dev);
  return check_cuda(cudaGetLastError(), "ewma_weights_kernel");
}

int launch_ewma(const float* y, int S, int T, int k, const float* w_dev, float* out, cudaStream_t st) {
  const size_t smem = (size_t)(k + 256 + k) * sizeof(float);
  if (smem > 200 * 1024) {
    set_error("ewma: window k=%d too large", k);
    return VOLT_ERR_ARG;
  }
  if (smem > 48 * 1024) {
    int s = check_cuda(cudaFuncSetAttribute(ewma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                       "cudaFuncSetAttribute(ewma_kernel)");
    if (s) return s;
  }
  for (int s0 = 0; s0 < S; s0 += 65535) {
    const int ns = min(65535, S - s0);
    dim3 grid((T + 1 + 255) / 256, ns);
    ewma_kernel<<<grid, 256, smem, st>>>(y + (size_t)s0 * T, T, k, w_dev, out + (size_t)s0 * (T + 1));
  }
  return check_cuda(cudaGetLastError(), "ewma_kernel");
}

}  // namespace volt

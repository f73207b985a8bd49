// Batched dense Cholesky / exact-MLL kernel, one CTA per series (persistent over the batch).
//
// Replaces, for B independent series of length T (fp32, dense, exact):
//   [GPyTorch] ExactMarginalLogLikelihood -> MultivariateNormal.log_prob -> psd_safe_cholesky / triangular solve /
//   log-det (call sites voltron/train_utils.py:89,136,249) and the autograd backward of the same chain
//   (train_utils.py:90,137,250), plus torch.linalg.cholesky_ex as used by voltron/rollout_utils.py:35.
//
// Algorithm per series (Tp = T rounded up to 64, padded with an identity block):
//   Phase A  left-looking blocked potrf, NB = 64 columns per step, rows processed in 128-row chunks:
//              S = A[rows, j] - L[rows, 0:j] L[j, 0:j]^T          (register-tiled fp32 GEMM, operands staged in smem)
//              L_jj = chol(S_jj) (shared memory), Linv_jj = L_jj^-1 (recursive doubling, shared memory)
//              L[rows>j, j] = S[rows>j] Linv_jj^T                  (same GEMM micro-kernel)
//              z_j = Linv_jj (r_j - L[j, 0:j] z_0:j)               (forward substitution fused into the sweep)
//   Phase B  in-place row-oriented trtri into the UPPER triangle of the same scratch (U = X^T, X = L^-1):
//              U[0:i, i] = -(U[0:i, 0:i] L[i, 0:i]^T) Linv_ii^T    (both products "TN": K-contiguous operands)
//            accumulating tr(A^-1) = ||X||_F^2 and alpha = X^T z on the fly.
//   Outputs  MLL = -1/2 (z.z + 2 sum log L_ii + T log 2pi)/T, dMLL/dnoise = 1/2 (alpha.alpha - tr A^-1)/T, alpha, info.
// The matrix A is generated on the fly (vol kernel: V[min(i,j)], BM kernel: s*min(x_i,x_j)) or read from a dense
// buffer, so in the fused path K never exists in HBM; L/U live in a per-CTA scratch that stays L2-resident.
#include "chol_dev.cuh"

#include <cmath>

namespace volt {

// ---------------------------------------------------------------------------------------------- the kernel
__global__ void __launch_bounds__(NT, 2) mll_batched_kernel(MllParams p) {
  extern __shared__ __align__(16) float smem_f[];
  Smem sm;
  sm.As = smem_f;
  sm.Bs = sm.As + 2 * BK * AS_LD;
  sm.Ct = sm.Bs + 2 * BK * BS_LD;
  sm.LiT = sm.Ct + NB * CT_LD;
  sm.Vs = sm.LiT + NB * LI_LD;
  sm.z = sm.Vs + p.Tp;
  sm.al = sm.z + p.Tp;
  sm.z2 = sm.al + p.Tp;
  sm.diagl = sm.z2 + p.Tp;
  sm.tmp = sm.diagl + NB;
  sm.red = sm.tmp + 2 * NB;
  sm.flag = reinterpret_cast<int*>(sm.red + 32);

  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int T = p.T, Tp = p.Tp, nb = p.nb, ld = p.Tp;
  float* S = p.scratch + (size_t)blockIdx.x * Tp * Tp;
  float* dinv = p.dinv + (size_t)blockIdx.x * nb * NB * NB;

  for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
    // ---- per-series vectors
    for (int i = tid; i < Tp; i += NT) {
      float v = 0.f;
      if (i < T) {
        if (p.kind == KIND_VOL) v = p.V[(size_t)b * T + i];
        else if (p.kind == KIND_BM) v = p.x[i];
      }
      sm.Vs[i] = v;
    }
    const float sc = (p.kind == KIND_BM) ? p.scale[(size_t)b * p.scale_stride] : 1.f;
    const float dadd0 = p.diag_add ? p.diag_add[(size_t)b * p.diag_stride] : 0.f;
    const float* rb = p.resid ? p.resid + (size_t)b * T : nullptr;
    const float* rb2 = p.resid2 ? p.resid2 + (size_t)b * T : nullptr;

    int fail = 0;          // 1-based index of the failing leading minor, 0 = success
    float jit_used = 0.f;
    float logdet_part = 0.f;
    for (int attempt = 0;; ++attempt) {
      const float dadd = dadd0 + jit_used;
      logdet_part = 0.f;
      if (tid == 0) *sm.flag = -1;
      for (int i = tid; i < Tp; i += NT) { sm.z[i] = 0.f; sm.al[i] = 0.f; sm.z2[i] = 0.f; }
      __syncthreads();
      fail = 0;
      // =============================== Phase A: potrf + forward substitution
      for (int j = 0; j < nb; ++j) {
        const int R0 = j * NB;
        const int nch = (Tp - R0 + CM - 1) / CM;
        for (int ch = 0; ch < nch; ++ch) {
          const int r_base = R0 + ch * CM;
          float acc[8][4];
#pragma unroll
          for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[i][q] = 0.f;
          gemm_tn<false>(acc, S, ld, r_base, Tp, R0, 0, R0, nullptr, sm.As, sm.Bs);
          // epilogue: Ct[c][r] = gen - acc
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int gc = R0 + tx * 4 + q;
            float o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int gr = r_base + ty * 8 + i;
              o[i] = (gr < Tp) ? gen_entry(p, b, gr, gc, sm.Vs, sc, dadd) - acc[i][q] : 0.f;
            }
            float* dst = sm.Ct + (tx * 4 + q) * CT_LD + ty * 8;
            *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<float4*>(dst + 4) = make_float4(o[4], o[5], o[6], o[7]);
          }
          __syncthreads();
          if (ch == 0) {
            potrf64<CT_LD>(sm.Ct, sm.diagl, sm.flag, R0);
            trtri64<CT_LD>(sm.Ct, sm.LiT, sm.As);
            if (tid < NB && R0 + tid < T) logdet_part += logf(sm.diagl[tid]);
            // L_jj -> scratch (lower part), LiT -> dinv[j]
            for (int idx = tid; idx < NB * NB; idx += NT) {
              const int r = idx >> 6, c = idx & 63;
              S[(size_t)(R0 + r) * ld + R0 + c] = (c <= r) ? sm.Ct[c * CT_LD + r] : 0.f;
              dinv[((size_t)j * NB + r) * NB + c] = sm.LiT[r * LI_LD + c];
            }
            if (rb) {
              // z_j = Linv_jj (r_j - L[j,0:j] z)   (and the same for the optional second right-hand side)
              const int c = tid >> 2, part = tid & 3;
              float s = 0.f, s2 = 0.f;
              const float* Lrow = S + (size_t)(R0 + c) * ld;
              for (int k = part * 4; k < R0; k += 16) {
                const float4 lv = *reinterpret_cast<const float4*>(Lrow + k);
                s = fmaf(lv.x, sm.z[k], s); s = fmaf(lv.y, sm.z[k + 1], s);
                s = fmaf(lv.z, sm.z[k + 2], s); s = fmaf(lv.w, sm.z[k + 3], s);
                if (rb2) {
                  s2 = fmaf(lv.x, sm.z2[k], s2); s2 = fmaf(lv.y, sm.z2[k + 1], s2);
                  s2 = fmaf(lv.z, sm.z2[k + 2], s2); s2 = fmaf(lv.w, sm.z2[k + 3], s2);
                }
              }
              s += __shfl_xor_sync(0xffffffffu, s, 1);
              s += __shfl_xor_sync(0xffffffffu, s, 2);
              s2 += __shfl_xor_sync(0xffffffffu, s2, 1);
              s2 += __shfl_xor_sync(0xffffffffu, s2, 2);
              if (part == 0) {
                sm.tmp[c] = ((R0 + c < T) ? rb[R0 + c] : 0.f) - s;
                sm.tmp[NB + c] = ((rb2 && R0 + c < T) ? rb2[R0 + c] : 0.f) - s2;
              }
              __syncthreads();
              if (tid < 2 * NB) {
                const int cc = tid & 63, which = tid >> 6;
                float zz = 0.f;
                for (int k = 0; k <= cc; ++k) zz = fmaf(sm.LiT[k * LI_LD + cc], sm.tmp[which * NB + k], zz);
                (which ? sm.z2 : sm.z)[R0 + cc] = zz;
              }
            }
            __syncthreads();
          }
          // panel rows: out = Ct^T . LiT  (= S Linv^T); in chunk 0 the first 64 rows are the diagonal block itself
          if (!(ch == 0 && ty < 8)) {
            float out[8][4];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
              for (int q = 0; q < 4; ++q) out[i][q] = 0.f;
            mma_tile(out, sm.Ct, CT_LD, sm.LiT, LI_LD, NB, ty, tx);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int gr = r_base + ty * 8 + i;
              if (gr < Tp)
                *reinterpret_cast<float4*>(S + (size_t)gr * ld + R0 + tx * 4) = make_float4(out[i][0], out[i][1], out[i][2], out[i][3]);
            }
          }
          __syncthreads();
        }
      }
      const int fcol = *sm.flag;
      __syncthreads();
      if (fcol < 0) break;
      fail = fcol + 1;
      if (attempt >= p.max_tries || !(p.jitter > 0.f)) break;
      jit_used = p.jitter * __powf(10.f, (float)attempt);
    }

    float tr_part = 0.f;
    if (p.do_inverse) {
      // =============================== Phase B: U = (L^-1)^T, tr(A^-1), alpha = X^T z
      for (int i = 0; i < nb; ++i) {
        const int R0 = i * NB;
        for (int idx = tid; idx < NB * NB; idx += NT) {
          const int r = idx >> 6, c = idx & 63;
          sm.LiT[r * LI_LD + c] = dinv[((size_t)i * NB + r) * NB + c];
        }
        __syncthreads();
        // diagonal block X_ii = Linv_ii: X[c][m] = LiT[m][c]
        for (int idx = tid; idx < NB * NB; idx += NT) {
          const int m = idx >> 6, c = idx & 63;
          if (R0 + m < T && R0 + c < T) {
            const float v = sm.LiT[m * LI_LD + c];
            tr_part = fmaf(v, v, tr_part);
          }
        }
        if (tid < NB) {
          float a = 0.f;
          for (int c = tid; c < NB; ++c) a = fmaf(sm.LiT[tid * LI_LD + c], sm.z[R0 + c], a);
          sm.al[R0 + tid] += a;
        }
        const int nch = (R0 + CM - 1) / CM;
        for (int ch = 0; ch < nch; ++ch) {
          const int m_base = ch * CM;
          float acc[8][4];
#pragma unroll
          for (int ii = 0; ii < 8; ++ii)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[ii][q] = 0.f;
          gemm_tn<true>(acc, S, ld, m_base, R0, R0, m_base, R0, dinv, sm.As, sm.Bs);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float* dst = sm.Ct + (tx * 4 + q) * CT_LD + ty * 8;
            *reinterpret_cast<float4*>(dst) = make_float4(-acc[0][q], -acc[1][q], -acc[2][q], -acc[3][q]);
            *reinterpret_cast<float4*>(dst + 4) = make_float4(-acc[4][q], -acc[5][q], -acc[6][q], -acc[7][q]);
          }
          __syncthreads();
          float out[8][4];
#pragma unroll
          for (int ii = 0; ii < 8; ++ii)
#pragma unroll
            for (int q = 0; q < 4; ++q) out[ii][q] = 0.f;
          mma_tile(out, sm.Ct, CT_LD, sm.LiT, LI_LD, NB, ty, tx);
          const float4 zv = *reinterpret_cast<const float4*>(sm.z + R0 + tx * 4);
#pragma unroll
          for (int ii = 0; ii < 8; ++ii) {
            const int m = m_base + ty * 8 + ii;
            float dot = 0.f;
            if (m < R0) {
              *reinterpret_cast<float4*>(S + (size_t)m * ld + R0 + tx * 4) = make_float4(out[ii][0], out[ii][1], out[ii][2], out[ii][3]);
              if (m < T) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                  if (R0 + tx * 4 + q < T) tr_part = fmaf(out[ii][q], out[ii][q], tr_part);
              }
              dot = out[ii][0] * zv.x + out[ii][1] * zv.y + out[ii][2] * zv.z + out[ii][3] * zv.w;
            }
            // reduce over the 16 tx lanes that share this row
            dot += __shfl_xor_sync(0xffffffffu, dot, 8);
            dot += __shfl_xor_sync(0xffffffffu, dot, 4);
            dot += __shfl_xor_sync(0xffffffffu, dot, 2);
            dot += __shfl_xor_sync(0xffffffffu, dot, 1);
            if (tx == 0 && m < R0) sm.al[m] += dot;
          }
          __syncthreads();
        }
        __syncthreads();
      }
    }

    // =============================== reductions and outputs
    float zz = 0.f, aa = 0.f, ar = 0.f, z22 = 0.f, z12 = 0.f;
    for (int i = tid; i < T; i += NT) {
      const float zi = sm.z[i], ai = sm.al[i], z2i = sm.z2[i];
      zz = fmaf(zi, zi, zz);
      z22 = fmaf(z2i, z2i, z22);
      z12 = fmaf(zi, z2i, z12);
      if (p.z_out) { p.z_out[((size_t)b * 2) * T + i] = zi; p.z_out[((size_t)b * 2 + 1) * T + i] = z2i; }
      aa = fmaf(ai, ai, aa);
      if (rb) ar = fmaf(ai, rb[i], ar);
      if (p.alpha && p.do_inverse) p.alpha[(size_t)b * T + i] = ai;
    }
    const float inv_quad = block_sum(zz, sm.red);
    const float logdet = 2.f * block_sum(logdet_part, sm.red);
    const float tr_inv = block_sum(tr_part, sm.red);
    const float alal = block_sum(aa, sm.red);
    const float alr = block_sum(ar, sm.red);
    const float sz22 = block_sum(z22, sm.red);
    const float sz12 = block_sum(z12, sm.red);
    if (tid == 0) {
      if (p.scalars) {
        float* o = p.scalars + (size_t)b * 16;
        const float Tf = (float)T;
        o[0] = -0.5f * (inv_quad + logdet + Tf * 1.8378770664093453f) / Tf;
        o[1] = 0.5f * (alal - tr_inv) / Tf;
        o[2] = logdet; o[3] = inv_quad; o[4] = tr_inv; o[5] = alal; o[6] = alr; o[7] = jit_used;
        o[8] = sz22; o[9] = sz12;
        for (int q = 10; q < 16; ++q) o[q] = 0.f;
      }
      if (p.info) p.info[b] = fail;
    }
    if (p.L_out) {
      float* Lo = p.L_out + (size_t)b * p.L_bstride;
      for (int idx = tid; idx < T * T; idx += NT) {
        const int r = idx / T, c = idx - r * T;
        Lo[(size_t)r * p.ldl + c] = (c <= r) ? S[(size_t)r * ld + c] : 0.f;
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------- host launcher
static size_t mll_smem_bytes(int Tp) {
  return sizeof(float) * (size_t)(2 * BK * AS_LD + 2 * BK * BS_LD + NB * CT_LD + NB * LI_LD + 4 * Tp + NB + 2 * NB + 32 + 4);
}

int launch_mll_batched_simt(MllParams p, cudaStream_t st) {
  p.Tp = (p.T + NB - 1) / NB * NB;
  p.nb = p.Tp / NB;
  const size_t smem = mll_smem_bytes(p.Tp);
  if (smem > 227 * 1024) {
    set_error("mll_batched: T=%d needs %zu bytes of shared memory (max 227 KB)", p.T, smem);
    return VOLT_ERR_ARG;
  }
  static size_t attr_smem_dev[16] = {};   // function attributes are per device
  size_t& attr_smem = attr_smem_dev[device_slot()];
  if (smem > attr_smem) {
    int s = check_cuda(cudaFuncSetAttribute(mll_batched_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                       "cudaFuncSetAttribute(mll_batched_kernel)");
    if (s) return s;
    attr_smem = smem;
  }
  int per_sm = (int)((227 * 1024) / (smem + 1024));
  if (per_sm > 2) per_sm = 2;
  if (per_sm < 1) per_sm = 1;
  int grid = sm_count() * per_sm;
  if (grid > p.B) grid = p.B;
  if (grid < 1) grid = 1;
  const size_t per_cta = (size_t)p.Tp * p.Tp + (size_t)p.nb * NB * NB;
  void* ws = nullptr;
  int s = get_workspace(per_cta * grid * sizeof(float), &ws, 0, st);
  if (s) return s;
  p.scratch = reinterpret_cast<float*>(ws);
  p.dinv = p.scratch + (size_t)grid * p.Tp * p.Tp;
  mll_batched_kernel<<<grid, NT, smem, st>>>(p);
  return check_cuda(cudaGetLastError(), "mll_batched_kernel");
}

}  // namespace volt

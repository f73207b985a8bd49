// Moving-average mean paths and the Monte-Carlo forecast rollout.
//
//   ma_paths_kernel   voltron/means/EWMA.py:20-135   EWMA / DEWMA / TEWMA / mean-reverting EMA paths (length T+1)
//   rollout_kernel    voltron/rollout_utils.py:6-93  GeneratePrediction + Rollouts (autoregressive) and the one-shot
//                     multi-point draw (rollout_utils.py:6-53 with H test points, VoltMagpie.py:67-99)
//
// Rollout algebra.  At horizon step idx the reference factors, for every draw s, the dense (n+idx)x(n+idx) matrix
// K_tr[i,j] = V_s[min(i,j)] and solves two systems with it (rollout_utils.py:35-44).  All S matrices of a series share
// their leading n x n block and the first n entries of every appended column (V[0:n]).  Writing
// L = [[L11, 0], [1 u^T, L22]] with u = L11^-1 V[0:n], the shared part (L11, u, z1 = L11^-1 r[0:n]) is computed ONCE per
// series by the batched potrf kernel (chol_batched.cu); the per-draw part is the bordered update
//      d_a = (V_s[n+a] + jitter - u.u) - sum_{t<a} f_t^2,  l_a = sqrt(d_a),
//      f_a = (V_s[n+a] - u.u - sum_{t<a} f_t^2) / l_a   (= L22[a', a] for every a' > a; = d_a / l_a without jitter),
//      w_a = (r_a - u.z1 - sum_{t<a} f_t w_t) / l_a,   q_a = (V_s[n+a] - u.u - sum_{t<a} f_t q_t) / l_a,
//      mean = u.z1 + sum q_a w_a + m_test,   cov = V_test - u.u - sum q_a^2,
// i.e. exactly the rows a dense left-looking Cholesky would append (every entry of column a of L22 is the same
// floating-point expression, so only running sums are kept).  No O(T) closed form is used.
#include "params.cuh"

namespace volt {


// ---------------------------------------------------------------------------------------------- ma_paths
// One CTA per series.  e = EWMA(y) (T+1); ee = EWMA(e)[:-1] (T+1); eee = EWMA(ee)[:-1] (T+1).
// out_kind (T+1) = e | 2e-ee | 3e-3ee+eee | e[j] - theta (e[j-1] - latent) (j >= 1).
// Optionally also writes e and ee (the rollout needs their tails) and resid = y - out[:-1].
__global__ void __launch_bounds__(256) ma_paths_kernel(const float* __restrict__ y, int T, int k, const float* __restrict__ w,
                                                       int kind, float theta, const float* __restrict__ latent,
                                                       float* __restrict__ out, float* __restrict__ e_out,
                                                       float* __restrict__ ee_out, float* __restrict__ resid_out) {
  extern __shared__ float sm[];
  float* sw = sm;              // k
  float* a0 = sw + k;          // T+1  series being filtered (y, then e, then ee)
  float* a1 = a0 + (T + 2);    // T+1  e
  float* a2 = a1 + (T + 2);    // T+1  ee
  const int s = blockIdx.x;
  const float* ys = y + (size_t)s * T;
  for (int t = threadIdx.x; t < k; t += blockDim.x) sw[t] = w[t];
  for (int i = threadIdx.x; i < T; i += blockDim.x) a0[i] = ys[i];
  __syncthreads();
  // filter(src of length L) -> dst[0..L]: dst[j] = sum_t w[t] * P[j+t], P[u] = src[0] (u<k) else src[u-k]
  auto filt = [&](const float* src, int L, float* dst) {
    for (int j = threadIdx.x; j <= L; j += blockDim.x) {
      float acc = 0.f;
      for (int t = 0; t < k; ++t) {
        const int u = j + t - k;
        acc = fmaf(sw[t], src[u < 0 ? 0 : u], acc);
      }
      dst[j] = acc;
    }
    __syncthreads();
  };
  filt(a0, T, a1);                       // e   (T+1)
  if (kind == MA_DEWMA || kind == MA_TEWMA) {
    filt(a1, T + 1, a2);                 // EWMA(e) has T+2 entries; [:-1] keeps T+1 -> a2[0..T] (a2[T+1] unused)
  }
  float* a3 = a0;                        // eee reuses the y staging buffer
  const float lat = (kind == MA_MEANREVERT && latent) ? latent[s] : 0.f;
  if (kind == MA_TEWMA) filt(a2, T + 1, a3);  // eee[0..T]; overwrites the staged y (resid re-reads y from global)
  for (int j = threadIdx.x; j <= T; j += blockDim.x) {
    float v = a1[j];
    if (kind == MA_DEWMA) v = 2.f * a1[j] - a2[j];
    else if (kind == MA_TEWMA) v = 3.f * a1[j] - 3.f * a2[j] + a3[j];
    else if (kind == MA_MEANREVERT) v = (j >= 1) ? a1[j] - theta * (a1[j - 1] - lat) : a1[j];
    out[(size_t)s * (T + 1) + j] = v;
    if (e_out) e_out[(size_t)s * (T + 1) + j] = a1[j];
    if (ee_out) ee_out[(size_t)s * (T + 1) + j] = (kind == MA_DEWMA || kind == MA_TEWMA) ? a2[j] : 0.f;
    if (resid_out && j < T) resid_out[(size_t)s * T + j] = ys[j] - v;
  }
}

// ---------------------------------------------------------------------------------------------- Philox + Box-Muller
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
  c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
}
__device__ __forceinline__ float philox_normal(unsigned long long seed, uint32_t a, uint32_t b, uint32_t c3) {
  uint32_t c[4] = {a, b, c3, 0x5eedu};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  const float u1 = ((float)c[0] + 1.0f) * 2.3283064365386963e-10f;  // (0,1]
  const float u2 = (float)c[1] * 2.3283064365386963e-10f;
  return sqrtf(-2.f * logf(u1)) * cospif(2.f * u2);
}

// four standard normals from ONE Philox4x32-10 block (both Box-Muller outputs of both uniform pairs): the autoregressive
// rollout draws one block per four horizon steps (counter = (series, draw, idx / 4)), a quarter of the generator work
__device__ __forceinline__ void philox_normal4(unsigned long long seed, uint32_t a, uint32_t b, uint32_t c3, float (&z)[4]) {
  uint32_t c[4] = {a, b, c3, 0x5eed4u};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const float u1 = ((float)c[2 * h] + 1.0f) * 2.3283064365386963e-10f;  // (0,1]
    const float u2 = (float)c[2 * h + 1] * 2.3283064365386963e-10f;
    const float rad = sqrtf(-2.f * logf(u1));
    float sn, cs;
    sincospif(2.f * u2, &sn, &cs);
    z[2 * h] = rad * cs;
    z[2 * h + 1] = rad * sn;
  }
}

// ---------------------------------------------------------------------------------------------- rollout
// value of a grown per-draw series at absolute index j: shared part (tail kept in smem) or the draw's own history
__device__ __forceinline__ float grown_at(int j, int nsh, const float* tail, int tail_len, float first, const float* hist) {
  if (j < 0) return first;
  if (j < nsh) {
    const int q = j - (nsh - tail_len);
    return q >= 0 ? tail[q] : first;  // q < 0 only when j == 0 region was requested beyond the tail: caller guarantees tail covers the window
  }
  return hist[j - nsh];
}

__global__ void __launch_bounds__(128) rollout_kernel(RolloutParams p) {
  extern __shared__ float sm[];
  const int TS = blockDim.x;
  const int H = p.H, Hp = p.Hp, k = p.k, n = p.n;
  const int tl = min(k + 1, n + 1);     // tail length kept for e/ee (indices n+1-tl .. n), and min(k, n) for y
  const int tly = min(k, n);
  float* sw = sm;                       // k
  float* ytail = sw + k;                // k
  float* etail = ytail + k;             // k+1
  float* eetail = etail + (k + 1);      // k+1
  // staging tiles: only the ones this mean family / eps mode needs are allocated (launch_rollout sizes smem alike)
  const bool ma = p.mean_kind != MA_GIVEN;
  const bool need_ee = (p.mean_kind == MA_DEWMA || p.mean_kind == MA_TEWMA);
  const bool need_e = need_ee || p.mean_kind == MA_MEANREVERT;
  // train-part prefixes of the MA windows, one per horizon step (identical for every draw of the series): the window sum
  // of step idx starts with the k - idx terms that come from the training tail, so that part of the fma chain is
  // computed once per CTA and every draw only continues it over its own idx generated values (same summation order)
  float* Pm = eetail + (k + 1);         // Hp
  float* Pe = Pm + Hp;                  // Hp
  float* Pee = Pe + Hp;                 // Hp
  float* t_pv = Pee + Hp;               // TS*Hp
  float* t_out = t_pv + TS * Hp;
  float* nxt = t_out + TS * Hp;
  // base normals (absent with in-kernel Philox).  Autoregressive mode reads eps[idx] exactly once, right before it writes
  // sample idx, so the normals are staged in the OUTPUT tile (no tile of their own: 7 instead of 4 resident CTAs per SM
  // for the EWMA mean); the one-shot draw may re-read them on a jitter retry and keeps a separate tile.
  const bool eps_tile = p.eps && p.joint;
  float* t_eps = eps_tile ? nxt : t_out;
  if (eps_tile) nxt += TS * Hp;
  float* t_e = nxt;                     // e history  (per draw, index a -> e[n+1+a])
  if (need_e) nxt += TS * Hp;
  float* t_ee = nxt;                    // ee history
  const int b = blockIdx.y;
  const int s0 = blockIdx.x * TS;
  const int tid = threadIdx.x;
  if (p.series_flag) {
    // launched as a programmatic dependent of the prep kernel: wait for THIS series' shared factor only (the prep kernel's
    // CTAs are all resident before any CTA of this grid can be, so the wait cannot starve them)
    if (tid == 0) {
      int f = 0;
      for (;;) {
        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(f) : "l"(p.series_flag + b) : "memory");
        if (f != 0) break;
        __nanosleep(100);
      }
    }
    __syncthreads();
  }

  for (int t = tid; t < k; t += TS) sw[t] = p.w ? p.w[t] : 0.f;
  const float* yb = p.ytrain + (size_t)b * n;
  for (int t = tid; t < tly; t += TS) ytail[t] = yb[n - tly + t];
  if (ma) {
    const float* eb = p.e_train + (size_t)b * (n + 1);
    for (int t = tid; t < tl; t += TS) etail[t] = eb[n + 1 - tl + t];
    if (need_ee) {
      const float* eeb = p.ee_train + (size_t)b * (n + 1);
      for (int t = tid; t < tl; t += TS) eetail[t] = eeb[n + 1 - tl + t];
    }
  }
  // coalesced tile loads
  const size_t base = ((size_t)b * p.S + s0) * H;
  const int ns = min(TS, p.S - s0);
  // (draw, step) of a flat tile index advance by a fixed (TS / H, TS % H) per iteration: one division per thread, not one per element
  const int dq = TS / H, dr = TS - dq * H;
  {
    int s = tid / H, h = tid - s * H;
    for (int idx = tid; idx < ns * H; idx += TS) {
      t_pv[s * Hp + h] = p.pred_vol[base + idx];
      if (p.eps) t_eps[s * Hp + h] = p.eps[base + idx];
      s += dq; h += dr;
      if (h >= H) { h -= H; ++s; }
    }
  }
  __syncthreads();

  const float y_first = yb[0];
  const float e_first = ma ? p.e_train[(size_t)b * (n + 1)] : 0.f;
  const float ee_first = need_ee ? p.ee_train[(size_t)b * (n + 1)] : 0.f;
  if (ma && !p.joint) {
    for (int idx = tid; idx < 1; idx += TS) {   // only step 0 starts from the training-tail sums (see the recurrence below)
      const int m = n + idx;
      const int tc = min(k, max(0, k - idx));        // window terms with absolute index < n      (training y)
      float acc = 0.f;
      for (int t = 0; t < tc; ++t) acc = fmaf(sw[t], grown_at(m - k + t, n, ytail, tly, y_first, nullptr), acc);
      Pm[idx] = acc;
      if (need_ee) {
        const int tc1 = min(k, max(0, k - idx + 1)); // window terms with absolute index < n + 1  (training e / ee)
        float a1 = 0.f, a2 = 0.f;
        for (int t = 0; t < tc1; ++t) {
          a1 = fmaf(sw[t], grown_at(m - k + t, n + 1, etail, tl, e_first, nullptr), a1);
          a2 = fmaf(sw[t], grown_at(m - k + t, n + 1, eetail, tl, ee_first, nullptr), a2);
        }
        Pe[idx] = a1;
        Pee[idx] = a2;
      }
    }
    __syncthreads();
  }
  const float* ser = p.series + (size_t)b * NSERIES;
  const float c0 = ser[0], uz = ser[1], Vn1 = ser[2], dx = ser[3], jit_s = ser[4];
  const float latent = (p.use_theta && p.latent) ? p.latent[b] : 0.f;
  const float mr_lat = (p.mean_kind == MA_MEANREVERT && p.mr_latent) ? p.mr_latent[b] : 0.f;

  if (tid < ns) {
    const int s = s0 + tid;
    float* pv = t_pv + tid * Hp;
    float* ep = t_eps + tid * Hp;
    float* out = t_out + tid * Hp;
    float* eh = t_e + tid * Hp;    // eh[a]  = e[n+1+a]   (EWMA path of the grown series, per draw)
    float* eeh = t_ee + tid * Hp;  // eeh[a] = ee[n+1+a]
    int flags = (p.series_info && p.series_info[b] != 0) ? 1 : 0;
    if (!p.joint) {
      // ================= autoregressive rollout (rollout_utils.py:57-93)
      double accV = (double)Vn1;   // trapezoid integral up to the last conditioning point
      float F2 = 0.f, FW = 0.f, FQ = 0.f, QW = 0.f, QQ = 0.f;
      float r_prev = 0.f;          // residual y - mean of the most recently appended point
      float zq[4] = {0.f, 0.f, 0.f, 0.f};   // in-kernel normals: one Philox block per four steps
      float e_run = 0.f, ee_run = 0.f, eee_run = 0.f;   // running window sums of the moving-average paths
      const float rho = 1.f - 2.f / (float)(k + 1), w_first = ma ? sw[0] : 0.f, w_last = ma ? sw[k - 1] : 0.f;
      for (int idx = 0; idx < H; ++idx) {
        const int m = n + idx;     // number of conditioning points at this step
        const float pvi = pv[idx];
        if (idx >= 1) {
          // append row a = idx-1 of the bordered factor (history vol is exp(log(pred_vol)), rollout_utils.py:77)
          const float pvh = expf(logf(pv[idx - 1]));
          accV += (double)(dx * (pvh * pvh));
          const float Ca = (float)accV - c0;
          const float d = (Ca + jit_s) - F2;  // jit_s: jitter the shared block needed (whole-diagonal, as psd_safe_cholesky)
          if (!(d > 0.f)) flags |= 1;
          const float l = sqrtf(d);
          const float f = (Ca - F2) / l;   // off-diagonal entries of column a carry no jitter (f == d / l when jit_s == 0)
          const float wa = (r_prev - uz - FW) / l;
          const float qa = (Ca - FQ) / l;
          F2 = fmaf(f, f, F2);
          FW = fmaf(f, wa, FW);
          FQ = fmaf(f, qa, FQ);
          QW = fmaf(qa, wa, QW);
          QQ = fmaf(qa, qa, QQ);
        }
        const float Vtest = (float)(accV + (double)((dx * 0.5f) * (pvi * pvi)));
        float cov = Vtest - c0 - QQ;
        // ---- test mean = last element of the MA path over the grown series (EWMA.py:48-50 and twins)
        float m_test;
        if (ma) {
          // window y[m-k .. m-1].  Step 0 is the full training-tail sum (Pm[0]); afterwards the window slides by one:
          // the weights are geometric (w[t-1] = rho w[t], rho = 1 - 2/(k+1)), so
          //   e[m] = rho (e[m-1] - w[0] y[m-1-k]) + w[k-1] y[m-1]
          // -- three operations per step instead of a k-term sum (the same filter; rounding differs at the 1e-7 level)
          float e_m;
          if (idx == 0) e_m = Pm[0];
          else e_m = fmaf(rho, e_run - w_first * grown_at(m - 1 - k, n, ytail, tly, y_first, out), w_last * out[idx - 1]);
          e_run = e_m;
          m_test = e_m;
          if (p.mean_kind == MA_MEANREVERT) {
            if (m >= 1) {
              const float e_prev = (idx == 0) ? (tl >= 2 ? etail[tl - 2] : e_first) : (idx == 1 ? etail[tl - 1] : eh[idx - 2]);
              m_test = e_m - p.mr_theta * (e_prev - mr_lat);
            }
          } else if (need_ee) {
            float ee_m;            // window e[m-k .. m-1] of the e path (training tail, then eh[a] = e[n+1+a]); same recurrence
            if (idx == 0) ee_m = Pe[0];
            else ee_m = fmaf(rho, ee_run - w_first * grown_at(m - 1 - k, n + 1, etail, tl, e_first, eh),
                             w_last * grown_at(m - 1, n + 1, etail, tl, e_first, eh));
            ee_run = ee_m;
            if (p.mean_kind == MA_DEWMA) {
              m_test = 2.f * e_m - ee_m;
            } else {
              float eee_m;         // window ee[m-k .. m-1] of the ee path (eeh[a] = ee[n+1+a])
              if (idx == 0) eee_m = Pee[0];
              else eee_m = fmaf(rho, eee_run - w_first * grown_at(m - 1 - k, n + 1, eetail, tl, ee_first, eeh),
                                w_last * grown_at(m - 1, n + 1, eetail, tl, ee_first, eeh));
              eee_run = eee_m;
              m_test = 3.f * e_m - 3.f * ee_m + eee_m;
            }
            if (idx >= 1) eeh[idx - 1] = ee_m;  // ee[n+idx]
          }
          if (need_e && idx >= 1) eh[idx - 1] = e_m;      // e[n+idx]
        } else {
          m_test = p.mean_test ? p.mean_test[(size_t)b * H + idx] : 0.f;
        }
        float mean = uz + QW + m_test;
        if (p.use_theta) mean -= p.theta * (mean - latent);
        // ---- psd_safe_cholesky(pred_cov, jitter) on the 1x1 matrix (rollout_utils.py:46)
        if (!(cov > 0.f)) {
          flags |= 2;
          float jit = p.jitter, c2 = cov;
          int tries = 0;
          for (; tries < 3; ++tries) {
            c2 = cov + jit;
            if (c2 > 0.f) break;
            jit *= 10.f;
          }
          if (tries == 3) flags |= 4;
          cov = c2;
        }
        float e_n;
        if (p.eps) {
          e_n = ep[idx];
        } else {
          if ((idx & 3) == 0) philox_normal4(p.seed, (uint32_t)(p.b_offset + b), (uint32_t)s, (uint32_t)(idx >> 2), zq);
          const int sel = idx & 3;
          e_n = sel == 0 ? zq[0] : sel == 1 ? zq[1] : sel == 2 ? zq[2] : zq[3];
        }
        const float sample = fmaf(sqrtf(cov), e_n, mean);
        out[idx] = sample;
        r_prev = sample - m_test;
      }
    } else {
      // ================= one-shot draw at H test points (rollout_utils.py:6-53 / VoltMagpie.py:67-99):
      // pred_cov[h,g] = Vfull[n+min(h,g)] - u.u (last trapezoid weight halved); sample = mean + chol(pred_cov) eps,
      // psd_safe_cholesky retry applied to the whole H x H block.
      float jit_total = 0.f;
      int attempt = 0;
      for (;;) {
        bool bad = false;
        double acc = (double)Vn1;
        float f2 = 0.f, fe = 0.f;  // sum f_t^2, sum f_t eps_t
        for (int h = 0; h < H; ++h) {
          const float pvh = pv[h];
          const float wgt = (h == H - 1) ? dx * 0.5f : dx;
          acc += (double)(wgt * (pvh * pvh));
          const float C0 = (float)acc - c0;
          const float d = (C0 + jit_total) - f2;
          if (!(d > 0.f)) bad = true;
          const float l = sqrtf(d);
          const float e_n = p.eps ? ep[h] : philox_normal(p.seed, (uint32_t)(p.b_offset + b), (uint32_t)s, (uint32_t)h);
          float mean = uz + (p.mean_test ? p.mean_test[(size_t)b * H + h] : 0.f);
          if (p.use_theta) mean -= p.theta * (mean - latent);
          out[h] = mean + fe + l * e_n;
          const float f = (C0 - f2) / l;  // every entry below the diagonal in column h
          f2 = fmaf(f, f, f2);
          fe = fmaf(f, e_n, fe);
        }
        if (!bad) break;
        flags |= 2;
        if (attempt >= 3) { flags |= 4; break; }
        jit_total = p.jitter * __powf(10.f, (float)attempt);
        ++attempt;
      }
    }
    if (p.info) p.info[(size_t)b * p.S + s] = flags;
  }
  __syncthreads();
  {
    int s = tid / H, h = tid - s * H;
    for (int idx = tid; idx < ns * H; idx += TS) {
      p.samples[base + idx] = t_out[s * Hp + h];
      s += dq; h += dr;
      if (h >= H) { h -= H; ++s; }
    }
  }
}

// ---------------------------------------------------------------------------------------------- launchers
int launch_ma_paths(const float* y, int S, int T, int k, const float* w, int kind, float theta, const float* latent, float* out,
                    float* e_out, float* ee_out, float* resid_out, cudaStream_t st) {
  const size_t smem = sizeof(float) * ((size_t)k + 3 * (size_t)(T + 2));
  if (smem > 220 * 1024) {
    set_error("ma_paths: T=%d, k=%d exceeds the shared-memory staged path", T, k);
    return VOLT_ERR_ARG;
  }
  static size_t attr_dev[16] = {};   // function attributes are per device
  size_t& attr = attr_dev[device_slot()];
  if (smem > 48 * 1024 && smem > attr) {
    int s = check_cuda(cudaFuncSetAttribute(ma_paths_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                       "cudaFuncSetAttribute(ma_paths_kernel)");
    if (s) return s;
    attr = smem;
  }
  ma_paths_kernel<<<S, 256, smem, st>>>(y, T, k, w, kind, theta, latent, out, e_out, ee_out, resid_out);
  return check_cuda(cudaGetLastError(), "ma_paths_kernel");
}

// The base normals the rollout kernel draws for (series, draw, step) when no eps is given -- the same Philox counters, so a
// caller can re-run single draws (the per-draw psd_safe_cholesky fallback of volt_b200.ops.rollout) with the exact numbers
// the in-kernel generator used.
__global__ void __launch_bounds__(256) rollout_normals_kernel(unsigned long long seed, int b_offset, long long total, int S, int H, int joint,
                                                              float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int h = (int)(i % H);
  const long long bs = i / H;
  const int s = (int)(bs % S), b = (int)(bs / S);
  if (joint) {
    out[i] = philox_normal(seed, (uint32_t)(b_offset + b), (uint32_t)s, (uint32_t)h);
  } else {
    float zq[4];
    philox_normal4(seed, (uint32_t)(b_offset + b), (uint32_t)s, (uint32_t)(h >> 2), zq);
    const int sel = h & 3;
    out[i] = sel == 0 ? zq[0] : sel == 1 ? zq[1] : sel == 2 ? zq[2] : zq[3];
  }
}
int launch_rollout_normals(unsigned long long seed, int b_offset, int B, int S, int H, int joint, float* out, cudaStream_t st) {
  const long long total = (long long)B * S * H;
  if (total == 0) return VOLT_OK;
  rollout_normals_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(seed, b_offset, total, S, H, joint, out);
  return check_cuda(cudaGetLastError(), "rollout_normals_kernel");
}

int launch_rollout(RolloutParams p, cudaStream_t st) {
  p.Hp = p.H | 1;
  int TS = 128;
  const bool need_ee = (p.mean_kind == MA_DEWMA || p.mean_kind == MA_TEWMA);
  const bool need_e = need_ee || p.mean_kind == MA_MEANREVERT;
  const int ntiles = 2 + ((p.eps && p.joint) ? 1 : 0) + (need_e ? 1 : 0) + (need_ee ? 1 : 0);   // as carved in the kernel
  auto smem_for = [&](int ts) { return sizeof(float) * ((size_t)4 * p.k + 2 + 3 * (size_t)p.Hp + (size_t)ntiles * ts * p.Hp); };
  while (TS > 32 && smem_for(TS) > 56 * 1024) TS >>= 1;   // aim for >= 4 resident CTAs per SM
  const size_t smem = smem_for(TS);
  if (smem > 220 * 1024) {
    set_error("rollout: H=%d, k=%d exceeds shared memory", p.H, p.k);
    return VOLT_ERR_ARG;
  }
  static size_t attr_dev[16] = {};   // function attributes are per device
  size_t& attr = attr_dev[device_slot()];
  if (smem > 48 * 1024 && smem > attr) {
    int s = check_cuda(cudaFuncSetAttribute(rollout_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                       "cudaFuncSetAttribute(rollout_kernel)");
    if (s) return s;
    attr = smem;
  }
  for (int b0 = 0; b0 < p.B; b0 += 65535) {
    RolloutParams q = p;
    const int nb = min(65535, p.B - b0);
    q.B = nb;
    q.b_offset = p.b_offset + b0;   // Philox counters are keyed on the GLOBAL series index: chunks draw distinct normals
    q.ytrain = p.ytrain + (size_t)b0 * p.n;
    if (p.e_train) q.e_train = p.e_train + (size_t)b0 * (p.n + 1);
    if (p.ee_train) q.ee_train = p.ee_train + (size_t)b0 * (p.n + 1);
    if (p.mean_test) q.mean_test = p.mean_test + (size_t)b0 * p.H;
    q.series = p.series + (size_t)b0 * NSERIES;
    if (p.series_info) q.series_info = p.series_info + b0;
    q.pred_vol = p.pred_vol + (size_t)b0 * p.S * p.H;
    if (p.eps) q.eps = p.eps + (size_t)b0 * p.S * p.H;
    if (p.latent) q.latent = p.latent + b0;
    if (p.mr_latent) q.mr_latent = p.mr_latent + b0;
    q.samples = p.samples + (size_t)b0 * p.S * p.H;
    if (p.info) q.info = p.info + (size_t)b0 * p.S;
    if (p.series_flag) q.series_flag = p.series_flag + b0;
    dim3 grid((p.S + TS - 1) / TS, nb);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3((unsigned)TS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = p.series_flag ? 1 : 0;     // only when the producer raises per-series flags
    int s = check_cuda(cudaLaunchKernelEx(&cfg, rollout_kernel, q), "cudaLaunchKernelEx(rollout_kernel)");
    if (s) return s;
  }
  return check_cuda(cudaGetLastError(), "rollout_kernel");
}

}  // namespace volt

// Batched exact-MLL kernel, tensor-core variant: same algorithm and outputs as chol_batched.cu (left-looking blocked
// potrf + in-place trtri, one CTA per series), but every GEMM-shaped product runs on the 5th-generation tensor cores:
//
//   tcgen05.mma.cta_group::1.kind::tf32, M=128, N=64, K=8 per instruction, accumulators in TMEM (fp32),
//   operands in shared memory in the canonical K-major SWIZZLE_128B layout, read back with tcgen05.ld.32x32b.x32.
//
// fp32 accuracy is kept with the 3xTF32 split: every operand tile is stored twice, hi = a & 0xffffe000 (exactly
// representable in TF32) and lo = a - hi, and each product is issued as hi*hi + hi*lo + lo*hi (SURVEY.md section 7,
// hard part 1: plain TF32 misses the 1e-3 posterior bar on the noise-free rollout matrix).
//
// Products on the tensor pipe, per 128-row chunk (all "TN": both operands K-contiguous in the scratch):
//   Phase A   S   = A[rows, j] - L[rows, 0:j] L[j, 0:j]^T          K = 64 j
//             L[rows, j] = S Linv_jj^T                              K = 64   (TRSM by the inverted diagonal block)
//   Phase B   G^T = U[0:i, 0:i] L[i, 0:i]^T                         K = 64 i (U = (L^-1)^T, upper triangle of the scratch;
//             control-warp instance: kept as L^-1 IN PLACE of the dead rows of L, so that a series touches half a square)
//             U[0:i, i] = -G^T Linv_ii^T                            K = 64
// SIMT work that remains: the 64x64 diagonal potrf / trtri (chol_dev.cuh), the hi/lo split while staging operand
// tiles, the forward substitution for z, and the reductions.
#include <cuda.h>

#include <cstring>

#include "chol_tc_dev.cuh"

namespace volt {
namespace tc {

// -DVOLT_PROFILE (tools/build.sh --profile -> libvolt_prof.so): thread 0 of CTA 0 accumulates clock64() per kernel
// segment and drops the 12 counters into the first floats of `alpha`; read by tools/seg_probe.py.  Never on in the
// shipped library.
#ifdef VOLT_PROFILE
#define TICK(i) do { if (threadIdx.x == 0) { const long long _n = clock64(); seg[i] += _n - tlast; tlast = _n; } } while (0)
#else
#define TICK(i) do { } while (0)
#endif

// TRI = true: three CTAs per SM (128 TMEM columns, ~71 KB shared memory, single-buffered stages; chol_tc_dev.cuh);
// TRI = false: two CTAs per SM with double-buffered stages (series too long for the small shared-memory map).
// HOSTIN = true: the instance behind volt_mll_grad_vol_host -- prefix sums built in-kernel from the raw volatility path and
// an arrival flag for series whose inputs are still being copied (params.cuh); kept out of the device-pointer instance,
// whose register allocation it would disturb (measured: +2 % on the c2 kernel time when compiled in unconditionally).
// Whole per-series prologue of that instance as ONE out-of-line call (arrival gate, then the prefix sums into shared
// memory): the main kernel body only sees a call at a point where nothing but kernel-lifetime values are live.
// Returns false when the inputs never arrived (the CTA stops, the host re-runs the batch ungated).
static __device__ __noinline__ bool hostin_prologue(const int* ready, int* ready_timeout, long long ready_spins, const float* xs,
                                                    const float* vs, int T, int Tp, int mode, float* Vs, int* sflag) {
  const int tid = threadIdx.x;
  if (ready) {
    // This series' inputs were still being copied when the kernel was launched; the copy stream writes the flag after
    // them (pure DMA, so it cannot wait for an SM that this CTA is holding).  Where streams are serialised (a profiler
    // replaying this kernel, CUDA_LAUNCH_BLOCKING) the flag cannot arrive while the kernel runs: after ready_spins
    // polls the CTA reports a timeout and stops.
    if (tid == 0) {
      int f = 0;
      for (long long spins = 0; spins < ready_spins; ++spins) {
        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(f) : "l"(ready) : "memory");
        if (f != 0) break;
        __nanosleep(200);
      }
      if (f == 0) *ready_timeout = 1;
      *sflag = f;
    }
    wsync();
    const int arrived = *sflag;
    wsync();
    if (arrived == 0) return false;
  }
  if (tid < 32) cumtrapz_warp(xs, vs, T, mode, 1, Vs, tid);
  for (int i = T + tid; i < Tp; i += NT) Vs[i] = 0.f;
  return true;
}

// TMA = true: the "W" instance (two CTAs per SM): operand tiles of every GEMM loop arrive through the TMA unit (gemm_tma,
// chol_tc_dev.cuh); shared-memory map and accumulator parking as in the three-CTA instance, 256 TMEM columns, one-pass TRSM.
// TMA = 2: the same with two control warps (warp 8: MMA issuer, warp 9: TMA producer; gemm_w2_worker / w2_control): 320 threads.
constexpr int W2_THREADS = NT + 64;

// Control warps of the TMA = 2 instance: they walk the same deterministic schedule of GEMM calls as the workers (series loop,
// psd_safe_cholesky attempts, phase A block steps x row chunks, phase B) and feed / issue every k-tile of every call.
template <bool HOSTIN>
__device__ void w2_control(const MllParams& p, Ctx& c, TmaPipe& tp, const CUtensorMap* tmA, const CUtensorMap* tmB, const CUtensorMap* tmAt,
                           int wu) {
  const int Tp = p.Tp, nb = p.nb;
  const uint32_t xb = s_u32(c.X), tmem_u = make_uniform(c.tmem);
  const int sq_row0 = (int)blockIdx.x * Tp;
  const bool is_tma = (wu == 9);
  for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
    if (HOSTIN) {   // host-buffer entry: the workers report whether this series' inputs arrived (they stop if not)
      asm volatile("bar.sync 2, %0;" ::"n"(W2_THREADS) : "memory");
      const int arrived = *reinterpret_cast<volatile int*>(c.flag + 1);
      asm volatile("bar.sync 2, %0;" ::"n"(W2_THREADS) : "memory");
      if (!arrived) break;
    }
    for (int attempt = 0;; ++attempt) {
      bool wait_ring = true;                               // the diagonal block of step 0 uses the ring region as scratch
      for (int j = 1; j < nb; ++j) {                       // block step 0 has no earlier columns: no GEMM call
        const int R0 = j * NB, nk = R0 / 16;
        const int nch = (Tp - R0 + CM - 1) / CM;
        for (int ch = 0; ch < nch; ++ch) {
          if (is_tma) w2_tma_call(tp, tmA, tmB, xb, sq_row0 + R0 + ch * CM, sq_row0 + R0, 0, nk, wait_ring, ch == 0, nk - 4);
          else w2_mma_call(tp, tmem_u, xb, nk);
          wait_ring = (ch == 0);                           // ... and so does the first-chunk epilogue of every step
        }
      }
      if (is_tma && wait_ring) {                           // the last step's release (keeps the phase count in step)
        if (elect_one()) mbar_wait(tp.ringfree, tp.rf_n & 1u);
        __syncwarp();
        ++tp.rf_n;
      }
      // the workers publish the outcome of the factorisation (first failing column or -1) between two CTA-wide barriers
      asm volatile("bar.sync 2, %0;" ::"n"(W2_THREADS) : "memory");
      const int fcol = *reinterpret_cast<volatile int*>(c.flag);
      asm volatile("bar.sync 2, %0;" ::"n"(W2_THREADS) : "memory");
      if (fcol < 0) break;
      if (attempt >= p.max_tries || !(p.jitter > 0.f)) break;
    }
    if (p.do_inverse) {
      for (int i = 1; i < nb; ++i) {
        const int R0 = i * NB;
        const int nch = (R0 + CM - 1) / CM;
        for (int ch = 0; ch < nch; ++ch) {
          const int m_base = ch * CM, nk = (R0 - m_base) / 16;
          // the preamble of every step stages the inverse diagonal block through the ring region
          if (is_tma) w2_tma_call(tp, tmAt, tmB, xb, sq_row0, sq_row0 + R0, m_base, nk, ch == 0, ch == 0, nk - 4, true, m_base);
          else w2_mma_call(tp, tmem_u, xb, nk);
        }
      }
    }
  }
}

template <bool TRI, bool HOSTIN = false, int TMA = 0>
__global__ void __launch_bounds__(TMA == 2 ? W2_THREADS : NT, TRI ? 3 : 2)
    mll_batched_tc_kernel(MllParams p, const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                          const __grid_constant__ CUtensorMap tmAt) {
  static_assert(!(TRI && TMA), "the TMA instances are two-CTA instances");
  static_assert(!(HOSTIN && TMA == 1), "the host-buffer entry uses the register-staged or the control-warp instance");
  constexpr bool PARK = TRI || TMA != 0;   // chunk-0 panel rows wait in the accumulator columns during the diagonal factorisation
  constexpr uint32_t LOFF = TMA ? W_L_OFF : (TRI ? Y_L_OFF : L_OFF), CTOFF = TMA ? W_CT_OFF : (TRI ? Y_CT_OFF : CT_OFF);
  constexpr uint32_t VECOFF = TMA ? W_VEC_OFF : (TRI ? Y_VEC_OFF : VEC_OFF);
  constexpr uint32_t XTMP = TRI ? Y_TMP : X_TMP, TCOLS = TRI ? T3_COLS : TM_COLS;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = smem_raw;
  if ((s_u32(smem_raw) & 1023u) != 0u) __trap();  // SWIZZLE_128B operand tiles need a 1024-byte aligned base
  Ctx c;
  c.X = base;
  c.Lr = base + LOFF;
  c.Ct = reinterpret_cast<float*>(base + CTOFF);
  c.Vs = reinterpret_cast<float*>(base + VECOFF);
  c.z = c.Vs + p.Tp;
  c.al = c.z + p.Tp;
  const bool has2 = (p.resid2 != nullptr);   // the second right-hand side (rollout prep) gets its own vector
  c.z2 = has2 ? c.al + p.Tp : c.z;
  float* rs_stage = c.al + (has2 ? 2 : 1) * p.Tp;   // p.stage_in: this series' residual, read from mapped host memory once
  c.diagl = rs_stage + (p.stage_in ? p.Tp : 0);
  c.tmp = c.diagl + NB;
  c.red = c.tmp + 2 * NB;
  c.flag = reinterpret_cast<int*>(c.red + 32);
  c.bar = reinterpret_cast<uint64_t*>(c.red + 36);
  uint32_t* s_tmem_p = reinterpret_cast<uint32_t*>(c.red + 40);   // c.bar holds two mbarriers (16 bytes)
  c.phase = 0;
  float* LiT = reinterpret_cast<float*>(c.X + X_LIT);
  float* tmpbuf = reinterpret_cast<float*>(c.X + XTMP);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = 32 * (warp & 3) + lane;   // accumulator row (TMEM lane) owned by this thread
  const int half_id = warp >> 2;            // which 32-column half of the 64 accumulator columns
  const int c0 = half_id * 32;
  const int T = p.T, Tp = p.Tp, nb = p.nb, ld = p.Tp;
  float* S = p.scratch + (size_t)blockIdx.x * Tp * Tp;
  float* dinv = p.dinv + (size_t)blockIdx.x * nb * NB * NB;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(s_tmem_p)), "n"(TCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (p.series_flag) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the rollout kernel may be scheduled behind us
  TmaPipe tp;
  tp.full = reinterpret_cast<uint64_t*>(c.red + 44);
  tp.done = tp.full + W_RING;
  tp.ready = tp.done + W_RING;
  tp.ringfree = tp.ready + W_RING;
  tp.depready = tp.ringfree + 1;
  tp.g = 0;
  tp.rf_n = tp.dep_n = 0;
#ifdef VOLT_PROFILE
  for (int i = 0; i < 8; ++i) tp.prof[i] = 0;
#endif
  if (tid == 0) {
    mbar_init(c.bar, 1);
    mbar_init(c.bar + 1, 1);
    if constexpr (TMA != 0) {
      for (int i = 0; i < W_RING; ++i) { mbar_init(tp.full + i, 1); mbar_init(tp.done + i, 1); mbar_init(tp.ready + i, NT / 32); }
      mbar_init(tp.ringfree, 1);
      mbar_init(tp.depready, 1);
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
      if constexpr (TMA == 2) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmAt) : "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();     // every thread of the CTA, control warps included
  tc_fence_after();
  c.tmem = *s_tmem_p;
  if constexpr (TMA == 2) {
    const int wu = uniform_warp_id();
    if (wu >= NT / 32) {
      w2_control<HOSTIN>(p, c, tp, &tmA, &tmB, &tmAt, wu);
      tc_fence_before();
      __syncthreads();   // pairs with the workers' barrier before the TMEM deallocation
      return;
    }
  }
  const uint32_t t_lane = (uint32_t)(32 * (warp & 3)) << 16;
  const int sq_row0 = (int)blockIdx.x * p.Tp;   // first row of this CTA's scratch square in the tensor maps

#ifdef VOLT_PROFILE
  long long seg[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  long long tlast = clock64();
#endif
  for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
    if (HOSTIN) {
      const float* xs = p.x_in + (p.x_batched ? (size_t)b * T : 0);
      const float* vs = p.vol_in + (size_t)b * T;
      if (p.stage_in) {
        // the inputs are the caller's pinned host buffers, mapped: ONE round trip over PCIe per series (12 T bytes, all
        // threads), then the prefix sums and the residual are served from shared memory (z / al are free until the
        // factorisation starts)
        const float* rsrc = p.resid + (size_t)b * T;
        for (int i = tid; i < T; i += NT) {
          c.al[i] = __ldcs(xs + i);
          c.z[i] = __ldcs(vs + i);
          rs_stage[i] = __ldcs(rsrc + i);
        }
        wsync();
        xs = c.al;
        vs = c.z;
      }
      const bool arrived = hostin_prologue((p.ready && b >= p.ready_from) ? p.ready : nullptr, p.ready_timeout, p.ready_spins, xs, vs, T, Tp,
                                           p.vol_mode, c.Vs, c.flag);
      if (p.stage_in) wsync();   // warp 0 has read al / z before they are cleared below
      if constexpr (TMA == 2) {   // tell the control warps (w2_control)
        if (tid == 0) c.flag[1] = arrived ? 1 : 0;
        asm volatile("bar.sync 2, %0;" ::"n"(W2_THREADS) : "memory");
        asm volatile("bar.sync 2, %0;" ::"n"(W2_THREADS) : "memory");
      }
      if (!arrived) break;
    } else {
      for (int i = tid; i < Tp; i += NT) {
        float v = 0.f;
        if (i < T) {
          if (p.kind == KIND_VOL) v = p.V[(size_t)b * T + i];
          else if (p.kind == KIND_BM) v = p.x[i];
        }
        c.Vs[i] = v;
      }
    }
    const float sc = (p.kind == KIND_BM) ? p.scale[(size_t)b * p.scale_stride] : 1.f;
    const float dadd0 = p.raw_noise ? noise_from_raw_dev(p.raw_noise[(size_t)b * p.raw_stride])
                                    : (p.diag_add ? p.diag_add[(size_t)b * p.diag_stride] : 0.f);
    const float* rb = (HOSTIN && p.stage_in) ? rs_stage : (p.resid ? p.resid + (size_t)b * T : nullptr);
    const float* rb2 = p.resid2 ? p.resid2 + (size_t)b * T : nullptr;

    int fail = 0;
    float jit_used = 0.f, logdet_part = 0.f;
    for (int attempt = 0;; ++attempt) {
      const float dadd = dadd0 + jit_used;
      logdet_part = 0.f;
      if (tid == 0) *c.flag = -1;
      for (int i = tid; i < Tp; i += NT) { c.z[i] = 0.f; c.al[i] = 0.f; if (has2) c.z2[i] = 0.f; }
      wsync();
      fail = 0;
      // =============================== Phase A
      for (int j = 0; j < nb; ++j) {
        const int R0 = j * NB;
        const int nch = (Tp - R0 + CM - 1) / CM;
        for (int ch = 0; ch < nch; ++ch) {
          const int r_base = R0 + ch * CM;
          const int gr = r_base + row;
          TICK(0);
          bool have;
          float zacc[2] = {0.f, 0.f};   // TMA = 2, first chunk: partial sums of L[j, 0:R0] z (and z2), taken from the B tiles
          if constexpr (TMA == 2)
            have = gemm_w2_worker<false>(c, tp, r_base, Tp, 0, R0, ch == 0, (ch == 0 && rb) ? c.z : nullptr,
                                         (ch == 0 && rb2) ? c.z2 : nullptr, zacc);
          else if constexpr (TMA == 1) have = gemm_tma<false>(c, tp, &tmA, &tmB, sq_row0, r_base, Tp, R0, 0, R0);
          else if constexpr (TRI) have = gemm_tc1<false>(c, S, ld, r_base, Tp, R0, 0, R0, nullptr);
          else have = gemm_tc<false>(c, S, ld, r_base, Tp, R0, 0, R0, nullptr);
          TICK(1);
          float s[32];
          if (have) {
            tmem_ld32(c.tmem + t_lane + (uint32_t)c0, s);
            tc_fence_before();
          } else {
#pragma unroll
            for (int q = 0; q < 32; ++q) s[q] = 0.f;
          }
          if (gr < Tp) {
            gen_sub_row32(p, b, gr, R0 + c0, c.Vs, sc, dadd, s);
          } else {
#pragma unroll
            for (int q = 0; q < 32; ++q) s[q] = 0.f;
          }
          TICK(2);
          if (ch == 0) {
            // rows < 64 are the diagonal block (-> Ct); the other rows park their 32 values in the free tail of X so
            // that no accumulator registers stay live across the warp-level factorisation below
            float4* stash = reinterpret_cast<float4*>(c.X + X_STASH);
            const int slot = (row - NB) + NB * half_id;
            if (row < NB) {
#pragma unroll
              for (int q = 0; q < 8; ++q)
                *reinterpret_cast<float4*>(c.Ct + row * CLD + c0 + 4 * q) = make_float4(s[4 * q], s[4 * q + 1], s[4 * q + 2], s[4 * q + 3]);
              if (half_id == 0) c.tmp[row] = gen_entry(p, b, gr, gr, c.Vs, sc, dadd);   // original A_ii for the pivot test
            } else if constexpr (PARK) {
              uint32_t u[32];   // park the rows in the accumulator columns (warp-uniform branch: rows >= 64 <=> (warp & 3) >= 2)
#pragma unroll
              for (int q = 0; q < 32; ++q) u[q] = __float_as_uint(s[q]);
              tmem_st32(c.tmem + t_lane + (uint32_t)c0, u);
              asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            } else {
#pragma unroll
              for (int q = 0; q < 8; ++q) stash[q * 128 + slot] = make_float4(s[4 * q], s[4 * q + 1], s[4 * q + 2], s[4 * q + 3]);
            }
            wsync();
            TICK(3);
            diag64_block_v2<CLD>(c.Ct, LiT, tmpbuf, c.diagl, c.tmp, c.flag, R0);
            TICK(4);
            if (tid < NB && R0 + tid < T) logdet_part += logf(c.diagl[tid]);
            for (int idx = tid; idx < NB * NB; idx += NT) {
              const int r = idx >> 6, cc = idx & 63;
              // TMA instance: the diagonal block of the scratch holds the block of U = (L^-1)^T instead of L_jj (nothing reads
              // L_jj from the scratch again; the launcher keeps callers that want L out of this instance), so that phase B's
              // operand tiles are plain rectangles of the scratch
              S[(size_t)(R0 + r) * ld + R0 + cc] =
                  TMA == 2 ? LiT[cc * CLD + r] : (TMA != 0 ? LiT[r * CLD + cc] : ((cc <= r) ? c.Ct[r * CLD + cc] : 0.f));
              dinv[((size_t)j * NB + r) * NB + cc] = LiT[r * CLD + cc];
            }
            if constexpr (PARK) wsync();   // D aliases the Linv operand: every read of L_jj precedes the staging
            stage_linv_from_lit(c, LiT);
            if (rb) {
              const int cz = tid >> 2, part = tid & 3;
              float a1 = zacc[0], a2 = zacc[1];
              const float* Lrow = S + (size_t)(R0 + cz) * ld;
              for (int k = part * 4; k < (TMA == 2 ? 0 : R0); k += 16) {
                const float4 lv = *reinterpret_cast<const float4*>(Lrow + k);
                a1 = fmaf(lv.x, c.z[k], a1); a1 = fmaf(lv.y, c.z[k + 1], a1);
                a1 = fmaf(lv.z, c.z[k + 2], a1); a1 = fmaf(lv.w, c.z[k + 3], a1);
                if (rb2) {
                  a2 = fmaf(lv.x, c.z2[k], a2); a2 = fmaf(lv.y, c.z2[k + 1], a2);
                  a2 = fmaf(lv.z, c.z2[k + 2], a2); a2 = fmaf(lv.w, c.z2[k + 3], a2);
                }
              }
              a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
              a1 += __shfl_xor_sync(0xffffffffu, a1, 2);
              a2 += __shfl_xor_sync(0xffffffffu, a2, 1);
              a2 += __shfl_xor_sync(0xffffffffu, a2, 2);
              if (part == 0) {
                c.tmp[cz] = ((R0 + cz < T) ? rb[R0 + cz] : 0.f) - a1;
                c.tmp[NB + cz] = ((rb2 && R0 + cz < T) ? rb2[R0 + cz] : 0.f) - a2;
              }
              wsync();
              {
                // z_j = Linv_jj t: four threads per entry (k = part, part + 4, ...), combined by shuffles in a fixed order
                float zz = 0.f, zz2 = 0.f;
#pragma unroll
                for (int it = 0; it < 16; ++it) {
                  const int k = part + 4 * it;
                  if (k <= cz) {
                    const float lv = LiT[k * CLD + cz];
                    zz = fmaf(lv, c.tmp[k], zz);
                    if (has2) zz2 = fmaf(lv, c.tmp[NB + k], zz2);
                  }
                }
                zz += __shfl_xor_sync(0xffffffffu, zz, 1);
                zz += __shfl_xor_sync(0xffffffffu, zz, 2);
                zz2 += __shfl_xor_sync(0xffffffffu, zz2, 1);
                zz2 += __shfl_xor_sync(0xffffffffu, zz2, 2);
                if (part == 0) {
                  c.z[R0 + cz] = zz;
                  if (has2) c.z2[R0 + cz] = zz2;
                }
              }
            }
            if (row >= NB) {
              if constexpr (PARK) {
                tmem_ld32(c.tmem + t_lane + (uint32_t)c0, s);
              } else {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                  const float4 v = stash[q * 128 + slot];
                  s[4 * q] = v.x; s[4 * q + 1] = v.y; s[4 * q + 2] = v.z; s[4 * q + 3] = v.w;
                }
              }
            } else {
#pragma unroll
              for (int q = 0; q < 32; ++q) s[q] = 0.f;
            }
            // LiT / tmp / stash (aliasing X) are dead from here on; Linv operand staged
            if constexpr (TMA == 2) w2_release_ring(tp);
            else wsync();
          }
          TICK(5);
          float o[32];
          if constexpr (TRI) trsm_tc1(c, s, o, row, half_id);
          else trsm_tc(c, s, o, row, half_id);
          TICK(6);
          {
            // warp-uniform: the warp's 32 rows start at a multiple of 32 and Tp, R0 are multiples of 64
            const int g0 = r_base + 32 * (warp & 3);
            if (!(ch == 0 && (warp & 3) < 2) && g0 < Tp) {
              if constexpr (TRI) store_block32_2p(reinterpret_cast<float*>(c.X) + warp * 640, o, S + (size_t)g0 * ld + R0 + c0, ld, lane);
              else if constexpr (TMA == 2) store_block32_bl(c.X + W_BL, warp, o, S + (size_t)g0 * ld + R0 + c0, ld, lane);
              else store_block32(reinterpret_cast<float*>(c.X) + warp * 1152, o, S + (size_t)g0 * ld + R0 + c0, ld, lane);
            }
          }
          wsync();
        }
      }
      const int fcol = *c.flag;
      if constexpr (TMA == 2) {   // the control warps read the outcome between these two barriers (w2_control)
        asm volatile("bar.sync 2, %0;" ::"n"(W2_THREADS) : "memory");
        asm volatile("bar.sync 2, %0;" ::"n"(W2_THREADS) : "memory");
      }
      wsync();
      if (fcol < 0) break;
      fail = fcol + 1;
      if (attempt >= p.max_tries || !(p.jitter > 0.f)) break;
      jit_used = p.jitter * __powf(10.f, (float)attempt);
    }

    float tr_part = 0.f;
    if (p.do_inverse) {
      // =============================== Phase B
      for (int i = 0; i < nb; ++i) {
        const int R0 = i * NB;
        const float* Di = dinv + (size_t)i * NB * NB;
        TICK(7);
        // bring the inverted diagonal block back into shared memory once (coalesced), then stage the TRSM operand and
        // take its own contribution to tr(A^-1) and alpha from there
        for (int idx = tid; idx < NB * NB / 4; idx += NT) {
          const int r = idx >> 4, c4 = (idx & 15) * 4;
          *reinterpret_cast<float4*>(LiT + r * CLD + c4) = *reinterpret_cast<const float4*>(Di + r * NB + c4);
        }
        wsync();
        stage_linv_from_lit(c, LiT);
        for (int idx = tid; idx < NB * NB; idx += NT) {
          const int m = idx >> 6, cc = idx & 63;
          if (R0 + m < T && R0 + cc < T) {
            const float v = LiT[m * CLD + cc];
            tr_part = fmaf(v, v, tr_part);
          }
        }
        {
          // alpha_i += Linv_ii^T z_i: four threads per row (cc = part, part + 4, ...), combined by shuffles in a fixed order
          const int ar = tid >> 2, part = tid & 3;
          float a = 0.f;
#pragma unroll
          for (int it = 0; it < 16; ++it) {
            const int cc = part + 4 * it;
            if (cc >= ar) a = fmaf(LiT[ar * CLD + cc], c.z[R0 + cc], a);
          }
          a += __shfl_xor_sync(0xffffffffu, a, 1);
          a += __shfl_xor_sync(0xffffffffu, a, 2);
          if (part == 0) c.al[R0 + ar] += a;
        }
        if (TMA == 2 && i >= 1) w2_release_ring(tp);   // LiT (ring region) is dead: the producer may load this step's tiles
        else wsync();
        TICK(12);
        const int nch = (R0 + CM - 1) / CM;
        for (int ch = 0; ch < nch; ++ch) {
          const int m_base = ch * CM;
          const int m = m_base + row;
          TICK(7);
          if constexpr (TMA == 2) gemm_w2_worker<true>(c, tp, m_base, R0, m_base, R0, ch == 0);
          else if constexpr (TMA == 1) gemm_tma<true>(c, tp, &tmA, &tmB, sq_row0, m_base, R0, R0, m_base, R0);
          else if constexpr (TRI) gemm_tc1<true>(c, S, ld, m_base, R0, R0, m_base, R0, dinv);
          else gemm_tc<true>(c, S, ld, m_base, R0, R0, m_base, R0, dinv);
          TICK(8);
          float s[32], o[32];
          tmem_ld32(c.tmem + t_lane + (uint32_t)c0, s);
          tc_fence_before();
#pragma unroll
          for (int q = 0; q < 32; ++q) s[q] = -s[q];
          if constexpr (TRI) trsm_tc1(c, s, o, row, half_id);
          else trsm_tc(c, s, o, row, half_id);
          TICK(9);
          float hdot = 0.f;
          {
            const int g0 = m_base + 32 * (warp & 3);
            if (g0 < R0) {
              if constexpr (TRI) store_block32_2p(reinterpret_cast<float*>(c.X) + warp * 640, o, S + (size_t)g0 * ld + R0 + c0, ld, lane);
              else if constexpr (TMA == 2) {
                // in place: U[m][R0 + c] = Linv[R0 + c][m] goes where L[i][m] was (dead once this step's calls have read their B
                // tiles): the warp's 32 rows m are 128 contiguous bytes of row R0 + c -- no transposition tile needed
                float* dst = S + (size_t)(R0 + c0) * ld + g0 + lane;
#pragma unroll
                for (int q = 0; q < 32; ++q) dst[(size_t)q * ld] = o[q];
              }
              else store_block32(reinterpret_cast<float*>(c.X) + warp * 1152, o, S + (size_t)g0 * ld + R0 + c0, ld, lane);
            }
          }
          TICK(13);
          if (m < R0) {
            float dot = 0.f;
#pragma unroll
            for (int q = 0; q < 32; ++q) {
              if (m < T && R0 + c0 + q < T) tr_part = fmaf(o[q], o[q], tr_part);
              dot = fmaf(o[q], c.z[R0 + c0 + q], dot);
            }
            if (half_id) c.tmp[row] = dot;   // the two column halves of a row are combined in a fixed order
            else hdot = dot;
          }
          wsync();
          if (m < R0 && half_id == 0) c.al[m] += hdot + c.tmp[row];
          wsync();
        }
      }
    }

    TICK(10);
    // =============================== reductions and outputs (identical to the SIMT kernel)
    float zz = 0.f, aa = 0.f, ar = 0.f, z22 = 0.f, z12 = 0.f;
    for (int i = tid; i < T; i += NT) {
      const float zi = c.z[i], ai = c.al[i], z2i = has2 ? c.z2[i] : 0.f;
      zz = fmaf(zi, zi, zz);
      z22 = fmaf(z2i, z2i, z22);
      z12 = fmaf(zi, z2i, z12);
      if (p.z_out) { p.z_out[((size_t)b * 2) * T + i] = zi; p.z_out[((size_t)b * 2 + 1) * T + i] = z2i; }
      aa = fmaf(ai, ai, aa);
      if (rb) ar = fmaf(ai, rb[i], ar);
      if (p.alpha && p.do_inverse) p.alpha[(size_t)b * T + i] = ai;
    }
    const float inv_quad = block_sum_w(zz, c.red);
    const float logdet = 2.f * block_sum_w(logdet_part, c.red);
    const float tr_inv = block_sum_w(tr_part, c.red);
    const float alal = block_sum_w(aa, c.red);
    const float alr = block_sum_w(ar, c.red);
    const float sz22 = block_sum_w(z22, c.red);
    const float sz12 = block_sum_w(z12, c.red);
    if (tid == 0) {
      if (p.scalars) {
        float* o = p.scalars + (size_t)b * NSCALARS;
        const float Tf = (float)T;
        o[0] = -0.5f * (inv_quad + logdet + Tf * 1.8378770664093453f) / Tf;
        o[1] = 0.5f * (alal - tr_inv) / Tf;
        o[2] = logdet; o[3] = inv_quad; o[4] = tr_inv; o[5] = alal; o[6] = alr; o[7] = jit_used;
        o[8] = sz22; o[9] = sz12;
        o[10] = p.raw_noise ? o[1] * sigmoid_dev(p.raw_noise[(size_t)b * p.raw_stride]) : 0.f;   // dMLL/draw_noise (softplus' = sigmoid)
        o[11] = p.raw_noise ? noise_from_raw_dev(p.raw_noise[(size_t)b * p.raw_stride]) : 0.f;
        for (int q = 12; q < NSCALARS; ++q) o[q] = 0.f;
      }
      if (p.info) p.info[b] = fail;
      if (p.pack_out) {   // what the rollout kernel needs of this series (rollout.cu); V[n-1] is still in shared memory
        float* o = p.pack_out + (size_t)b * NSERIES;
        o[0] = sz22; o[1] = sz12; o[2] = c.Vs[T - 1]; o[3] = p.pack_x[1] - p.pack_x[0]; o[4] = jit_used;
        o[5] = o[6] = o[7] = 0.f;
      }
      if (p.series_flag) {
        __threadfence();
        asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p.series_flag + b), "r"(1) : "memory");
      }
    }
    if (p.U_out && p.do_inverse) {
      // (L^-1)^T: strictly-upper 64-blocks live in the scratch (control-warp instance: transposed, in place of L), the
      // diagonal blocks in dinv (dinv[j][r][c] = Linv_jj[c][r])
      float* Uo = p.U_out + (size_t)b * T * T;
      for (int idx = tid; idx < T * T; idx += NT) {
        const int r = idx / T, cc = idx - r * T;
        const int rb = r >> 6, cb = cc >> 6;
        float v = 0.f;
        if (rb < cb) v = (TMA == 2) ? S[(size_t)cc * ld + r] : S[(size_t)r * ld + cc];
        else if (rb == cb) v = dinv[((size_t)rb * NB + (r & 63)) * NB + (cc & 63)];
        Uo[idx] = v;
      }
    }
    if (p.L_out) {
      float* Lo = p.L_out + (size_t)b * p.L_bstride;
      for (int idx = tid; idx < T * T; idx += NT) {
        const int r = idx / T, cc = idx - r * T;
        Lo[(size_t)r * p.ldl + cc] = (cc <= r) ? S[(size_t)r * ld + cc] : 0.f;
      }
    }
    wsync();
  }

  if (p.loss_out) {
    // scalar loss of the training step, -sum_b MLL_b, without a second launch: the last CTA to get here sums the per-series
    // values in a fixed order (thread t takes b = t, t + 256, ...; then the block reduction), so the result does not
    // depend on which CTA happens to be last.
    __threadfence();
    wsync();
    if (tid == 0) {
      const unsigned int ticket = atomicAdd(p.done_counter, 1u);
      *c.flag = (ticket == gridDim.x - 1) ? 1 : 0;
    }
    wsync();
    if (*c.flag) {
      __threadfence();
      float acc = 0.f;
      for (int b = tid; b < p.B; b += NT) acc += __ldcg(p.scalars + (size_t)b * NSCALARS);
      const float tot = block_sum_w(acc, c.red);
      if (tid == 0) {
        p.loss_out[0] = -tot;
        *p.done_counter = 0u;   // ready for the next launch on this stream
      }
      if ((p.ex.peers || p.ex.totals) && tid < 32) exchange_partial_warp(p.ex, -tot, tid);

    }
  }
#ifdef VOLT_PROFILE
  TICK(11);
  if (tid == 0 && blockIdx.x == 0 && p.z_out == nullptr && p.alpha) {
    for (int i = 0; i < 12; ++i) p.alpha[i] = (float)seg[i];
    for (int i = 0; i < 8; ++i) { p.alpha[12 + i] = (float)g_diag_prof[i]; g_diag_prof[i] = 0; }
    if constexpr (TMA != 0) for (int i = 0; i < 8; ++i) p.alpha[20 + i] = (float)tp.prof[i];
    if constexpr (TMA == 2) for (int i = 0; i < 8; ++i) { p.alpha[28 + i] = (float)g_w2_prof[i]; g_w2_prof[i] = 0; }
    for (int i = 0; i < 4; ++i) p.alpha[36 + i] = (float)seg[12 + i];
  }
#endif
  tc_fence_before();
  __syncthreads();     // every thread of the CTA (the control warps of the TMA = 2 instance wait here too)
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(c.tmem), "n"(TCOLS) : "memory");
}

}  // namespace tc

// Tensor maps of the TMA instance: the scratch arena as a 2-D fp32 tensor (grid * Tp rows of Tp floats); boxes of
// 16 floats x 128 rows (A operand) and 16 floats x 64 rows (B operand), SWIZZLE_64B.  Encoded on the host through the
// driver entry point (no link-time dependency on libcuda).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int encode_scratch_map(CUtensorMap* map, float* base, int Tp, long long rows, int box_rows, int box_cols = 16) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || !f) {
      set_error("cuTensorMapEncodeTiled is not available from this driver");
      return VOLT_ERR_CUDA;
    }
    fn = reinterpret_cast<EncodeTiledFn>(f);
  }
  const cuuint64_t dims[2] = {(cuuint64_t)Tp, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)Tp * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1u, 1u};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        box_cols == 16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (Tp=%d rows=%lld box_rows=%d)", (int)r, Tp, rows, box_rows);
    return VOLT_ERR_CUDA;
  }
  return VOLT_OK;
}

template <bool TRI, bool HOSTIN, int TMA = 0>
static int launch_tc(MllParams p, cudaStream_t st, size_t smem, int per_sm) {
  static size_t attr_smem_dev[16] = {};   // function attributes are per device
  size_t& attr_smem = attr_smem_dev[device_slot()];
  if (smem > attr_smem) {
    int s = check_cuda(cudaFuncSetAttribute(tc::mll_batched_tc_kernel<TRI, HOSTIN, TMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                       "cudaFuncSetAttribute(mll_batched_tc_kernel)");
    if (s) return s;
    attr_smem = smem;
  }
  int grid = sm_count() * per_sm;
  if (grid > p.B) grid = p.B;
  if (grid < 1) grid = 1;
  const size_t per_cta = (size_t)p.Tp * p.Tp + (size_t)p.nb * NB * NB;
  void* ws = nullptr;
  int s = get_workspace(per_cta * grid * sizeof(float), &ws, 0, st);
  if (s) return s;
  p.scratch = reinterpret_cast<float*>(ws);
  p.dinv = p.scratch + (size_t)grid * p.Tp * p.Tp;
  CUtensorMap tmA, tmB, tmAt;
  memset(&tmA, 0, sizeof(tmA));
  memset(&tmB, 0, sizeof(tmB));
  memset(&tmAt, 0, sizeof(tmAt));
  if (TMA) {
    s = encode_scratch_map(&tmA, p.scratch, p.Tp, (long long)grid * p.Tp, 128);
    if (s) return s;
    s = encode_scratch_map(&tmB, p.scratch, p.Tp, (long long)grid * p.Tp, 64);
    if (s) return s;
    if (TMA == 2) {   // phase B's A operand, stored transposed: 128 floats x 16 rows, dense
      s = encode_scratch_map(&tmAt, p.scratch, p.Tp, (long long)grid * p.Tp, 16, 128);
      if (s) return s;
    }
  }
  tc::mll_batched_tc_kernel<TRI, HOSTIN, TMA><<<grid, TMA == 2 ? tc::W2_THREADS : NT, smem, st>>>(p, tmA, tmB, tmAt);
  return check_cuda(cudaGetLastError(), "mll_batched_tc_kernel");
}

// resident CTAs (= series in flight) of the batched kernel for series of length T: the first wave of a batch
int mll_tc_resident_ctas(int T, int two_rhs) {
  const int Tp = (T + NB - 1) / NB * NB;
  const size_t vec = sizeof(float) * (size_t)((two_rhs ? 4 : 3) * Tp + NB + 2 * NB + 32 + 12 + 32);
  const size_t smem3 = tc::Y_VEC_OFF + vec;
  if (3 * (smem3 + 1024) <= 233472) return 3 * sm_count();
  const size_t smem = tc::VEC_OFF + vec;
  int per_sm = (int)(233472 / (smem + 1024));
  return sm_count() * (per_sm > 2 ? 2 : (per_sm < 1 ? 1 : per_sm));
}

// Default choice between the register-staged instances and the TMA-fed one, from measurements on the B200.
// B200, one launch (ms), control-warp TMA instance vs the register-staged instances (two / three CTAs per SM):
// 1024 x 512: 1.75 vs 1.81; 256 x 1024: 2.27 vs 2.84; 296 x 512: 0.514 vs 0.565; 148 x 512: 0.397 vs 0.431.
// Short series are dominated by the diagonal blocks, where a third resident CTA per SM helps more than a faster GEMM loop:
// 2048 x 256: 0.97 vs 0.89; 4096 x 128: 0.76 vs 0.64 -- those stay on the three-CTA register-staged instance.
static bool use_tma_default(int B, int T, int sms) {
  (void)B; (void)sms;
  return (T + NB - 1) / NB * NB >= 448;
}

int launch_mll_batched_tc(MllParams p, cudaStream_t st) {
  p.Tp = (p.T + NB - 1) / NB * NB;
  p.nb = p.Tp / NB;
  const size_t vec = sizeof(float) * (size_t)(((p.resid2 ? 4 : 3) + (p.stage_in ? 1 : 0)) * p.Tp + NB + 2 * NB + 32 + 12 + 32);
  const size_t smem_total = 233472, smem_cta_reserved = 1024;   // sm_100: 228 KB per SM, 1 KB reserved per resident CTA
  // three resident CTAs per SM when the small shared-memory map fits three times (T <= 832); VOLT_TC_CTAS=2 / 3 forces the
  // double-buffered two-CTA / the three-CTA kernel (A/B timing)
  static const int forced = [] { const char* e = getenv("VOLT_TC_CTAS"); return e ? atoi(e) : 0; }();   // 2 / 3: A/B timing
  const int force2 = (forced == 2);
  const size_t smem3 = tc::Y_VEC_OFF + vec;
  const bool hostin = (p.vol_in != nullptr);
  // The double-buffered two-CTA instance has the shorter dependency chain per series (T = 512: 0.423 vs 0.445 ms for one
  // wave at one CTA per SM, 0.551 vs 0.583 ms at two); the three-CTA instance wins once a third chain per SM can be
  // overlapped (0.757 ms for 444 series vs 0.937).  So: batches of at most two series per SM, and batches whose last
  // wave would be a single straggler per SM (3 < B / SMs <= 4), take the two-CTA instance.
  const int sms = sm_count();
  {
    static const int tma_first = [] { const char* e = getenv("VOLT_TC_TMA"); return e ? atoi(e) : -1; }();
    const size_t smem_w0 = tc::W_VEC_OFF + vec;
    const bool eligible = !p.L_out && 2 * (smem_w0 + smem_cta_reserved) <= smem_total;
    if (eligible && forced == 0 && (tma_first == 2 || (tma_first < 0 && use_tma_default(p.B, p.T, sms))))   // VOLT_TC_TMA=0 / VOLT_TC_CTAS: A/B timing
      return hostin ? launch_tc<false, true, 2>(p, st, smem_w0, 2) : launch_tc<false, false, 2>(p, st, smem_w0, 2);
    if (tma_first == 3 && !hostin && !p.L_out && 2 * (smem_w0 + smem_cta_reserved) <= smem_total)
      return launch_tc<false, false, 1>(p, st, smem_w0, 2);      // VOLT_TC_TMA=3: the in-line TMA instance
  }
  const bool prefer2 = forced != 3 && ((p.B <= 2 * sms) || (p.B > 3 * sms && p.B <= 4 * sms));
  if (!force2 && !prefer2 && 3 * (smem3 + smem_cta_reserved) <= smem_total)
    return hostin ? launch_tc<true, true>(p, st, smem3, 3) : launch_tc<true, false>(p, st, smem3, 3);
  // TMA-fed instance (operand tiles through cp.async.bulk.tensor, two CTAs per SM): device-pointer entry, callers that do
  // not ask for the factor itself (its scratch keeps the inverse blocks on the diagonal).  VOLT_TC_TMA=0 disables it.
  static const int tma_env = [] { const char* e = getenv("VOLT_TC_TMA"); return e ? atoi(e) : -1; }();
  const size_t smem_w = tc::W_VEC_OFF + vec;
  const bool tma_ok = !hostin && !p.L_out && 2 * (smem_w + smem_cta_reserved) <= smem_total;
  const bool want_tma = tma_env == 2 || (tma_env != 0 && use_tma_default(p.B, p.T, sms));
  if (tma_ok && want_tma && (forced != 2 || tma_env == 2)) return launch_tc<false, false, 2>(p, st, smem_w, 2);
  const size_t smem = tc::VEC_OFF + vec;
  if (smem > 227 * 1024) {
    set_error("mll_batched_tc: T=%d needs %zu bytes of shared memory (max 227 KB)", p.T, smem);
    return VOLT_ERR_ARG;
  }
  int per_sm = (int)(smem_total / (smem + smem_cta_reserved));
  if (per_sm > 2) per_sm = 2;
  if (per_sm < 1) per_sm = 1;
  return hostin ? launch_tc<false, true>(p, st, smem, per_sm) : launch_tc<false, false>(p, st, smem, per_sm);
}

}  // namespace volt

// hodlr_leaf.cuh — K4, tensor-pipe version: leaf build + blocked LDL^T with DMMA trailing updates.
// Replaces get_exact_matrix + Eigen LDLT (hodlr.h:122-133, 225-227, 87-89).  (A 32-column blocked leaf SOLVE was tried
// and measured slower than leaf_solve_kernel for the 1..160 right-hand sides of this path, so it was dropped.)
//
// One CTA per leaf, 32-column panels:
//   (1) the 32x32 diagonal block is factorised by ONE warp with shuffles only (no block barriers),
//   (2) the panel below it is solved one row per thread and also staged in shared memory as [k][row],
//   (3) the trailing update A22 -= L21 D L21^T runs on the FP64 tensor pipe (mma.sync.m8n8k4.f64): 32x32 output tiles of
//       the lower triangle are dealt to the 8 warps, both operands come from the same shared panel (the D scaling is
//       applied to the B fragment on the fly), leading dimension = 4 (mod 16) doubles -> conflict-free fragment loads.
#pragma once

#include "gemm_dmma.cuh"
#include "hodlr_kernels.cuh"

namespace bgp {

constexpr int LF_THREADS = 256;
constexpr int LF_NB = 32;

__host__ __device__ inline int lf_panel_ld(int max_m) { return ((max_m + 15) / 16) * 16 + 4; }

__global__ void __launch_bounds__(LF_THREADS) leaf_factor_dmma_kernel(const DevProgram* __restrict__ gprog,
                                                                      const double* __restrict__ x,
                                                                      const double* __restrict__ diag,
                                                                      const LeafDesc* __restrict__ leaves,
                                                                      double* __restrict__ Lbuf,
                                                                      double* __restrict__ leaf_logdet, int ldp) {
  extern __shared__ __align__(16) double lf_panel[];  // [LF_NB][ldp]: L21(i, k) at lf_panel[k*ldp + i]
  __shared__ DevProgram P;
  __shared__ double red[32];
  __shared__ double dblk[LF_NB][LF_NB + 1];
  __shared__ double dd[LF_NB], dinv[LF_NB];
  stage_program(&P, gprog);
  __syncthreads();
  const LeafDesc lf = leaves[blockIdx.x];
  const int m = lf.size, nd = P.ndim;
  double* A = Lbuf + lf.off;
  const double* xs = x + (int64_t)lf.start * nd;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  for (int j = 0; j < m; ++j) {
    for (int i = j + threadIdx.x; i < m; i += LF_THREADS) {
      double v = kernel_value(P, xs + (int64_t)i * nd, xs + (int64_t)j * nd);
      if (i == j) v += diag[lf.start + i];
      A[(int64_t)j * m + i] = v;
    }
  }
  __syncthreads();

  double logdet = 0.0;
  for (int k0 = 0; k0 < m; k0 += LF_NB) {
    const int nb = min(LF_NB, m - k0);
    for (int t = threadIdx.x; t < LF_NB * LF_NB; t += LF_THREADS) {
      const int i = t % LF_NB, j = t / LF_NB;
      dblk[i][j] = (i < nb && j < nb && i >= j) ? A[(int64_t)(k0 + j) * m + k0 + i] : ((i == j) ? 1.0 : 0.0);
    }
    __syncthreads();
    // (1) warp 0: LDL^T of the diagonal block, lane = row
    if (warp == 0) {
      for (int k = 0; k < nb; ++k) {
        const double d = dblk[k][k];
        double l = 0.0;
        if (lane > k && lane < nb) { l = dblk[lane][k] / d; dblk[lane][k] = l; }
        __syncwarp();
        if (lane > k && lane < nb) {
          const double ld = l * d;
          for (int j = k + 1; j <= lane; ++j) dblk[lane][j] -= ld * dblk[j][k];
        }
        __syncwarp();
      }
      if (lane < nb) {
        const double d = dblk[lane][lane];
        dd[lane] = d;
        dinv[lane] = 1.0 / d;
        logdet += log(fabs(d));
      } else {
        dd[lane] = 0.0;
        dinv[lane] = 0.0;
      }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < nb * nb; t += LF_THREADS) {
      const int i = t % nb, j = t / nb;
      if (i >= j) A[(int64_t)(k0 + j) * m + k0 + i] = dblk[i][j];
    }
    const int rem = m - k0 - nb;
    if (rem <= 0) break;
    // (2) panel: L21 = A21 L11^-T D^-1, one row per thread; result to global and to the shared panel
    const int rem_pad = ((rem + 31) / 32) * 32;
    for (int i = threadIdx.x; i < rem_pad; i += LF_THREADS) {
      if (i < rem) {
        double* row = A + k0 + nb + i;
        double w[LF_NB];
#pragma unroll
        for (int j = 0; j < LF_NB; ++j) w[j] = (j < nb) ? row[(int64_t)(k0 + j) * m] : 0.0;
#pragma unroll
        for (int j = 0; j < LF_NB; ++j) {
          if (j < nb) {
            double s = w[j];
#pragma unroll
            for (int q = 0; q < j; ++q) s -= w[q] * dblk[j][q];
            w[j] = s;
          }
        }
#pragma unroll
        for (int j = 0; j < LF_NB; ++j) {
          const double l = (j < nb) ? w[j] * dinv[j] : 0.0;
          if (j < nb) row[(int64_t)(k0 + j) * m] = l;
          lf_panel[j * ldp + i] = l;
        }
      } else {
#pragma unroll
        for (int j = 0; j < LF_NB; ++j) lf_panel[j * ldp + i] = 0.0;
      }
    }
    __syncthreads();
    // (3) trailing update on the tensor pipe
    {
      const int T = rem_pad / 32;
      const int n_tiles = T * (T + 1) / 2;
      const int lr = lane >> 2, lc = lane & 3;
      double* C = A + (int64_t)(k0 + nb) * m + k0 + nb;
      for (int t = warp; t < n_tiles; t += LF_THREADS / 32) {
        // t -> (ti >= tj) in the lower triangle
        int ti = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
        while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
        while (ti * (ti + 1) / 2 > t) --ti;
        const int tj = t - ti * (ti + 1) / 2;
        double acc[4][4][2];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) { acc[a][b][0] = 0.0; acc[a][b][1] = 0.0; }
#pragma unroll
        for (int kk = 0; kk < LF_NB / 4; ++kk) {
          const int k = kk * 4 + lc;
          const double dk = dd[k];
          double af[4], bf[4];
#pragma unroll
          for (int a = 0; a < 4; ++a) af[a] = lf_panel[k * ldp + ti * 32 + a * 8 + lr];
#pragma unroll
          for (int b = 0; b < 4; ++b) bf[b] = lf_panel[k * ldp + tj * 32 + b * 8 + lr] * dk;
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) dmma884(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
        }
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const int i = ti * 32 + a * 8 + lr;
#pragma unroll
          for (int b = 0; b < 4; ++b)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int j = tj * 32 + b * 8 + 2 * lc + e;
              if (i < rem && j < rem && i >= j) C[(int64_t)j * m + i] -= acc[a][b][e];
            }
        }
      }
    }
    __syncthreads();
  }
  logdet = block_sum(logdet, red);
  if (threadIdx.x == 0) leaf_logdet[blockIdx.x] = logdet;
}

}  // namespace bgp

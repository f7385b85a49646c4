// gemm_dmma.cuh — batched FP64 GEMM on the tensor pipe (mma.sync.aligned.m8n8k4.f64 -> SASS DMMA.8x8x4).
//
// tcgen05 has no f64 kind, so the dense contractions of this solver (Cholesky trailing updates, HODLR Gram / update
// products, triangular-solve updates) run on DMMA; measured issue-bound peak on B200: 37.2 TFLOP/s
// (tools/fp64_peaks.cu, profiles/fp64_peaks_r01.txt; cuBLAS dgemm reaches 35.4).
//
//   C (M x N, column-major ldc)  op=  A' (M x K) * B' (K x N)
//   A' element (m, k):  A_KCONTIG ? A[m*lda + k] : A[k*lda + m]      (i.e. A stored K x M   or   M x K, column-major)
//   B' element (k, n):  B_KCONTIG ? B[n*ldb + k] : B[k*ldb + n]      (i.e. B stored K x N   or   N x K, column-major)
//   op: GD_SUB  C -= A'B'   |  GD_ATOMIC_ADD  atomicAdd(C, A'B') (split-K)   |  lower: only entries with m >= n
//
// 128 x 128 x 16 CTA tile, 256 threads = 8 warps (2 x 4), 64 x 32 warp tile = 8 x 4 DMMA tiles, 3-stage cp.async
// pipeline.  Shared-memory leading dimensions are = 4 (mod 16) doubles so that the (lane/4, lane%4) fragment pattern
// of a half-warp touches 16 distinct 8-byte banks.
#pragma once

#include "common.cuh"

namespace bgp {

constexpr int GD_BM = 128, GD_BN = 128, GD_BK = 16, GD_THREADS = 256, GD_STAGES = 3;
constexpr int GD_LDK = GD_BK + 4;    // [mn][k] layout
constexpr int GD_LDM = GD_BM + 4;    // [k][mn] layout
constexpr int GD_TILE_ELEMS = (GD_BM * GD_LDK > GD_BK * GD_LDM) ? GD_BM * GD_LDK : GD_BK * GD_LDM;  // 2560 doubles
constexpr size_t GD_SMEM_BYTES = sizeof(double) * 2 * GD_STAGES * GD_TILE_ELEMS;                      // 122880

enum { GD_SUB = 0, GD_ATOMIC_ADD = 1 };

struct GemmDesc {
  const double* A;
  const double* B;
  double* C;
  int M, N, K, mode;   // mode: GD_SUB | GD_ATOMIC_ADD, bit 8: lower-triangular output only
  int64_t lda, ldb, ldc;
};
constexpr int GD_LOWER = 256;

__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, bool valid) {
  const uint32_t s = smem_u32(smem);
  const int sz = valid ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ double gd_ld_global(const double* p) {
  double v;
  asm volatile("ld.global.f64 %0, [%1];\n" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void gd_st_global(double* p, double v) {
  asm volatile("st.global.f64 [%0], %1;\n" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ void gd_red_add_global(double* p, double v) {
  asm volatile("red.global.add.f64 [%0], %1;\n" ::"l"(p), "d"(v) : "memory");
}

// Loader of one operand's BK-slabs (128 rows/cols x 16 k) into shared memory.  Everything that does not depend on the
// k-slab — the thread's running global pointer, its shared-memory offset and the row-validity bits of its 8 copies — is
// computed ONCE per tile; per slab only pointer increments and the k-bound test remain (the straightforward version
// spent ~275 integer / predicate instructions per slab and warp between the barrier and the first DMMA, with the
// tensor pipe idle: cuobjdump of the previous build).
template <bool KCONTIG>
struct GdLoader {
  const double* p;  // running pointer: element (mn0 + mn_t, k of the NEXT slab + k_t) of this thread's first copy
  int64_t gstep;    // global stride between the thread's consecutive copies i -> i + 1
  int64_t kstep;    // global stride of one BK-slab
  int soff;         // shared-memory offset of the first copy
  int kleft;        // K - (k of the thread's first copy in the next slab): copy i is inside K iff its k offset < kleft
  unsigned okm;     // bit i: row/col of copy i is inside the matrix
  static constexpr int NCOPY = (GD_BM * GD_BK) / GD_THREADS;                                             // 8
  static constexpr int DSTEP = KCONTIG ? (GD_THREADS / GD_BK) * GD_LDK : (GD_THREADS / GD_BM) * GD_LDM;  // smem stride
  __device__ __forceinline__ void init(const double* G, int64_t ld, int mn0, int mn_max, int K) {
    okm = 0;
    if (KCONTIG) {
      // global element (mn, k) at G[mn*ld + k]; shared sm[mn*GD_LDK + k]; copy i: k = tid % 16, mn = tid / 16 + 16 i
      const int k = threadIdx.x % GD_BK, mn = threadIdx.x / GD_BK;
      p = G + (int64_t)(mn0 + mn) * ld + k;
      gstep = (int64_t)(GD_THREADS / GD_BK) * ld;
      kstep = GD_BK;
      soff = mn * GD_LDK + k;
      kleft = K - k;
#pragma unroll
      for (int i = 0; i < NCOPY; ++i)
        if (mn0 + mn + i * (GD_THREADS / GD_BK) < mn_max) okm |= 1u << i;
    } else {
      // global element (mn, k) at G[k*ld + mn]; shared sm[k*GD_LDM + mn]; copy i: mn = tid % 128, k = tid / 128 + 2 i
      const int mn = threadIdx.x % GD_BM, k = threadIdx.x / GD_BM;
      p = G + (int64_t)k * ld + mn0 + mn;
      gstep = (int64_t)(GD_THREADS / GD_BM) * ld;
      kstep = (int64_t)GD_BK * ld;
      soff = k * GD_LDM + mn;
      kleft = K - k;
      if (mn0 + mn < mn_max) okm = 0xffu;
    }
    // keep the loop-invariant state in registers: without this the compiler re-derives okm / soff from
    // %ctaid / %tid inside the k-loop (rematerialisation), ~100 extra instructions per slab
    asm volatile("" : "+r"(okm), "+r"(soff));
  }
  // copy the NEXT slab into `sm` and advance.  A masked-off copy passes its (possibly out-of-range) address with
  // src-size 0: nothing is read, the destination is zero-filled.
  __device__ __forceinline__ void load_next(double* sm) {
    const double* q = p;
    double* d = sm + soff;
    // validity of the 8 copies in one mask: rows/cols from okm, k from kleft (all-ones except in the last slab)
    unsigned m = okm;
    if (KCONTIG) {
      if (kleft <= 0) m = 0;
    } else if (kleft < 2 * NCOPY) {
      const int nv = kleft <= 0 ? 0 : (kleft + 1) / 2;  // copies i with 2 i < kleft
      m &= (1u << nv) - 1u;
    }
#pragma unroll
    for (int i = 0; i < NCOPY; ++i) {
      cp_async8(d + i * DSTEP, q, (m >> i) & 1u);
      q += gstep;
    }
    p += kstep;
    kleft -= GD_BK;
  }
};

template <bool A_KCONTIG, bool B_KCONTIG>
__global__ void __launch_bounds__(GD_THREADS) gemm_dmma_kernel(const GemmDesc* __restrict__ descs, const int* info) {
  if (info && *info != 0) return;
  const GemmDesc d = descs[blockIdx.z];
  const int m0 = blockIdx.x * GD_BM, n0 = blockIdx.y * GD_BN;
  if (m0 >= d.M || n0 >= d.N) return;
  const bool lower = (d.mode & GD_LOWER) != 0;
  if (lower && m0 + GD_BM <= n0) return;  // tile strictly above the diagonal
  extern __shared__ __align__(16) double gd_smem[];
  double* sA = gd_smem;
  double* sB = gd_smem + GD_STAGES * GD_TILE_ELEMS;

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wm = warp >> 2, wn = warp & 3;  // 2 x 4 warps
  const int lr = lane >> 2, lc = lane & 3;

  double acc[8][4][2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

  const int nk = (d.K + GD_BK - 1) / GD_BK;
  GdLoader<A_KCONTIG> ldA;
  GdLoader<B_KCONTIG> ldB;
  ldA.init(d.A, d.lda, m0, d.M, d.K);
  ldB.init(d.B, d.ldb, n0, d.N, d.K);
  // prologue
#pragma unroll
  for (int s = 0; s < GD_STAGES - 1; ++s) {
    if (s < nk) {
      ldA.load_next(sA + s * GD_TILE_ELEMS);
      ldB.load_next(sB + s * GD_TILE_ELEMS);
    }
    cp_async_commit();
  }
  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<GD_STAGES - 2>();
    __syncthreads();
    {
      const int nxt = kt + GD_STAGES - 1;
      if (nxt < nk) {
        const int s = nxt % GD_STAGES;
        ldA.load_next(sA + s * GD_TILE_ELEMS);
        ldB.load_next(sB + s * GD_TILE_ELEMS);
      }
      cp_async_commit();
    }
    const double* a_s = sA + (kt % GD_STAGES) * GD_TILE_ELEMS;
    const double* b_s = sB + (kt % GD_STAGES) * GD_TILE_ELEMS;
#pragma unroll
    for (int kk = 0; kk < GD_BK / 4; ++kk) {
      double af[8], bf[4];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int m = wm * 64 + i * 8 + lr, k = kk * 4 + lc;
        af[i] = A_KCONTIG ? a_s[m * GD_LDK + k] : a_s[k * GD_LDM + m];
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = wn * 32 + j * 8 + lr, k = kk * 4 + lc;
        bf[j] = B_KCONTIG ? b_s[n * GD_LDK + k] : b_s[k * GD_LDM + n];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
  }
  cp_async_wait<0>();

  // epilogue: thread holds C(m = .. + lr, n = .. + 2*lc + {0,1}).  C is read-modify-written with explicit .global
  // accesses (the descriptor's pointer is generic to the compiler, which would otherwise serialise 64 dependent
  // LD -> ST round trips per thread): 16 loads in flight per batch, fire-and-forget reductions in the split-K mode.
  const int op = d.mode & 0xff;
#pragma unroll
  for (int i2 = 0; i2 < 8; i2 += 2) {
    double cv[2][4][2];
    if (op == GD_SUB) {
#pragma unroll
      for (int ii = 0; ii < 2; ++ii) {
        const int m = m0 + wm * 64 + (i2 + ii) * 8 + lr;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int n = n0 + wn * 32 + j * 8 + 2 * lc + e;
            const bool ok = m < d.M && n < d.N && !(lower && m < n);
            cv[ii][j][e] = ok ? gd_ld_global(d.C + (int64_t)n * d.ldc + m) : 0.0;
          }
        }
      }
    }
#pragma unroll
    for (int ii = 0; ii < 2; ++ii) {
      const int m = m0 + wm * 64 + (i2 + ii) * 8 + lr;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int n = n0 + wn * 32 + j * 8 + 2 * lc + e;
          const bool ok = m < d.M && n < d.N && !(lower && m < n);
          if (!ok) continue;
          double* c = d.C + (int64_t)n * d.ldc + m;
          if (op == GD_SUB) gd_st_global(c, cv[ii][j][e] - acc[i2 + ii][j][e]);
          else gd_red_add_global(c, acc[i2 + ii][j][e]);
        }
      }
    }
  }
}

// host launcher: all descriptors of a batch share the grid; maxM / maxN bound the tile grid
template <bool A_KCONTIG, bool B_KCONTIG>
inline int gemm_dmma_launch(const GemmDesc* d_descs, int batch, int maxM, int maxN, const int* info, cudaStream_t s) {
  if (batch <= 0 || maxM <= 0 || maxN <= 0) return BGP_OK;
  // (the attribute is per device / context: set it on every call, it is cheap)
  cudaFuncSetAttribute(gemm_dmma_kernel<A_KCONTIG, B_KCONTIG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GD_SMEM_BYTES);
  for (int b0 = 0; b0 < batch; b0 += 65535) {
    const int nb = batch - b0 < 65535 ? batch - b0 : 65535;
    dim3 grid((maxM + GD_BM - 1) / GD_BM, (maxN + GD_BN - 1) / GD_BN, nb);
    gemm_dmma_kernel<A_KCONTIG, B_KCONTIG><<<grid, GD_THREADS, GD_SMEM_BYTES, s>>>(d_descs + b0, info);
    BGP_LAUNCH_CHECK();
  }
  return BGP_OK;
}

}  // namespace bgp

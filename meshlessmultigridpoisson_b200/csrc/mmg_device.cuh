// Device-side helpers shared by the kernel translation units (sm_100a).
#pragma once
#include <cuda_runtime.h>

#include "mmg_internal.hpp"

namespace mmg {

constexpr int kBlock = 256;
constexpr unsigned long long kSentinelBits = 0xFFF8DEADBEEF0001ull;  // quiet NaN with a payload no computation produces

enum { OP_SPMV = 0, OP_RESID = 1, OP_PROLONG = 2, OP_RESTRICT = 3 };

__device__ __forceinline__ const double* row_val(const HybView& A, int r) {
  return reinterpret_cast<const double*>(A.chunks + (size_t)r * A.chunk_bytes);
}
__device__ __forceinline__ const int* row_col(const HybView& A, int r) {
  return reinterpret_cast<const int*>(A.chunks + (size_t)r * A.chunk_bytes + (size_t)A.W * 8);
}
__device__ __forceinline__ int ovf_find(const HybView& A, int row) {  // index into ovf_rows, rows with len>W only
  int lo = 0, hi = A.n_ovf - 1;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (A.ovf_rows[mid] < row) lo = mid + 1; else hi = mid;
  }
  return lo;
}
template <int LPR>
__device__ __forceinline__ double group_sum(double v, unsigned mask) {
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) v = __dadd_rn(v, __shfl_xor_sync(mask, v, o));
  return v;
}
template <int LPR>
__device__ __forceinline__ unsigned group_mask(int lane) {
  if (LPR == 32) return 0xffffffffu;
  return ((1u << LPR) - 1u) << ((lane / LPR) * LPR);
}
__device__ __forceinline__ double ld_relaxed(const double* p) {
  double v;
  asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_relaxed_sys(const double* p) {   // polls of words a peer GPU stores with st.relaxed.sys
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_sys(double* p, double v) {
  asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed(double* p, double v) {
  asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ bool is_sentinel(double v) { return (unsigned long long)__double_as_longlong(v) == kSentinelBits; }

// block-wide sum of two doubles; result valid in thread 0
__device__ __forceinline__ void block_sum2(double& a, double& b) {
  __shared__ double sa[kBlock / 32], sb[kBlock / 32];
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { sa[w] = a; sb[w] = b; }
  __syncthreads();
  if (w == 0) {
    a = l < (blockDim.x >> 5) ? sa[l] : 0.0;
    b = l < (blockDim.x >> 5) ? sb[l] : 0.0;
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
  }
}

// L2 residency hints: the matrix stream is read once (evict_first), the gathered vector is re-read by every row that
// references it (evict_last) so that the 8*N-byte x stays L2 resident under the 12*nnz-byte stream.
__device__ __forceinline__ unsigned long long policy_evict_last() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ unsigned long long policy_evict_first() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ double ldg_keep(const double* p, unsigned long long pol) {   // gathered vector: keep in L2
  double v;
  asm volatile("ld.global.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ double ldg_stream_f64(const double* p, unsigned long long pol) {   // matrix stream: read once
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ int ldg_stream_s32(const int* p, unsigned long long pol) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
  return v;
}


}  // namespace mmg

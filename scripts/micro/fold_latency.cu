// Micro-benchmark behind DESIGN.md §5: cost per entry of the in-order row fold (a serial fp64 add chain)
// fed (a) from registers, (b) by warp shuffles, (c) by shared-memory broadcast loads.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_chain(double* out, double s, double p, long long* cyc) {
  long long t0 = clock64();
#pragma unroll 64
  for (int i = 0; i < 4096; i++) s = __dsub_rn(s, p);
  long long t1 = clock64();
  if (threadIdx.x == 0) { *out = s; cyc[0] = t1 - t0; }
}
__global__ void k_shfl(double* out, double s, double p0, long long* cyc) {
  double p = p0 + threadIdx.x;
  long long t0 = clock64();
  for (int r = 0; r < 128; r++) {
#pragma unroll
    for (int g8 = 0; g8 < 4; g8++) {
      double q[8];
#pragma unroll
      for (int l = 0; l < 8; l++) q[l] = __shfl_sync(0xffffffffu, p, g8 * 8 + l);
#pragma unroll
      for (int l = 0; l < 8; l++) s = __dsub_rn(s, q[l]);
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) { *out = s; cyc[1] = t1 - t0; }
}
__global__ void k_lds(double* out, double s, double p0, long long* cyc) {
  __shared__ double buf[32];
  double p = p0 + threadIdx.x;
  long long t0 = clock64();
  for (int r = 0; r < 128; r++) {
    buf[threadIdx.x] = p + r;
    __syncwarp();
    const double2* b2 = reinterpret_cast<const double2*>(buf);
    double2 q[16];
#pragma unroll
    for (int l = 0; l < 16; l++) q[l] = b2[l];
#pragma unroll
    for (int l = 0; l < 16; l++) { s = __dsub_rn(s, q[l].x); s = __dsub_rn(s, q[l].y); }
    __syncwarp();
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) { *out = s; cyc[2] = t1 - t0; }
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 8); cudaMallocManaged(&cyc, 32);
  for (int rep = 0; rep < 2; rep++) {
    k_chain<<<1, 32>>>(out, 1.0, 1e-9, cyc); k_shfl<<<1, 32>>>(out, 1.0, 1e-9, cyc); k_lds<<<1, 32>>>(out, 1.0, 1e-9, cyc);
    cudaDeviceSynchronize();
  }
  printf("dependent DSUB chain: %.1f cycles/op\nshuffle-fed fold: %.1f cycles/entry\nshared-memory-fed fold: %.1f cycles/entry\n", cyc[0] / 4096.0, cyc[1] / 4096.0, cyc[2] / 4096.0);
  return 0;
}

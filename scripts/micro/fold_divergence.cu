// Does a formally diverged warp run the shuffle-fed fold slowly?  (BRA.DIV slow path in front of every shuffle group)
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double ld_relaxed(const double* p) { double v; asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ double fold(double s, double p, int cnt, bool sync) {
  if (sync) __syncwarp();
#pragma unroll
  for (int g8 = 0; g8 < 4; g8++) {
    if (cnt > g8 * 8) {
      double q[8];
#pragma unroll
      for (int l = 0; l < 8; l++) q[l] = __shfl_sync(0xffffffffu, p, g8 * 8 + l);
#pragma unroll
      for (int l = 0; l < 8; l++) s = __dsub_rn(s, q[l]);
    }
  }
  return s;
}
__global__ void k(int variant, const double* g, int cnt, double* out, long long* cyc) {
  const int lane = threadIdx.x;
  double p = 1e-9 + lane, s = 1.0;
  // lane-dependent waiting loop, like the watch phase
  const double* watch = lane == 0 ? g : lane == 1 ? g + 64 : nullptr;
  bool waiting = watch != nullptr;
  int spins = 0;
  while (__any_sync(0xffffffffu, waiting)) {
    if (waiting) waiting = ld_relaxed(watch) != 0.0 && spins < 3;
    spins++;
  }
  if (variant == 2) { if (lane & 1) p += 1e-12; else p -= 1e-12; }   // a divergent if before the fold
  long long t0 = clock64();
  for (int r = 0; r < 8; r++) s = fold(s, p, cnt, variant == 1);
  long long t1 = clock64();
  if (lane == 0) { *out = s; cyc[0] = t1 - t0; }
}
int main() {
  double *g, *out; long long* cyc;
  cudaMalloc(&g, 8000); cudaMemset(g, 0xff, 8000); cudaMalloc(&out, 8); cudaMallocManaged(&cyc, 8);
  const char* names[] = {"after a lane-dependent loop, no syncwarp", "same, __syncwarp() before the fold", "plus a divergent if"};
  for (int v = 0; v < 3; v++) for (int cnt : {32, 4}) {
    k<<<1, 32>>>(v, g, cnt, out, cyc); cudaDeviceSynchronize();
    printf("%-44s cnt %2d: %.0f cycles per fold call\n", names[v], cnt, cyc[0] / 8.0);
  }
  return 0;
}

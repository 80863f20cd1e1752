// What slows the in-order fold inside the lexicographic kernels?  Warp 0 times the shuffle-fed fold while the other
// warps of the CTA (and other CTAs on the SM) run one of the waiting behaviours the sweep kernel uses.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double ld_relaxed(const double* p) { double v; asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory"); return v; }
__global__ void k(int mode, int nwait_warps, double* g, double* out, long long* cyc, int* stop) {
  __shared__ unsigned long long bar;
  __shared__ double sm[64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned bar_a = (unsigned)__cvta_generic_to_shared(&bar);
  if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_a), "r"(1) : "memory");
  __syncthreads();
  if (warp == 0 && blockIdx.x == 0) {
    double p = 1e-9 + lane, s = 1.0;
    for (volatile int d = 0; d < 20000; d++) {}
    long long t0 = clock64();
    for (int r = 0; r < 32; r++) {
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
    if (lane == 0) { *out = s; cyc[0] = t1 - t0; atomicExch(stop, 1); }
    if (lane == 0) { unsigned long long st; asm volatile("mbarrier.arrive.shared::cta.b64 %0, [%1];" : "=l"(st) : "r"(bar_a) : "memory"); }
  } else if (warp <= nwait_warps) {
    if (mode == 1) {            // per-lane L2 polling
      while (!*(volatile int*)stop) { double v = ld_relaxed(g + (threadIdx.x * 37 + blockIdx.x * 4099) % 100000); if (v == 12345.0) break; }
    } else if (mode == 2) {     // two-lane watch polling
      while (!*(volatile int*)stop) { if (lane < 2) { double v = ld_relaxed(g + (threadIdx.x * 37 + blockIdx.x * 4099) % 100000); if (v == 12345.0) break; } __syncwarp(); }
    } else if (mode == 3) {     // mbarrier try_wait with suspend hint
      unsigned done = 0;
      while (!done && !*(volatile int*)stop) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar_a), "r"(0u), "r"(0x989680u) : "memory");
      }
    } else if (mode == 4) {     // shared-memory spin
      volatile double* s = sm;
      while (!*(volatile int*)stop) { if (s[lane] == 12345.0) break; }
    } else if (mode == 5) {     // nanosleep loop
      while (!*(volatile int*)stop) __nanosleep(200);
    }
  }
}
int main() {
  double *g, *out; long long* cyc; int* stop;
  cudaMalloc(&g, 800000); cudaMemset(g, 0, 800000); cudaMalloc(&out, 8); cudaMallocManaged(&cyc, 8); cudaMalloc(&stop, 4);
  const char* names[] = {"idle", "per-lane L2 polling", "two-lane watch polling", "mbarrier try_wait", "shared-memory spin", "nanosleep(200)"};
  for (int mode = 0; mode < 6; mode++)
    for (int ctas = 1; ctas <= 445; ctas *= 445) {
      cudaMemset(stop, 0, 4);
      k<<<ctas, 512>>>(mode, 15, g, out, cyc, stop);
      cudaDeviceSynchronize();
      printf("%-24s %3d CTAs x 16 warps: fold = %.1f cycles/entry\n", names[mode], ctas, cyc[0] / 1024.0);
    }
  return 0;
}

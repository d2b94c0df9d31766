// Microbenchmark (round 1): is the FP64 tensor-core path (DMMA) worth using for the Jacobian pair update on B200?
// Measures per-SM cycles per warp-instruction for mma.sync m8n8k4 / m16n8k8 f64 (independent and dependent accumulators),
// DADD alone, and DMMA interleaved with DADD (do they share the FP64 pipe?).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_probe dmma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 2048

__device__ __forceinline__ void mma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void mma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

template <int MODE>
__global__ void probe(double* out, long long* cyc) {
  const int lane = threadIdx.x & 31;
  double c[8][4];
  for (int q = 0; q < 8; ++q) for (int r = 0; r < 4; ++r) c[q][r] = lane * 1e-3 + q + r;
  double a[4] = {1.0 + lane * 1e-6, 0.5, 0.25, 0.125}, b[2] = {0.999, 1e-3};
  double s0 = lane, s1 = lane + 1, s2 = lane + 2, s3 = lane + 3, s4 = lane + 4, s5 = lane + 5, s6 = lane + 6, s7 = lane + 7;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < ITERS; ++i) {
    if (MODE == 0) {  // m8n8k4, 8 independent accumulators
#pragma unroll
      for (int q = 0; q < 8; ++q) mma884(c[q][0], c[q][1], a[0], b[0]);
    } else if (MODE == 1) {  // m8n8k4, dependent chain (4 per iter)
#pragma unroll
      for (int q = 0; q < 4; ++q) mma884(c[0][0], c[0][1], a[0], b[0]);
    } else if (MODE == 2) {  // m16n8k8, 8 independent accumulators
#pragma unroll
      for (int q = 0; q < 8; ++q) mma1688(c[q], a, b);
    } else if (MODE == 3) {  // m16n8k8 dependent chain (4 per iter)
#pragma unroll
      for (int q = 0; q < 4; ++q) mma1688(c[0], a, b);
    } else if (MODE == 4) {  // 8 DADD only
      s0 = __dadd_rn(s0, 1e-9); s1 = __dadd_rn(s1, 1e-9); s2 = __dadd_rn(s2, 1e-9); s3 = __dadd_rn(s3, 1e-9);
      s4 = __dadd_rn(s4, 1e-9); s5 = __dadd_rn(s5, 1e-9); s6 = __dadd_rn(s6, 1e-9); s7 = __dadd_rn(s7, 1e-9);
    } else if (MODE == 5) {  // 2 x m8n8k4 + 16 DADD per iter: the instruction mix of a DMMA pair update
      mma884(c[0][0], c[0][1], a[0], b[0]); mma884(c[1][0], c[1][1], a[1], b[1]);
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        s0 = __dadd_rn(s0, 1e-9); s1 = __dadd_rn(s1, 1e-9); s2 = __dadd_rn(s2, 1e-9); s3 = __dadd_rn(s3, 1e-9);
        s4 = __dadd_rn(s4, 1e-9); s5 = __dadd_rn(s5, 1e-9); s6 = __dadd_rn(s6, 1e-9); s7 = __dadd_rn(s7, 1e-9);
      }
    } else if (MODE == 6) {  // 1 x m16n8k8 + 32 DADD per iter
      mma1688(c[0], a, b);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        s0 = __dadd_rn(s0, 1e-9); s1 = __dadd_rn(s1, 1e-9); s2 = __dadd_rn(s2, 1e-9); s3 = __dadd_rn(s3, 1e-9);
        s4 = __dadd_rn(s4, 1e-9); s5 = __dadd_rn(s5, 1e-9); s6 = __dadd_rn(s6, 1e-9); s7 = __dadd_rn(s7, 1e-9);
      }
    } else if (MODE == 7) {  // 16 DADD only (baseline for mode 5)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        s0 = __dadd_rn(s0, 1e-9); s1 = __dadd_rn(s1, 1e-9); s2 = __dadd_rn(s2, 1e-9); s3 = __dadd_rn(s3, 1e-9);
        s4 = __dadd_rn(s4, 1e-9); s5 = __dadd_rn(s5, 1e-9); s6 = __dadd_rn(s6, 1e-9); s7 = __dadd_rn(s7, 1e-9);
      }
    }
  }
  long long t1 = clock64();
  double acc = s0 + s1 + s2 + s3 + s4 + s5 + s6 + s7;
  for (int q = 0; q < 8; ++q) for (int r = 0; r < 4; ++r) acc += c[q][r];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE> void run(const char* name, int per_iter, int threads, double flop_per_instr) {
  double* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 148 * 8);
  probe<MODE><<<148, threads>>>(out, cyc);
  probe<MODE><<<148, threads>>>(out, cyc);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
  int warps = threads / 32;
  double cpi = avg / ((double)ITERS * per_iter * warps);
  printf("%-44s warps/SM=%2d  cycles/warp-instr(per SM)=%7.3f  per-warp latency view %7.2f cyc", name, warps, cpi, avg / ((double)ITERS * per_iter));
  if (flop_per_instr > 0) printf("   -> %.1f flop/clk/SM", flop_per_instr / cpi);
  printf("   [iter cycles per SM-warp %.1f]\n", avg / ((double)ITERS * warps));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int threads : {128, 256, 512, 1024}) {
    run<0>("DMMA m8n8k4 8 indep accumulators", 8, threads, 512);
    run<1>("DMMA m8n8k4 dependent chain", 4, threads, 512);
    run<2>("DMMA m16n8k8 8 indep accumulators", 8, threads, 2048);
    run<3>("DMMA m16n8k8 dependent chain", 4, threads, 2048);
    run<4>("DADD 8 indep chains", 8, threads, 32);
    run<7>("16 DADD", 16, threads, 32);
    run<5>("2 DMMA m8n8k4 + 16 DADD (per-iter instr=18)", 18, threads, 0);
    run<6>("1 DMMA m16n8k8 + 32 DADD (per-iter instr=33)", 33, threads, 0);
    printf("\n");
  }
  return 0;
}

// Microbenchmark (round 1): what bounds the Jacobian kernel -- FP64 pipe, shared-memory broadcast delivery or shuffles?
// Measures per-SM cycles per warp-instruction for LDS (distinct / uniform, 64/128-bit), SHFL and DFMA at several
// occupancies.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o smem_fp64_probe smem_fp64_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 2048
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int MODE>
__global__ void probe(double* out, long long* cyc, int stride_mode) {
  extern __shared__ double sm[];
  for (int q = threadIdx.x; q < 4096; q += blockDim.x) sm[q] = q * 1e-3;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = lane, a5 = lane + 1, a6 = lane + 2, a7 = lane + 3;
  unsigned addr;
  if (MODE == 0) addr = smem_u32(sm) + threadIdx.x % 32 * 8;        // LDS.64 distinct
  if (MODE == 1) addr = smem_u32(sm);                                // LDS.64 uniform
  if (MODE == 2) addr = smem_u32(sm);                                // LDS.128 uniform
  if (MODE == 3) addr = smem_u32(sm) + (lane & 1) * 256;             // LDS.128 two addresses (lane parity)
  if (MODE == 4) addr = smem_u32(sm) + lane * 16;                    // LDS.128 distinct
  if (MODE == 8) addr = smem_u32(sm) + (lane >> 4) * 256;            // LDS.128 two addresses (half warps)
  __syncthreads();
  long long t0 = clock64();
  if (MODE <= 4 || MODE == 8) {
#pragma unroll 1
    for (int i = 0; i < ITERS; ++i) {
      unsigned a = addr + (i & 15) * 512;
      if (MODE == 0 || MODE == 1) {
        double x0, x1, x2, x3;
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(x0) : "r"(a));
        asm volatile("ld.shared.f64 %0, [%1+1024];" : "=d"(x1) : "r"(a));
        asm volatile("ld.shared.f64 %0, [%1+2048];" : "=d"(x2) : "r"(a));
        asm volatile("ld.shared.f64 %0, [%1+3072];" : "=d"(x3) : "r"(a));
        a0 += x0; a1 += x1; a2 += x2; a3 += x3;
      } else {
        double x0, x1, x2, x3, x4, x5, x6, x7;
        asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(x0), "=d"(x1) : "r"(a));
        asm volatile("ld.shared.v2.f64 {%0,%1}, [%2+1024];" : "=d"(x2), "=d"(x3) : "r"(a));
        asm volatile("ld.shared.v2.f64 {%0,%1}, [%2+2048];" : "=d"(x4), "=d"(x5) : "r"(a));
        asm volatile("ld.shared.v2.f64 {%0,%1}, [%2+3072];" : "=d"(x6), "=d"(x7) : "r"(a));
        a0 += x0; a1 += x1; a2 += x2; a3 += x3; a4 += x4; a5 += x5; a6 += x6; a7 += x7;
      }
    }
  } else if (MODE == 5) {  // SHFL.32 x4 per iter
    int s0 = lane, s1 = lane + 1, s2 = lane + 2, s3 = lane + 3;
#pragma unroll 1
    for (int i = 0; i < ITERS; ++i) {
      s0 = __shfl_xor_sync(0xffffffffu, s0, 1); s1 = __shfl_xor_sync(0xffffffffu, s1, 1);
      s2 = __shfl_xor_sync(0xffffffffu, s2, 1); s3 = __shfl_xor_sync(0xffffffffu, s3, 1);
    }
    a0 = s0 + s1 + s2 + s3;
  } else if (MODE == 6) {  // DFMA 8 independent chains (8 per iter)
#pragma unroll 1
    for (int i = 0; i < ITERS; ++i) {
      a0 = fma(a0, 0.9999, 1e-9); a1 = fma(a1, 0.9999, 1e-9); a2 = fma(a2, 0.9999, 1e-9); a3 = fma(a3, 0.9999, 1e-9);
      a4 = fma(a4, 0.9999, 1e-9); a5 = fma(a5, 0.9999, 1e-9); a6 = fma(a6, 0.9999, 1e-9); a7 = fma(a7, 0.9999, 1e-9);
    }
  } else if (MODE == 7) {  // DFMA dependent chain (4 per iter)
#pragma unroll 1
    for (int i = 0; i < ITERS; ++i) {
      a0 = fma(a0, 0.9999, 1e-9); a0 = fma(a0, 0.9999, 1e-9); a0 = fma(a0, 0.9999, 1e-9); a0 = fma(a0, 0.9999, 1e-9);
    }
  } else if (MODE == 9) {  // DADD dependent-free x8 (Kahan adds are DADD, same pipe?)
#pragma unroll 1
    for (int i = 0; i < ITERS; ++i) {
      a0 = __dadd_rn(a0, 1e-9); a1 = __dadd_rn(a1, 1e-9); a2 = __dadd_rn(a2, 1e-9); a3 = __dadd_rn(a3, 1e-9);
      a4 = __dadd_rn(a4, 1e-9); a5 = __dadd_rn(a5, 1e-9); a6 = __dadd_rn(a6, 1e-9); a7 = __dadd_rn(a7, 1e-9);
    }
  }
  long long t1 = clock64();
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE> void run(const char* name, int per_iter, int threads) {
  double* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 148 * 8);
  probe<MODE><<<148, threads, 4096 * 8>>>(out, cyc, 0);
  probe<MODE><<<148, threads, 4096 * 8>>>(out, cyc, 0);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
  int warps = threads / 32;
  printf("%-34s warps/SM=%2d  cycles/warp-instr(per SM)=%.3f   (per-warp latency view: %.2f cyc/instr)\n", name, warps,
         avg / ((double)ITERS * per_iter * warps), avg / ((double)ITERS * per_iter));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int threads : {128, 256, 512, 1024}) {
    run<0>("LDS.64 distinct", 4, threads);
    run<1>("LDS.64 uniform (broadcast)", 4, threads);
    run<2>("LDS.128 uniform (broadcast)", 4, threads);
    run<3>("LDS.128 2 addr (lane parity)", 4, threads);
    run<8>("LDS.128 2 addr (half warps)", 4, threads);
    run<4>("LDS.128 distinct", 4, threads);
    run<5>("SHFL.32 xor", 4, threads);
    run<6>("DFMA 8 indep chains", 8, threads);
    run<7>("DFMA dependent chain", 4, threads);
    run<9>("DADD 8 indep chains", 8, threads);
    printf("\n");
  }
  return 0;
}

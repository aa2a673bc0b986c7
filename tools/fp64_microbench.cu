// FP64 issue / latency microbenchmark for one SM of a B200 (what bounds the lattice kernel's dependent chains).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/fp64_microbench tools/fp64_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void chain(double* out, long long* cycles, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u)
#pragma unroll
      for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// mixed: DFMA chain interleaved with integer / FP32-select instructions (like the kernel's FSEL / IMAD mix)
__global__ void mixed(double* out, long long* cycles, int iters, double a, double b) {
  double x = threadIdx.x * 1e-3;
  int k = threadIdx.x;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      x = fma(x, a, b);
      k = k * 3 + (int)__double2hiint(x);
      x = (k & 1) ? x : -x;
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = x + k;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int ILP>
void run(int warps, double* out, long long* cyc) {
  const int iters = 2000;
  chain<ILP><<<1, warps * 32>>>(out, cyc, iters, 0.999, 1e-3);
  cudaDeviceSynchronize();
  chain<ILP><<<1, warps * 32>>>(out, cyc, iters, 0.999, 1e-3);
  cudaDeviceSynchronize();
  long long c;
  cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  const double n = (double)iters * 16 * ILP;
  printf("warps=%2d ilp=%d  cycles/DFMA/warp=%6.2f  warp-DFMA/clk/SM=%5.3f\n", warps, ILP, c / n, n * warps / c);
}

int main() {
  double* out;
  long long* cyc;
  cudaMalloc(&out, 1 << 20);
  cudaMalloc(&cyc, 1 << 12);
  for (int w : {1, 4, 8, 16, 24, 32}) {
    run<1>(w, out, cyc);
    run<2>(w, out, cyc);
    run<4>(w, out, cyc);
  }
  for (int w : {1, 8, 24}) {
    mixed<<<1, w * 32>>>(out, cyc, 2000, 0.999, 1e-3);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("mixed warps=%2d cycles per (DFMA+IMAD+FSELx2) group = %6.2f\n", w, c / (2000.0 * 16));
  }
  return 0;
}

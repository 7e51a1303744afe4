// B200 fp64 pipe microbenchmark: latency (1 dependent chain) and throughput (ILP chains x warps per SM) of DFMA, F2F.F64.F32, MUFU.RCP64H.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_pipe fp64_pipe.cu ; run on one GPU.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITERS = 4096;
template <int ILP>
__global__ void dfma_kernel(double* out, long long* cyc, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS / 8; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
  }
  const long long t1 = clock64();
  double s = 0; for (int i = 0; i < ILP; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int ILP>
__global__ void f2f_kernel(double* out, long long* cyc, const float* in) {
  float v[ILP]; double acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { v[i] = in[threadIdx.x + 32 * i]; acc[i] = 0; }
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS / 8; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int i = 0; i < ILP; ++i) { const double d = (double)v[i]; v[i] = __double2float_rn(d) ; v[i] = __int_as_float(__float_as_int(v[i]) ^ 1); acc[i] = d; }
  }
  const long long t1 = clock64();
  double s = 0; for (int i = 0; i < ILP; ++i) s += acc[i] + v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int ILP>
__global__ void f2f_only_kernel(double* out, long long* cyc, const float* in) {
  // independent conversions consumed by cheap integer xor on the high word (keeps them alive without fp64 math)
  float v[ILP]; unsigned acc = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) v[i] = in[threadIdx.x + 32 * i];
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS / 8; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int i = 0; i < ILP; ++i) { const double d = (double)v[i]; const unsigned h = __double2hiint(d); acc ^= h; v[i] = __int_as_float(__float_as_int(v[i]) + (h & 1)); }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int ILP>
__global__ void rcp_kernel(double* out, long long* cyc, double a) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = 1.0 + threadIdx.x * 1e-3 + i;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS / 8; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int i = 0; i < ILP; ++i) { double y; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x[i])); x[i] = y; }
  }
  const long long t1 = clock64();
  double s = 0; for (int i = 0; i < ILP; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <typename K, typename... A>
void run(const char* name, K k, int ilp, int threads, A... args) {
  long long* cyc; cudaMalloc(&cyc, 8 * 148);
  k<<<148, threads>>>(args..., cyc);   // placeholder (unused)
}
int main() {
  double* out; long long* cyc; float* in;
  cudaMalloc(&out, 8 * 148 * 1024); cudaMalloc(&cyc, 8 * 148); cudaMalloc(&in, 4 * 1024 * 64); cudaMemset(in, 0x3f, 4 * 1024 * 64);
  long long h[148];
  auto report = [&](const char* name, int ilp, int threads) {
    cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, 8 * 148, cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
    const double per_inst_warp = c / ITERS;                        // cycles per loop step of one warp (ILP instr)
    const double lanes_per_clk_sm = (double)threads * ilp * ITERS / c;  // thread-instr per cycle per SM
    printf("%-10s ilp=%d warps/SM=%2d : %.2f cyc per %d-wide step  -> %.1f thread-inst/clk/SM\n", name, ilp, threads / 32, per_inst_warp, ilp, lanes_per_clk_sm);
  };
#define RUN(K, NAME, ILP, ...) for (int th : {32, 128, 256, 512, 1024}) { K<ILP><<<148, th>>>(out, cyc, __VA_ARGS__); report(NAME, ILP, th); }
  RUN(dfma_kernel, "DFMA", 1, 1.0000001, 1e-9)
  RUN(dfma_kernel, "DFMA", 2, 1.0000001, 1e-9)
  RUN(dfma_kernel, "DFMA", 4, 1.0000001, 1e-9)
  RUN(dfma_kernel, "DFMA", 8, 1.0000001, 1e-9)
  RUN(f2f_only_kernel, "F2F", 1, in)
  RUN(f2f_only_kernel, "F2F", 4, in)
  RUN(f2f_only_kernel, "F2F", 8, in)
  RUN(rcp_kernel, "RCP64H", 1, 1.0)
  RUN(rcp_kernel, "RCP64H", 4, 1.0)
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}

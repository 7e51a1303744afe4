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
// DFMA chains (ILP of them) with NF float->double conversions per step feeding one of the chains: do the XU conversions steal
// fp64-pipe issue slots?  MODE 0: no conversion, 1: F2F.F64.F32 (XU), 2: integer bit conversion (ALU/FMA pipes)
__device__ __forceinline__ double cvt_bits(float f) {   // exact for normal floats; 0 maps to 2^-127
  const unsigned b = __float_as_uint(f);
  const unsigned hi = (((b >> 3) & 0x0fffffffu) + 0x38000000u) | (b & 0x80000000u);
  return __hiloint2double((int)hi, (int)(b << 29));
}
template <int ILP, int NF, int MODE>
__global__ void mix_kernel(double* out, long long* cyc, const float* in, double a, double b) {
  double x[ILP]; float v[NF > 0 ? NF : 1];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
#pragma unroll
  for (int i = 0; i < NF; ++i) v[i] = in[threadIdx.x + 32 * i];
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS / 8; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
      if (MODE != 0) {
#pragma unroll
        for (int i = 0; i < NF; ++i) {
          const double d = MODE == 1 ? (double)v[i] : cvt_bits(v[i]);
          x[i % ILP] += d;                                        // consumed by the fp64 chains (one extra DADD per conversion)
          v[i] = __uint_as_float(__float_as_uint(v[i]) + 2u);   // next input differs (integer op, no fp pipe)
        }
      } else {
#pragma unroll
        for (int i = 0; i < NF; ++i) x[i % ILP] += 1.5;
      }
    }
  }
  const long long t1 = clock64();
  double s = 0; for (int i = 0; i < ILP; ++i) s += x[i];
  for (int i = 0; i < NF; ++i) s += v[i];
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
  // K3-like mix: 10 fp64 ops (8 DFMA + 2 DADD) per 2 conversions
  auto report_mix = [&](const char* name, int fp64_ops, int threads) {
    cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, 8 * 148, cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
    printf("%-22s warps/SM=%2d : %.2f cyc per step -> %.1f fp64 thread-inst/clk/SM\n", name, threads / 32, c / ITERS, (double)threads * fp64_ops * ITERS / c);
  };
  for (int th : {128, 256, 512, 640, 1024}) { mix_kernel<8, 2, 0><<<148, th>>>(out, cyc, in, 1.0000001, 1e-9); report_mix("8 DFMA+2 DADD", 10, th); }
  for (int th : {128, 256, 512, 640, 1024}) { mix_kernel<8, 2, 1><<<148, th>>>(out, cyc, in, 1.0000001, 1e-9); report_mix("8 DFMA+2 DADD+2 F2F", 10, th); }
  for (int th : {128, 256, 512, 640, 1024}) { mix_kernel<8, 2, 2><<<148, th>>>(out, cyc, in, 1.0000001, 1e-9); report_mix("8 DFMA+2 DADD+2 icvt", 10, th); }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}

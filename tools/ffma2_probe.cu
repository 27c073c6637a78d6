// Probe of Blackwell's packed fp32 instructions (FFMA2 / FMUL2 / FADD2, PTX fma.rn.f32x2) on sm_100a:
// pipe throughput, dependent latency, and whether a packed instruction frees issue slots for ALU work beside it.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ffma2_probe tools/ffma2_probe.cu && ./ffma2_probe
#include <cstdio>
#include <cuda_runtime.h>

#define CHAINS 8
// mode 0: scalar FFMA, 2*CHAINS chains.  mode 1: FFMA2, CHAINS packed chains (same flop count).
// mode 2 / 3: the same with one independent integer op (LOP3/IADD3 on the ALU pipe) per scalar-FMA-equivalent pair.
// mode 4 / 5: ONE dependent chain (latency): scalar / packed.
template <int MODE> __global__ void __launch_bounds__(512) probe(float* out, int iters, float a, float b) {
  float2 x[CHAINS];
  unsigned v[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) { x[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f - i); v[i] = threadIdx.x * 77u + i; }
  const float2 A = make_float2(a, a), B = make_float2(b, b);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      if (MODE == 4) { x[0].x = fmaf(x[0].x, a, b); }
      else if (MODE == 5) { x[0] = __ffma2_rn(x[0], A, B); }
      else {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) {
          if (MODE == 0 || MODE == 2) { x[i].x = fmaf(x[i].x, a, b); x[i].y = fmaf(x[i].y, a, b); }
          else x[i] = __ffma2_rn(x[i], A, B);
          if (MODE == 2 || MODE == 3) { v[i] = (v[i] ^ (v[i] >> 3)) + 0x9e3779b9u; }
        }
      }
    }
  }
  float s = 0.f;
  unsigned t = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) { s += x[i].x + x[i].y; t ^= v[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)t;
}

template <int MODE> void run(const char* name, float* out, int sms) {
  const int iters = 2000, blocks = sms * 4, threads = 512;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  probe<MODE><<<blocks, threads>>>(out, 10, 0.999f, 1e-3f);
  cudaEventRecord(e0);
  probe<MODE><<<blocks, threads>>>(out, iters, 0.999f, 1e-3f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double fma_per_thread = (MODE >= 4 ? 1.0 : 2.0 * CHAINS) * 16.0 * iters * (MODE == 5 ? 2.0 : 1.0);
  const double tflops = 2.0 * fma_per_thread * blocks * threads / (ms * 1e-3) / 1e12;
  // cycles per loop trip per warp-scheduler: 4 blocks x 16 warps / 4 schedulers = 16 warps per scheduler
  printf("%-44s %8.3f ms  %7.2f TFLOP/s\n", name, ms, tflops);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  float* out;
  cudaMalloc(&out, (size_t)p.multiProcessorCount * 4 * 512 * sizeof(float));
  printf("%s, %d SMs, %d MHz\n", p.name, p.multiProcessorCount, p.clockRate / 1000);
  run<0>("scalar FFMA, 16 chains", out, p.multiProcessorCount);
  run<1>("FFMA2, 8 packed chains (same flops)", out, p.multiProcessorCount);
  run<2>("scalar FFMA + 1 ALU pair per 2 FMAs", out, p.multiProcessorCount);
  run<3>("FFMA2 + 1 ALU pair per packed FMA", out, p.multiProcessorCount);
  run<4>("one dependent scalar chain (latency)", out, p.multiProcessorCount);
  run<5>("one dependent packed chain (latency)", out, p.multiProcessorCount);
  return 0;
}

// ubench.cu -- FP32 issue/register-bank microbenchmarks for sm_100a (developer tool).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench tools/ubench.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096

// A: x = fma(x, a, b): two operands shared by every instruction (reuse cache friendly)
__global__ void __launch_bounds__(256) kA(float a, float b, float* out) {
  float x[16];
  for (int k = 0; k < 16; ++k) x[k] = threadIdx.x * 1e-3f + k;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int k = 0; k < 16; ++k) x[k] = fmaf(x[k], a, b);
  }
  float s = 0;
  for (int k = 0; k < 16; ++k) s += x[k];
  if (s == 123.456f) out[0] = s;
}
// B: three distinct registers per FFMA, no sharing between consecutive instructions
__global__ void __launch_bounds__(256) kB(const float* in, float* out) {
  float x[16], y[16], z[16];
  for (int k = 0; k < 16; ++k) { x[k] = in[k] + threadIdx.x; y[k] = in[16 + k] + threadIdx.x * 1e-6f; z[k] = in[32 + k] + threadIdx.x * 1e-7f; }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int k = 0; k < 16; ++k) x[k] = fmaf(x[k], y[k], z[k]);
  }
  float s = 0;
  for (int k = 0; k < 16; ++k) s += x[k];
  if (s == 123.456f) out[0] = s;
}
// C: one operand shared across consecutive instructions (source-value pattern of the matvec)
__global__ void __launch_bounds__(256) kC(const float* in, float* out) {
  float x[16], y[16];
  for (int k = 0; k < 16; ++k) { x[k] = in[k] + threadIdx.x; y[k] = in[16 + k] + threadIdx.x * 1e-6f; }
  float a = in[40] + threadIdx.x * 1e-7f;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int k = 0; k < 16; ++k) x[k] = fmaf(y[k], a, x[k]);
  }
  float s = 0;
  for (int k = 0; k < 16; ++k) s += x[k];
  if (s == 123.456f) out[0] = s;
}
// D: packed fp32x2 FMA, three distinct 64-bit operands
__global__ void __launch_bounds__(256) kD(const float* in, float* out) {
  unsigned long long x[8], y[8], z[8];
  for (int k = 0; k < 8; ++k) {
    float t = threadIdx.x * 1e-6f;
    float2 a = make_float2(in[k] + threadIdx.x, in[k + 1]), b = make_float2(in[16 + k] + t, in[17 + k] + t), c = make_float2(in[32 + k] + t, in[33 + k] - t);
    x[k] = *reinterpret_cast<unsigned long long*>(&a);
    y[k] = *reinterpret_cast<unsigned long long*>(&b);
    z[k] = *reinterpret_cast<unsigned long long*>(&c);
  }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int k = 0; k < 8; ++k) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[k]) : "l"(y[k]), "l"(z[k]));
  }
  unsigned long long s = 0;
  for (int k = 0; k < 8; ++k) s ^= x[k];
  if (s == 123456ull) out[0] = (float)s;
}
// E: packed fp32x2 FMA with two shared operands (peak of the packed form)
__global__ void __launch_bounds__(256) kE(const float* in, float* out) {
  unsigned long long x[8], a, b;
  for (int k = 0; k < 8; ++k) {
    float2 v = make_float2(in[k] + threadIdx.x, in[k + 1]);
    x[k] = *reinterpret_cast<unsigned long long*>(&v);
  }
  float2 va = make_float2(in[40] + threadIdx.x * 1e-6f, in[41]), vb = make_float2(in[42] - threadIdx.x * 1e-6f, in[43]);
  a = *reinterpret_cast<unsigned long long*>(&va);
  b = *reinterpret_cast<unsigned long long*>(&vb);
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int k = 0; k < 8; ++k) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[k]) : "l"(a), "l"(b));
  }
  unsigned long long s = 0;
  for (int k = 0; k < 8; ++k) s ^= x[k];
  if (s == 123456ull) out[0] = (float)s;
}
// F: FFMA with a constant-bank operand + 2 distinct registers
__global__ void __launch_bounds__(256) kF(const float* in, float* out, float c0) {
  float x[16], y[16];
  for (int k = 0; k < 16; ++k) { x[k] = in[k] + threadIdx.x; y[k] = in[16 + k] + threadIdx.x * 1e-6f; }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int k = 0; k < 16; ++k) x[k] = fmaf(x[k], c0, y[k]);
  }
  float s = 0;
  for (int k = 0; k < 16; ++k) s += x[k];
  if (s == 123.456f) out[0] = s;
}
// G: A plus one MUFU.RSQ per 32 FFMA (the matvec's ratio)
__global__ void __launch_bounds__(256) kG(float a, float b, float* out) {
  float x[16], m = threadIdx.x + 1.0f;
  for (int k = 0; k < 16; ++k) x[k] = threadIdx.x * 1e-3f + k;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int k = 0; k < 16; ++k) x[k] = fmaf(x[k], a, b);
      if ((u & 1) == 0) { asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(m)); m += 1.5f; }
    }
  }
  float s = m;
  for (int k = 0; k < 16; ++k) s += x[k];
  if (s == 123.456f) out[0] = s;
}
// H: FMUL with 2 distinct registers
__global__ void __launch_bounds__(256) kH(const float* in, float* out) {
  float x[16], y[16];
  for (int k = 0; k < 16; ++k) { x[k] = in[k] + threadIdx.x; y[k] = in[16 + k] + threadIdx.x * 1e-6f; }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int k = 0; k < 16; ++k) x[k] = x[k] * y[k];
  }
  float s = 0;
  for (int k = 0; k < 16; ++k) s += x[k];
  if (s == 123.456f) out[0] = s;
}

template <typename F>
static void run(const char* name, F launch, double inst_per_thread_iter) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  launch();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  launch();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  int dev, sms, khz;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  const double warps = (double)sms * 8 * 256 / 32;
  const double winst = warps * ITERS * inst_per_thread_iter;
  const double per_clk = winst / (ms * 1e-3 * khz * 1e3) / (sms * 4);
  printf("%-44s %8.3f ms  %6.3f warp-inst/clk/SMSP (at %d MHz nominal)  err=%s\n", name, ms, per_clk, khz / 1000,
         cudaGetErrorString(cudaGetLastError()));
}

int main() {
  float *in, *out;
  cudaMalloc(&in, 4096);
  cudaMalloc(&out, 4096);
  float h[64];
  for (int i = 0; i < 64; ++i) h[i] = 0.5f + 0.01f * i;
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  int dev, sms;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int g = sms * 8;
  run("A ffma x=fma(x,a,b) shared a,b", [&] { kA<<<g, 256>>>(0.999f, 1e-3f, out); }, 64);
  run("B ffma 3 distinct regs", [&] { kB<<<g, 256>>>(in, out); }, 64);
  run("C ffma x=fma(y,a,x) one shared", [&] { kC<<<g, 256>>>(in, out); }, 64);
  run("D ffma2 (f32x2) 3 distinct pairs [x2 flops]", [&] { kD<<<g, 256>>>(in, out); }, 32);
  run("E ffma2 (f32x2) shared a,b       [x2 flops]", [&] { kE<<<g, 256>>>(in, out); }, 32);
  run("F ffma reg,const,reg", [&] { kF<<<g, 256>>>(in, out, 0.999f); }, 64);
  run("G A + 1 MUFU per 32 FFMA", [&] { kG<<<g, 256>>>(0.999f, 1e-3f, out); }, 66);
  run("H fmul 2 distinct regs", [&] { kH<<<g, 256>>>(in, out); }, 64);
  return 0;
}

// matvec_lab.cu -- developer lab: inner-loop variants of the RPY pair kernel on synthetic
// tiles (no scheduling, no fix-up), reporting issue cycles per pair.  Not part of the product.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o matvec_lab tools/matvec_lab.cu
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
#include "../rigid_body_light_b200/csrc/rbl_pair.cuh"

using namespace rbl;
constexpr int NSRC = 2048;
// v2: fp32 "register-port friendly" formulation: wall polynomials in the monomial basis
// {V, EV, V^2, EV^2} (V = a^2 W) so every FMA has a literal coefficient and at most two fresh
// register operands; p = z_i z_j W replaced by (1 - r^2 W)/4 (R^2 - r^2 = 4 z_i z_j).
template <bool NEAR>
__device__ __forceinline__ void pair_v2(const PairConsts<float>& C, float xi, float yi, float zi, float xj, float yj,
                                        float zj, float fx, float fy, float fz, float z2j, float nzz4j, float& ux,
                                        float& uy, float& uz) {
  const float dx = xi - xj, dy = yi - yj, dz = zi - zj;
  const float q = fmaf(dy, dy, fmaf(dx, dx, C.tiny));
  const float r2 = fmaf(dz, dz, q);
  const float s = fmaf(dy, fy, dx * fx);
  const float df = fmaf(dz, fz, s);
  const float invr = rsqrt_fast(r2);
  const float i2 = invr * invr;
  const float i3 = invr * i2;
  float c1 = fmaf(i3, C.c23a2, invr);
  float c2 = fmaf(i2 * i3, C.m2a2, i3);
  if (NEAR) {
    const float r = r2 * invr;
    const float c1n = fmaf(r, C.n1, C.n0);
    const float c2n = invr * C.n2;
    const bool nr = r2 < C.four_a2;
    c1 = nr ? c1n : c1;
    c2 = nr ? c2n : c2;
  }
  const float t = c2 * df;
  const float Z = zi + zj;
  const float Z2 = Z * Z;
  const float R2 = q + Z2;
  const float w = rsqrt_fast(R2);
  const float W = w * w;
  const float g = fmaf(Z, fz, s);
  const float E = Z2 * W;
  const float rW = r2 * W;
  const float V = W * C.a * C.a;
  const float EV = E * V;
  const float V2 = V * V;
  const float EV2 = E * V2;
  const float a1n = fmaf(EV2, -10.f / 3.f, fmaf(V2, 2.f / 3.f, fmaf(EV, 2.f, fmaf(V, -2.f / 3.f, fmaf(rW, 0.5f, -1.5f)))));
  const float a2n = fmaf(EV2, 70.f / 3.f, fmaf(V2, -10.f / 3.f, fmaf(EV, -10.f, fmaf(V, 2.f, fmaf(rW, -1.5f, 0.5f)))));
  const float in3 = fmaf(EV2, 140.f / 3.f, fmaf(V2, -40.f / 3.f, fmaf(EV, -20.f, V * 4.f)));
  const float ZW = Z * W;
  const float b = fmaf(zi * ZW, -6.f, 1.f);
  const float a3 = fmaf(-Z, in3, z2j * b);
  const float a4 = fmaf(Z * V2, -20.f / 3.f, z2j);
  const float in5 = fmaf(EV, -20.f, fmaf(V, 8.f / 3.f, E * 4.f));
  const float a5n = fmaf(in5, -C.a * C.a, nzz4j);
  const float wW = w * W;
  const float cF = fmaf(w, a1n, c1);
  const float A = wW * fmaf(a3, fz, a2n * g);
  const float Bz = wW * fmaf(a5n, fz, a4 * g);
  const float txy = t + A;
  ux = fmaf(cF, fx, ux); ux = fmaf(txy, dx, ux);
  uy = fmaf(cF, fy, uy); uy = fmaf(txy, dy, uy);
  uz = fmaf(cF, fz, uz); uz = fmaf(t, dz, uz); uz = fmaf(A, Z, uz); uz += Bz;
}

__constant__ float4 c_src[2 * NSRC];  // 64 KB: x y z fx | fy fz 2z 4z^2

template <int T, int NT, bool NEAR, int UNROLL, int VER = 0>
__global__ void __launch_bounds__(NT) k_smem(const float4* __restrict__ src, const float* __restrict__ tgt,
                                             float* __restrict__ out, int reps, PairConsts<float> C) {
  __shared__ float4 sb[2 * 256];
  float xi[T], yi[T], zi[T], ux[T], uy[T], uz[T];
  for (int t = 0; t < T; ++t) {
    int i = (blockIdx.x * NT + threadIdx.x) * T + t;
    xi[t] = tgt[3 * i]; yi[t] = tgt[3 * i + 1]; zi[t] = tgt[3 * i + 2];
    ux[t] = uy[t] = uz[t] = 0;
  }
  for (int r = 0; r < reps; ++r)
    for (int tile = 0; tile < NSRC / 256; ++tile) {
      __syncthreads();
      for (int k = threadIdx.x; k < 512; k += NT) sb[k] = src[tile * 512 + k];
      __syncthreads();
#pragma unroll UNROLL
      for (int j = 0; j < 256; ++j) {
        const float4 p = sb[2 * j], q = sb[2 * j + 1];
#pragma unroll
        for (int t = 0; t < T; ++t)
          if (VER == 0) pair<float, true, NEAR>(C, xi[t], yi[t], zi[t], p.x, p.y, p.z, p.w, q.x, q.y, q.z, q.w, ux[t], uy[t], uz[t]);
          else pair_v2<NEAR>(C, xi[t], yi[t], zi[t], p.x, p.y, p.z, p.w, q.x, q.y, q.z, -q.w, ux[t], uy[t], uz[t]);
      }
    }
  for (int t = 0; t < T; ++t) {
    int i = (blockIdx.x * NT + threadIdx.x) * T + t;
    out[3 * i] = ux[t]; out[3 * i + 1] = uy[t]; out[3 * i + 2] = uz[t];
  }
}

template <int T, int NT, bool NEAR, int UNROLL, int VER = 0>
__global__ void __launch_bounds__(NT) k_const(const float* __restrict__ tgt, float* __restrict__ out, int reps,
                                              PairConsts<float> C) {
  float xi[T], yi[T], zi[T], ux[T], uy[T], uz[T];
  for (int t = 0; t < T; ++t) {
    int i = (blockIdx.x * NT + threadIdx.x) * T + t;
    xi[t] = tgt[3 * i]; yi[t] = tgt[3 * i + 1]; zi[t] = tgt[3 * i + 2];
    ux[t] = uy[t] = uz[t] = 0;
  }
  for (int r = 0; r < reps; ++r) {
#pragma unroll UNROLL
    for (int j = 0; j < NSRC; ++j) {
      const float4 p = c_src[2 * j], q = c_src[2 * j + 1];
#pragma unroll
      for (int t = 0; t < T; ++t)
        if (VER == 0) pair<float, true, NEAR>(C, xi[t], yi[t], zi[t], p.x, p.y, p.z, p.w, q.x, q.y, q.z, q.w, ux[t], uy[t], uz[t]);
        else pair_v2<NEAR>(C, xi[t], yi[t], zi[t], p.x, p.y, p.z, p.w, q.x, q.y, q.z, -q.w, ux[t], uy[t], uz[t]);
    }
  }
  for (int t = 0; t < T; ++t) {
    int i = (blockIdx.x * NT + threadIdx.x) * T + t;
    out[3 * i] = ux[t]; out[3 * i + 1] = uy[t]; out[3 * i + 2] = uz[t];
  }
}


template <int T, int NT, bool NEAR, int UNROLL>
__global__ void __launch_bounds__(NT) k_sym(const float4* __restrict__ src, const float4* __restrict__ tgtrec,
                                            float* __restrict__ out, float* __restrict__ raw, int reps,
                                            PairConsts<float> C) {
  __shared__ float4 sb[2 * 256];
  float xi[T], yi[T], zi[T], fxi[T], fyi[T], fzi[T], z2i[T], nz4i[T], ux[T], uy[T], uz[T];
  const int lane = threadIdx.x & 31;
  for (int t = 0; t < T; ++t) {
    int i = (blockIdx.x * NT + threadIdx.x) * T + t;
    float4 p = tgtrec[2 * (i % NSRC)], q = tgtrec[2 * (i % NSRC) + 1];
    xi[t] = p.x + 20.f + 0.001f * (i % 97); yi[t] = p.y; zi[t] = p.z; fxi[t] = p.w; fyi[t] = q.x; fzi[t] = q.y; z2i[t] = q.z; nz4i[t] = -q.w;
    ux[t] = uy[t] = uz[t] = 0;
  }
  for (int r = 0; r < reps; ++r)
    for (int tile = 0; tile < NSRC / 256; ++tile) {
      __syncthreads();
      for (int k = threadIdx.x; k < 512; k += NT) sb[k] = src[tile * 512 + k];
      __syncthreads();
      for (int j0 = 0; j0 < 256; j0 += 32) {
        float rx = 0, ry = 0, rz = 0;
#pragma unroll UNROLL
        for (int jj = 0; jj < 32; ++jj) {
          const float4 p = sb[2 * (j0 + jj)], q = sb[2 * (j0 + jj) + 1];
          float ax = 0, ay = 0, az = 0;
#pragma unroll
          for (int t = 0; t < T; ++t)
            pair_sym<float, true, NEAR>(C, xi[t], yi[t], zi[t], fxi[t], fyi[t], fzi[t], z2i[t], nz4i[t], p.x, p.y, p.z, p.w,
                                        q.x, q.y, q.z, -q.w, ux[t], uy[t], uz[t], ax, ay, az);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            ax += __shfl_xor_sync(0xffffffffu, ax, o);
            ay += __shfl_xor_sync(0xffffffffu, ay, o);
            az += __shfl_xor_sync(0xffffffffu, az, o);
          }
          if (lane == jj) { rx = ax; ry = ay; rz = az; }
        }
        const int jg = tile * 256 + j0 + lane;
        atomicAdd(raw + 3 * jg, rx);
        atomicAdd(raw + 3 * jg + 1, ry);
        atomicAdd(raw + 3 * jg + 2, rz);
      }
    }
  for (int t = 0; t < T; ++t) {
    int i = (blockIdx.x * NT + threadIdx.x) * T + t;
    out[3 * i] = ux[t]; out[3 * i + 1] = uy[t]; out[3 * i + 2] = uz[t];
  }
}

static std::vector<float> g_ref;
template <typename F>
static void run(const char* name, int T, int NT, int ctas_per_sm, int reps, float* out, F launch) {
  int dev, sms, khz;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  const int grid = sms * ctas_per_sm;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch(grid, 1);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  launch(grid, reps);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double pairs = (double)grid * NT * T * NSRC * reps;
  const double cyc = ms * 1e-3 * khz * 1e3 * sms * 4 / (pairs / 32);
  std::vector<float> h(64);
  cudaMemcpy(h.data(), out, 64 * sizeof(float), cudaMemcpyDeviceToHost);
  double d = 0;
  if (g_ref.empty()) g_ref = h;
  for (int i = 0; i < 64; ++i) d = fmax(d, fabs(h[i] - g_ref[i]) / (fabs(g_ref[i]) + 1e-30));
  printf("%-34s T=%d NT=%d occ=%d  %8.3f ms  %6.2f cyc/pair/SMSP  %7.1f Gpairs/s  frac(127 flop)=%.3f  relchk=%.1e  %s\n", name, T, NT,
         ctas_per_sm, ms, cyc, pairs / ms / 1e6, 127.0 / 2 / cyc, d, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  const double a = 0.131, eta = 1.0;
  PairConsts<float> C = make_pair_consts<float>(a, eta);
  // sources: a slab of blobs far (>= 2a) from the targets so NEAR=false is legal
  std::vector<float4> src(2 * NSRC);
  for (int j = 0; j < NSRC; ++j) {
    float x = 0.3f * (j % 45), y = 0.3f * ((j / 45) % 46), z = 1.0f + 0.01f * (j % 7);
    float fx = sinf(j * 0.37f), fy = cosf(j * 0.11f), fz = sinf(j * 0.73f + 1);
    src[2 * j] = make_float4(x, y, z, fx);
    src[2 * j + 1] = make_float4(fy, fz, 2 * z, 4 * z * z);
  }
  const int max_tgt = 148 * 8 * 256 * 8;
  std::vector<float> tgt(3 * (size_t)max_tgt);
  for (int i = 0; i < max_tgt; ++i) { tgt[3 * i] = 20.f + 0.01f * (i % 1000); tgt[3 * i + 1] = 0.02f * (i % 777); tgt[3 * i + 2] = 1.5f + 0.001f * (i % 333); }
  float4* dsrc; float *dtgt, *dout;
  cudaMalloc(&dsrc, src.size() * sizeof(float4));
  cudaMalloc(&dtgt, tgt.size() * sizeof(float));
  cudaMalloc(&dout, tgt.size() * sizeof(float));
  cudaMemcpy(dsrc, src.data(), src.size() * sizeof(float4), cudaMemcpyHostToDevice);
  cudaMemcpy(dtgt, tgt.data(), tgt.size() * sizeof(float), cudaMemcpyHostToDevice);
  cudaMemcpyToSymbol(c_src, src.data(), src.size() * sizeof(float4));
  const int reps = 24;
#define SMEM(T, NT, OCC, U) run("smem far unroll" #U, T, NT, OCC, reps, dout, [&](int g, int r) { k_smem<T, NT, false, U><<<g, NT>>>(dsrc, dtgt, dout, r, C); })
#define CONST(T, NT, OCC, U) run("const(UR) far unroll" #U, T, NT, OCC, reps, dout, [&](int g, int r) { k_const<T, NT, false, U><<<g, NT>>>(dtgt, dout, r, C); })
  SMEM(4, 256, 2, 4);
  SMEM(4, 256, 3, 4);
  SMEM(2, 256, 3, 4);
  SMEM(8, 128, 3, 2);
  SMEM(4, 256, 2, 2);
  SMEM(4, 256, 2, 8);
  SMEM(1, 256, 4, 8);
  CONST(4, 256, 2, 4);
  CONST(4, 256, 3, 4);
  CONST(2, 256, 3, 4);
  CONST(2, 256, 4, 8);
  CONST(8, 128, 3, 2);
  CONST(4, 256, 2, 2);
  CONST(4, 256, 2, 8);
  CONST(1, 256, 4, 8);
  CONST(1, 256, 6, 16);
  run("v2 smem far unroll4", 4, 256, 2, reps, dout, [&](int g, int r) { k_smem<4, 256, false, 4, 1><<<g, 256>>>(dsrc, dtgt, dout, r, C); });
  run("v2 smem far unroll4", 4, 256, 3, reps, dout, [&](int g, int r) { k_smem<4, 256, false, 4, 1><<<g, 256>>>(dsrc, dtgt, dout, r, C); });
  run("v2 smem far unroll2", 4, 256, 3, reps, dout, [&](int g, int r) { k_smem<4, 256, false, 2, 1><<<g, 256>>>(dsrc, dtgt, dout, r, C); });
  run("v2 smem far unroll4 T2", 2, 256, 3, reps, dout, [&](int g, int r) { k_smem<2, 256, false, 4, 1><<<g, 256>>>(dsrc, dtgt, dout, r, C); });
  run("v2 smem far unroll8 T1", 1, 256, 4, reps, dout, [&](int g, int r) { k_smem<1, 256, false, 8, 1><<<g, 256>>>(dsrc, dtgt, dout, r, C); });
  run("v2 const far unroll4", 4, 256, 3, reps, dout, [&](int g, int r) { k_const<4, 256, false, 4, 1><<<g, 256>>>(dtgt, dout, r, C); });
  run("v2 const far unroll4 T2", 2, 256, 3, reps, dout, [&](int g, int r) { k_const<2, 256, false, 4, 1><<<g, 256>>>(dtgt, dout, r, C); });
  run("v2 smem NEAR unroll4", 4, 256, 2, reps, dout, [&](int g, int r) { k_smem<4, 256, true, 4, 1><<<g, 256>>>(dsrc, dtgt, dout, r, C); });
  float* draw; cudaMalloc(&draw, 3 * NSRC * sizeof(float)); cudaMemset(draw, 0, 3 * NSRC * sizeof(float));
#define SYM(T, NT, OCC, U) run("SYM far unroll" #U " (cyc per UNORDERED pair)", T, NT, OCC, reps, dout, [&](int g, int r) { k_sym<T, NT, false, U><<<g, NT>>>(dsrc, dsrc, dout, draw, r, C); })
  SYM(4, 256, 2, 2);
  SYM(4, 256, 1, 2);
  SYM(2, 256, 2, 2);
  SYM(2, 256, 3, 2);
  SYM(2, 256, 3, 4);
  SYM(2, 128, 4, 2);
  SYM(1, 256, 3, 4);
  SYM(4, 128, 3, 2);
  SYM(4, 128, 4, 1);
  run("smem NEAR unroll4", 4, 256, 2, reps, dout, [&](int g, int r) { k_smem<4, 256, true, 4><<<g, 256>>>(dsrc, dtgt, dout, r, C); });
  run("const NEAR unroll4", 4, 256, 2, reps, dout, [&](int g, int r) { k_const<4, 256, true, 4><<<g, 256>>>(dtgt, dout, r, C); });
  return 0;
}

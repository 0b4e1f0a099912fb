// util.cu -- measurement helpers exported through the C-ABI (no part of the evaluation path).
//
// pda_measure_fp64_peak: the FP64 denominator of the roofline.  MEASURED_PEAKS.json (driver-written) holds HBM and
// bf16-tensor peaks only; the WENO/Rusanov velocity kernels are bound by the FP64 pipe, so bench.py measures a pure
// DFMA loop on the same device, in the same process, right before the timed region.
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/pda_b200.h"
#include "kernels_reforder.hpp"

namespace {

// the same DFMA loop, bracketed per CTA by clock64 (SM cycles) and globaltimer (ns): rec[2*b] = cycles, rec[2*b+1] = ns
__global__ void __launch_bounds__(256) k_dfma_peak_clocked(double* out, int iters, double a, double b, long long* rec) {
  double x0 = threadIdx.x * 1e-9, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  __syncthreads();
  long long c0 = 0, t0 = 0;
  if (threadIdx.x == 0) { c0 = clock64(); asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0)); }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (s == 123.456) out[0] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long c1 = clock64(), t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    rec[2 * blockIdx.x] = c1 - c0;
    rec[2 * blockIdx.x + 1] = t1 - t0;
  }
}

__global__ void __launch_bounds__(256) k_dfma_peak16(double* out, int iters, double a, double b) {
  double x[16], m[16], c[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { x[i] = threadIdx.x * 1e-9 + i; m[i] = a - 1e-9 * i; c[i] = b * (i + 1); }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = fma(x[i], m[i], c[i]);
    }
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  if (s == 123.456) out[0] = s;
}

__global__ void __launch_bounds__(256) k_dfma_peak(double* out, int iters, double a, double b) {
  // 8 independent dependency chains per thread: enough ILP to cover the DFMA latency at 8 warps/scheduler
  double x0 = threadIdx.x * 1e-9, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (s == 123.456) out[0] = s;   // never true; keeps the loop alive
}

}  // namespace

extern "C" pda_status pda_measure_fp64_peak(int device, double* tflops, double* sm_mhz_hint) {
  if (!tflops) return PDA_ERR_INVALID;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) { cudaGetLastError(); return PDA_ERR_NO_DEVICE; }
  if (cudaSetDevice(device) != cudaSuccess) return PDA_ERR_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return PDA_ERR_CUDA;
  double* d = nullptr;
  if (cudaMalloc(&d, 8) != cudaSuccess) return PDA_ERR_CUDA;
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0.0;
  for (int rep = 0; rep < 6; ++rep) {
    cudaEventRecord(e0);
    k_dfma_peak<<<blocks, threads>>>(d, iters, 0.999999, 1e-9);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(d); return PDA_ERR_CUDA; }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 64.0 * iters * (double)blocks * threads;   // 64 FMA per iteration per thread
    const double tf = flops / (ms * 1e-3) * 1e-12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(d);
  *tflops = best;
  if (sm_mhz_hint) *sm_mhz_hint = prop.clockRate * 1e-3;
  return PDA_OK;
}

extern "C" pda_status pda_measure_fp64_peak_ex(int device, double* tflops, double* sm_mhz_under_probe, double* dfma_per_sm_clk) {
  if (!tflops) return PDA_ERR_INVALID;
  double hint = 0.0;
  const pda_status st = pda_measure_fp64_peak(device, tflops, &hint);
  if (st != PDA_OK) return st;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return PDA_ERR_CUDA;
  // one wave of resident CTAs (8 per SM) so that every CTA runs from start to end concurrently with all others
  const int perSm = 8, blocks = prop.multiProcessorCount * perSm, threads = 256, iters = 8192;
  double* d = nullptr;
  long long* rec = nullptr;
  if (cudaMalloc(&d, 8) != cudaSuccess || cudaMalloc(&rec, sizeof(long long) * 2 * blocks) != cudaSuccess) return PDA_ERR_CUDA;
  double mhz = 0.0, rate = 0.0;
  for (int rep = 0; rep < 3; ++rep) {
    k_dfma_peak_clocked<<<blocks, threads>>>(d, iters, 0.999999, 1e-9, rec);
    if (cudaDeviceSynchronize() != cudaSuccess) { cudaFree(d); cudaFree(rec); return PDA_ERR_CUDA; }
    long long* h = new long long[2 * blocks];
    cudaMemcpy(h, rec, sizeof(long long) * 2 * blocks, cudaMemcpyDeviceToHost);
    double cyc = 0.0, ns = 0.0;
    for (int b = 0; b < blocks; ++b) { cyc += (double)h[2 * b]; ns += (double)h[2 * b + 1]; }
    delete[] h;
    cyc /= blocks; ns /= blocks;
    mhz = cyc / ns * 1e3;   // SM cycles per microsecond while the DFMA loop runs
  }
  cudaFree(d); cudaFree(rec);
  // second probe shape: 16 independent chains with DISTINCT multiplier / addend registers per chain (no operand shared
  // between neighbouring DFMAs); the better of the two is the measured peak
  {
    double* d2 = nullptr;
    if (cudaMalloc(&d2, 8) == cudaSuccess) {
      const int blocks2 = prop.multiProcessorCount * 4, it2 = 4096;
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        k_dfma_peak16<<<blocks2, 256>>>(d2, it2, 0.999999, 1e-9);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double tf = 2.0 * 16.0 * 8.0 * it2 * (double)blocks2 * 256 / (ms * 1e-3) * 1e-12;
        if (rep > 0 && tf > *tflops) *tflops = tf;
      }
      cudaEventDestroy(e0); cudaEventDestroy(e1);
      cudaFree(d2);
    }
  }
  // DFMA lanes issued per SM and cycle = measured throughput / 2 flop / SMs / the clock the probe ran at
  rate = (*tflops) * 1e12 / 2.0 / prop.multiProcessorCount / (mhz * 1e6);
  if (sm_mhz_under_probe) *sm_mhz_under_probe = mhz;
  if (dfma_per_sm_clk) *dfma_per_sm_clk = rate;
  return PDA_OK;
}

extern "C" pda_status pda_test_glibc_pow(int device, const double* x, double y, double* out, int64_t n) {
  if (!x || !out || n < 0) return PDA_ERR_INVALID;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) { cudaGetLastError(); return PDA_ERR_NO_DEVICE; }
  if (cudaSetDevice(device) != cudaSuccess) return PDA_ERR_CUDA;
  double *dx = nullptr, *dout = nullptr;
  if (cudaMalloc(&dx, 8 * (size_t)(n + 1)) != cudaSuccess || cudaMalloc(&dout, 8 * (size_t)(n + 1)) != cudaSuccess) return PDA_ERR_CUDA;
  cudaMemcpy(dx, x, 8 * (size_t)n, cudaMemcpyHostToDevice);
  pda::dev::launchGlibcPow(dx, y, dout, n, nullptr);
  const cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(out, dout, 8 * (size_t)n, cudaMemcpyDeviceToHost);
  cudaFree(dx); cudaFree(dout);
  return e == cudaSuccess ? PDA_OK : PDA_ERR_CUDA;
}

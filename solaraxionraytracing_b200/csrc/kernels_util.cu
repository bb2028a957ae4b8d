// kernels_util.cu — measurement helpers: FMA issue-rate micro-benchmarks.
//
// MEASURED_PEAKS.json records HBM and bf16 tensor peaks only; the ray-tracing kernels are bound by FP64 / FP32
// CUDA-core issue (no dense contraction on this path), so bench.py measures those two denominators itself, in
// the same run as the throughput number, with these kernels.
#include <cuda_runtime.h>

#include "kernels.h"
#include "sart_internal.h"

namespace sart {

// Folds the image replicas of a launch (fast_params.h: FastTables::nImgRep) into the handle's image and clears them.
__global__ void __launch_bounds__(256) k_fold_replicas(double* __restrict__ rep, double* __restrict__ rep2, int nRep,
                                                       size_t stride, size_t plane, double* __restrict__ image,
                                                       double* __restrict__ imageW2) {
  const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= plane) return;
  double a = 0.0, b = 0.0;
  for (int r = 0; r < nRep; ++r) {
    const double x = rep[size_t(r) * stride + i], y = rep2[size_t(r) * stride + i];
    if (x != 0.0 || y != 0.0) { a += x; b += y; rep[size_t(r) * stride + i] = 0.0; rep2[size_t(r) * stride + i] = 0.0; }
  }
  if (a != 0.0 || b != 0.0) { image[i] += a; imageW2[i] += b; }
}
cudaError_t launch_fold_replicas(double* rep, double* rep2, int nRep, size_t stride, size_t plane, double* image,
                                 double* imageW2, cudaStream_t s) {
  k_fold_replicas<<<unsigned((plane + 255) / 256), 256, 0, s>>>(rep, rep2, nRep, stride, plane, image, imageW2);
  return cudaGetLastError();
}

// Mass scan: adds the mass-major accumulators acc[bin][SART_MAX_MASSES] of a launch to the images [mass][bin] and clears them.
__global__ void __launch_bounds__(256) k_fold_mass_acc(double* __restrict__ acc, double* __restrict__ acc2, int nMasses,
                                                       size_t plane, double* __restrict__ image, double* __restrict__ imageW2) {
  const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;   // = bin * SART_MAX_MASSES + mass
  if (i >= plane * SART_MAX_MASSES) return;
  const int m = int(i % SART_MAX_MASSES);
  const size_t bin = i / SART_MAX_MASSES;
  const double a = acc[i], b = acc2[i];
  if (a != 0.0 || b != 0.0) {
    acc[i] = 0.0; acc2[i] = 0.0;
    if (m < nMasses) { image[size_t(m) * plane + bin] += a; imageW2[size_t(m) * plane + bin] += b; }
  }
}
cudaError_t launch_fold_mass_acc(double* acc, double* acc2, int nMasses, size_t plane, double* image, double* imageW2,
                                 cudaStream_t s) {
  const size_t n = plane * SART_MAX_MASSES;
  k_fold_mass_acc<<<unsigned((n + 255) / 256), 256, 0, s>>>(acc, acc2, nMasses, plane, image, imageW2);
  return cudaGetLastError();
}

template <typename T, int kChains>
__global__ void __launch_bounds__(256) k_fma_peak(T* out, T a, T b, int iters) {
  T acc[kChains];
#pragma unroll
  for (int c = 0; c < kChains; ++c) acc[c] = T(threadIdx.x + c);
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < kChains; ++c) acc[c] = acc[c] * a + b;
  }
  T s = T(0);
#pragma unroll
  for (int c = 0; c < kChains; ++c) s += acc[c];
  if (s == T(-1.2345)) out[blockIdx.x * blockDim.x + threadIdx.x] = s;  // never true: keeps the loop alive
}

template <typename T>
static int measure(int device, double* tflops) {
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return fail(SART_ERR_CUDA, "cudaGetDeviceProperties failed");
  constexpr int kChains = 8;
  const int block = 256, grid = prop.multiProcessorCount * 8, iters = 20000;
  T* d = nullptr;
  if (cudaMalloc(&d, size_t(grid) * block * sizeof(T)) != cudaSuccess) return fail(SART_ERR_CUDA, "cudaMalloc failed");
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    k_fma_peak<T, kChains><<<grid, block>>>(d, T(1.0000001), T(1e-7), iters);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(d); return fail(SART_ERR_CUDA, "fma benchmark failed: %s", cudaGetErrorString(cudaGetLastError())); }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flop = 2.0 * double(kChains) * iters * double(grid) * block;
    const double tf = flop / (double(ms) * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(d);
  *tflops = best;
  return SART_OK;
}

}  // namespace sart

extern "C" int sart_measure_fma_peak(int device, int fp64, double* tflops) {
  if (!tflops) return sart::fail(SART_ERR_ARG, "tflops is NULL");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) { cudaGetLastError(); return sart::fail(SART_ERR_CUDA, "no such CUDA device %d", device); }
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(device);
  const int rc = fp64 ? sart::measure<double>(device, tflops) : sart::measure<float>(device, tflops);
  cudaSetDevice(prev);
  return rc;
}

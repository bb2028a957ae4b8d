// kernels_f32x2.cu — the plain run of the single-precision fused Monte Carlo kernel with two rays per thread
// (trace_f32x2.cuh: their FP32 arithmetic in the packed two-lane instructions of sm_100a), sm_100a.
//
// Thread t of the grid traces the rays (2 j, 2 j + 1) of the launch for j = t, t + T, t + 2 T, ... (T threads in the grid):
// consecutive Philox counters, consecutive re-trace queue entries. The per-block tables, the sink (image replicas, exit
// counters per warp, sums per thread) and the flush are those of k_trace_mc_f32 (kernels_f32.cu), and a ray's outcome is
// the same bits, so the two kernels are interchangeable launch by launch (launch_mc_image_f32 chooses).

#include "trace_f32x2.cuh"

namespace sart {
namespace fast {

template <bool kMargins>
__global__ void __launch_bounds__(kBlockX2, 1)
k_trace_mc_f32x2(const __grid_constant__ FastParams P, const __grid_constant__ Geo32 G, const __grid_constant__ FastTables T,
                 double mAxion2, uint64_t first, uint64_t nRays, const __grid_constant__ PhiloxKeys K, double* __restrict__ image,
                 double* __restrict__ imageW2, sart_counters_t* __restrict__ counters) {
  extern __shared__ __align__(16) unsigned char smem[];
  Smem32 S;
  unsigned char* tail;
  smem_layout32<false>(P, smem, S, tail);
  WarpCounters* wc = reinterpret_cast<WarpCounters*>(tail);
  smem_fill32<false>(P, T, S);
  for (int i = threadIdx.x; i < kWarpsX2 * int(sizeof(WarpCounters) / 4); i += kBlockX2) reinterpret_cast<unsigned int*>(wc)[i] = 0u;
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned int nPassed = 0, nTill = 0, nIter = 0;
  double sumW = 0.0, sumW2 = 0.0, sumX = 0.0, sumY = 0.0, sumR = 0.0;
  const uint32_t rep = T.nImgRep > 1 ? (blockIdx.x % unsigned(T.nImgRep)) * uint32_t(T.imgRepStride) : 0u;
  ImageSinkT<false> sink{T, mAxion2, image, imageW2, rep, wc[warp], nPassed, nTill, sumW, sumW2, sumX, sumY, sumR};
  const uint64_t nPairs = (nRays + 1) >> 1;
  const uint64_t stride = uint64_t(gridDim.x) * kBlockX2;   // in pairs
  const uint64_t j0 = uint64_t(blockIdx.x) * kBlockX2 + threadIdx.x;
  const unsigned nTrips = j0 < nPairs ? unsigned((nPairs - 1 - j0) / stride) + 1u : 0u;
  uint64_t ray = first + 2 * j0;
  uint32_t id = uint32_t(2 * j0);
  const uint32_t stride32 = uint32_t(2 * stride);
  for (unsigned k = nTrips; k != 0u; --k, ray += 2 * stride, id += stride32) {
    const bool has1 = uint64_t(id) + 1 < nRays;   // false only for the last pair of an odd launch
    trace_pair32<kMargins>(P, G, T, S, K, ray, ray + 1, id, id + 1u, has1, sink);
    nIter += has1 ? 2u : 1u;
  }
  for (int o = 16; o > 0; o >>= 1) {
    nPassed += __shfl_down_sync(0xffffffffu, nPassed, o);
    nTill += __shfl_down_sync(0xffffffffu, nTill, o);
    nIter += __shfl_down_sync(0xffffffffu, nIter, o);
    sumW += __shfl_down_sync(0xffffffffu, sumW, o);
    sumW2 += __shfl_down_sync(0xffffffffu, sumW2, o);
    sumX += __shfl_down_sync(0xffffffffu, sumX, o);
    sumY += __shfl_down_sync(0xffffffffu, sumY, o);
    sumR += __shfl_down_sync(0xffffffffu, sumR, o);
  }
  __syncwarp();
  if (lane == 0) flush_counters(counters, wc[warp], nIter, nPassed, nTill, sumW, sumW2, sumX, sumY, sumR);
}

}  // namespace fast

#ifndef SART_NO_LAUNCHERS
// One launch of at most kMaxRaysPerLaunch rays (the caller cuts longer runs).
cudaError_t launch_mc_image_f32x2(const fast::FastParams& P, const fast::Geo32& G, const fast::FastTables& T, double mAxion2,
                                  uint64_t first, uint64_t n, const PhiloxKeys& keys, double* image, double* imageW2,
                                  sart_counters_t* counters, int smCount, bool margins, cudaStream_t s) {
  const size_t smem = fast::smem_bytes32(P, fast::kWarpsX2, false);
  auto kern = margins ? fast::k_trace_mc_f32x2<true> : fast::k_trace_mc_f32x2<false>;
  cudaError_t e = fast::set_smem(kern, smem);
  if (e != cudaSuccess) return e;
  int perSM = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kern, fast::kBlockX2, smem);
  if (e != cudaSuccess) return e;
  if (perSM < 1) perSM = 1;
  const uint64_t cap = uint64_t(smCount) * perSM;
  const uint64_t want = ((n + 1) / 2 + fast::kBlockX2 - 1) / fast::kBlockX2;
  const unsigned grid = unsigned(want < cap ? want : cap);
  kern<<<grid, fast::kBlockX2, smem, s>>>(P, G, T, mAxion2, first, n, keys, image, imageW2, counters);
  return cudaGetLastError();
}
#endif

}  // namespace sart

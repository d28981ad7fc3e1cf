// microbench.cu -- measured denominators for the gather roofline:
// an L2-resident random-row gather, the matching vector RED scatter, and the
// latency of the software grid barrier used by den_kernel.cu.

#include "chain_kernels.h"

namespace tb {
namespace {

__device__ __forceinline__ unsigned hash32(unsigned x) {
  x ^= x >> 16;
  x *= 0x7feb352dU;
  x ^= x >> 15;
  x *= 0x846ca68bU;
  x ^= x >> 16;
  return x;
}

// Each warp reads `iters` x 8 random rows of `row_f4` float4 per lane-group.
__global__ void __launch_bounds__(512) gather_kernel(const float4* __restrict__ table, unsigned nrows, int row_f4,
                                                     int iters, float* sink) {
  const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int lanes_per_row = row_f4;             // 32 for 512-byte rows
  const int rows_per_ld = 32 / lanes_per_row;   // rows fetched by one warp-wide load
  const int sub = lane / lanes_per_row, l = lane % lanes_per_row;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  unsigned seed = warp * 9781u + 17u;
  for (int it = 0; it < iters; it++) {
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const unsigned r = hash32(seed + (unsigned)(it * 8 + u) * 2654435761u + (unsigned)sub * 40503u) % nrows;
      v[u] = __ldcg(table + (size_t)r * row_f4 + l);
    }
#pragma unroll
    for (int u = 0; u < 8; u++) {
      acc.x += v[u].x;
      acc.y += v[u].y;
      acc.z += v[u].z;
      acc.w += v[u].w;
    }
    (void)rows_per_ld;
  }
  if (acc.x + acc.y + acc.z + acc.w == 123.456f) *sink = acc.x;
}

__global__ void __launch_bounds__(512) red_kernel(float4* table, unsigned nrows, int row_f4, int iters) {
  const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int lanes_per_row = row_f4;
  const int sub = lane / lanes_per_row, l = lane % lanes_per_row;
  unsigned seed = warp * 9781u + 17u;
  for (int it = 0; it < iters * 8; it++) {
    const unsigned r = hash32(seed + (unsigned)it * 2654435761u + (unsigned)sub * 40503u) % nrows;
    float* ptr = reinterpret_cast<float*>(table + (size_t)r * row_f4 + l);
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(ptr), "f"(1.0f), "f"(1.0f), "f"(1.0f),
                 "f"(1.0f)
                 : "memory");
  }
}

__global__ void __launch_bounds__(512) barrier_kernel(unsigned* counter, int iters) {
  unsigned epoch = 0;
  for (int i = 0; i < iters; i++) {
    __syncthreads();
    if (threadIdx.x == 0) {
      epoch += gridDim.x;
      __threadfence();
      atomicAdd(counter, 1u);
      unsigned v;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
      } while (v < epoch);
    }
    __syncthreads();
  }
}

// Emulates the forward recurrence's access pattern: `warps` warps per CTA (one CTA per SM), each
// keeping `depth` 512-byte row gathers in flight, alternating between a 5 MB and a 1.5 MB table.
template <int DEPTH>
__global__ void __launch_bounds__(1024) pattern_kernel(const float4* __restrict__ ta, unsigned na,
                                                       const float4* __restrict__ te, unsigned ne, int iters,
                                                       float* sink) {
  const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  unsigned seed = warp * 9781u + 17u;
  for (int it = 0; it < iters; it++) {
    float4 v[DEPTH];
#pragma unroll
    for (int u = 0; u < DEPTH; u += 2) {
      const unsigned r = hash32(seed + (unsigned)(it * DEPTH + u) * 2654435761u);
      v[u] = __ldcg(ta + (size_t)(r % na) * 32 + lane);
      v[u + 1] = __ldcg(te + (size_t)((r >> 7) % ne) * 32 + lane);
    }
#pragma unroll
    for (int u = 0; u < DEPTH; u++) {
      acc.x += v[u].x;
      acc.y += v[u].y;
      acc.z += v[u].z;
      acc.w += v[u].w;
    }
  }
  if (acc.x + acc.y + acc.z + acc.w == 123.456f) *sink = acc.x;
}

}  // namespace

double microbench_pattern(int device, int warps, int depth, int iters) {
  if (cudaSetDevice(device) != cudaSuccess) return -1.0;
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  const unsigned na = 10000, ne = 2928;
  float4 *ta = nullptr, *te = nullptr;
  float* sink = nullptr;
  cudaMalloc(&ta, (size_t)na * 512);
  cudaMalloc(&te, (size_t)ne * 512);
  cudaMalloc(&sink, 4);
  cudaMemset(ta, 0, (size_t)na * 512);
  cudaMemset(te, 0, (size_t)ne * 512);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    cudaEventRecord(e0);
    if (depth == 2) pattern_kernel<2><<<sms, warps * 32>>>(ta, na, te, ne, iters, sink);
    else if (depth == 4) pattern_kernel<4><<<sms, warps * 32>>>(ta, na, te, ne, iters, sink);
    else if (depth == 6) pattern_kernel<6><<<sms, warps * 32>>>(ta, na, te, ne, iters, sink);
    else if (depth == 8) pattern_kernel<8><<<sms, warps * 32>>>(ta, na, te, ne, iters, sink);
    else if (depth == 12) pattern_kernel<12><<<sms, warps * 32>>>(ta, na, te, ne, iters, sink);
    else if (depth == 16) pattern_kernel<16><<<sms, warps * 32>>>(ta, na, te, ne, iters, sink);
    else pattern_kernel<24><<<sms, warps * 32>>>(ta, na, te, ne, iters, sink);
    count_launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0) best = std::min(best, ms);
  }
  const bool ok = cudaGetLastError() == cudaSuccess;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(ta);
  cudaFree(te);
  cudaFree(sink);
  if (!ok) return -1.0;
  const int d = depth == 2 || depth == 4 || depth == 6 || depth == 8 || depth == 12 || depth == 16 ? depth : 24;
  return (double)sms * warps * (double)iters * d * 512.0 / (best * 1e-3) / 1e9;
}

double microbench_l2_gather(int device, int64_t table_bytes, int row_bytes, int iters) {
  if (cudaSetDevice(device) != cudaSuccess) return -1.0;
  if (row_bytes < 16 || row_bytes > 512 || (row_bytes & (row_bytes - 1))) return -1.0;
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  const int row_f4 = row_bytes / 16;
  const unsigned nrows = (unsigned)(table_bytes / row_bytes);
  float4* table = nullptr;
  float* sink = nullptr;
  if (cudaMalloc(&table, (size_t)nrows * row_bytes) != cudaSuccess) return -1.0;
  cudaMalloc(&sink, 4);
  cudaMemset(table, 0, (size_t)nrows * row_bytes);
  const int grid = sms * 2, block = 512;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  gather_kernel<<<grid, block>>>(table, nrows, row_f4, iters, sink);  // warm-up (fills L2)
  count_launch();
  float best = 1e30f;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0);
    gather_kernel<<<grid, block>>>(table, nrows, row_f4, iters, sink);
    count_launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    best = std::min(best, ms);
  }
  const bool ok = cudaGetLastError() == cudaSuccess;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(table);
  cudaFree(sink);
  if (!ok) return -1.0;
  const double bytes = (double)grid * (block / 32) * (double)iters * 8.0 * 32.0 * 16.0;
  return bytes / (best * 1e-3) / 1e9;
}

double microbench_l2_red(int device, int64_t table_bytes, int row_bytes, int iters) {
  if (cudaSetDevice(device) != cudaSuccess) return -1.0;
  if (row_bytes < 16 || row_bytes > 512 || (row_bytes & (row_bytes - 1))) return -1.0;
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  const int row_f4 = row_bytes / 16;
  const unsigned nrows = (unsigned)(table_bytes / row_bytes);
  float4* table = nullptr;
  if (cudaMalloc(&table, (size_t)nrows * row_bytes) != cudaSuccess) return -1.0;
  cudaMemset(table, 0, (size_t)nrows * row_bytes);
  const int grid = sms * 2, block = 512;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  red_kernel<<<grid, block>>>(table, nrows, row_f4, iters);
  count_launch();
  float best = 1e30f;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0);
    red_kernel<<<grid, block>>>(table, nrows, row_f4, iters);
    count_launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    best = std::min(best, ms);
  }
  const bool ok = cudaGetLastError() == cudaSuccess;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(table);
  if (!ok) return -1.0;
  const double bytes = (double)grid * (block / 32) * (double)iters * 8.0 * 32.0 * 16.0;
  return bytes / (best * 1e-3) / 1e9;
}

double microbench_grid_barrier(int device, int iters) {
  if (cudaSetDevice(device) != cudaSuccess) return -1.0;
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  unsigned* counter = nullptr;
  if (cudaMalloc(&counter, 4) != cudaSuccess) return -1.0;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    cudaMemset(counter, 0, 4);
    void* args[] = {&counter, &iters};
    cudaEventRecord(e0);
    cudaError_t e = cudaLaunchCooperativeKernel((void*)barrier_kernel, dim3(sms), dim3(512), args, 0, 0);
    count_launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    if (e != cudaSuccess) {
      best = -1.f;
      break;
    }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0) best = std::min(best, ms);
  }
  const bool ok = cudaGetLastError() == cudaSuccess && best > 0.f;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(counter);
  if (!ok) return -1.0;
  return (double)best * 1e3 / iters;
}

}  // namespace tb

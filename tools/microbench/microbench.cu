// microbench.cu -- libtorchain_microbench.so: measured denominators for the gather roofline
//  and feasibility probes.  BENCH-ONLY: bench.py and tools/ load it; nothing in
// torchain_b200/ links or loads it, and none of its symbols is part of the drop-in boundary
// (include/torchain_b200.h).
//   * L2-resident random-row gather / vector RED scatter / software grid barrier latency
//   * the forward recurrence's gather pattern (two tables, 512 B rows) with register loads,
//     with an LDGSTS (cp.async) ring in shared memory, and with a TMA bulk-copy ring
//   * the exp(y) tile held in distributed shared memory of a cluster
//   * the sequence-sliced recurrence ruled out in round 1

#include "microbench.h"

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include <algorithm>
#include <cstring>
#include <vector>

namespace tb {
namespace {

std::atomic<long long> g_mb_launches{0};
void count_launch(int n = 1) { g_mb_launches += n; }

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


// ---------------------------------------------------------------------------
// Feasibility probe for the sequence-sliced recurrence: a CTA owns S = 8 or
// 16 sequences and 1/M of the arcs (G = 128/S groups of M = 148/G CTAs); exp(y_t) of its sequences lives in
// shared memory ([pdf][S], re-staged every frame), alpha / beta rows are gathered from L2 as ONE 4*S-byte
// segment per arc, arc records are streamed from global memory (interleaved so that a warp step is one
// coalesced load), 4 lanes (S = 16) or 2 lanes (S = 8) walk one arc stream with float4 vectors, i.e. a warp
// advances 8 / 16 streams at once.  MODE 1 adds the backward pass's occupancy accumulation as shared-memory
// atomics into a [pdf][S] slice that is RED-ed into the global tile at the end of the frame (S = 8 only:
// two slices must fit).  Synthetic graph of config 2's shape; the numbers mean nothing, the time does.
// ---------------------------------------------------------------------------
struct SlicedArgs {
  const float* table;   // [2][N][128] alpha (or beta) ping-pong
  float* table_w;
  const float* y;       // [frames][128][P]
  const int* src;       // [M][steps*streams_per_cta] row index of the other state, stream-interleaved per warp
  const int4* rec;      // same order: {pdf, prob | end-of-state flag, this state, pi bits}
  float* gtile;         // [P][128] occupancy tile (MODE 1)
  unsigned* barriers;   // [G][32]
  int N, P, frames, steps, M;
  int stage_once;  // 1: exp(y) is staged for frame 0 only (separates the staging cost from the gather phase)
};

template <int S, int DEPTH, int MODE>
__global__ void __launch_bounds__(512, 1) sliced_kernel(const __grid_constant__ SlicedArgs a) {
  extern __shared__ __align__(16) float sm[];
  constexpr int L = S / 4, STREAMS = 32 / L, G = 128 / S;
  float* esm = sm;                 // [P][S]
  float* gsm = sm + (size_t)a.P * S;  // [P][S] (MODE 1)
  if ((int)blockIdx.x >= G * a.M) return;
  const int g = blockIdx.x % G, m = blockIdx.x / G;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int li = lane % L;
  const int nwarps = blockDim.x >> 5;
  // element (step k, this lane's stream) of this warp's interleaved arc arrays
  const size_t wbase = ((size_t)m * nwarps + warp) * a.steps * STREAMS + lane / L;
  unsigned epoch = 0;
  float4 keep = make_float4(0.f, 0.f, 0.f, 0.f);
  if (MODE >= 1)
    for (int i = threadIdx.x; i < a.P * S; i += blockDim.x) gsm[i] = 0.f;
  for (int t = 0; t < a.frames; t++) {
    // stage exp(y_t) of sequences [g*S, g*S+S): coalesced along pdf, 16-byte chunks swizzled by pdf
    for (int i = threadIdx.x; i < ((a.stage_once && t > 0) ? 0 : a.P * S); i += blockDim.x) {
      const int s = i / a.P, pdf = i - s * a.P;
      const float v = __ldg(a.y + ((size_t)t * 128 + g * S + s) * a.P + pdf);
      esm[pdf * S + (((s >> 2) ^ (pdf & (L - 1))) << 2) + (s & 3)] = __expf(v * 0.01f);
    }
    __syncthreads();
    const float* tr = a.table + (size_t)(t & 1) * a.N * 128 + g * S + li * 4;
    float* tw = a.table_w + (size_t)((t + 1) & 1) * a.N * 128 + g * S + li * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const float lam = 1e-3f, occ = 0.5f;
    float4 av[2][DEPTH];
    int4 rc[2][DEPTH];
    int sn[2][DEPTH];  // row indices, fetched two batches ahead of their gathers
    auto issue = [&](int b, int k) {  // gathers + records of batch k (its row indices are in sn[b])
#pragma unroll
      for (int u = 0; u < DEPTH; u++) {
        av[b][u] = __ldcg(reinterpret_cast<const float4*>(tr + (size_t)sn[b][u] * 128));
        rc[b][u] = __ldg(a.rec + wbase + (size_t)(k + u) * STREAMS);
      }
    };
    auto fetch_src = [&](int b, int k) {
#pragma unroll
      for (int u = 0; u < DEPTH; u++) sn[b][u] = (k + u < a.steps) ? __ldg(a.src + wbase + (size_t)(k + u) * STREAMS) : 0;
    };
    auto consume = [&](int b) {
#pragma unroll
      for (int u = 0; u < DEPTH; u++) {
        const int4 r = rc[b][u];
        const float pr = fabsf(__int_as_float(r.y)), pi = __int_as_float(r.w);
        const float4 e = *reinterpret_cast<const float4*>(esm + r.x * S + ((li ^ (r.x & (L - 1))) << 2));
        const float4 x = av[b][u];
        if (MODE == 0) {
          acc.x = fmaf(fmaf(pi, lam, x.x) * pr, e.x, acc.x);
          acc.y = fmaf(fmaf(pi, lam, x.y) * pr, e.y, acc.y);
          acc.z = fmaf(fmaf(pi, lam, x.z) * pr, e.z, acc.z);
          acc.w = fmaf(fmaf(pi, lam, x.w) * pr, e.w, acc.w);
        } else {
          const float v0 = pr * (x.x + lam) * e.x, v1 = pr * (x.y + lam) * e.y, v2 = pr * (x.z + lam) * e.z,
                      v3 = pr * (x.w + lam) * e.w;
          acc.x += v0; acc.y += v1; acc.z += v2; acc.w += v3;
          float* gp = gsm + r.x * S + ((li ^ (r.x & (L - 1))) << 2);
          if (MODE == 1) {  // fp32 shared atomics: ATOMS.CAST.SPIN loops on sm_100a
            atomicAdd(gp + 0, v0 * occ);
            atomicAdd(gp + 1, v1 * occ);
            atomicAdd(gp + 2, v2 * occ);
            atomicAdd(gp + 3, v3 * occ);
          } else {  // MODE 2: 2^-30 fixed point (occupancies are probabilities), native integer ATOMS.ADD
            int* ip = reinterpret_cast<int*>(gp);
            const float fx = 1073741824.f * occ;
            atomicAdd(ip + 0, __float2int_rn(v0 * fx));
            atomicAdd(ip + 1, __float2int_rn(v1 * fx));
            atomicAdd(ip + 2, __float2int_rn(v2 * fx));
            atomicAdd(ip + 3, __float2int_rn(v3 * fx));
          }
        }
        if (r.y < 0) {  // last arc of this state
          *reinterpret_cast<float4*>(tw + (size_t)r.z * 128) = acc;
          keep.x += acc.x;
          acc = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    };
    fetch_src(0, 0);
    fetch_src(1, DEPTH);
    issue(0, 0);
    fetch_src(0, 2 * DEPTH);
    for (int k = 0; k < a.steps; k += 2 * DEPTH) {  // steps is a multiple of 2*DEPTH
      issue(1, k + DEPTH);
      fetch_src(1, k + 3 * DEPTH);
      consume(0);
      if (k + 2 * DEPTH < a.steps) issue(0, k + 2 * DEPTH);
      fetch_src(0, k + 4 * DEPTH);
      consume(1);
    }
    if (MODE >= 1) {  // this CTA's partial occupancies of its S sequences -> global tile, slice cleared
      __syncthreads();
      for (int i = threadIdx.x; i < a.P * L; i += blockDim.x) {
        const int pdf = i / L, c = i - pdf * L;
        float4* sp = reinterpret_cast<float4*>(gsm + pdf * S + ((c ^ (pdf & (L - 1))) << 2));
        const float4 v = *sp;
        *sp = make_float4(0.f, 0.f, 0.f, 0.f);
        float* dst = a.gtile + (size_t)pdf * 128 + g * S + c * 4;
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                     : "memory");
      }
    }
    // barrier over the M CTAs of this sequence group
    __syncthreads();
    if (threadIdx.x == 0) {
      epoch += a.M;
      __threadfence();
      atomicAdd(a.barriers + g * 32, 1u);
      unsigned v;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(a.barriers + g * 32) : "memory");
      } while (v < epoch);
    }
    __syncthreads();
  }
  if (keep.x == 123.456f) a.table_w[0] = keep.x;
}

template <int S, int DEPTH, int MODE>
cudaError_t launch_sliced(const SlicedArgs& a, int sms, size_t smem) {
  auto kern = sliced_kernel<S, DEPTH, MODE>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  void* args[] = {const_cast<SlicedArgs*>(&a)};
  e = cudaLaunchCooperativeKernel((void*)kern, dim3(sms), dim3(512), args, smem, 0);
  count_launch();
  return e;
}

}  // namespace

// us per frame of the sliced probe; S = 8 | 16 sequences per CTA, depth = 4 | 8 arcs per batch (two batches in
// flight), mode 0 = forward-like, 1 = backward-like (S = 8).  < 0 on error.
double microbench_sliced(int device, int S, int depth, int mode, int frames) {
  if (cudaSetDevice(device) != cudaSuccess) return -1.0;
  const int stage_once = mode >= 4 ? 1 : 0;  // mode + 4: stage exp(y) once
  mode &= 3;
  if ((S != 8 && S != 16) || (depth != 4 && depth != 8) || (mode < 0 || mode > 2) || (mode >= 1 && S != 8) || frames < 1)
    return -1.0;
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  const int N = 10000, P = 2928, deg = 30, A = N * deg;
  const int G = 128 / S, M = sms / G, L = S / 4, STREAMS = 32 / L, nwarps = 16;
  if (M < 1) return -1.0;
  const int streams_cta = nwarps * STREAMS;
  const int per_cta = (A + M - 1) / M, per_stream = (per_cta + streams_cta - 1) / streams_cta;
  const int steps = (per_stream + 2 * depth - 1) / (2 * depth) * (2 * depth);
  std::vector<int> src((size_t)M * streams_cta * steps, 0);
  std::vector<int4> rec((size_t)M * streams_cta * steps, make_int4(0, 0, 0, 0));
  unsigned rng = 12345u;
  auto rnd = [&]() {
    rng = rng * 1664525u + 1013904223u;
    return rng >> 8;
  };
  for (int m = 0; m < M; m++)
    for (int w = 0; w < nwarps; w++)
      for (int st = 0; st < STREAMS; st++)
        for (int k = 0; k < steps; k++) {
          const int sidx = w * STREAMS + st;
          const int arc = m * per_cta + sidx * per_stream + k;  // global arc id (destination-sorted)
          const size_t o = (((size_t)m * nwarps + w) * steps + k) * STREAMS + st;
          if (k < per_stream && arc < A) {
            const int state = arc / deg;
            float pr = 1.0f / deg, pi = 1e-4f;
            int prb, pib;
            std::memcpy(&prb, &pr, 4);
            std::memcpy(&pib, &pi, 4);
            if (arc % deg == deg - 1 || k == per_stream - 1) prb |= (int)0x80000000u;
            src[o] = (int)(rnd() % N);
            rec[o] = make_int4((int)(rnd() % P), prb, state, pib);
          }
        }
  float *table = nullptr, *y = nullptr, *gtile = nullptr;
  int* dsrc = nullptr;
  int4* drec = nullptr;
  unsigned* bar = nullptr;
  const size_t tbytes = (size_t)2 * N * 128 * 4, ybytes = (size_t)frames * 128 * P * 4;
  bool ok = cudaMalloc(&table, tbytes) == cudaSuccess && cudaMalloc(&y, ybytes) == cudaSuccess &&
            cudaMalloc(&gtile, (size_t)P * 128 * 4) == cudaSuccess && cudaMalloc(&dsrc, src.size() * 4) == cudaSuccess &&
            cudaMalloc(&drec, rec.size() * 16) == cudaSuccess && cudaMalloc(&bar, G * 32 * 4) == cudaSuccess;
  double result = -1.0;
  if (ok) {
    cudaMemset(table, 0, tbytes);
    cudaMemset(y, 0, ybytes);
    cudaMemset(gtile, 0, (size_t)P * 128 * 4);
    cudaMemcpy(dsrc, src.data(), src.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(drec, rec.data(), rec.size() * 16, cudaMemcpyHostToDevice);
    SlicedArgs a{table, table, y, dsrc, drec, gtile, bar, N, P, frames, steps, M, stage_once};
    const size_t smem = (size_t)P * S * 4 * (mode >= 1 ? 2 : 1);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4 && ok; rep++) {
      cudaMemset(bar, 0, G * 32 * 4);
      cudaEventRecord(e0);
      cudaError_t e;
      if (S == 16) e = depth == 4 ? launch_sliced<16, 4, 0>(a, sms, smem) : launch_sliced<16, 8, 0>(a, sms, smem);
      else if (mode == 0) e = depth == 4 ? launch_sliced<8, 4, 0>(a, sms, smem) : launch_sliced<8, 8, 0>(a, sms, smem);
      else if (mode == 1) e = depth == 4 ? launch_sliced<8, 4, 1>(a, sms, smem) : launch_sliced<8, 8, 1>(a, sms, smem);
      else e = depth == 4 ? launch_sliced<8, 4, 2>(a, sms, smem) : launch_sliced<8, 8, 2>(a, sms, smem);
      cudaEventRecord(e1);
      ok = e == cudaSuccess && cudaEventSynchronize(e1) == cudaSuccess;
      float ms = 0.f;
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep > 0) best = std::min(best, ms);
    }
    ok = ok && cudaGetLastError() == cudaSuccess;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (ok) result = (double)best * 1e3 / frames;
  }
  cudaFree(table);
  cudaFree(y);
  cudaFree(gtile);
  cudaFree(dsrc);
  cudaFree(drec);
  cudaFree(bar);
  return result;
}

namespace {

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


namespace {
// ---------------------------------------------------------------------------
// The same two-table gather with the rows landing in SHARED memory instead of registers.
// Register loads cap a warp at ~2 rows in flight whatever the unroll depth (pattern_kernel: the
// throughput follows the warp count, not the depth); asynchronous copies are not tied to the
// warp's scoreboards, so a few warps can keep many rows in flight.
//   MODE 0: LDGSTS (cp.async.cg 16 B per lane, one 512 B row per warp instruction), groups
//           committed per arc, cp.async.wait_group DEPTH-1 before an arc is consumed
//   MODE 1: TMA bulk copies (cp.async.bulk, one elected lane per row, completion on an mbarrier
//           per ring slot)
// Each warp owns a ring of DEPTH slots x 2 rows (alpha row + exp row) = DEPTH KB.
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int DEPTH, int MODE>
__global__ void __launch_bounds__(1024) pattern_async_kernel(const float4* __restrict__ ta, unsigned na,
                                                             const float4* __restrict__ te, unsigned ne, int iters,
                                                             float* sink) {
  extern __shared__ __align__(128) unsigned char smem_async[];
  const unsigned warp_in_cta = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  float4* ring = reinterpret_cast<float4*>(smem_async) + (size_t)warp_in_cta * DEPTH * 64;  // [DEPTH][2][32] float4
  unsigned long long* bars =
      reinterpret_cast<unsigned long long*>(smem_async + (size_t)nwarps * DEPTH * 1024) + warp_in_cta * DEPTH;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const unsigned seed = warp * 9781u + 17u;
  if (MODE == 1) {
    if (lane == 0)
      for (int u = 0; u < DEPTH; u++)
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bars + u)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
  }
  auto issue = [&](int k) {  // arc k -> slot k % DEPTH
    const unsigned r = hash32(seed + (unsigned)k * 2654435761u);
    const float4* ga = ta + (size_t)(r % na) * 32;
    const float4* ge = te + (size_t)((r >> 7) % ne) * 32;
    float4* slot = ring + (size_t)(k % DEPTH) * 64;
    if (MODE == 0) {
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(slot + lane)), "l"(ga + lane) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(slot + 32 + lane)), "l"(ge + lane)
                   : "memory");
      asm volatile("cp.async.commit_group;" ::: "memory");
    } else if (lane == 0) {
      const unsigned bar = smem_u32(bars + k % DEPTH);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 1024;" ::"r"(bar) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 512, [%2];" ::"r"(
                       smem_u32(slot)),
                   "l"(ga), "r"(bar)
                   : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 512, [%2];" ::"r"(
                       smem_u32(slot + 32)),
                   "l"(ge), "r"(bar)
                   : "memory");
    }
  };
#pragma unroll
  for (int u = 0; u < DEPTH; u++) issue(u);
  for (int k = 0; k < iters; k++) {
    if (MODE == 0) {
      asm volatile("cp.async.wait_group %0;" ::"n"(DEPTH - 1) : "memory");
    } else {
      const unsigned bar = smem_u32(bars + k % DEPTH), parity = (unsigned)(k / DEPTH) & 1u;
      unsigned done = 0;
      while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
      }
    }
    const float4* slot = ring + (size_t)(k % DEPTH) * 64;
    const float4 a = slot[lane], e = slot[32 + lane];
    acc.x = fmaf(a.x, e.x, acc.x);
    acc.y = fmaf(a.y, e.y, acc.y);
    acc.z = fmaf(a.z, e.z, acc.z);
    acc.w = fmaf(a.w, e.w, acc.w);
    if (MODE == 1) __syncwarp();  // every lane has read the slot before lane 0 re-arms it
    issue(k + DEPTH);             // (the tail issues DEPTH rows nobody consumes; drained below)
  }
  if (MODE == 0) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  } else {
    for (int k = iters; k < iters + DEPTH; k++) {
      const unsigned bar = smem_u32(bars + k % DEPTH), parity = (unsigned)(k / DEPTH) & 1u;
      unsigned done = 0;
      while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
      }
    }
  }
  if (acc.x + acc.y + acc.z + acc.w == 123.456f) *sink = acc.x;
}

template <int DEPTH, int MODE>
cudaError_t launch_pattern_async(int sms, int warps, const float4* ta, unsigned na, const float4* te, unsigned ne,
                                 int iters, float* sink) {
  auto kern = pattern_async_kernel<DEPTH, MODE>;
  const size_t smem = (size_t)warps * DEPTH * 1024 + (size_t)warps * DEPTH * 8;
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kern<<<sms, warps * 32, smem>>>(ta, na, te, ne, iters, sink);
  count_launch();
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// exp(y) tile in DISTRIBUTED shared memory: a cluster of CS CTAs holds the
// [P][128] fp32 tile (1.5 MB at config 2) across its shared memories, P/CS rows per CTA; every warp
// alternates an L2 gather of an alpha row with a 512 B row read from the shared memory of the CTA
// that owns the pdf (ld.shared::cluster.v4 through mapa).  MODE 0: both halves (the mixed pattern),
// MODE 1: DSMEM reads only, MODE 2: L2 reads only on the same launch geometry.
// ---------------------------------------------------------------------------
template <int DEPTH>
__global__ void __launch_bounds__(1024) dsmem_pattern_kernel(const float4* __restrict__ ta, unsigned na, unsigned ne,
                                                             int cs, int mode, int iters, float* sink) {
  extern __shared__ __align__(128) unsigned char smem_ds[];
  float4* mine = reinterpret_cast<float4*>(smem_ds);
  const unsigned rows_per_cta = (ne + cs - 1) / cs;
  for (unsigned i = threadIdx.x; i < rows_per_cta * 32; i += blockDim.x) mine[i] = make_float4(1.f, 1.f, 1.f, 1.f);
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const unsigned base = smem_u32(mine);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const unsigned seed = warp * 9781u + 17u;
  for (int it = 0; it < iters; it++) {
    float4 v[DEPTH];
#pragma unroll
    for (int u = 0; u < DEPTH; u += 2) {
      const unsigned r = hash32(seed + (unsigned)(it * DEPTH + u) * 2654435761u);
      const unsigned pdf = (r >> 7) % ne, pdf2 = (r >> 3) % ne;
      unsigned remote, remote2;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;"
                   : "=r"(remote)
                   : "r"(base + ((pdf / (unsigned)cs) * 32 + lane) * 16), "r"(pdf % (unsigned)cs));
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;"
                   : "=r"(remote2)
                   : "r"(base + ((pdf2 / (unsigned)cs) * 32 + lane) * 16), "r"(pdf2 % (unsigned)cs));
      if (mode != 1) v[u] = __ldcg(ta + (size_t)(r % na) * 32 + lane);
      else
        asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w)
                     : "r"(remote2));
      if (mode != 2)
        asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(v[u + 1].x), "=f"(v[u + 1].y), "=f"(v[u + 1].z), "=f"(v[u + 1].w)
                     : "r"(remote));
      else v[u + 1] = __ldcg(ta + (size_t)((r >> 3) % na) * 32 + lane);
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
  // nobody may exit while a peer can still read its shared memory
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

}  // namespace

// GB/s of row bytes delivered (both tables), scaled to nothing: the launch covers every SM.
// mode 0 = LDGSTS ring, 1 = TMA bulk ring.
double microbench_pattern_async(int device, int warps, int depth, int mode, int iters) {
  if (cudaSetDevice(device) != cudaSuccess) return -1.0;
  if (warps < 1 || warps > 32 || mode < 0 || mode > 1) return -1.0;
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
  bool ok = true;
  for (int rep = 0; rep < 4 && ok; rep++) {
    cudaEventRecord(e0);
    cudaError_t e = cudaErrorInvalidValue;
#define TB_PA(D)                                                                                             \
  if (depth == D)                                                                                            \
    e = mode == 0 ? launch_pattern_async<D, 0>(sms, warps, ta, na, te, ne, iters, sink)                      \
                  : launch_pattern_async<D, 1>(sms, warps, ta, na, te, ne, iters, sink);
    TB_PA(2) TB_PA(4) TB_PA(6) TB_PA(8) TB_PA(12) TB_PA(16)
#undef TB_PA
    cudaEventRecord(e1);
    ok = e == cudaSuccess && cudaEventSynchronize(e1) == cudaSuccess;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0) best = std::min(best, ms);
  }
  ok = ok && cudaGetLastError() == cudaSuccess;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(ta);
  cudaFree(te);
  cudaFree(sink);
  if (!ok) return -1.0;
  return (double)sms * warps * (double)iters * 1024.0 / (best * 1e-3) / 1e9;
}

// GB/s of row bytes delivered PER 148 SMs (the launch may cover fewer: clusters of 8 / 16 do not
// tile every GPC; *ctas_out receives the CTAs that ran).  mode 0 = L2 + DSMEM, 1 = DSMEM only, 2 = L2 only.
double microbench_dsmem_pattern(int device, int cluster_size, int warps, int depth, int mode, int iters,
                                int* ctas_out) {
  if (cudaSetDevice(device) != cudaSuccess) return -1.0;
  if ((cluster_size != 2 && cluster_size != 4 && cluster_size != 8 && cluster_size != 16) || warps < 1 || warps > 32 ||
      mode < 0 || mode > 2)
    return -1.0;
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  const unsigned na = 10000, ne = 2928;
  const size_t smem = (size_t)((ne + cluster_size - 1) / cluster_size) * 512;
  if (smem > 227 * 1024) return -1.0;
  void (*kern)(const float4*, unsigned, unsigned, int, int, int, float*) = nullptr;
  if (depth == 2) kern = dsmem_pattern_kernel<2>;
  else if (depth == 4) kern = dsmem_pattern_kernel<4>;
  else if (depth == 8) kern = dsmem_pattern_kernel<8>;
  else return -1.0;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1.0;
  if (cluster_size > 8 &&
      cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess)
    return -1.0;
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster_size;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(warps * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cfg.gridDim = dim3(cluster_size);
  int nclusters = 0;
  if (cudaOccupancyMaxActiveClusters(&nclusters, kern, &cfg) != cudaSuccess || nclusters < 1) {
    (void)cudaGetLastError();
    return -1.0;
  }
  nclusters = std::min(nclusters, sms / cluster_size);
  const int ctas = nclusters * cluster_size;
  cfg.gridDim = dim3(ctas);
  if (ctas_out) *ctas_out = ctas;
  float4* ta = nullptr;
  float* sink = nullptr;
  cudaMalloc(&ta, (size_t)na * 512);
  cudaMalloc(&sink, 4);
  cudaMemset(ta, 0, (size_t)na * 512);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  bool ok = true;
  for (int rep = 0; rep < 4 && ok; rep++) {
    cudaEventRecord(e0);
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, (const float4*)ta, na, ne, cluster_size, mode, iters, sink);
    count_launch();
    cudaEventRecord(e1);
    ok = e == cudaSuccess && cudaEventSynchronize(e1) == cudaSuccess;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0) best = std::min(best, ms);
  }
  ok = ok && cudaGetLastError() == cudaSuccess;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(ta);
  cudaFree(sink);
  if (!ok) return -1.0;
  return (double)sms * warps * (double)iters * depth * 512.0 / (best * 1e-3) / 1e9;
}

}  // namespace tb

extern "C" {
double tb_microbench_l2_gather(int device, int64_t table_bytes, int row_bytes, int iters) {
  return tb::microbench_l2_gather(device, table_bytes, row_bytes, iters);
}
double tb_microbench_l2_red(int device, int64_t table_bytes, int row_bytes, int iters) {
  return tb::microbench_l2_red(device, table_bytes, row_bytes, iters);
}
double tb_microbench_grid_barrier(int device, int iters) { return tb::microbench_grid_barrier(device, iters); }
double tb_microbench_pattern(int device, int warps, int depth, int iters) {
  return tb::microbench_pattern(device, warps, depth, iters);
}
double tb_microbench_pattern_async(int device, int warps, int depth, int mode, int iters) {
  return tb::microbench_pattern_async(device, warps, depth, mode, iters);
}
double tb_microbench_dsmem_pattern(int device, int cluster_size, int warps, int depth, int mode, int iters,
                                   int* ctas_out) {
  return tb::microbench_dsmem_pattern(device, cluster_size, warps, depth, mode, iters, ctas_out);
}
double tb_microbench_sliced(int device, int seqs_per_cta, int depth, int mode, int frames) {
  return tb::microbench_sliced(device, seqs_per_cta, depth, mode, frames);
}
long long tb_microbench_launch_count(void) { return tb::g_mb_launches.load(); }
}

// layout_kernel.cu -- (B, P, T) <-> (T*B, P) relayout of the network output / its gradient for sm_100a.
//
// Replaces, for the hot path only, `to2d` of the reference
// (/root/reference/torchain/functions.py:118-125: x.permute(2, 0, 1).contiguous().view(-1, C)) and the
// permuted gradient copy its autograd backward produces.  Both are pure HBM traffic (read + write of
// the matrix once each); the generic strided copy kernel behind permute().contiguous() reaches well
// under 1 TB/s on this shape (T = 50 floats innermost), these tiled kernels run at HBM speed.
//
// One CTA moves a (sequence, 32 pdfs) block: in (B, P, T) that is ONE contiguous run of 32*T floats,
// in (T*B, P) it is T row segments of 128 bytes.  The block is staged in shared memory with an odd row
// pitch, so both the contiguous side and the row-segment side are conflict-free and fully coalesced.

#include "chain_kernels.h"

namespace tb {
namespace {

constexpr int kLayThreads = 256;
constexpr int kLayPdfs = 32;

// TO2D = true : x (B, P, T) contiguous -> y rows (t*B + s), pdf innermost, row stride `stride`
// TO2D = false: y rows -> x (B, P, T), multiplied by `scale`
template <bool TO2D>
__global__ void __launch_bounds__(kLayThreads) relayout_kernel(float* __restrict__ x, float* __restrict__ y, int B,
                                                               int P, int T, int64_t stride, float scale) {
  extern __shared__ float tile[];  // [kLayPdfs][pitch]
  const int pitch = T | 1;
  const int s = blockIdx.y, p0 = blockIdx.x * kLayPdfs;
  const int np = min(kLayPdfs, P - p0);
  const int n = np * T;
  float* chunk = x + ((int64_t)s * P + p0) * T;  // np*T contiguous floats
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // (p, t) of contiguous element i = threadIdx.x + k*kLayThreads, advanced without divisions
  const int dq = kLayThreads / T, dr = kLayThreads % T;
  int p = threadIdx.x / T, t = threadIdx.x % T;
  constexpr int kU = 8;  // loads in flight per thread
  constexpr int kWarps = kLayThreads / 32;
  if (TO2D) {
    for (int i0 = threadIdx.x; i0 < n; i0 += kU * kLayThreads) {
      float v[kU];
#pragma unroll
      for (int u = 0; u < kU; u++) v[u] = (i0 + u * kLayThreads < n) ? __ldcs(chunk + i0 + u * kLayThreads) : 0.f;
#pragma unroll
      for (int u = 0; u < kU; u++) {
        if (i0 + u * kLayThreads < n) tile[p * pitch + t] = v[u];
        p += dq;
        t += dr;
        if (t >= T) {
          t -= T;
          p++;
        }
      }
    }
    __syncthreads();
    if (lane < np)
      for (int tt = warp; tt < T; tt += kWarps) y[((int64_t)tt * B + s) * stride + p0 + lane] = tile[lane * pitch + tt];
  } else {
    if (lane < np) {
      for (int t0 = warp; t0 < T; t0 += kU * kWarps) {
        float v[kU];
#pragma unroll
        for (int u = 0; u < kU; u++) {
          const int tt = t0 + u * kWarps;
          v[u] = (tt < T) ? __ldcs(y + ((int64_t)tt * B + s) * stride + p0 + lane) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < kU; u++) {
          const int tt = t0 + u * kWarps;
          if (tt < T) tile[lane * pitch + tt] = v[u];
        }
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += kLayThreads) {
      __stcs(chunk + i, scale * tile[p * pitch + t]);
      p += dq;
      t += dr;
      if (t >= T) {
        t -= T;
        p++;
      }
    }
  }
}

template <bool TO2D>
cudaError_t launch_relayout(float* x, float* y, int B, int P, int T, int64_t stride, float scale,
                            cudaStream_t stream) {
  if (B <= 0 || P <= 0 || T <= 0) return cudaSuccess;
  const size_t smem = sizeof(float) * kLayPdfs * (size_t)(T | 1);
  if (smem > 200 * 1024 || B > 65535) return cudaErrorInvalidValue;
  auto kern = relayout_kernel<TO2D>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kern<<<dim3((P + kLayPdfs - 1) / kLayPdfs, B), kLayThreads, smem, stream>>>(x, y, B, P, T, stride, scale);
  count_launch();
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_to2d(const float* x_bpt, float* y_rows, int B, int P, int T, int64_t stride, cudaStream_t stream) {
  return launch_relayout<true>(const_cast<float*>(x_bpt), y_rows, B, P, T, stride, 1.0f, stream);
}

cudaError_t launch_from2d(const float* y_rows, float* x_bpt, int B, int P, int T, int64_t stride, float scale,
                          cudaStream_t stream) {
  return launch_relayout<false>(x_bpt, const_cast<float*>(y_rows), B, P, T, stride, scale, stream);
}

}  // namespace tb

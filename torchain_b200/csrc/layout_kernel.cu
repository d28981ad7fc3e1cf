// layout_kernel.cu -- (B, P, T) <-> (T*B, P) relayout of the network output / its gradient for sm_100a.
//
// Replaces, for the hot path only, `to2d` of the reference
// (/root/reference/torchain/functions.py:118-125: x.permute(2, 0, 1).contiguous().view(-1, C)) and the
// permuted gradient copy its autograd backward produces.  Both are pure HBM traffic (read + write of
// the matrix once each); the generic strided copy kernel behind permute().contiguous() reaches well
// under 1 TB/s on this shape (T = 50 floats innermost), these tiled kernels run at HBM speed.
//
// One CTA moves a (sequence, 32 pdfs) block: in (B, P, T) that is ONE contiguous run of 32*T floats,
// in (T*B, P) it is T row segments of 128 bytes.  Two implementations: TMA copies on both sides with an
// in-shared-memory transpose (relayout_tma_kernel; needs 16-byte aligned rows / blocks and T <= 256), and
// a plain one for every other shape (relayout_kernel: the block is staged in shared memory with an odd row
// pitch, so both the contiguous side and the row-segment side are conflict-free and fully coalesced).

#include "chain_kernels.h"

#include <cuda.h>  // CUtensorMap and its enums (the encoder is fetched through cudaGetDriverEntryPoint: no -lcuda)

#include <atomic>

namespace tb {
namespace {

// ---------------------------------------------------------------------------------------------------------
// TMA version (used whenever its alignment rules hold).  One CTA moves a (sequence, 32 pdfs) block, as below,
// but the threads never touch global memory: the (B, P, T) side is ONE bulk copy of the block's contiguous
// 128*T bytes (cp.async.bulk, UBLKCP), the (T*B, P) side ONE tensor copy of a {32 pdfs, 1 sequence, T frames}
// box of the matrix seen as a 3-D tensor (cp.async.bulk.tensor.3d, UTMALDG / UTMASTG); the threads transpose
// [pdf][frame] <-> [frame][pdf] between two shared-memory images.  16 resident CTAs per SM keep ~200 KB in
// flight per SM, which is what a pure tile mover needs to approach the HBM copy rate.
// ---------------------------------------------------------------------------------------------------------
constexpr int kTmaThreads = 128;
constexpr int kLayPdfsTma = 32;

__device__ __forceinline__ unsigned lay_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  unsigned done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done)
                 : "r"(bar), "r"(parity)
                 : "memory");
}

template <bool TO2D>
__global__ void __launch_bounds__(kTmaThreads) relayout_tma_kernel(const __grid_constant__ CUtensorMap ymap, float* __restrict__ x,
                                                                  int P, int T, float scale) {
  extern __shared__ __align__(128) unsigned char lay_smem[];
  // the tensor copy wants its shared-memory image 128-byte aligned; the launch reserves the slack
  float* A = reinterpret_cast<float*>(lay_smem + ((128u - (lay_smem_u32(lay_smem) & 127u)) & 127u));  // [32][T]: as in (B, P, T)
  float* Bm = A + ((kLayPdfsTma * T + 31) & ~31);                 // [T][32]: the block as the tensor box lays it out
  __shared__ __align__(8) unsigned long long bar_storage;
  const unsigned bar = lay_smem_u32(&bar_storage);
  const int s = blockIdx.y, p0 = blockIdx.x * kLayPdfsTma;
  const int np = min(kLayPdfsTma, P - p0);
  const unsigned nbytes = (unsigned)(np * T) * 4u;
  float* chunk = x + ((int64_t)s * P + p0) * T;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (TO2D) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(nbytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(lay_smem_u32(A)),
                   "l"(chunk), "r"(nbytes), "r"(bar)
                   : "memory");
    } else {
      // a box clipped by the tensor's extent still delivers its full size (out-of-bounds elements arrive as zeros)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((unsigned)(kLayPdfsTma * T) * 4u)
                   : "memory");
      asm volatile(
          "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
              lay_smem_u32(Bm)),
          "l"(&ymap), "r"(p0), "r"(s), "r"(0), "r"(bar)
          : "memory");
    }
  }
  mbar_wait(bar, 0);
  // transpose between the two images; index o = t*32 + p runs over the tensor-side image (conflict-free), the
  // other side is read / written with stride T (two-way conflicts at even T)
  for (int o = threadIdx.x; o < kLayPdfsTma * T; o += kTmaThreads) {
    const int t = o >> 5, p = o & 31;
    if (p < np) {
      if (TO2D) Bm[o] = A[p * T + t];
      else A[p * T + t] = scale * Bm[o];
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the bulk-copy engine
  __syncthreads();
  if (threadIdx.x == 0) {
    if (TO2D) {
      asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(&ymap), "r"(p0),
                   "r"(s), "r"(0), "r"(lay_smem_u32(Bm))
                   : "memory");
    } else {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(chunk), "r"(lay_smem_u32(A)), "r"(nbytes)
                   : "memory");
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // shared memory must outlive the engine's reads
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn tensor_map_encoder() {
  static std::atomic<void*> cached{nullptr};
  static std::atomic<int> tried{0};
  if (!tried.load()) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult st;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &st) != cudaSuccess ||
        st != cudaDriverEntryPointSuccess)
      fn = nullptr;
    (void)cudaGetLastError();
    cached.store(fn);
    tried.store(1);
  }
  return reinterpret_cast<EncodeTiledFn>(cached.load());
}

// true if the TMA kernels took the call, false if the caller has to use the plain ones
template <bool TO2D>
bool launch_relayout_tma(float* x, float* y, int B, int P, int T, int64_t stride, float scale, cudaStream_t stream,
                         cudaError_t* err) {
  const int np_last = P % kLayPdfsTma;
  if (T > 256 || (stride & 3) || (((int64_t)P * T) & 3) || ((np_last * T) & 3) || B > 65535 ||
      (reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(y) & 15))
    return false;
  EncodeTiledFn encode = tensor_map_encoder();
  if (!encode) return false;
  CUtensorMap map;
  const cuuint64_t dims[3] = {(cuuint64_t)P, (cuuint64_t)B, (cuuint64_t)T};
  const cuuint64_t strides[2] = {(cuuint64_t)stride * 4, (cuuint64_t)B * (cuuint64_t)stride * 4};
  const cuuint32_t box[3] = {(cuuint32_t)kLayPdfsTma, 1, (cuuint32_t)T};
  const cuuint32_t estr[3] = {1, 1, 1};
  if (strides[1] >= (1ull << 40)) return false;
  if (encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, y, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return false;
  const size_t smem = sizeof(float) * (size_t)(((kLayPdfsTma * T + 31) & ~31) + kLayPdfsTma * T) + 128;
  auto kern = relayout_tma_kernel<TO2D>;
  *err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (*err != cudaSuccess) return true;
  kern<<<dim3((P + kLayPdfsTma - 1) / kLayPdfsTma, B), kTmaThreads, smem, stream>>>(map, x, P, T, scale);
  count_launch();
  *err = cudaGetLastError();
  return true;
}

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
  cudaError_t tma_err = cudaSuccess;
  if (launch_relayout_tma<TO2D>(x, y, B, P, T, stride, scale, stream, &tma_err)) return tma_err;
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

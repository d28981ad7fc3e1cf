// den_kernel.cu -- persistent denominator forward (alpha) and backward
// (beta + derivative) recurrences of the LF-MMI loss for sm_100a.
//
// Replaces, for the hot path only, Kaldi's DenominatorComputation
// (kaldi@22fbdd:src/chain/chain-denominator.cc: AlphaFirstFrame, AlphaDash,
// AlphaGeneralFrame, ComputeTotLogLike, BetaDashLastFrame, Beta,
// BetaDashGeneralFrame, BetaGeneralFrameDebug) and its two per-frame kernels
// in chain-kernels.cu, which the reference reaches through
// /root/reference/src/my_lib_chain.cpp:129-131.  Algorithm restated in
// SURVEY.md Appendix A.2; nothing here is derived from Kaldi's CUDA code.
//
// Design:
//  * TWO launches per pass, forward (alpha) and backward (beta + derivative); each runs all its
//    frames inside, one CTA per SM, frames separated by a software grid barrier (cooperative
//    launch guarantees co-residency).  The forward launch leaves 8 K registers per SM free so the
//    numerator kernel (num_kernel.cu) runs next to it.
//  * HMM states are dealt to CTA x warp bins by a host-side LPT schedule (balanced by arc count,
//    re-balanced from measured per-warp busy cycles); each CTA keeps its arc/task slice in
//    shared memory.  States with very many arcs are split into RED-accumulating sub-tasks.
//  * Sequences are innermost: a lane group handles one state for 4 x (lanes in the group)
//    sequences, so every gather of alpha / beta / exp(y) is one coalesced row segment.
//  * exp(y) is produced on the fly, one frame ahead, by "aux" warps that transpose frame t of
//    nnet_output into an L2-resident [pdf][seq] tile; in the backward launch the same warps
//    subtract the numerator occupancies from the occupancy tile and transpose the finished tile
//    of frame t+1 out to nnet_output_deriv (scaled, with the l2 term): one write per element.
//  * Leaky-HMM mixing is applied by the consumer (alpha' = alpha + leaky*pi*S), so a frame needs
//    ONE grid barrier; the per-frame normalisers S_t and the leaky beta sums are CTA-reduced
//    and then accumulated with fp64 atomics.
//  * Denominator occupancies are accumulated with vectorised red.global.add.f32 into the
//    L2-resident tile: no separate scatter kernel, no transposed derivative copy.

#include "chain_kernels.h"

#include <cstdio>

namespace tb {
namespace {

// ------------------------------------------------------------------ helpers
template <int V>
__device__ __forceinline__ void ld_cg(const float* ptr, float (&v)[V]) {
  if constexpr (V == 4) {
    float4 t = __ldcg(reinterpret_cast<const float4*>(ptr));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else if constexpr (V == 2) {
    float2 t = __ldcg(reinterpret_cast<const float2*>(ptr));
    v[0] = t.x; v[1] = t.y;
  } else {
    v[0] = __ldcg(ptr);
  }
}

template <int V>
__device__ __forceinline__ void ld_cs(const float* ptr, float (&v)[V]) {
  if constexpr (V == 4) {
    float4 t = __ldcs(reinterpret_cast<const float4*>(ptr));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else if constexpr (V == 2) {
    float2 t = __ldcs(reinterpret_cast<const float2*>(ptr));
    v[0] = t.x; v[1] = t.y;
  } else {
    v[0] = __ldcs(ptr);
  }
}

template <int V>
__device__ __forceinline__ void st_vec(float* ptr, const float (&v)[V]) {
  if constexpr (V == 4) {
    *reinterpret_cast<float4*>(ptr) = make_float4(v[0], v[1], v[2], v[3]);
  } else if constexpr (V == 2) {
    *reinterpret_cast<float2*>(ptr) = make_float2(v[0], v[1]);
  } else {
    *ptr = v[0];
  }
}

// Fire-and-forget fp32 reduction at L2 (no return value): RED.E.ADD.F32[.vN].
// No "memory" clobber on purpose: the targets (occupancy tile, rows of split states) are never
// read back by this thread before the next barrier, and a clobber would pin the gathers of the
// next batch behind the RED.
template <int V>
__device__ __forceinline__ void red_add(float* ptr, const float (&v)[V]) {
  if constexpr (V == 4) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(ptr), "f"(v[0]), "f"(v[1]), "f"(v[2]),
                 "f"(v[3]));
  } else if constexpr (V == 2) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(ptr), "f"(v[0]), "f"(v[1]));
  } else {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(ptr), "f"(v[0]));
  }
}

// Per-sequence normalisers (S_t, b_t: one row of BP doubles per frame) are read by EVERY main warp of EVERY CTA at the
// start of a frame.  With ld.global.cg all 2072 warps send their requests for the same 8 lines to the same L2 slices
// at the same moment: 2 MB through one or two slices = ~4 us in which the warps only wait (measured on a graph with
// almost no arcs: profiles/r02_state_cost.txt).  A plain (L1-allocating) load turns the 14 requests of a CTA into
// one.  It is coherent here: the row was completed before the grid barrier the CTA has just left (thread 0's
// ld.acquire.gpu + bar.sync order every thread's later weak loads after those writes), and nothing writes it again.
#ifndef DEN_S_CACHED
#define DEN_S_CACHED 1
#endif
__device__ __forceinline__ double ld_norm(const double* ptr) {
#if DEN_S_CACHED
  return *ptr;
#else
  return __ldcg(ptr);
#endif
}

// exp(y) rows of the frame's tile (P rows, 1.5 MB at config 2) gathered through L1 (DEN_F_E_CA / DEN_B_E_CA = 1) or past
// it (0).  The tile is complete before the barrier the CTA has just left and is not written during the frame, so a plain
// load is coherent for the same reason as ld_norm; the state rows (5 MB per frame) stay on ld.global.cg so that they do
// not evict the tile from L1.  Measured (profiles/r02_etile_l1_probe.txt): forward frame 20.7 -> 20.3 us (on), backward
// 37.4 -> 40.4 us (off: the backward pass's REDs go to the L2 lines of the occupancy tile, and its L1 has the alpha rows).
#ifndef DEN_F_E_CA
#define DEN_F_E_CA 1
#endif
#ifndef DEN_B_E_CA
#define DEN_B_E_CA 0
#endif
template <int V, bool CA>
__device__ __forceinline__ void ld_tile(const float* ptr, float (&v)[V]) {
  if constexpr (CA && V == 4) {
    const float4 t = *reinterpret_cast<const float4*>(ptr);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else {
    ld_cg<V>(ptr, v);
  }
}

// Address of row `idx` of a [rows][BP] fp32 table whose column offset is already folded into
// `base`: one IMAD.WIDE.U32 instead of the ~7 instructions of generic 64-bit index arithmetic
// (the gather loops are close to issue-bound at 3.5 warps per scheduler).
__device__ __forceinline__ const float* row_ptr(const float* base, int idx, unsigned rowbytes) {
  unsigned long long addr;
  asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(addr) : "r"(idx), "r"(rowbytes), "l"(base));
  return reinterpret_cast<const float*>(addr);
}
__device__ __forceinline__ float* row_ptr(float* base, int idx, unsigned rowbytes) {
  unsigned long long addr;
  asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(addr) : "r"(idx), "r"(rowbytes), "l"(base));
  return reinterpret_cast<float*>(addr);
}

// Packed fp32 pairs (Blackwell FFMA2 / FMUL2 / FADD2 through the PTX .f32x2 forms): the gather loops are
// sensitive to their instruction count, and four sequences per lane are two natural pairs.
#ifndef DEN_PACKED_MATH
#define DEN_PACKED_MATH 1
#endif
using u64 = unsigned long long;
__device__ __forceinline__ u64 pk2(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk2(u64 p, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p));
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
  u64 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
  u64 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
  u64 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

__device__ __forceinline__ unsigned ld_acquire(const unsigned* ptr) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
  return v;
}

// Flags written by the copy engine (host-buffer entry) are released at system scope.
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* ptr) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
  return v;
}
// Publishes a progress word to the host (mapped pinned memory): everything this thread has observed --
// including other CTAs' derivative rows, through the grid barrier it sits behind -- is visible to a host
// thread (and to copies that thread enqueues afterwards) that reads the new value.
__device__ __forceinline__ void st_release_sys(unsigned* ptr, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(ptr), "r"(v) : "memory");
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- asynchronous gather ring (LDGSTS) -------------------------------------------------------
// Register loads cap a warp at about two 512-byte rows in flight whatever the unroll depth, so the
// gather throughput of a CTA follows its WARP count (tools/gather_probe.py, profiles/r02_gather_probe.txt:
// 14 warps 12.6 TB/s ... 32 warps 19.9 TB/s), and warps are what the 112 / 128 registers per thread of
// these loops make scarce.  cp.async (LDGSTS) is not tied to the warp's scoreboards: rows land in a
// per-warp ring in shared memory (kRing arcs x 2 rows x 512 B), every lane copies and later reads back
// its OWN 16 bytes (no cross-lane dependency, no barrier), and the same probe reaches 17.1 TB/s with 12
// warps and 18.9 TB/s with 16 at four arcs in flight per warp.
// Measured on config 2 (profiles/r02_variants_async_gather.txt): SLOWER than the register gathers in the real
// kernels -- forward 27.6 vs 25.5 us per frame, backward 49.6 vs 41.4 -- although the bare gather pattern is
// faster (17 vs 12.6 TB/s at 14 warps): every row crosses the shared-memory pipe twice (LDGSTS write + LDS
// read, 18 wavefronts per arc against 2 for the register path), next to the arc records the same pipe
// serves.  Kept as a compile-time option for the record; off in the product.
#ifndef DEN_ASYNC_F
#define DEN_ASYNC_F 0
#endif
#ifndef DEN_ASYNC_B
#define DEN_ASYNC_B 0
#endif
#ifndef DEN_RING
#define DEN_RING 4
#endif
constexpr int kRing = DEN_RING;  // arcs in flight per warp; divides kStreamPad, so a stream is whole ring turns
constexpr size_t kRingBytes = (DEN_ASYNC_F || DEN_ASYNC_B) ? sizeof(float4) * 64 * kRing * kDenMainWarps : 0;
static_assert(kStreamPad % kRing == 0, "a stream is a whole number of ring turns");
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(unsigned dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void lds128(unsigned addr, float (&v)[4]) {
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(addr));
}

struct Smem {
  float* red;     // [kDenMainWarps][BP]
  float* tile;    // [kDenAuxWarps][32][33]
  float4* ring;   // [kDenMainWarps][kRing][2][32]  gather ring of the main warps
  int4* tasks;    // [max_tasks_cta]   (only if staged)
  int4* arcs;     // [max_arcs_cta]    (only if staged)
};

__device__ __forceinline__ Smem carve(unsigned char* base, const DenParams& p) {
  Smem s;
  s.red = reinterpret_cast<float*>(base);
  size_t off = sizeof(float) * kDenMainWarps * p.subs * p.BP;
  s.tile = reinterpret_cast<float*>(base + off);
  off += sizeof(float) * kDenAuxWarps * 32 * 33;
  off = (off + 127) & ~size_t(127);
  s.ring = reinterpret_cast<float4*>(base + off);
  off += kRingBytes;
  int max_tasks = max(p.dir[0].max_tasks_cta, p.dir[1].max_tasks_cta);
  s.tasks = reinterpret_cast<int4*>(base + off);
  off += sizeof(int4) * max_tasks;
  s.arcs = reinterpret_cast<int4*>(base + off);
  return s;
}

// CTA-level reduction of the main warps' per-sequence partial sums followed by
// one fp64 atomic per sequence (148-way contention per address per frame).
__device__ __forceinline__ void cta_reduce_to(const float* red, int BP, int rows, double* dst) {
  __syncthreads();
  for (int c = threadIdx.x; c < BP; c += kDenThreads) {
    float acc = 0.f;
    if (rows == kDenMainWarps) {
#pragma unroll
      for (int w = 0; w < kDenMainWarps; w++) acc += red[w * BP + c];
    } else {
      for (int w = 0; w < rows; w++) acc += red[w * BP + c];
    }
    atomicAdd(dst + c, (double)acc);
  }
}

// ------------------------------------------------- aux-warp work: staging
// Both jobs work on 32 (sequences) x 32 (pdfs) tiles, one warp per tile, with all 32
// (or 64) global loads of a tile in flight at once: the aux warps are latency-bound,
// and they must finish inside the frame time of the main warps.

// Transpose frame t of nnet_output (rows frame*Btot+seq, pdf innermost) into
// etile[pdf][seq] = exp(y) through padded shared memory.
// nnet_output is read with ld.global.cg (L2 only), never ld.global.nc: in the host-buffer entry the
// copy engine is still writing later frames of the same buffer while this launch runs, and a 32-byte
// sector can straddle two frames (B*P not a multiple of 8), so the matrix is not read-only for the
// lifetime of the kernel and a line cached in L1 could go stale.  Every element is read once per
// launch, so nothing is lost by skipping L1.
__device__ __forceinline__ void stage_exp_tile(const DenParams& p, int t, float* etile, float* tile, int i, int lane,
                                               bool want_sq, double& sumsq) {
  const int BP = p.BP;
  const int nts = BP >> 5;
  const int s0 = (i % nts) << 5, p0 = (i / nts) << 5;
  const int pdf = p0 + lane;
  const float* ycol = p.y + (int64_t)t * p.yl.t + (int64_t)(p.seq0 + s0) * p.yl.s + (int64_t)pdf * p.yl.p;
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; j++) v[j] = (s0 + j < p.B && pdf < p.P) ? __ldcg(ycol + (int64_t)j * p.yl.s) : 0.f;
#pragma unroll
  for (int j = 0; j < 32; j++) {
    if (want_sq) sumsq += (double)v[j] * (double)v[j];
    tile[j * 33 + lane] = expf(v[j]);  // padded sequences / pdfs: exp(0) = 1, never read back out
  }
  __syncwarp();
#pragma unroll 8
  for (int k = 0; k < 32; k++) {
    const int pd = p0 + k;
    if (pd < p.P) etile[(size_t)pd * BP + s0 + lane] = tile[lane * 33 + k];
  }
  __syncwarp();
}

// Transpose the finished occupancy tile of `frame` out to nnet_output_deriv:
//   deriv(row, pdf) = scale_gamma * gamma(pdf, seq) + scale_y * y(row, pdf)
// and reset the tile to zero for reuse two frames later.
__device__ __forceinline__ void flush_gamma_tile(const DenParams& p, int frame, float* gtile, float* tile, int i,
                                                 int lane, bool want_sum, double& gsum) {
  const int BP = p.BP;
  const int nts = BP >> 5;
  const int s0 = (i % nts) << 5, p0 = (i / nts) << 5;
  const int pdf = p0 + lane;
  const float* ycol = p.y + (int64_t)frame * p.yl.t + (int64_t)(p.seq0 + s0) * p.yl.s + (int64_t)pdf * p.yl.p;
  float* dcol = p.deriv + (int64_t)frame * p.dl.t + (int64_t)(p.seq0 + s0) * p.dl.s + (int64_t)pdf * p.dl.p;
  float g[32], yv[32];
#pragma unroll
  for (int k = 0; k < 32; k++) g[k] = (p0 + k < p.P) ? __ldcg(gtile + (size_t)(p0 + k) * BP + s0 + lane) : 0.f;
  const bool want_y = p.scale_y != 0.f;
#pragma unroll
  for (int j = 0; j < 32; j++)
    yv[j] = (want_y && s0 + j < p.B && pdf < p.P) ? __ldcg(ycol + (int64_t)j * p.yl.s) : 0.f;
#pragma unroll
  for (int k = 0; k < 32; k++) {
    if (p0 + k < p.P) gtile[(size_t)(p0 + k) * BP + s0 + lane] = 0.f;
    tile[k * 33 + lane] = g[k];
  }
  __syncwarp();
  // pdf-innermost rows are written whole (streaming, evict-first); with any other layout a 32 B
  // sector is completed over several frames and has to stay in L2 until then
  const bool stream_out = p.dl.p == 1;
#pragma unroll
  for (int j = 0; j < 32; j++) {
    if (s0 + j < p.B && pdf < p.P) {
      const float gg = tile[lane * 33 + j];
      const float d = fmaf(p.scale_y, yv[j], p.scale_gamma * gg);
      if (stream_out) __stcs(dcol + (int64_t)j * p.dl.s, d);
      else dcol[(int64_t)j * p.dl.s] = d;
      if (want_sum) gsum += (double)gg;
    }
  }
  __syncwarp();
}

// One frame's aux work: flush occupancies of `flush_frame` (if >= 0) and stage exp(y) of
// `stage_frame` (if >= 0), as ONE list of tile jobs dealt round-robin to all aux warps.
// forceinline on purpose: as a real call (ABI stack frame, by-reference doubles in local
// memory) the aux warps became the critical path of the backward pass (+25 us/frame).
__device__ __forceinline__ void aux_work(const DenParams& p, int flush_frame, int stage_frame, float* tile, int auxw,
                                         int lane, bool want_sq, double& sumsq, bool want_gsum, double& gsum,
                                         int num_frame = -1) {
  const int BP = p.BP;
  if (stage_frame >= 0 && p.y_ready != nullptr) {  // host-buffer entry: wait until this frame's chunk has landed
    if (lane == 0) {
      const unsigned* flag = p.y_ready + stage_frame / p.y_chunk_frames;
      while (ld_acquire_sys(flag) == 0u) {
      }
    }
    __syncwarp();
  }
  const int ntiles = (BP >> 5) * ((p.P + 31) >> 5);
  const int nflush = flush_frame >= 0 ? ntiles : 0, nstage = stage_frame >= 0 ? ntiles : 0;
  for (int j = blockIdx.x + gridDim.x * auxw; j < nflush + nstage; j += gridDim.x * kDenAuxWarps) {
    if (j < nflush)
      flush_gamma_tile(p, flush_frame, p.gtile + (size_t)(flush_frame & 1) * p.P * BP, tile, j, lane, want_gsum, gsum);
    else
      stage_exp_tile(p, stage_frame, p.etile + (size_t)(stage_frame & 1) * p.P * BP, tile, j - nflush, lane, want_sq,
                     sumsq);
  }
  if (num_frame >= 0 && p.num_frame_off != nullptr) {
    // subtract the numerator occupancies of `num_frame` from its (already zeroed) tile; the REDs
    // of the main warps add the denominator occupancies to the same tile during this frame
    float* gt = p.gtile + (size_t)(num_frame & 1) * p.P * BP;
    const int a1 = p.num_frame_off[num_frame + 1];
    for (int j = p.num_frame_off[num_frame] + (blockIdx.x + gridDim.x * auxw) * 32 + lane; j < a1;
         j += gridDim.x * kDenAuxWarps * 32) {
      const int a = p.num_frame_arcs[j];
      const int s = p.num_arc_seq[a] - p.seq0;
      if (s >= 0 && s < p.B) {
        const float v[1] = {-p.num_arc_post[a]};
        red_add<1>(gt + (size_t)p.num_arc_pdf[a] * BP + s, v);
      }
    }
  }
}

// ------------------------------------------------- main-warp work: recursions
// A warp walks its contiguous arc list as ONE stream; the last arc of every state carries the
// end-of-state flag (bit 31 of its pdf field), states without arcs get one zero-probability
// arc.  Measured on B200: round 1 (profiles/r01_summary.md) had the forward pass at 6 arcs in flight as a
// 3-stage ring of 2; round 2 moved it to 2 stages of 4 with a branch-free fast path for stages that hold
// no end of state (DEN_F_FAST below).  The backward pass is best with ONE stage of 4: every deeper
// variant (more gathers in flight next to the REDs) is slower (44.4 us -> 48..52 us/frame).
static_assert(kStreamPad % DEN_F_SA_DEFAULT == 0 && kStreamPad % DEN_B_SA_DEFAULT == 0, "stream padding");
#ifndef DEN_F_SA
#define DEN_F_SA DEN_F_SA_DEFAULT
#endif
#ifndef DEN_F_STAGES
#define DEN_F_STAGES 2
#endif
#ifndef DEN_B_SA
#define DEN_B_SA DEN_B_SA_DEFAULT
#endif
static_assert(kStreamPad % DEN_F_SA == 0 && kStreamPad % DEN_B_SA == 0, "arc streams are padded to kStreamPad arcs");
#ifndef DEN_B_STAGES
#define DEN_B_STAGES 1
#endif
// Round 2 (profiles/r02_fast_path_variants.txt): a stage of arcs WITHOUT an end of state (87 % of the stages of 4 at ~30
// arcs per state) runs as straight-line code behind one test -- the warps are bound by the latency of their own
// instruction streams (ncu: ~43 % of the stall samples are fixed-latency waits, issue slots 31 % used), so a branch per
// arc costs more than its instruction.  Forward: 2 stages x 4 arcs with the fast path 24.0 us per frame against 25.4 us
// for 3 x 2 without (25.7 us for 2 x 4 without the fast path).  Backward: the same change was SLOWER (42.8 vs 41.2 us).
#ifndef DEN_F_FAST
#define DEN_F_FAST 1
#endif
#ifndef DEN_B_FAST
#define DEN_B_FAST 0
#endif

template <int V, int SA>
struct GatherBuf {
  float a[SA][V];  // alpha_{t-1}(src) / beta'_{t+1}(dst)
  float e[SA][V];  // exp(y)(pdf)
};

// alpha_t(h,s) = (1/S_{t-1}(s)) * sum_{arcs i->h} (alpha_{t-1}(i,s) + leaky*pi(i)*S_{t-1}(s)) * p * e_{t-1}(pdf,s)
// SPLIT = true is the variant for graphs in which some state has more arcs than the split cap
// (real den graphs are heavy-tailed): such a state is dealt out as several sub-tasks (tk.z = 2 for
// the first, 1 for the others) that accumulate with REDs into a pre-zeroed row.  Everything is
// linear in the arc sums, so nothing else changes.  It is a separate instantiation because the
// mere presence of the RED path in the end-of-state block costs ~1.5 us per frame.
template <int V, int SUBS, bool SPLIT>
__device__ __forceinline__ void alpha_step(const DenParams& p, const int4* __restrict__ arcs,
                                           const int4* __restrict__ tasks, int abeg, int aend, int tb, int t,
                                           float* red, int warp, int lane) {
  const int BP = p.BP;
  const float* __restrict__ alpha_prev = p.alpha + (size_t)(t - 1) * p.N * BP;
  float* __restrict__ alpha_cur = p.alpha + (size_t)t * p.N * BP;
  const float* __restrict__ et = p.etile + (size_t)((t - 1) & 1) * p.P * BP;
  const double* Sprev = p.Ssum + (size_t)(t - 1) * BP;
  const int4* __restrict__ a = arcs + abeg;
  const int n = aend - abeg;
  const unsigned rowbytes = (unsigned)BP * 4u;
  for (int col = (lane % (32 / SUBS)) * V; col < BP; col += (32 / SUBS) * V) {
    float lamS[V], invS[V], part[V], acc[V];
    double Sd[V];
#pragma unroll
    for (int v = 0; v < V; v++) Sd[v] = ld_norm(Sprev + col + v);
    int task = tb;
    constexpr int kSA = DEN_F_SA, kStages = DEN_F_STAGES;
    GatherBuf<V, kSA> ring[kStages];
    // warp-uniform 64-bit bases + per-lane 32-bit byte offsets (arc fields are pre-scaled by the host)
    const char* const ua = reinterpret_cast<const char*>(alpha_prev + col);
    const char* const ue = reinterpret_cast<const char*>(et + col);
    // every stream is padded by the host to a multiple of kPad arcs with zero-probability
    // arcs, so a stage needs ONE bounds check instead of one per arc
    auto load = [&](GatherBuf<V, kSA>& buf, int k) {
      if (k < n) {
#pragma unroll
        for (int u = 0; u < kSA; u++) {
          const uint2 ar = *reinterpret_cast<const uint2*>(a + k + u);  // {source state, pdf} as row byte offsets
          ld_cg<V>(reinterpret_cast<const float*>(ua + ar.x), buf.a[u]);
          ld_tile<V, DEN_F_E_CA != 0>(reinterpret_cast<const float*>(ue + ar.y), buf.e[u]);
        }
      }
    };
    auto consume = [&](GatherBuf<V, kSA>& buf, int k) {
      if constexpr (DEN_F_FAST && DEN_PACKED_MATH && V == 4) {
        // a stage without an end of state (states have ~30 arcs): straight-line code, one test for the stage
        int2 rec[kSA];
        int ends = 0;
#pragma unroll
        for (int u = 0; u < kSA; u++) {
          rec[u] = *(reinterpret_cast<const int2*>(a + k + u) + 1);
          ends |= rec[u].x;
        }
        if (ends >= 0) {
#pragma unroll
          for (int u = 0; u < kSA; u++) {
            const float pr = __int_as_float(rec[u].x), pi = __int_as_float(rec[u].y);
            const u64 pr2 = pk2(pr, pr), pi2 = pk2(pi, pi);
#pragma unroll
            for (int h = 0; h < 2; h++) {
              const u64 ad = fma2(pi2, pk2(lamS[2 * h], lamS[2 * h + 1]), pk2(buf.a[u][2 * h], buf.a[u][2 * h + 1]));
              const u64 r = fma2(mul2(ad, pr2), pk2(buf.e[u][2 * h], buf.e[u][2 * h + 1]), pk2(acc[2 * h], acc[2 * h + 1]));
              upk2(r, acc[2 * h], acc[2 * h + 1]);
            }
          }
          return;
        }
      }
#pragma unroll
      for (int u = 0; u < kSA; u++) {
        {
          const int2 ar = *(reinterpret_cast<const int2*>(a + k + u) + 1);  // {prob | end flag, pi[source]}
          const float pr = fabsf(__int_as_float(ar.x)), pi = __int_as_float(ar.y);
          if constexpr (DEN_PACKED_MATH && V == 4) {
            const u64 pr2 = pk2(pr, pr), pi2 = pk2(pi, pi);
#pragma unroll
            for (int h = 0; h < 2; h++) {
              const u64 ad = fma2(pi2, pk2(lamS[2 * h], lamS[2 * h + 1]), pk2(buf.a[u][2 * h], buf.a[u][2 * h + 1]));
              const u64 r = fma2(mul2(ad, pr2), pk2(buf.e[u][2 * h], buf.e[u][2 * h + 1]), pk2(acc[2 * h], acc[2 * h + 1]));
              upk2(r, acc[2 * h], acc[2 * h + 1]);
            }
          } else {
#pragma unroll
            for (int v = 0; v < V; v++) {
              const float ad = fmaf(pi, lamS[v], buf.a[u][v]);
              acc[v] = fmaf(ad * pr, buf.e[u][v], acc[v]);
            }
          }
          if (ar.x < 0) {  // sign bit of the probability = last arc of this state (warp-uniform branch)
            const int4 tk = tasks[task++];
            float out[V];
#pragma unroll
            for (int v = 0; v < V; v++) {
              out[v] = acc[v] * invS[v];
              part[v] += out[v];
              acc[v] = 0.f;
            }
            float* const dst = row_ptr(alpha_cur + col, tk.x, rowbytes);
            if (SPLIT && tk.z != 0) red_add<V>(dst, out);
            else st_vec<V>(dst, out);
          }
        }
      }
    };
    // first gathers are in flight before the normalisers are consumed
#pragma unroll
    for (int st = 0; st < kStages; st++) load(ring[st], st * kSA);
#pragma unroll
    for (int v = 0; v < V; v++) {
      const float S = (float)Sd[v];
      lamS[v] = p.leaky * S;
      invS[v] = 1.0f / S;
      part[v] = 0.f;
      acc[v] = 0.f;
    }
    for (int k = 0; k < n;) {
#pragma unroll
      for (int st = 0; st < kStages; st++) {
        if (k < n) {
          consume(ring[st], k);
          load(ring[st], k + kStages * kSA);
          k += kSA;
        }
      }
    }
#pragma unroll
    for (int v = 0; v < V; v++) red[(warp * SUBS + lane / (32 / SUBS)) * BP + col + v] = part[v];
  }
}

// beta'_t(h,s) = (1/S_t(s)) * sum_{arcs h->j} p * e_t(pdf,s) * (beta'_{t+1}(j,s) + b_{t+1}(s))
// gamma_t(pdf,s) += that term * alpha'_t(h,s) / S_t(s)
template <int V, int SUBS, bool SPLIT>
__device__ __forceinline__ void beta_step(const DenParams& p, const int4* __restrict__ arcs,
                                          const int4* __restrict__ tasks, int abeg, int aend, int tb, int te, int t,
                                          float* red, int warp, int lane, double& ab) {
  const int BP = p.BP;
  const float* __restrict__ alpha_t = p.alpha + (size_t)t * p.N * BP;
  // beta' rotates through three buffers: written at step t, gathered at step t-1, and (for
  // split states, which accumulate with REDs) re-zeroed at step t-2
  const float* __restrict__ beta_next = p.beta + (size_t)((t + 1) % 3) * p.N * BP;
  float* __restrict__ beta_cur = p.beta + (size_t)(t % 3) * p.N * BP;
  float* __restrict__ beta_zero = p.beta + (size_t)((t + 2) % 3) * p.N * BP;
  const float* __restrict__ et = p.etile + (size_t)(t & 1) * p.P * BP;
  float* __restrict__ gt = p.gtile + (size_t)(t & 1) * p.P * BP;
  const double* St = p.Ssum + (size_t)t * BP;
  const double* Bn = p.Bsum + (size_t)(t + 1) * BP;
  const int4* __restrict__ a = arcs + abeg;
  const int n = aend - abeg;
  const unsigned rowbytes = (unsigned)BP * 4u;
  for (int col = (lane % (32 / SUBS)) * V; col < BP; col += (32 / SUBS) * V) {
    float lamS[V], invS[V], bv[V], part[V], X[V], occ[V], aln[V], al0[V];
    double Sd[V], Bd[V];
#pragma unroll
    for (int v = 0; v < V; v++) {
      Sd[v] = ld_norm(St + col + v);
      Bd[v] = ld_norm(Bn + col + v);
      aln[v] = 0.f;
      al0[v] = 0.f;
    }
    int task = tb;
    float pi_h = 0.f;
    constexpr int kSA = DEN_B_SA, kStages = DEN_B_STAGES;
    GatherBuf<V, kSA> ring[kStages];
    const char* const ub = reinterpret_cast<const char*>(beta_next + col);
    const char* const ue = reinterpret_cast<const char*>(et + col);
    char* const ug = reinterpret_cast<char*>(gt + col);
    auto load = [&](GatherBuf<V, kSA>& buf, int k) {
      if (k < n) {  // streams are padded to a multiple of kPad arcs (see alpha_step)
#pragma unroll
        for (int u = 0; u < kSA; u++) {
          const uint2 ar = *reinterpret_cast<const uint2*>(a + k + u);  // {destination state, pdf} as row byte offsets
          ld_cg<V>(reinterpret_cast<const float*>(ub + ar.x), buf.a[u]);
          ld_tile<V, DEN_B_E_CA != 0>(reinterpret_cast<const float*>(ue + ar.y), buf.e[u]);
        }
      }
    };
    auto consume = [&](GatherBuf<V, kSA>& buf, int k) {
      if constexpr (DEN_B_FAST && DEN_PACKED_MATH && V == 4) {
        int2 rec[kSA];  // {pdf row offset, prob | end flag}
        int ends = 0;
#pragma unroll
        for (int u = 0; u < kSA; u++) {
          const int4 r = a[k + u];
          rec[u] = make_int2(r.y, r.z);
          ends |= r.z;
        }
        if (ends >= 0) {  // no end of state in this stage: straight-line code
#pragma unroll
          for (int u = 0; u < kSA; u++) {
            const float pr = __int_as_float(rec[u].y);
            const u64 pr2 = pk2(pr, pr);
            float g[V];
#pragma unroll
            for (int h = 0; h < 2; h++) {
              const u64 sum = add2(pk2(buf.a[u][2 * h], buf.a[u][2 * h + 1]), pk2(bv[2 * h], bv[2 * h + 1]));
              const u64 vf = mul2(mul2(pr2, sum), pk2(buf.e[u][2 * h], buf.e[u][2 * h + 1]));
              upk2(add2(pk2(X[2 * h], X[2 * h + 1]), vf), X[2 * h], X[2 * h + 1]);
              upk2(mul2(vf, pk2(occ[2 * h], occ[2 * h + 1])), g[2 * h], g[2 * h + 1]);
            }
            red_add<V>(reinterpret_cast<float*>(ug + (unsigned)rec[u].x), g);
          }
          return;
        }
      }
#pragma unroll
      for (int u = 0; u < kSA; u++) {
        {
          const int4 ar = a[k + u];
          const float pr = fabsf(__int_as_float(ar.z));
          float g[V];
          if constexpr (DEN_PACKED_MATH && V == 4) {
            const u64 pr2 = pk2(pr, pr);
#pragma unroll
            for (int h = 0; h < 2; h++) {
              const u64 sum = add2(pk2(buf.a[u][2 * h], buf.a[u][2 * h + 1]), pk2(bv[2 * h], bv[2 * h + 1]));
              const u64 vf = mul2(mul2(pr2, sum), pk2(buf.e[u][2 * h], buf.e[u][2 * h + 1]));
              upk2(add2(pk2(X[2 * h], X[2 * h + 1]), vf), X[2 * h], X[2 * h + 1]);
              upk2(mul2(vf, pk2(occ[2 * h], occ[2 * h + 1])), g[2 * h], g[2 * h + 1]);
            }
          } else {
#pragma unroll
            for (int v = 0; v < V; v++) {
              const float vf = pr * (buf.a[u][v] + bv[v]) * buf.e[u][v];
              X[v] += vf;
              g[v] = vf * occ[v];
            }
          }
          red_add<V>(reinterpret_cast<float*>(ug + (unsigned)ar.y), g);
          if (ar.z < 0) {  // sign bit of the probability = last arc of this state
            const int4 tk = tasks[task];
            float bd[V];
#pragma unroll
            for (int v = 0; v < V; v++) {
              bd[v] = X[v] * invS[v];
              part[v] = fmaf(pi_h, bd[v], part[v]);
              if (t == 0 && col + v < p.B) ab += (double)(occ[v] * (lamS[v] / p.leaky)) * (double)bd[v];
              X[v] = 0.f;
            }
            float* const dst = row_ptr(beta_cur + col, tk.x, rowbytes);
            if (SPLIT && tk.z != 0) {
              red_add<V>(dst, bd);
              if (tk.z == 2) {  // the state's first sub-task also clears the row used at step t-1
                float zero[V];
#pragma unroll
                for (int v = 0; v < V; v++) zero[v] = 0.f;
                st_vec<V>(row_ptr(beta_zero + col, tk.x, rowbytes), zero);
              }
            } else {
              st_vec<V>(dst, bd);
            }
            task++;
            if (task < te) {
              pi_h = __int_as_float(tasks[task].w);
#pragma unroll
              for (int v = 0; v < V; v++) occ[v] = fmaf(pi_h, lamS[v], aln[v]) * invS[v];
              if (task + 1 < te) ld_cs<V>(row_ptr(alpha_t + col, tasks[task + 1].x, rowbytes), aln);
            }
          }
        }
      }
    };
    // everything the first arcs need is requested before anything is consumed
    if (tb < te) {
      const int4 tk0 = tasks[tb];
      pi_h = __int_as_float(tk0.w);
      ld_cs<V>(row_ptr(alpha_t + col, tk0.x, rowbytes), al0);
    }
#pragma unroll
    for (int st = 0; st < kStages; st++) load(ring[st], st * kSA);
    if (tb + 1 < te) ld_cs<V>(row_ptr(alpha_t + col, tasks[tb + 1].x, rowbytes), aln);
#pragma unroll
    for (int v = 0; v < V; v++) {
      const float S = (float)Sd[v];
      lamS[v] = p.leaky * S;
      invS[v] = 1.0f / S;
      bv[v] = p.leaky * (float)Bd[v];
      part[v] = 0.f;
      X[v] = 0.f;
      occ[v] = fmaf(pi_h, lamS[v], al0[v]) * invS[v];
    }
    for (int k = 0; k < n;) {
#pragma unroll
      for (int st = 0; st < kStages; st++) {
        if (k < n) {
          consume(ring[st], k);
          load(ring[st], k + kStages * kSA);
          k += kSA;
        }
      }
    }
#pragma unroll
    for (int v = 0; v < V; v++) red[(warp * SUBS + lane / (32 / SUBS)) * BP + col + v] = part[v];
  }
}


// ---- the same recurrence with the gathers going through the shared-memory ring

template <int SUBS, bool SPLIT>
__device__ __forceinline__ void alpha_step_async(const DenParams& p, const int4* __restrict__ arcs,
                                                 const int4* __restrict__ tasks, int abeg, int aend, int tb, int t,
                                                 float* red, float4* ring, int warp, int lane) {
  constexpr int V = 4;
  const int BP = p.BP;  // = (32 / SUBS) * V: a lane group covers the whole (padded) batch
  const int col = (lane % (32 / SUBS)) * V;
  const float* __restrict__ alpha_prev = p.alpha + (size_t)(t - 1) * p.N * BP;
  float* __restrict__ alpha_cur = p.alpha + (size_t)t * p.N * BP;
  const float* __restrict__ et = p.etile + (size_t)((t - 1) & 1) * p.P * BP;
  const double* Sprev = p.Ssum + (size_t)(t - 1) * BP;
  const int4* __restrict__ a = arcs + abeg;
  const int n = aend - abeg;  // a multiple of kStreamPad
  const unsigned rowbytes = (unsigned)BP * 4u;
  const char* const ua = reinterpret_cast<const char*>(alpha_prev + col);
  const char* const ue = reinterpret_cast<const char*>(et + col);
  const unsigned slot0 = smem_u32(ring + lane);  // this lane's 16 bytes of slot 0's first row
  // arc k -> slot k % kRing: state row at +0, exp row at +512; exactly ONE group is committed per arc
  // (an empty one past the end of the stream), so "at most kRing-1 groups pending" == "arc k has landed"
  auto issue = [&](int k, int u) {
    if (k < n) {
      const uint2 ar = *reinterpret_cast<const uint2*>(a + k);  // {source state, pdf} as row byte offsets
      cp_async16(slot0 + u * 1024, ua + ar.x);
      cp_async16(slot0 + u * 1024 + 512, ue + ar.y);
    }
    cp_async_commit();
  };
#pragma unroll
  for (int u = 0; u < kRing; u++) issue(u, u);
  float lamS[V], invS[V], part[V], acc[V];
#pragma unroll
  for (int v = 0; v < V; v++) {
    const float S = (float)__ldcg(Sprev + col + v);
    lamS[v] = p.leaky * S;
    invS[v] = 1.0f / S;
    part[v] = 0.f;
    acc[v] = 0.f;
  }
  int task = tb;
  for (int k = 0; k < n; k += kRing) {
#pragma unroll
    for (int u = 0; u < kRing; u++) {
      cp_async_wait<kRing - 1>();
      float av[V], ev[V];
      lds128(slot0 + u * 1024, av);
      lds128(slot0 + u * 1024 + 512, ev);
      const int2 ar = *(reinterpret_cast<const int2*>(a + k + u) + 1);  // {prob | end flag, pi[source]}
      issue(k + u + kRing, u);  // the slot is free again: this lane has read its 16 bytes of both rows
      const float pr = fabsf(__int_as_float(ar.x)), pi = __int_as_float(ar.y);
      const u64 pr2 = pk2(pr, pr), pi2 = pk2(pi, pi);
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const u64 ad = fma2(pi2, pk2(lamS[2 * h], lamS[2 * h + 1]), pk2(av[2 * h], av[2 * h + 1]));
        const u64 r = fma2(mul2(ad, pr2), pk2(ev[2 * h], ev[2 * h + 1]), pk2(acc[2 * h], acc[2 * h + 1]));
        upk2(r, acc[2 * h], acc[2 * h + 1]);
      }
      if (ar.x < 0) {  // sign bit of the probability = last arc of this state (uniform within the lane group)
        const int4 tk = tasks[task++];
        float out[V];
#pragma unroll
        for (int v = 0; v < V; v++) {
          out[v] = acc[v] * invS[v];
          part[v] += out[v];
          acc[v] = 0.f;
        }
        float* const dst = row_ptr(alpha_cur + col, tk.x, rowbytes);
        if (SPLIT && tk.z != 0) red_add<V>(dst, out);
        else st_vec<V>(dst, out);
      }
    }
  }
  cp_async_wait<0>();
#pragma unroll
  for (int v = 0; v < V; v++) red[(warp * SUBS + lane / (32 / SUBS)) * BP + col + v] = part[v];
}

template <int SUBS, bool SPLIT>
__device__ __forceinline__ void beta_step_async(const DenParams& p, const int4* __restrict__ arcs,
                                                const int4* __restrict__ tasks, int abeg, int aend, int tb, int te,
                                                int t, float* red, float4* ring, int warp, int lane, double& ab) {
  constexpr int V = 4;
  const int BP = p.BP;
  const int col = (lane % (32 / SUBS)) * V;
  const float* __restrict__ alpha_t = p.alpha + (size_t)t * p.N * BP;
  const float* __restrict__ beta_next = p.beta + (size_t)((t + 1) % 3) * p.N * BP;
  float* __restrict__ beta_cur = p.beta + (size_t)(t % 3) * p.N * BP;
  float* __restrict__ beta_zero = p.beta + (size_t)((t + 2) % 3) * p.N * BP;
  const float* __restrict__ et = p.etile + (size_t)(t & 1) * p.P * BP;
  float* __restrict__ gt = p.gtile + (size_t)(t & 1) * p.P * BP;
  const double* St = p.Ssum + (size_t)t * BP;
  const double* Bn = p.Bsum + (size_t)(t + 1) * BP;
  const int4* __restrict__ a = arcs + abeg;
  const int n = aend - abeg;
  const unsigned rowbytes = (unsigned)BP * 4u;
  const char* const ub = reinterpret_cast<const char*>(beta_next + col);
  const char* const ue = reinterpret_cast<const char*>(et + col);
  char* const ug = reinterpret_cast<char*>(gt + col);
  const unsigned slot0 = smem_u32(ring + lane);
  auto issue = [&](int k, int u) {
    if (k < n) {
      const uint2 ar = *reinterpret_cast<const uint2*>(a + k);  // {destination state, pdf} as row byte offsets
      cp_async16(slot0 + u * 1024, ub + ar.x);
      cp_async16(slot0 + u * 1024 + 512, ue + ar.y);
    }
    cp_async_commit();
  };
  float al0[V], aln[V];
#pragma unroll
  for (int v = 0; v < V; v++) al0[v] = aln[v] = 0.f;
  int task = tb;
  float pi_h = 0.f;
  if (tb < te) {
    const int4 tk0 = tasks[tb];
    pi_h = __int_as_float(tk0.w);
    ld_cs<V>(row_ptr(alpha_t + col, tk0.x, rowbytes), al0);
  }
#pragma unroll
  for (int u = 0; u < kRing; u++) issue(u, u);
  if (tb + 1 < te) ld_cs<V>(row_ptr(alpha_t + col, tasks[tb + 1].x, rowbytes), aln);
  float lamS[V], invS[V], bv[V], part[V], X[V], occ[V];
#pragma unroll
  for (int v = 0; v < V; v++) {
    const float S = (float)__ldcg(St + col + v);
    lamS[v] = p.leaky * S;
    invS[v] = 1.0f / S;
    bv[v] = p.leaky * (float)ld_norm(Bn + col + v);
    part[v] = 0.f;
    X[v] = 0.f;
    occ[v] = fmaf(pi_h, lamS[v], al0[v]) * invS[v];
  }
  for (int k = 0; k < n; k += kRing) {
#pragma unroll
    for (int u = 0; u < kRing; u++) {
      cp_async_wait<kRing - 1>();
      float bn[V], ev[V], g[V];
      lds128(slot0 + u * 1024, bn);
      lds128(slot0 + u * 1024 + 512, ev);
      const int4 ar = a[k + u];
      issue(k + u + kRing, u);
      const float pr = fabsf(__int_as_float(ar.z));
      const u64 pr2 = pk2(pr, pr);
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const u64 sum = add2(pk2(bn[2 * h], bn[2 * h + 1]), pk2(bv[2 * h], bv[2 * h + 1]));
        const u64 vf = mul2(mul2(pr2, sum), pk2(ev[2 * h], ev[2 * h + 1]));
        upk2(add2(pk2(X[2 * h], X[2 * h + 1]), vf), X[2 * h], X[2 * h + 1]);
        upk2(mul2(vf, pk2(occ[2 * h], occ[2 * h + 1])), g[2 * h], g[2 * h + 1]);
      }
      red_add<V>(reinterpret_cast<float*>(ug + (unsigned)ar.y), g);
      if (ar.z < 0) {  // sign bit of the probability = last arc of this state
        const int4 tk = tasks[task];
        float bd[V];
#pragma unroll
        for (int v = 0; v < V; v++) {
          bd[v] = X[v] * invS[v];
          part[v] = fmaf(pi_h, bd[v], part[v]);
          if (t == 0 && col + v < p.B) ab += (double)(occ[v] * (lamS[v] / p.leaky)) * (double)bd[v];
          X[v] = 0.f;
        }
        float* const dst = row_ptr(beta_cur + col, tk.x, rowbytes);
        if (SPLIT && tk.z != 0) {
          red_add<V>(dst, bd);
          if (tk.z == 2) {  // the state's first sub-task also clears the row used at step t-1
            float zero[V];
#pragma unroll
            for (int v = 0; v < V; v++) zero[v] = 0.f;
            st_vec<V>(row_ptr(beta_zero + col, tk.x, rowbytes), zero);
          }
        } else {
          st_vec<V>(dst, bd);
        }
        task++;
        if (task < te) {
          pi_h = __int_as_float(tasks[task].w);
#pragma unroll
          for (int v = 0; v < V; v++) occ[v] = fmaf(pi_h, lamS[v], aln[v]) * invS[v];
          if (task + 1 < te) ld_cs<V>(row_ptr(alpha_t + col, tasks[task + 1].x, rowbytes), aln);
        }
      }
    }
  }
  cp_async_wait<0>();
#pragma unroll
  for (int v = 0; v < V; v++) red[(warp * SUBS + lane / (32 / SUBS)) * BP + col + v] = part[v];
}

template <bool STAGED, int SUBS>
__device__ __forceinline__ void load_dir(const DenParams& p, int d, const Smem& sm, const int4*& arcs,
                                         const int4*& tasks, int& tb, int& te, int& abeg, int& aend, int warp,
                                         int lane) {
  const DenDir& D = p.dir[d];
  const int task0 = D.cta_task_off[blockIdx.x], ntask = D.cta_task_off[blockIdx.x + 1] - task0;
  const int arc0 = D.cta_arc_off[blockIdx.x], narc = D.cta_arc_off[blockIdx.x + 1] - arc0;
  tb = te = abeg = aend = 0;
  if (warp < kDenMainWarps) {  // one bin per lane group
    const int bin = (blockIdx.x * kDenMainWarps + warp) * SUBS + lane / (32 / SUBS);
    tb = D.warp_task_off[bin] - task0;
    te = D.warp_task_off[bin + 1] - task0;
    abeg = D.warp_arc_off[bin] - arc0;
    aend = D.warp_arc_off[bin + 1] - arc0;
  }
  if constexpr (STAGED) {
    for (int i = threadIdx.x; i < ntask; i += kDenThreads) sm.tasks[i] = D.tasks[task0 + i];
    for (int i = threadIdx.x; i < narc; i += kDenThreads) sm.arcs[i] = D.arcs[arc0 + i];
    arcs = sm.arcs;
    tasks = sm.tasks;
    __syncthreads();
  } else {
    arcs = D.arcs + arc0;
    tasks = D.tasks + task0;
  }
}

// ------------------------------------------------------------------ kernel
// PHASE 0: forward launch (prologue + T alpha frames + total log-likelihood).
// PHASE 1: backward launch (beta prologue + T beta/derivative frames + last flush).
// SUBS = 1: a warp handles one state for 128 sequences at a time (BP = 128).  For smaller
// batches the 32 lanes are split into SUBS = 2 or 4 groups of 16 / 8 lanes (BP = 64 / 32), each
// walking its OWN arc stream with float4 vectors: the same instruction stream then covers SUBS
// states at once, which is what small batches need (they are bound by per-warp instruction
// latency, not by bandwidth).  All per-warp quantities below are simply per-lane-group values.
//
// NCTX = 2: the launch carries TWO independent sequence groups ("contexts", 128 sequences each,
// own alpha/beta/tile/sum buffers and own barrier counter) and alternates between them frame by
// frame: a CTA arrives at the barrier of context A's frame t, works on context B's frame t, and
// only then waits for A -- barrier latency and the other CTAs' finish-time spread are hidden
// behind useful work.  With NCTX = 1 arrive and wait are back to back (a plain grid barrier).
template <int SUBS, bool STAGED, int PHASE, bool SPLIT, int NCTX>
__device__ __forceinline__ void den_chain_body(const DenBatch& batch) {
  constexpr int V = 4;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const DenParams& p0 = batch.c[0];  // geometry, graph and options are common to all contexts
  const Smem sm = carve(smem_raw, p0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool is_aux = warp >= kDenMainWarps;
  const int auxw = warp - kDenMainWarps;
  float* tile = sm.tile + (is_aux ? auxw : 0) * 32 * 33;
  const int BP = p0.BP, T = p0.T;
  unsigned epoch[NCTX];
#pragma unroll
  for (int ci = 0; ci < NCTX; ci++) epoch[ci] = 0;
  double sumsq = 0.0, ab = 0.0, gsum = 0.0;
  long long busy = 0;  // SM cycles this warp spent on its per-frame work (lane 0 only)

  const int4* arcs;
  const int4* tasks;
  int tb, te, abeg, aend;
  unsigned long long* tl = (p0.timeline && blockIdx.x == 0 && threadIdx.x == 0) ? p0.timeline : nullptr;
  int tli = 0;
  auto stamp = [&]() {
    if (tl) {
      unsigned long long ns;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
      tl[tli++] = ns;
    }
  };
  // debug trace (torchain_profile_enable(2), single context only): slots = earliest / latest
  // main-warp finish, latest aux-warp finish, grid-barrier exit; combined per CTA in shared memory
  __shared__ unsigned long long s_trace[3];
  auto trace_warp_done = [&](int frame) {
    if (NCTX == 1 && p0.cta_trace && lane == 0) {
      unsigned long long ns;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
      // per-warp finish times live after the [2][T+1][grid][4] block
      p0.cta_trace[(size_t)2 * (T + 1) * gridDim.x * 4 +
                   (((size_t)PHASE * (T + 1) + frame) * gridDim.x + blockIdx.x) * kDenWarps + warp] = ns;
      if (is_aux) {
        atomicMax(&s_trace[2], ns);
      } else {
        atomicMin(&s_trace[0], ns);
        atomicMax(&s_trace[1], ns);
      }
    }
  };
  auto trace_frame_end = [&](int frame) {  // thread 0, after the grid barrier of `frame`
    if (NCTX == 1 && p0.cta_trace && threadIdx.x == 0) {
      unsigned long long ns;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
      unsigned long long* out = p0.cta_trace + (((size_t)PHASE * (T + 1) + frame) * gridDim.x + blockIdx.x) * 4;
      out[0] = s_trace[0];
      out[1] = s_trace[1];
      out[2] = s_trace[2];
      out[3] = ns;
      s_trace[0] = ~0ull;
      s_trace[1] = 0ull;
      s_trace[2] = 0ull;
    }
  };
  if (threadIdx.x == 0) {
    s_trace[0] = ~0ull;
    s_trace[1] = 0ull;
    s_trace[2] = 0ull;
  }
  // split grid barrier: arrive() publishes this CTA's work on context ci, wait() blocks until
  // every CTA has arrived there
  auto arrive = [&](int ci) {
    __syncthreads();
    if (threadIdx.x == 0) {
      epoch[ci] += gridDim.x;
      __threadfence();  // release everything this CTA wrote (cumulative over the bar.sync above)
      atomicAdd(batch.c[ci].barrier + 32 * PHASE, 1u);
    }
  };
  auto wait = [&](int ci) {
    if (threadIdx.x == 0) {
      const unsigned* ctr = batch.c[ci].barrier + 32 * PHASE;
      while (ld_acquire(ctr) < epoch[ci]) {
      }
    }
    __syncthreads();
  };
  if (tl) tl += PHASE * (T + 3);
  stamp();
  load_dir<STAGED, SUBS>(p0, PHASE, sm, arcs, tasks, tb, te, abeg, aend, warp, lane);

  if constexpr (PHASE == 0) {
    // ---- prologue: alpha_0 = pi, S_0; stage exp(y_0)
#pragma unroll
    for (int ci = 0; ci < NCTX; ci++) {
      const DenParams& p = batch.c[ci];
      if (is_aux) {
        aux_work(p, -1, 0, tile, auxw, lane, true, sumsq, false, gsum);
      } else {
        for (int col = (lane % (32 / SUBS)) * V; col < BP; col += (32 / SUBS) * V) {
          float part = 0.f;
          for (int task = tb; task < te; task++) {
            const int4 tk = tasks[task];
            const float pi_h = __int_as_float(tk.w);
            if (SPLIT && tk.z == 1) continue;  // secondary sub-task of a split state: its first one does the set-up
            float out[V];
#pragma unroll
            for (int v = 0; v < V; v++) out[v] = pi_h;
            st_vec<V>(p.alpha + tk.x * BP + col, out);
            part += pi_h;
            if (SPLIT && tk.z == 2) {  // rows of frames 1..T are accumulated with REDs: clear them
#pragma unroll
              for (int v = 0; v < V; v++) out[v] = 0.f;
              for (int tt = 1; tt <= T; tt++) st_vec<V>(p.alpha + ((size_t)tt * p.N + tk.x) * BP + col, out);
            }
          }
#pragma unroll
          for (int v = 0; v < V; v++) sm.red[(warp * SUBS + lane / (32 / SUBS)) * BP + col + v] = part;
        }
      }
      cta_reduce_to(sm.red, BP, kDenMainWarps * SUBS, p.Ssum);
      arrive(ci);
    }

    // ---- forward
    for (int t = 1; t <= T; t++) {
#pragma unroll
      for (int ci = 0; ci < NCTX; ci++) {
        const DenParams& p = batch.c[ci];
        wait(ci);  // frame t-1 of this context is complete everywhere
        if (ci == 0) {
          if (t > 1) trace_frame_end(t - 1);
          stamp();
        }
        if (is_aux) {
          if (t < T) aux_work(p, -1, t, tile, auxw, lane, true, sumsq, false, gsum);
        } else {
          const long long c0 = clock64();
          if constexpr (DEN_ASYNC_F)
            alpha_step_async<SUBS, SPLIT>(p, arcs, tasks, abeg, aend, tb, t, sm.red, sm.ring + (size_t)warp * kRing * 64,
                                          warp, lane);
          else
            alpha_step<V, SUBS, SPLIT>(p, arcs, tasks, abeg, aend, tb, t, sm.red, warp, lane);
          busy += clock64() - c0;
        }
        trace_warp_done(t);
        cta_reduce_to(sm.red, BP, kDenMainWarps * SUBS, p.Ssum + (size_t)t * BP);
        arrive(ci);
      }
    }
#pragma unroll
    for (int ci = 0; ci < NCTX; ci++) wait(ci);
    trace_frame_end(T);
    stamp();
    if (is_aux) {
      sumsq = warp_sum(sumsq);
      if (lane == 0 && sumsq != 0.0) atomicAdd(p0.scalars + 3, sumsq);
    }
    if (blockIdx.x == 0) {
      // ComputeTotLogLike: sum_s log(sum_h alpha'_T(h,s)) + sum_{t<T} sum_s log S_t(s)
      double acc = 0.0;
#pragma unroll
      for (int ci = 0; ci < NCTX; ci++) {
        const DenParams& p = batch.c[ci];
        for (int i = threadIdx.x; i < (T + 1) * BP; i += kDenThreads) {
          const int t = i / BP, s = i - t * BP;
          if (s < p.B) {
            float S = (float)__ldcg(p.Ssum + i);
            if (t == T) S = S * (1.0f + p.leaky * p.sum_pi);
            acc += log((double)S);
          }
        }
      }
      acc = warp_sum(acc);
      if (lane == 0) atomicAdd(p0.scalars + 2, acc);
    }
  } else {
    // ---- backward prologue: beta'_T(h,s) = 1 / sum_h alpha'_T(h,s) = 1 / (S_T (1 + leaky*sum_pi))
#pragma unroll
    for (int ci = 0; ci < NCTX; ci++) {
      const DenParams& p = batch.c[ci];
      if (!is_aux) {
        const double* ST = p.Ssum + (size_t)T * BP;
        float* betaT = p.beta + (size_t)(T % 3) * p.N * BP;
        float* betaTm1 = p.beta + (size_t)((T + 2) % 3) * p.N * BP;  // buffer of step T-1
        for (int col = (lane % (32 / SUBS)) * V; col < BP; col += (32 / SUBS) * V) {
          float inv[V], part[V];
#pragma unroll
          for (int v = 0; v < V; v++) {
            const float S = (float)__ldcg(ST + col + v);
            inv[v] = 1.0f / (S * (1.0f + p.leaky * p.sum_pi));
            part[v] = 0.f;
          }
          for (int task = tb; task < te; task++) {
            const int4 tk = tasks[task];
            const float pi_h = __int_as_float(tk.w);
            if (SPLIT && tk.z == 1) continue;
            st_vec<V>(betaT + tk.x * BP + col, inv);
#pragma unroll
            for (int v = 0; v < V; v++) part[v] = fmaf(pi_h, inv[v], part[v]);
            if (SPLIT && tk.z == 2) {
              float zero[V];
#pragma unroll
              for (int v = 0; v < V; v++) zero[v] = 0.f;
              st_vec<V>(betaTm1 + tk.x * BP + col, zero);
            }
          }
#pragma unroll
          for (int v = 0; v < V; v++) sm.red[(warp * SUBS + lane / (32 / SUBS)) * BP + col + v] = part[v];
        }
      }
      cta_reduce_to(sm.red, BP, kDenMainWarps * SUBS, p.Bsum + (size_t)T * BP);
      arrive(ci);
    }

    // ---- backward
    for (int t = T - 1; t >= 0; t--) {
#pragma unroll
      for (int ci = 0; ci < NCTX; ci++) {
        const DenParams& p = batch.c[ci];
        wait(ci);  // step t+1 of this context is complete everywhere
        if (ci == 0) {
          if (t < T - 1) trace_frame_end(t + 1);
          stamp();
        }
        if (ci == NCTX - 1 && t < T - 1 && p0.deriv_progress && blockIdx.x == 0 && warp == kDenMainWarps && lane == 0 &&
            ((p0.progress_mask[((T - 2 - t) >> 6) & 7] >> ((T - 2 - t) & 63)) & 1ull)) {
          // The derivative rows of frames > t+1 are complete in every context: every CTA wrote them before
          // it arrived at the barrier thread 0 of this CTA has just acquired, and the bar.sync inside wait()
          // orders that acquire before this store.  The release is at SYSTEM scope because the reader is
          // the host (which then enqueues copy-engine reads of those rows).  It is issued by an aux-warp
          // lane: on thread 0 -- the barrier thread -- the fence in front of the store held every frame
          // back by the PCIe round trip (65 us per call, profiles/r01_host_trace.txt).  The aux warps have less
          // slack than round 1 assumed: published every frame it cost 83 us per call, and behind the barrier
          // arrival 230 us (profiles/r02_host_trace_release_after_arrive.txt), so it is published only where the
          // host has a chunk boundary (progress_mask: ~16 times per call).
          st_release_sys(p0.deriv_progress, (unsigned)(T - 2 - t));
        }
        if (is_aux) {
#ifndef DEN_DEBUG_SKIP_AUX_B  // timing experiment only (wrong results): how much of a backward frame is the aux warps'?
          aux_work(p, t + 1 < T ? t + 1 : -1, t - 1, tile, auxw, lane, false, sumsq, false, gsum, t);
#endif
        } else {
          const long long c0 = clock64();
          if constexpr (DEN_ASYNC_B)
            beta_step_async<SUBS, SPLIT>(p, arcs, tasks, abeg, aend, tb, te, t, sm.red,
                                         sm.ring + (size_t)warp * kRing * 64, warp, lane, ab);
          else
            beta_step<V, SUBS, SPLIT>(p, arcs, tasks, abeg, aend, tb, te, t, sm.red, warp, lane, ab);
          busy += clock64() - c0;
        }
        trace_warp_done(t);
        cta_reduce_to(sm.red, BP, kDenMainWarps * SUBS, p.Bsum + (size_t)t * BP);
        arrive(ci);
      }
    }
#pragma unroll
    for (int ci = 0; ci < NCTX; ci++) wait(ci);
    trace_frame_end(0);
    stamp();
    // step 0 is complete everywhere: only frame 0 is still to be flushed (below), so the host may copy frame 1 out
    if (p0.deriv_progress && blockIdx.x == 0 && warp == kDenMainWarps && lane == 0)
      st_release_sys(p0.deriv_progress, (unsigned)(T - 1));

    // ---- epilogue: frame 0's derivative rows, sanity sums
    if (is_aux) {
#pragma unroll
      for (int ci = 0; ci < NCTX; ci++) aux_work(batch.c[ci], 0, -1, tile, auxw, lane, false, sumsq, true, gsum);
      gsum = warp_sum(gsum);
      if (lane == 0 && gsum != 0.0) atomicAdd(p0.scalars + 1, gsum);
    } else {
      ab = warp_sum(ab);
      if (lane == 0 && ab != 0.0) atomicAdd(p0.scalars + 0, ab);
    }
  }  // PHASE
  if (p0.warp_busy && lane == 0)
    p0.warp_busy[((size_t)PHASE * gridDim.x + blockIdx.x) * kDenWarps + warp] = (unsigned long long)busy;
  __syncthreads();
  stamp();
}

#ifndef DEN_F_MAXNREG_TIED
#define DEN_F_MAXNREG_TIED 112
#endif
// ONE stage of 4 records: 21.7 us per frame against 24.4 for 3 stages of 2 and 22.0 for 2 of 4
// (profiles/r02_tied_sa_probe.txt) -- the loop is bound by its own instruction stream, not by loads in flight
// phase A: ONE run in flight per warp.  Two cost the whole backward kernel 2 us per frame, four 15 us
// (profiles/r02_tied_phaseA_probe.txt): the kernel lives at its 128-register cap, and what phase A keeps in
// registers is spilled around phase B's loop.
#ifndef DEN_TIED_A_DEPTH
#define DEN_TIED_A_DEPTH 1
#endif
#ifndef DEN_TIED_B_SA
#define DEN_TIED_B_SA 4
#endif
#ifndef DEN_TIED_B_STAGES
#define DEN_TIED_B_STAGES 1  // one stage of 4: 36.8 us per backward step against 39.7 for two (profiles/r02_tied_phaseB_probe.txt)
#endif
static_assert(kStreamPad % DEN_TIED_B_SA == 0, "record streams are padded to kStreamPad");
#ifndef DEN_TIED_F_SA
#define DEN_TIED_F_SA 4
#endif
#ifndef DEN_TIED_F_STAGES
#define DEN_TIED_F_STAGES 1
#endif
static_assert(kStreamPad % DEN_TIED_F_SA == 0, "record streams are padded to kStreamPad");

// =====================================================================================================
// Tied-pdf ("run-based") recurrences.
//
// In a real den.fst the pdf on an arc belongs to the HMM state being entered, so the in-arcs of a state
// carry very few distinct pdfs: ~2 RUNS = (destination state, pdf) pairs per state for ~30 in-arcs.
// Grouping by run takes exp(y) and the occupancy RED out of the per-arc work:
//   forward   a_{t+1}(r) = e_t(pdf_r) / S_t * sum_{arcs i->j in r} alpha'_t(i) p     (one exp row per RUN)
//             alpha_{t+1}(j) = sum_{runs of j} a_{t+1}(r);  a_{t+1}(r) is kept (T x R x BP, instead of the
//             alpha store: the backward pass needs nothing else from the forward pass)
//   backward  phase A, per run:   full = beta'_{t+1}(j_r) + b_{t+1};  gamma_t(pdf_r) += a_{t+1}(r) * full  (one RED per RUN)
//                                 W_t(r) = full * e_t(pdf_r)
//             phase B, per arc:   beta'_t(i) = (1 / S_t) * sum_{arcs i->j} p * W_t(run of the arc)          (ONE gather per arc)
// L2 bytes per sequence-frame: 8 A + ~20 R against 20 A for the per-arc kernels (A arcs, R runs): 2.1x fewer at
// R = A / 15.  Same results: it is the same sum in a different order (tools/tied_pdf_study.py holds the algebra
// in numpy against the oracle).  The host picks these kernels when the graph has few runs per arc; phase A and
// phase B are separated by a second grid barrier per frame.
//
// Streams (16-byte records, staged in shared memory):
//   forward   arc record {alpha row offset of the source, prob bits (>= 0), pi[source] bits, -}
//             run record {exp row offset of the pdf, 0x80000000 | last-run-of-state, a_run row offset, -}
//             -> exactly ONE 512-byte gather per record, whatever its kind
//   phase A   run record {beta row offset of the destination, exp row offset, a_run / W row offset, -}
//   phase B   arc record {W row offset of the arc's run, -, prob bits | end-of-state sign, -}
// =====================================================================================================
struct TiedSmem {
  float* red;    // [kDenMainWarps][BP]
  float* tile;   // [kDenAuxWarps][32][33]
  int4* tasks;   // [max tasks per CTA over both directions]
  int4* arcs;    // [max records per CTA over both directions]
  int4* runs;    // [max runs per CTA]
};
__device__ __forceinline__ TiedSmem carve_tied(unsigned char* base, const DenParams& p) {
  TiedSmem s;
  s.red = reinterpret_cast<float*>(base);
  size_t off = sizeof(float) * kDenMainWarps * p.BP;
  s.tile = reinterpret_cast<float*>(base + off);
  off += sizeof(float) * kDenAuxWarps * 32 * 33;
  off = (off + 15) & ~size_t(15);
  s.tasks = reinterpret_cast<int4*>(base + off);
  off += sizeof(int4) * max(p.dir[0].max_tasks_cta, p.dir[1].max_tasks_cta);
  s.arcs = reinterpret_cast<int4*>(base + off);
  off += sizeof(int4) * max(p.dir[0].max_arcs_cta, p.dir[1].max_arcs_cta);
  s.runs = reinterpret_cast<int4*>(base + off);
  return s;
}

template <int SA>
struct RowBuf {
  float a[SA][4];
};

__device__ __forceinline__ void alpha_step_tied(const DenParams& p, const int4* __restrict__ recs,
                                                const int4* __restrict__ tasks, int abeg, int aend, int tb, int t,
                                                float* red, int warp, int lane) {
  constexpr int V = 4, kSA = DEN_TIED_F_SA, kStages = DEN_TIED_F_STAGES;
  const int BP = p.BP, col = lane * V;
  const float* __restrict__ alpha_prev = p.alpha + (size_t)((t - 1) & 1) * p.N * BP;
  float* __restrict__ alpha_cur = p.alpha + (size_t)(t & 1) * p.N * BP;
  const float* __restrict__ et = p.etile + (size_t)((t - 1) & 1) * p.P * BP;
  float* __restrict__ arun = p.arun + (size_t)(t - 1) * p.R * BP;
  const double* Sprev = p.Ssum + (size_t)(t - 1) * BP;
  const int4* __restrict__ a = recs + abeg;
  const int n = aend - abeg;
  const unsigned rowbytes = (unsigned)BP * 4u;
  const char* const ua = reinterpret_cast<const char*>(alpha_prev + col);
  const char* const ue = reinterpret_cast<const char*>(et + col);
  char* const ur = reinterpret_cast<char*>(arun + col);
  double Sd[V];
#pragma unroll
  for (int v = 0; v < V; v++) Sd[v] = ld_norm(Sprev + col + v);
  RowBuf<kSA> ring[kStages];
  auto load = [&](RowBuf<kSA>& buf, int k) {
    if (k < n) {
#pragma unroll
      for (int u = 0; u < kSA; u++) {
        const int2 r = *reinterpret_cast<const int2*>(a + k + u);  // {row byte offset, prob >= 0 | run marker}
        ld_cg<V>(reinterpret_cast<const float*>((r.y < 0 ? ue : ua) + (unsigned)r.x), buf.a[u]);
      }
    }
  };
  float lamS[V], invS[V], part[V], racc[V], sacc[V];
  int task = tb;
  auto arc = [&](const float (&row)[V], int y, int z) {  // racc += (alpha(src) + leaky * pi(src) * S) * p
    const float pr = __int_as_float(y), pi = __int_as_float(z);
    const u64 pr2 = pk2(pr, pr), pi2 = pk2(pi, pi);
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const u64 ad = fma2(pi2, pk2(lamS[2 * h], lamS[2 * h + 1]), pk2(row[2 * h], row[2 * h + 1]));
      upk2(fma2(ad, pr2, pk2(racc[2 * h], racc[2 * h + 1])), racc[2 * h], racc[2 * h + 1]);
    }
  };
  auto consume = [&](RowBuf<kSA>& buf, int k) {
    int4 r[kSA];
    int kinds = 0;
#pragma unroll
    for (int u = 0; u < kSA; u++) {
      r[u] = a[k + u];
      kinds |= r[u].y;
    }
    if (kinds >= 0) {  // a stage of arcs only (15 arcs per run): straight-line code the compiler can overlap
#pragma unroll
      for (int u = 0; u < kSA; u++) arc(buf.a[u], r[u].y, r[u].z);
      return;
    }
#pragma unroll
    for (int u = 0; u < kSA; u++) {
      if (r[u].y >= 0) {
        arc(buf.a[u], r[u].y, r[u].z);
      } else {  // end of a run: buf holds exp(y) of the run's pdf
        float out[V];
#pragma unroll
        for (int v = 0; v < V; v++) {
          out[v] = racc[v] * buf.a[u][v] * invS[v];
          sacc[v] += out[v];
          racc[v] = 0.f;
        }
        __stcs(reinterpret_cast<float4*>(ur + (unsigned)r[u].z), make_float4(out[0], out[1], out[2], out[3]));
        if (r[u].y & 1) {  // last run of this state
          const int4 tk = tasks[task++];
#pragma unroll
          for (int v = 0; v < V; v++) part[v] += sacc[v];
          st_vec<V>(row_ptr(alpha_cur + col, tk.x, rowbytes), sacc);
#pragma unroll
          for (int v = 0; v < V; v++) sacc[v] = 0.f;
        }
      }
    }
  };
#pragma unroll
  for (int st = 0; st < kStages; st++) load(ring[st], st * kSA);
#pragma unroll
  for (int v = 0; v < V; v++) {
    const float S = (float)Sd[v];
    lamS[v] = p.leaky * S;
    invS[v] = 1.0f / S;
    part[v] = racc[v] = sacc[v] = 0.f;
  }
  for (int k = 0; k < n;) {
#pragma unroll
    for (int st = 0; st < kStages; st++) {
      if (k < n) {
        consume(ring[st], k);
        load(ring[st], k + kStages * kSA);
        k += kSA;
      }
    }
  }
#pragma unroll
  for (int v = 0; v < V; v++) red[warp * BP + col + v] = part[v];
}

// phase A of backward step t: per run, occupancy RED and W_t(run)
__device__ __forceinline__ void runs_step_tied(const DenParams& p, const int4* __restrict__ runs, int rbeg, int rend,
                                               int t, int lane) {
  constexpr int V = 4, kD = DEN_TIED_A_DEPTH;
  const int BP = p.BP, col = lane * V;
  const char* const ub = reinterpret_cast<const char*>(p.beta + (size_t)((t + 1) & 1) * p.N * BP + col);
  const char* const ue = reinterpret_cast<const char*>(p.etile + (size_t)(t & 1) * p.P * BP + col);
  const char* const ur = reinterpret_cast<const char*>(p.arun + (size_t)t * p.R * BP + col);
  char* const ug = reinterpret_cast<char*>(p.gtile + (size_t)(t & 1) * p.P * BP + col);
  char* const uw = reinterpret_cast<char*>(p.wbuf + col);
  const double* Bn = p.Bsum + (size_t)(t + 1) * BP;
  float bv[V];
  float bn[kD][V], ev[kD][V], ar[kD][V];
  auto load = [&](int slot, int k) {
    if (k < rend) {
      const int4 r = runs[k];
      ld_cg<V>(reinterpret_cast<const float*>(ub + (unsigned)r.x), bn[slot]);
      ld_cg<V>(reinterpret_cast<const float*>(ue + (unsigned)r.y), ev[slot]);
      ld_cs<V>(reinterpret_cast<const float*>(ur + (unsigned)r.z), ar[slot]);
    }
  };
#pragma unroll
  for (int d = 0; d < kD; d++) load(d, rbeg + d);
#pragma unroll
  for (int v = 0; v < V; v++) bv[v] = p.leaky * (float)ld_norm(Bn + col + v);
  for (int k = rbeg; k < rend; k += kD) {
#pragma unroll
    for (int d = 0; d < kD; d++) {
      if (k + d < rend) {
        const int4 r = runs[k + d];
        float g[V], w[V];
#pragma unroll
        for (int v = 0; v < V; v++) {
          const float full = bn[d][v] + bv[v];
          g[v] = ar[d][v] * full;
          w[v] = full * ev[d][v];
        }
        red_add<V>(reinterpret_cast<float*>(ug + (unsigned)r.y), g);
        st_vec<V>(reinterpret_cast<float*>(uw + (unsigned)r.z), w);
        load(d, k + d + kD);
      }
    }
  }
}

// phase B of backward step t: beta'_t(i) = (1/S_t) sum_{arcs i->j} p W_t(run of the arc)
__device__ __forceinline__ void beta_step_tied(const DenParams& p, const int4* __restrict__ recs,
                                               const int4* __restrict__ tasks, int abeg, int aend, int tb, int te, int t,
                                               float* red, int warp, int lane, double& ab) {
  constexpr int V = 4, kSA = DEN_TIED_B_SA, kStages = DEN_TIED_B_STAGES;
  const int BP = p.BP, col = lane * V;
  float* __restrict__ beta_cur = p.beta + (size_t)(t & 1) * p.N * BP;
  const double* St = p.Ssum + (size_t)t * BP;
  const int4* __restrict__ a = recs + abeg;
  const int n = aend - abeg;
  const unsigned rowbytes = (unsigned)BP * 4u;
  const char* const uw = reinterpret_cast<const char*>(p.wbuf + col);
  double Sd[V];
#pragma unroll
  for (int v = 0; v < V; v++) Sd[v] = ld_norm(St + col + v);
  RowBuf<kSA> ring[kStages];
  auto load = [&](RowBuf<kSA>& buf, int k) {
    if (k < n) {
#pragma unroll
      for (int u = 0; u < kSA; u++) ld_cg<V>(reinterpret_cast<const float*>(uw + (unsigned)a[k + u].x), buf.a[u]);
    }
  };
  float invS[V], part[V], X[V];
  float S0[V];
  int task = tb;
  auto consume = [&](RowBuf<kSA>& buf, int k) {
#pragma unroll
    for (int u = 0; u < kSA; u++) {
      const int prb = a[k + u].z;
      const float pr = fabsf(__int_as_float(prb));
      const u64 pr2 = pk2(pr, pr);
#pragma unroll
      for (int h = 0; h < 2; h++)
        upk2(fma2(pr2, pk2(buf.a[u][2 * h], buf.a[u][2 * h + 1]), pk2(X[2 * h], X[2 * h + 1])), X[2 * h], X[2 * h + 1]);
      if (prb < 0) {  // last arc of this state
        const int4 tk = tasks[task++];
        const float pi_h = __int_as_float(tk.w);
        float bd[V];
#pragma unroll
        for (int v = 0; v < V; v++) {
          bd[v] = X[v] * invS[v];
          part[v] = fmaf(pi_h, bd[v], part[v]);
          // sanity sum at t = 0: alpha'_0(h) = pi(h) (1 + leaky S_0)
          if (t == 0 && col + v < p.B) ab += (double)(pi_h * (1.0f + p.leaky * S0[v])) * (double)bd[v];
          X[v] = 0.f;
        }
        st_vec<V>(row_ptr(beta_cur + col, tk.x, rowbytes), bd);
      }
    }
  };
#pragma unroll
  for (int st = 0; st < kStages; st++) load(ring[st], st * kSA);
#pragma unroll
  for (int v = 0; v < V; v++) {
    S0[v] = (float)Sd[v];
    invS[v] = 1.0f / S0[v];
    part[v] = X[v] = 0.f;
  }
  (void)te;
  for (int k = 0; k < n;) {
#pragma unroll
    for (int st = 0; st < kStages; st++) {
      if (k < n) {
        consume(ring[st], k);
        load(ring[st], k + kStages * kSA);
        k += kSA;
      }
    }
  }
#pragma unroll
  for (int v = 0; v < V; v++) red[warp * BP + col + v] = part[v];
}

template <int PHASE>
__device__ __forceinline__ void den_tied_body(const DenBatch& batch) {
  constexpr int V = 4;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const DenParams& p = batch.c[0];
  const TiedSmem sm = carve_tied(smem_raw, p);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool is_aux = warp >= kDenMainWarps;
  const int auxw = warp - kDenMainWarps;
  float* tile = sm.tile + (is_aux ? auxw : 0) * 32 * 33;
  const int BP = p.BP, T = p.T;
  unsigned epoch = 0;
  double sumsq = 0.0, ab = 0.0, gsum = 0.0;
  unsigned long long* tl = (p.timeline && blockIdx.x == 0 && threadIdx.x == 0) ? p.timeline + PHASE * (T + 3) : nullptr;
  int tli = 0;
  auto stamp = [&]() {
    if (tl) {
      unsigned long long ns;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
      tl[tli++] = ns;
    }
  };
  auto grid_barrier = [&]() {
    __syncthreads();
    if (threadIdx.x == 0) {
      epoch += gridDim.x;
      __threadfence();
      atomicAdd(p.barrier + 32 * PHASE, 1u);
      const unsigned* ctr = p.barrier + 32 * PHASE;
      while (ld_acquire(ctr) < epoch) {
      }
    }
    __syncthreads();
  };
  stamp();
  // stage this CTA's slices: records + tasks of this launch's direction, and (backward) its runs
  const DenDir& D = p.dir[PHASE];
  const int task0 = D.cta_task_off[blockIdx.x], ntask = D.cta_task_off[blockIdx.x + 1] - task0;
  const int arc0 = D.cta_arc_off[blockIdx.x], narc = D.cta_arc_off[blockIdx.x + 1] - arc0;
  int tb = 0, te = 0, abeg = 0, aend = 0, rbeg = 0, rend = 0;
  if (!is_aux) {
    const int bin = blockIdx.x * kDenMainWarps + warp;
    tb = D.warp_task_off[bin] - task0;
    te = D.warp_task_off[bin + 1] - task0;
    abeg = D.warp_arc_off[bin] - arc0;
    aend = D.warp_arc_off[bin + 1] - arc0;
  }
  for (int i = threadIdx.x; i < ntask; i += kDenThreads) sm.tasks[i] = D.tasks[task0 + i];
  for (int i = threadIdx.x; i < narc; i += kDenThreads) sm.arcs[i] = D.arcs[arc0 + i];
  if constexpr (PHASE == 1) {
    const int run0 = p.rdir.cta_arc_off[blockIdx.x], nrun = p.rdir.cta_arc_off[blockIdx.x + 1] - run0;
    for (int i = threadIdx.x; i < nrun; i += kDenThreads) sm.runs[i] = p.rdir.arcs[run0 + i];
    if (!is_aux) {
      const int bin = blockIdx.x * kDenMainWarps + warp;
      rbeg = p.rdir.warp_arc_off[bin] - run0;
      rend = p.rdir.warp_arc_off[bin + 1] - run0;
    }
  }
  __syncthreads();
  const int4* arcs = sm.arcs;
  const int4* tasks = sm.tasks;

  if constexpr (PHASE == 0) {
    // ---- prologue: alpha_0 = pi, S_0; stage exp(y_0)
    if (is_aux) {
      aux_work(p, -1, 0, tile, auxw, lane, true, sumsq, false, gsum);
    } else {
      const int col = lane * V;
      float part = 0.f;
      for (int task = tb; task < te; task++) {
        const int4 tk = tasks[task];
        const float pi_h = __int_as_float(tk.w);
        float out[V];
#pragma unroll
        for (int v = 0; v < V; v++) out[v] = pi_h;
        st_vec<V>(p.alpha + (size_t)tk.x * BP + col, out);
        part += pi_h;
      }
#pragma unroll
      for (int v = 0; v < V; v++) sm.red[warp * BP + col + v] = part;
    }
    cta_reduce_to(sm.red, BP, kDenMainWarps, p.Ssum);
    grid_barrier();
    for (int t = 1; t <= T; t++) {
      stamp();
      if (is_aux) {
        if (t < T) aux_work(p, -1, t, tile, auxw, lane, true, sumsq, false, gsum);
      } else {
        alpha_step_tied(p, arcs, tasks, abeg, aend, tb, t, sm.red, warp, lane);
      }
      cta_reduce_to(sm.red, BP, kDenMainWarps, p.Ssum + (size_t)t * BP);
      grid_barrier();
    }
    stamp();
    if (is_aux) {
      sumsq = warp_sum(sumsq);
      if (lane == 0 && sumsq != 0.0) atomicAdd(p.scalars + 3, sumsq);
    }
    if (blockIdx.x == 0) {  // ComputeTotLogLike
      double acc = 0.0;
      for (int i = threadIdx.x; i < (T + 1) * BP; i += kDenThreads) {
        const int t = i / BP, s = i - t * BP;
        if (s < p.B) {
          float S = (float)__ldcg(p.Ssum + i);
          if (t == T) S = S * (1.0f + p.leaky * p.sum_pi);
          acc += log((double)S);
        }
      }
      acc = warp_sum(acc);
      if (lane == 0) atomicAdd(p.scalars + 2, acc);
    }
  } else {
    // ---- backward prologue: beta'_T(h) = 1 / (S_T (1 + leaky*sum_pi))
    if (!is_aux) {
      const int col = lane * V;
      const double* ST = p.Ssum + (size_t)T * BP;
      float* betaT = p.beta + (size_t)(T & 1) * p.N * BP;
      float inv[V], part[V];
#pragma unroll
      for (int v = 0; v < V; v++) {
        const float S = (float)__ldcg(ST + col + v);
        inv[v] = 1.0f / (S * (1.0f + p.leaky * p.sum_pi));
        part[v] = 0.f;
      }
      for (int task = tb; task < te; task++) {
        const int4 tk = tasks[task];
        const float pi_h = __int_as_float(tk.w);
        st_vec<V>(betaT + (size_t)tk.x * BP + col, inv);
#pragma unroll
        for (int v = 0; v < V; v++) part[v] = fmaf(pi_h, inv[v], part[v]);
      }
#pragma unroll
      for (int v = 0; v < V; v++) sm.red[warp * BP + col + v] = part[v];
    }
    cta_reduce_to(sm.red, BP, kDenMainWarps, p.Bsum + (size_t)T * BP);
    grid_barrier();
#ifdef DEN_TIED_DEBUG_PHASES  // tools: per-phase time of the backward step (globaltimer accumulators, printed by a few warps)
    unsigned long long dbg[5] = {0, 0, 0, 0, 0}, dbg_t = 0;
    auto tick = [&](int slot) {
      unsigned long long ns;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
      if (slot >= 0) dbg[slot] += ns - dbg_t;
      dbg_t = ns;
    };
#else
    auto tick = [&](int) {};
#endif
    tick(-1);
    for (int t = T - 1; t >= 0; t--) {
      stamp();
      // phase A: runs (the aux warps only pass through the barrier: their work belongs to the long phase)
      if (!is_aux) runs_step_tied(p, sm.runs, rbeg, rend, t, lane);
      tick(0);
      grid_barrier();
      tick(1);
      if (t < T - 1 && p.deriv_progress && blockIdx.x == 0 && warp == kDenMainWarps && lane == 0 &&
          ((p.progress_mask[((T - 2 - t) >> 6) & 7] >> ((T - 2 - t) & 63)) & 1ull))
        st_release_sys(p.deriv_progress, (unsigned)(T - 2 - t));  // rows of frames > t+1 were complete a step ago
      // phase B: arcs.  First, the a_run rows this warp reads in the NEXT step's phase A are requested into L2:
      // they come from HBM (written a whole forward launch ago), and phase A is a short dependent chain of
      // ~10 runs per warp that would otherwise pay the DRAM latency five times over.
      if (!is_aux && t > 0) {
        const char* nxt = reinterpret_cast<const char*>(p.arun + (size_t)(t - 1) * p.R * BP);
        for (int k = rbeg + (lane >> 2); k < rend; k += 8)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + (unsigned)sm.runs[k].z + (lane & 3) * 128));
      }
#ifndef DEN_DEBUG_SKIP_AUX_B
      if (is_aux) aux_work(p, t + 1 < T ? t + 1 : -1, t - 1, tile, auxw, lane, false, sumsq, false, gsum, t);
#else
      if (is_aux) {
      }
#endif
      else beta_step_tied(p, arcs, tasks, abeg, aend, tb, te, t, sm.red, warp, lane, ab);
      tick(2);
      cta_reduce_to(sm.red, BP, kDenMainWarps, p.Bsum + (size_t)t * BP);
      tick(3);
      grid_barrier();
      tick(4);
    }
#ifdef DEN_TIED_DEBUG_PHASES
    if (p.timeline && lane == 0 && (blockIdx.x == 0 || blockIdx.x == 77) && (warp == 0 || warp == 7 || warp == kDenMainWarps))
      printf("tied bwd cta %d warp %d: per frame us: phase A %.2f | barrier1 %.2f | phase B %.2f | reduce %.2f | barrier2 %.2f\n",
             (int)blockIdx.x, warp, dbg[0] / 1e3 / T, dbg[1] / 1e3 / T, dbg[2] / 1e3 / T, dbg[3] / 1e3 / T, dbg[4] / 1e3 / T);
#endif
    stamp();
    if (p.deriv_progress && blockIdx.x == 0 && warp == kDenMainWarps && lane == 0)
      st_release_sys(p.deriv_progress, (unsigned)(T - 1));
    if (is_aux) {
      aux_work(p, 0, -1, tile, auxw, lane, false, sumsq, true, gsum);
      gsum = warp_sum(gsum);
      if (lane == 0 && gsum != 0.0) atomicAdd(p.scalars + 1, gsum);
    } else {
      ab = warp_sum(ab);
      if (lane == 0 && ab != 0.0) atomicAdd(p.scalars + 0, ab);
    }
  }
  __syncthreads();
  stamp();
}

__global__ void __maxnreg__(DEN_F_MAXNREG_TIED) den_tied_forward_kernel(const __grid_constant__ DenBatch batch) {
  den_tied_body<0>(batch);
}
__global__ void __launch_bounds__(kDenThreads, 1) den_tied_backward_kernel(const __grid_constant__ DenBatch batch) {
  den_tied_body<1>(batch);
}

// The forward launch is capped at 112 registers/thread so that 8 K registers per SM stay free:
// the numerator kernel (128 threads x 48 registers per CTA) can then run on the same SMs while
// the forward recurrence is in flight.  The backward launch takes the whole register file.
#ifndef DEN_F_MAXNREG
#define DEN_F_MAXNREG 112
#endif
template <int SUBS, bool STAGED, bool SPLIT, int NCTX>
__global__ void __maxnreg__(DEN_F_MAXNREG) den_forward_kernel(const __grid_constant__ DenBatch batch) {
  den_chain_body<SUBS, STAGED, 0, SPLIT, NCTX>(batch);
}
#ifdef DEN_B_MAXNREG
template <int SUBS, bool STAGED, bool SPLIT, int NCTX>
__global__ void __maxnreg__(DEN_B_MAXNREG) den_backward_kernel(
#else
template <int SUBS, bool STAGED, bool SPLIT, int NCTX>
__global__ void __launch_bounds__(kDenThreads, 1) den_backward_kernel(
#endif
    const __grid_constant__ DenBatch batch) {
  den_chain_body<SUBS, STAGED, 1, SPLIT, NCTX>(batch);
}

template <int SUBS, bool STAGED, bool SPLIT, int NCTX>
cudaError_t launch_t(const DenBatch& b, int phase, int grid, size_t smem, cudaStream_t stream) {
  void (*kern)(DenBatch) = phase == 0 ? den_forward_kernel<SUBS, STAGED, SPLIT, NCTX>
                                      : den_backward_kernel<SUBS, STAGED, SPLIT, NCTX>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  void* args[] = {const_cast<DenBatch*>(&b)};
  e = cudaLaunchCooperativeKernel((void*)kern, dim3(grid), dim3(kDenThreads), args, smem, stream);
  if (e == cudaSuccess) count_launch();
  return e;
}

template <int SUBS, bool STAGED>
cudaError_t launch_p(const DenBatch& b, int phase, int grid, size_t smem, cudaStream_t stream) {
  const bool split = b.c[0].has_split != 0;
  if constexpr (SUBS == 1) {
    if (b.nctx == 2)
      return split ? launch_t<1, STAGED, true, 2>(b, phase, grid, smem, stream)
                   : launch_t<1, STAGED, false, 2>(b, phase, grid, smem, stream);
  }
  return split ? launch_t<SUBS, STAGED, true, 1>(b, phase, grid, smem, stream)
               : launch_t<SUBS, STAGED, false, 1>(b, phase, grid, smem, stream);
}

}  // namespace

int den_subs(int BP) { return BP >= 128 ? 1 : (BP >= 64 ? 2 : 4); }

size_t den_smem_bytes(const DenParams& p) {
  size_t off = sizeof(float) * kDenMainWarps * p.subs * p.BP;
  off += sizeof(float) * kDenAuxWarps * 32 * 33;
  off = (off + 127) & ~size_t(127);
  off += kRingBytes;
  if (p.arcs_in_smem) {
    off += sizeof(int4) * (size_t)std::max(p.dir[0].max_tasks_cta, p.dir[1].max_tasks_cta);
    off += sizeof(int4) * (size_t)std::max(p.dir[0].max_arcs_cta, p.dir[1].max_arcs_cta);
  }
  return off;
}

size_t den_tied_smem_bytes(const DenParams& p) {
  size_t off = sizeof(float) * kDenMainWarps * p.BP;
  off += sizeof(float) * kDenAuxWarps * 32 * 33;
  off = (off + 15) & ~size_t(15);
  off += sizeof(int4) * (size_t)std::max(p.dir[0].max_tasks_cta, p.dir[1].max_tasks_cta);
  off += sizeof(int4) * (size_t)std::max(p.dir[0].max_arcs_cta, p.dir[1].max_arcs_cta);
  off += sizeof(int4) * (size_t)p.rdir.max_arcs_cta;
  return off;
}

cudaError_t launch_den_tied(const DenBatch& b, int phase, int grid, cudaStream_t stream) {
  const DenParams& p = b.c[0];
  if (b.nctx != 1 || p.BP != kCtxBP || p.subs != 1) return cudaErrorInvalidValue;
  const size_t smem = den_tied_smem_bytes(p);
  void (*kern)(DenBatch) = phase == 0 ? den_tied_forward_kernel : den_tied_backward_kernel;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  void* args[] = {const_cast<DenBatch*>(&b)};
  e = cudaLaunchCooperativeKernel((void*)kern, dim3(grid), dim3(kDenThreads), args, smem, stream);
  if (e == cudaSuccess) count_launch();
  return e;
}

int den_max_grid(int device) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  return sms;
}

cudaError_t launch_den(const DenBatch& b, int phase, int grid, cudaStream_t stream) {
  const DenParams& p = b.c[0];
  const size_t smem = den_smem_bytes(p);
  if (p.subs != den_subs(p.BP) || b.nctx < 1 || b.nctx > 2 || (b.nctx == 2 && p.subs != 1)) return cudaErrorInvalidValue;
  if (p.arcs_in_smem) {
    if (p.subs == 1) return launch_p<1, true>(b, phase, grid, smem, stream);
    if (p.subs == 2) return launch_p<2, true>(b, phase, grid, smem, stream);
    return launch_p<4, true>(b, phase, grid, smem, stream);
  }
  if (p.subs == 1) return launch_p<1, false>(b, phase, grid, smem, stream);
  if (p.subs == 2) return launch_p<2, false>(b, phase, grid, smem, stream);
  return launch_p<4, false>(b, phase, grid, smem, stream);
}

}  // namespace tb

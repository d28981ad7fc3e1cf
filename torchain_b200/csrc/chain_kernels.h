// chain_kernels.h -- host-visible launch interface of the sm_100a kernels.
// Internal to libtorchain_b200.so (the public boundary is include/torchain_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tb {

// Where element (frame t, sequence s, pdf c) of a caller matrix lives: base + t*st_t + s*st_s + c*st_p
// (in floats).  The reference's 2-D form (T*B rows, row = t*B + s, /root/reference/torchain/functions.py:118-125)
// is {B*row_stride, row_stride, 1}; a network output left in its native (B, P, T) shape is {1, P*T, T},
// which lets the kernels read it and write its gradient in place of to2d's permute+contiguous copy.
struct Layout3 {
  int64_t t, s, p;
};

// ---------------------------------------------------------------------------
// Denominator (persistent alpha / beta+derivative kernel)
// ---------------------------------------------------------------------------
#ifndef DEN_THREADS
#define DEN_THREADS 512
#endif
#ifndef DEN_AUX_WARPS
#define DEN_AUX_WARPS 2
#endif
constexpr int kDenThreads = DEN_THREADS;       // threads per CTA (one CTA per SM)
constexpr int kDenWarps = kDenThreads / 32;
constexpr int kDenAuxWarps = DEN_AUX_WARPS;                // staging (transpose+exp) / derivative write-out warps
constexpr int kDenMainWarps = kDenWarps - kDenAuxWarps;
// every lane group's arc stream is padded (zero-probability arcs) to a multiple of this many arcs,
// a common multiple of the stage widths of the two gather loops
#ifndef DEN_STREAM_PAD
#define DEN_STREAM_PAD 4
#endif
constexpr int kStreamPad = DEN_STREAM_PAD;
#define DEN_F_SA_DEFAULT 4
#define DEN_B_SA_DEFAULT 4
constexpr int kCtxBP = 128;                    // sequences per context (lane width of the full-size kernels)
constexpr int kMaxBP = 256;                    // sequences per launch pair: two contexts of kCtxBP

// One direction of the graph, partitioned over `grid` CTAs x kDenMainWarps warps.
//   dir 0 ("alpha"): tasks are destination states, arcs are their in-arcs
//   dir 1 ("beta") : tasks are source states,      arcs are their out-arcs
struct DenDir {
  const int4* arcs;          // {other_state * 4BP, pdf * 4BP (row byte offsets), float_bits(prob) | end-of-state flag in the sign bit, float_bits(pi[other_state])}
  const int4* tasks;         // {state, arc_begin (relative to the CTA's slice), split flag (0 whole state, 2 first / 1 other sub-task of a split state), float_bits(pi[state])}
  const int* cta_arc_off;    // [grid+1]
  const int* cta_task_off;   // [grid+1]
  const int* warp_task_off;  // [grid*kDenMainWarps*subs+1] absolute task indices, one bin per lane group
  const int* warp_arc_off;   // [grid*kDenMainWarps*subs+1] absolute arc indices
  int max_arcs_cta;
  int max_tasks_cta;
};

struct DenParams {
  DenDir dir[2];
  int N, P;
  int B;        // sequences in this pass
  int BP;       // padded to 32*V*chunks
  int Btot;     // sequences in the whole minibatch (row = frame*Btot + seq0 + s)
  int seq0;     // first sequence of this pass
  int T;
  const float* y;
  Layout3 yl;
  float* deriv;
  Layout3 dl;
  float* alpha;      // (T+1) x N x BP   un-mixed alphas (Kaldi stores alpha-dash; we mix on the fly)
  float* beta;       // 3 x N x BP       beta-dash rotation (write / gather / re-zero)
  float* etile;      // 2 x P x BP       exp(y) of one frame, sequence-innermost
  float* gtile;      // 2 x P x BP       occupancies of one frame, sequence-innermost (zero on entry)
  double* Ssum;      // (T+1) x BP       sum_h alpha_t(h,s)      (zero on entry)
  double* Bsum;      // (T+1) x BP       sum_h pi(h) beta'_t(h,s) (zero on entry)
  double* scalars;   // [0] alpha.beta at t=0, [1] sum gamma_0, [2] den logprob, [3] sum y^2  (accumulated)
  unsigned* barrier; // [2 x 32] one counter per phase (forward / backward launch), zero on entry
  float leaky;
  float sum_pi;
  float scale_gamma;  // deriv_scale * (-supervision.weight)
  float scale_y;      // deriv_scale * (-supervision.weight * l2_regularize)
  int arcs_in_smem;   // 1: arc slices staged in shared memory
  int subs;           // lane groups per warp: 1 (BP >= 128), 2 (BP = 64), 4 (BP = 32); the partition has grid*kDenMainWarps*subs bins
  int has_split;      // 1: some states are dealt out as several sub-tasks (selects the SPLIT kernels)
  // numerator occupancies, subtracted from the occupancy tile by the backward launch so that the
  // flush writes w*gamma_num - w*gamma_den in one go (arcs ordered by frame)
  const int* num_frame_off;    // [T+1] ranges into num_frame_arcs, or null (no numerator)
  const int* num_frame_arcs;   // arc ids
  const int* num_arc_pdf;      // [A]
  const int* num_arc_seq;      // [A] sequence index in the whole minibatch
  const float* num_arc_post;   // [A] written by num_fb_kernel
  // host-buffer entry only: nnet_output arrives in chunks of y_chunk_frames frames; y_ready[k] becomes
  // non-zero (written by the copy stream) once chunk k has landed.  null = everything is resident.
  const unsigned* y_ready;
  int y_chunk_frames;
  // host-buffer entry only: mapped pinned counter; after the grid barrier of backward step t the
  // derivative rows of frames > t are complete and the kernel publishes T-1-t here with a system-scope
  // release (at the start of step t-1, and once more after the last barrier), so the host can start
  // copying finished frame chunks out while the recurrence is still running.  null = unused.
  unsigned* deriv_progress;
  // bit v set: publish v when v trailing frames are complete (the host sets the bits of its chunk boundaries: a
  // system-scope release costs the publishing warp ~1.7 us, so it is paid per chunk, not per frame); T > 512: all ones
  unsigned long long progress_mask[8];
  // ---- tied-pdf ("run-based") kernels only (den_tied_*): a run is a (destination state, pdf) pair
  // dir[0] then holds the forward stream of arc + run records, dir[1] the backward stream of arcs that
  // point at their run's W row, and rdir the runs dealt evenly over the bins (arcs = run records, no tasks).
  DenDir rdir;
  int R;             // number of runs
  float* arun;       // T x R x BP   forward contribution of every run, a_{t+1}(run), written forward / read backward
  float* wbuf;       // R x BP       W_t(run) = (beta'_{t+1}(dst) + b_{t+1}) * e_t(pdf) of the current backward step
  unsigned long long* warp_busy;  // [2][grid][kDenWarps] SM cycles each warp spent in its frame work (feeds the partition autotuner)
  unsigned long long* cta_trace;  // optional [phase][frame][cta][4]: main warp 0 done, aux warp 0 done, CTA done, barrier exit
  unsigned long long* timeline;  // optional: globaltimer at every phase boundary (2T+5 entries), or null
};

// One launch carries one or two independent sequence groups ("contexts") of the same geometry;
// c[i] differ only in B, seq0 and their buffers (alpha, beta, etile, gtile, Ssum, Bsum, barrier).
struct DenBatch {
  DenParams c[2];
  int nctx;
};

size_t den_smem_bytes(const DenParams& p);
// Enqueues one persistent launch (cooperative, `grid` CTAs): phase 0 = forward (alpha)
// recurrence over all frames, phase 1 = backward (beta + derivative) recurrence.
cudaError_t launch_den(const DenBatch& b, int phase, int grid, cudaStream_t stream);
// The same two launches for a graph with tied pdfs (few runs per state): one context of kCtxBP sequences,
// arc / run slices in shared memory, no split states.  den_tied_smem_bytes(p) must fit the device.
size_t den_tied_smem_bytes(const DenParams& p);
cudaError_t launch_den_tied(const DenBatch& b, int phase, int grid, cudaStream_t stream);
int den_max_grid(int device);  // co-resident CTAs (one per SM)
int den_subs(int BP);          // lane groups per warp for a padded batch size

// ---------------------------------------------------------------------------
// Numerator (batched supervision-FST forward-backward) + final pass
// ---------------------------------------------------------------------------
struct NumParams {
  int num_chains;
  const int* chain_level_off;   // [num_chains+1] into level_state_off
  const int* level_state_off;   // per level: first state (global state index); one extra entry per chain
  const int* chain_arc_off;     // [num_chains+1] arc-id range of each chain
  const int* in_off;            // [S+1]  in-arc CSR per state
  const int2* in_arcs;          // {src_state, arc_id}
  const int* out_off;           // [S+1]  out-arc CSR per state
  const int2* out_arcs;         // {dst_state, arc_id}
  const float* final_cost;      // [S]
  const int* arc_row;           // [A] row of nnet_output (frame*B + seq)
  const int* arc_pdf;           // [A]
  const float* arc_cost;        // [A]
  int B;                        // sequences in the minibatch (row -> frame = row / B, seq = row % B)
  const float* y;
  Layout3 yl;
  double* arc_val;              // [A] scratch: y - cost
  double* log_alpha;            // [S]
  double* log_beta;             // [S]
  float* arc_post;              // [A] occupancy of each arc
  double* chain_logprob;        // [num_chains]
};
cudaError_t launch_numerator(const NumParams& p, cudaStream_t stream);

struct FinalParams {
  // scalars
  const double* den_scalars;    // [4] from the denominator kernel
  const double* chain_logprob;  // [num_chains]
  int num_chains;
  int B, T;
  float sup_weight;
  float l2;
  float deriv_scale;
  float* results_dev;           // [4]: objf, l2_term, weight, ok
  // derivative fix-up
  float* deriv;
  Layout3 dl;
  float* xent;                  // may be null
  Layout3 xl;
  const float* y;
  Layout3 yl;
  int64_t rows, cols;           // rows = B*T
  // numerator occupancies per arc (only needed for the optional xent output)
  int num_arcs;
  const int* arc_row;
  const int* arc_pdf;
  const float* arc_post;
};
cudaError_t launch_final(const FinalParams& p, cudaStream_t stream);
// x(t, s, c) = 0 over T x B x P elements of a strided matrix (the xent output accumulates with atomics)
cudaError_t launch_zero(float* x, Layout3 l, int T, int B, int P, cudaStream_t stream);

// ---------------------------------------------------------------------------
// Relayout of the network output / gradient (layout_kernel.cu): x (B, P, T) contiguous <-> rows t*B + s
// ---------------------------------------------------------------------------
cudaError_t launch_to2d(const float* x_bpt, float* y_rows, int B, int P, int T, int64_t stride, cudaStream_t stream);
cudaError_t launch_from2d(const float* y_rows, float* x_bpt, int B, int P, int T, int64_t stride, float scale,
                          cudaStream_t stream);

// kernels launched by this library so far (process-wide)
void count_launch(int n = 1);
int64_t launch_count();

}  // namespace tb

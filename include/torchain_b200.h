/* torchain_b200.h -- C ABI of libtorchain_b200.so
 *
 * Drop-in boundary for the LF-MMI ("chain") loss hot path of nttcslab-sp/torchain.
 * Every entry point is plain C: raw pointers, sizes, strides, a device index and
 * a CUDA stream handle (void* == cudaStream_t).  No torch / TH / THC / Kaldi types.
 *
 * Each declaration cites the reference interface it replaces
 * (paths are into the reference tree, /root/reference in the build container).
 *
 * Conventions
 *   - functions returning int: 1 = success, 0 = failure (reference convention,
 *     src/my_lib_chain.cpp:135, src/my_lib_example.cpp:42); after a failure
 *     torchain_last_error() returns a thread-local message.  Nothing throws
 *     across this boundary (the reference lets Kaldi abort the process instead).
 *   - functions returning void*: NULL = failure (src/my_lib_example.cpp:73).
 *   - numeric failure (NaN/inf objective or the alpha*beta sanity check off by
 *     more than 2) is NOT an error: objf = -10*weight and zero derivatives, as
 *     Kaldi's ComputeChainObjfAndDeriv does.
 *   - there is no CPU fallback: every compute entry needs a CUDA device.
 */
#ifndef TORCHAIN_B200_H_
#define TORCHAIN_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- errors / library info ------------------------------------------------ */

/* Thread-local description of the last failure on this thread ("" if none). */
const char* torchain_last_error(void);

/* ABI version of this library (bumped on incompatible change). */
int torchain_abi_version(void);

/* Number of kernels of THIS library launched by the calling process so far
 * (all threads).  bench.py reports the delta over its timed region. */
int64_t torchain_kernel_launch_count(void);

/* Replaces: void my_lib_set_kaldi_device(THCudaTensor* t)   src/my_lib.h:48,
 * src/my_lib_chain.cpp:34-36, src/common.hpp:120-132.  The reference overwrote
 * Kaldi's process-wide CuDevice singleton; this library has no global device
 * state (device and stream are per-call arguments), so this only validates that
 * `device` exists.  Kept so torchain.io.set_kaldi_device keeps working. */
int my_lib_set_kaldi_device(int device);

/* ---- denominator graph ---------------------------------------------------- */

/* Replaces: void* my_lib_denominator_graph_new(const char* rxfilename, int num_pdf)
 * src/my_lib.h:29, src/my_lib_example.cpp:129-134 (fst::ReadFstKaldi + new
 * kaldi::chain::DenominatorGraph).  Reads an OpenFst binary VectorFst<StdArc>
 * ("vector"/"standard") or an AT&T/Kaldi text FST, with no OpenFst dependency. */
void* my_lib_denominator_graph_new(const char* rxfilename, int num_pdf);

/* Same object from arrays (what DenominatorGraph's ctor consumes after
 * ReadFstKaldi): arc a goes src[a] -> dst[a] with ilabel[a] = pdf_id + 1 and
 * tropical cost[a] = -log(prob); final_cost[s] = +inf for non-final states
 * (final weights only enter the initial-probability estimate, as in Kaldi). */
void* torchain_denominator_graph_from_arrays(int32_t num_states, int32_t num_arcs,
                                             const int32_t* src, const int32_t* dst,
                                             const int32_t* ilabel, const float* cost,
                                             int32_t start, const float* final_cost,
                                             int32_t num_pdfs);

/* Replaces: void my_lib_denominator_graph_free(void*)   src/my_lib.h:30,
 * src/my_lib_example.cpp:136-138. */
void my_lib_denominator_graph_free(void* den_graph_ptr);

int torchain_denominator_graph_num_states(void* den_graph_ptr);
int torchain_denominator_graph_num_arcs(void* den_graph_ptr);
int torchain_denominator_graph_num_pdfs(void* den_graph_ptr);
/* Runs = distinct (destination state, pdf) pairs of the graph; a real den.fst, whose pdfs belong to the state
 * being entered, has ~2 per state.  is_tied: 1 if the run-based kernels will be used for this graph (few runs
 * per arc, no state with more arcs than the split cap), else 0 (per-arc kernels). */
int torchain_denominator_graph_num_runs(void* den_graph_ptr);
int torchain_denominator_graph_is_tied(void* den_graph_ptr);
/* Copies the N initial probabilities (Kaldi DenominatorGraph::InitialProbs()). */
int torchain_denominator_graph_initial_probs(void* den_graph_ptr, float* out_n);

/* ---- supervision ---------------------------------------------------------- */

/* Replaces: void* my_lib_supervision_new(void* reader_ptr) and
 * my_lib_supervision_rand_new   src/my_lib.h:19,21, src/my_lib_example.cpp:71-76,
 * src/my_lib_example_rand.cpp:244-249, i.e. `new kaldi::chain::Supervision(...)`.
 * The Kaldi egs readers that fed those are outside the hot path; here the
 * Supervision is built from the arrays of its merged FST: an epsilon-free
 * acceptor over num_sequences*frames_per_sequence frames, states topologically
 * sorted with state 0 the start, arcs of state s at [arc_begin[s], arc_begin[s+1]),
 * arc_label = pdf_id + 1, tropical costs, final_cost = +inf for non-final. */
void* torchain_supervision_from_arrays(float weight, int32_t num_sequences,
                                       int32_t frames_per_sequence, int32_t label_dim,
                                       int32_t num_states, const int32_t* arc_begin,
                                       const int32_t* arc_dst, const int32_t* arc_label,
                                       const float* arc_cost, const float* final_cost);

/* One binary Kaldi `Supervision` object (as chain::Supervision::Write(os, true) emits it, e.g. the
 * payload of an ark written by chain-get-supervision; kaldi@22fbdd:src/chain/chain-supervision.cc),
 * starting at the "<Supervision>" token.  *consumed (optional) receives the bytes parsed.  The
 * byte layout is restated from the published Kaldi / OpenFst sources and is unverified against
 * a Kaldi-written file (none is available in the build environment). */
void* torchain_supervision_from_kaldi_bytes(const void* data, int64_t nbytes, int64_t* consumed);

/* The Supervision of the next binary nnet3 chain example in `data` (an ark or a piece of one: "[key \0B]
 * <Nnet3ChainEg> ... <NnetChainSup> ... <Supervision> ... </Supervision> ... </Nnet3ChainEg>"), i.e. what the
 * reference obtains as `new Supervision(eg.outputs[0].supervision)` (src/my_lib_example.cpp:71-76,
 * src/my_lib_example_rand.cpp:244-249).  Examples are located by scanning for the <Nnet3ChainEg> token, as the
 * reference's own indexer does (example/chime5/ark2scp.py:24-40); features, i-vectors and derivative weights
 * are skipped.  *consumed (optional) = offset just past the example, so a caller can walk an ark.  Text-mode
 * egs are refused.  Combine the examples of a minibatch with torchain_supervision_merge. */
void* torchain_supervision_from_chain_egs(const void* data, int64_t nbytes, int64_t* consumed);

/* Minibatch assembly: the hot-path part of Kaldi's MergeChainExamples / AppendSupervision, which the
 * reference runs inside its egs reader (src/my_lib_example_rand.cpp:160).  Appends `n` supervisions
 * (each with one or more sequences) sequence-major into a new handle whose num_sequences is the sum;
 * like Kaldi it refuses (NULL) examples of different weight, frames_per_sequence or label_dim.  The
 * inputs stay valid and owned by the caller. */
void* torchain_supervision_merge(void* const* supervision_ptrs, int32_t n);

/* Starts the (asynchronous) upload of the supervision's device image on `stream` of `device`, so that a
 * data-loading thread can have minibatch k+1 resident while minibatch k is still being computed.  Optional:
 * the first ComputeChainObjfAndDeriv on a device uploads on its own stream otherwise.  Later calls on other
 * streams order themselves behind the upload.  Never synchronises the device: the image lives in pooled
 * device + pinned memory that supervision handles hand back when they are freed (a training loop builds one
 * Supervision per minibatch, /root/reference/example/chime5/train_faster.py:116-141). */
int torchain_supervision_prefetch(void* supervision_ptr, int device, void* stream);

/* Frees the pooled supervision blocks no handle is using (at most 8 are kept; like
 * torch.cuda.empty_cache() for this library's one pool).  Waits for the device. */
void torchain_supervision_pool_trim(void);

/* The reference's reader-based constructors; egs reading is out of scope, so
 * these always return NULL and set the error string. */
void* my_lib_supervision_new(void* reader_ptr);
void* my_lib_supervision_rand_new(void* reader_ptr);

/* Replace src/my_lib.h:22-25, src/my_lib_example.cpp:78-95. */
void my_lib_supervision_free(void* supervision_ptr);
int my_lib_supervision_num_pdf(void* supervision_ptr);
int my_lib_supervision_num_sequence(void* supervision_ptr);
int my_lib_supervision_num_frame(void* supervision_ptr);
/* Number of independent numerator chains the merged FST was split into
 * (num_sequences when it factorises per sequence, else 1). */
int torchain_supervision_num_chains(void* supervision_ptr);

/* ---- the chain loss ------------------------------------------------------- */

/* Replaces: int my_lib_ComputeChainObjfAndDeriv(void* den_graph_ptr,
 *   void* supervision_ptr, THCudaTensor* nnet_output_ptr, THFloatTensor* results,
 *   THCudaTensor* nnet_output_deriv_ptr, THCudaTensor* xent_output_deriv_ptr,
 *   float l2_regularize, float leaky_hmm_coefficient, float xent_regularize)
 * src/my_lib.h:33-42, src/my_lib_chain.cpp:104-136.  Same argument order and
 * meaning; each tensor becomes (device pointer, row stride in elements) the way
 * common::make_matrix viewed it (src/common.hpp:109-117); rows/cols are the
 * matrix shape and must equal (num_sequences*frames_per_sequence, num_pdfs);
 * rows are ordered frame-major: row = frame * num_sequences + sequence.
 *
 *   nnet_output        device, rows x cols fp32, read-only
 *   results            HOST float[3] = {objf, l2_term, weight}  (THFloatTensor)
 *   nnet_output_deriv  device, rows x cols fp32; need not be initialised
 *   xent_output_deriv  device or NULL; written only if xent_regularize != 0
 *                      (= supervision.weight * numerator posteriors)
 *   device, stream     CUDA device index and cudaStream_t the work is enqueued on
 *
 * Synchronises `stream` once, to deliver `results` to the host. */
int my_lib_ComputeChainObjfAndDeriv(void* den_graph_ptr, void* supervision_ptr,
                                    const float* nnet_output, int64_t rows, int64_t cols,
                                    int64_t nnet_output_stride,
                                    float* results,
                                    float* nnet_output_deriv, int64_t deriv_stride,
                                    float* xent_output_deriv, int64_t xent_stride,
                                    float l2_regularize, float leaky_hmm_coefficient,
                                    float xent_regularize,
                                    int device, void* stream);

/* Extended form used by the Python autograd op:
 *   deriv_scale   both derivative outputs are multiplied by this; -1 yields the
 *                 negated gradient torchain's backward returns
 *                 (torchain/functions.py:106-115) without a separate kernel.
 *   results_dev   if non-NULL, {objf, l2_term, weight} are also written to this
 *                 DEVICE float[3] (lets a caller defer the host read).
 *   results       HOST float[3] or NULL; if NULL the call does not synchronise. */
int torchain_ComputeChainObjfAndDeriv_ex(void* den_graph_ptr, void* supervision_ptr,
                                         const float* nnet_output, int64_t rows, int64_t cols,
                                         int64_t nnet_output_stride,
                                         float* results, float* results_dev,
                                         float* nnet_output_deriv, int64_t deriv_stride,
                                         float* xent_output_deriv, int64_t xent_stride,
                                         float l2_regularize, float leaky_hmm_coefficient,
                                         float xent_regularize, float deriv_scale,
                                         int device, void* stream);

/* Strided form: replaces the pair to2d + my_lib_ComputeChainObjfAndDeriv of
 * torchain/functions.py:118-125,128-138.  chain_loss permutes a (B, num_pdfs, T)
 * network output into a fresh (T*B, num_pdfs) copy before the call and autograd
 * permutes the gradient back; here each matrix is (pointer, three ELEMENT strides
 * {frame, sequence, pdf}), so the kernels read the network output and write its
 * gradient where they are.  The reference's 2-D form is {B*row_stride, row_stride, 1};
 * a contiguous (B, P, T) tensor is {1, P*T, T}.  num_frames / num_sequences /
 * num_pdfs must equal the supervision's and the den graph's; derivative strides
 * must be positive and describe non-overlapping elements.  Everything else as
 * torchain_ComputeChainObjfAndDeriv_ex. */
int torchain_ComputeChainObjfAndDeriv_strided(void* den_graph_ptr, void* supervision_ptr,
                                              const float* nnet_output, int64_t num_frames,
                                              int64_t num_sequences, int64_t num_pdfs,
                                              const int64_t* nnet_output_strides /* [3] */,
                                              float* results, float* results_dev,
                                              float* nnet_output_deriv, const int64_t* deriv_strides /* [3] */,
                                              float* xent_output_deriv, const int64_t* xent_strides /* [3] */,
                                              float l2_regularize, float leaky_hmm_coefficient,
                                              float xent_regularize, float deriv_scale,
                                              int device, void* stream);

/* Replace `to2d` (torchain/functions.py:118-125: x.permute(2, 0, 1).contiguous().view(-1, C))
 * and the permuted gradient copy its autograd backward makes: tiled relayout kernels between a
 * contiguous (B, P, T) device tensor and the (T*B, P) matrix (row = t*B + b, row stride in
 * elements).  torchain_from2d multiplies by `scale`.  Return 0 when T or B is too large for
 * the tiled kernel (T > 1599 frames, B > 65535); the caller then permutes some other way. */
int torchain_to2d(const float* x_bpt, int64_t num_sequences, int64_t num_pdfs, int64_t num_frames,
                  float* out_rows, int64_t out_stride, int device, void* stream);
int torchain_from2d(const float* rows, int64_t in_stride, int64_t num_sequences, int64_t num_pdfs,
                    int64_t num_frames, float* out_bpt, float scale, int device, void* stream);

/* Host-buffer form (what a non-torch host binding would call): nnet_output,
 * nnet_output_deriv (and xent_output_deriv, may be NULL) are HOST pointers,
 * ideally pinned; the call copies the input to the device, runs the loss and
 * copies derivatives and results back before returning.  Both copies are chunked
 * by whole frames and overlap the forward / backward recurrence; when the numeric
 * guard trips (objf = -10 * weight) the derivative delivered is the guard's
 * (l2 part only), as in the device form. */
int torchain_ComputeChainObjfAndDeriv_host(void* den_graph_ptr, void* supervision_ptr,
                                           const float* nnet_output_host, int64_t rows, int64_t cols,
                                           int64_t nnet_output_stride,
                                           float* results,
                                           float* nnet_output_deriv_host, int64_t deriv_stride,
                                           float* xent_output_deriv_host, int64_t xent_stride,
                                           float l2_regularize, float leaky_hmm_coefficient,
                                           float xent_regularize, int device);
/* The frame boundaries of the chunks the host-buffer form moves a minibatch of num_frames frames in
 * (about 16 chunks of whole frames, single frames at both ends so that the copies that cannot overlap
 * the recurrences are short): writes up to max_entries of [0, ..., num_frames] and returns the number
 * of boundaries (chunks + 1), 0 if num_frames < 1.  Host logic only; no device is touched. */
int torchain_host_chunk_bounds(int num_frames, int* bounds, int max_entries);

#ifdef __cplusplus
}
#endif
#endif /* TORCHAIN_B200_H_ */

// chain_oracle.cpp -- CPU restatement of the LF-MMI ("chain") loss hot path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product.  It
// may be imported/linked/executed only by tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs, and there only as the
// checker or the reported CPU baseline -- never as the path being shipped.
//
// PARITY UNPINNED: the arithmetic of the reference's hot path
// (/root/reference/src/my_lib_chain.cpp:129-131 calling
// kaldi::chain::ComputeChainObjfAndDeriv) lives in Kaldi, an un-vendored
// dependency named only by prose ("kaldi 22fbdd", /root/reference/README.md:3;
// cloned unpinned by /root/reference/Makefile:60-66).  Kaldi, OpenFst and the
// TH/THC binding layer are absent from this image, and the reference ships no
// golden vectors for this path (its test/test.py asserts no values).  This
// file restates the published Kaldi algorithm (CPU branches) and is validated
// by (a) the property assertions the reference's own vendored Kaldi test holds
// (/root/reference/src/chain-supervision-test.hpp:267-285,287-340,417-462,
// 135-149) and (b) an independent dense-matrix float64 autograd formulation in
// tests/dense_ref.py.  See DESIGN.md "Oracle".
//
// Two arithmetic modes:
//   Real=double  "fp64"       : everything in double; the truth for tolerances.
//   Real=float   "kaldi_fp32" : fp32 storage with double accumulators exactly
//                               where Kaldi's CPU branch has them; this is the
//                               timed "Kaldi-CPU-equivalent" baseline.
//
// Build: see oracle/Makefile (g++ -O2 -fopenmp -shared).

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

thread_local std::string g_err;

constexpr double kLogZero = -std::numeric_limits<double>::infinity();
constexpr double kMinLogDiffDouble = -36.04365338911715;  // log(DBL_EPSILON)

// kaldi@22fbdd:src/base/kaldi-math.h LogAdd(double,double)
inline double LogAdd(double x, double y) {
  double diff;
  if (x < y) {
    diff = x - y;
    x = y;
  } else {
    diff = y - x;
  }
  // diff is negative; x is the larger one.
  if (diff >= kMinLogDiffDouble) {
    return x + std::log1p(std::exp(diff));
  }
  return x;  // also covers diff = NaN from (-inf) - (-inf)
}

// ---------------------------------------------------------------------------
// Denominator graph.  kaldi@22fbdd:src/chain/chain-den-graph.{h,cc}
// (constructed by the reference at /root/reference/src/my_lib_example.cpp:129-134)
// ---------------------------------------------------------------------------
struct Transition {  // kaldi@22fbdd:src/chain/chain-datastruct.h DenominatorGraphTransition
  float transition_prob;
  int32_t pdf_id;
  int32_t hmm_state;
};

struct DenGraph {
  int32_t num_states = 0, num_pdfs = 0;
  // out-lists ("forward transitions", hmm_state = destination) and
  // in-lists ("backward transitions", hmm_state = source), CSR by state.
  std::vector<int32_t> fwd_begin, bwd_begin;  // size N+1
  std::vector<Transition> fwd, bwd;
  std::vector<float> initial_probs;            // size N
};

// SetTransitions + SetInitialProbs (chain-den-graph.cc).
bool BuildDenGraph(int32_t N, int32_t A, const int32_t* src, const int32_t* dst,
                   const int32_t* ilabel, const float* cost, int32_t start,
                   const float* final_cost, int32_t num_pdfs, DenGraph* g) {
  if (N <= 0 || A < 0 || start < 0 || start >= N) {
    g_err = "bad den graph sizes";
    return false;
  }
  g->num_states = N;
  g->num_pdfs = num_pdfs;
  std::vector<int32_t> outdeg(N, 0), indeg(N, 0);
  for (int32_t a = 0; a < A; a++) {
    if (src[a] < 0 || src[a] >= N || dst[a] < 0 || dst[a] >= N) {
      g_err = "den arc state out of range";
      return false;
    }
    int32_t pdf = ilabel[a] - 1;  // SetTransitions: pdf_id = arc.ilabel - 1
    if (pdf < 0 || pdf >= num_pdfs) {
      g_err = "den arc label out of range";
      return false;
    }
    outdeg[src[a]]++;
    indeg[dst[a]]++;
  }
  g->fwd_begin.assign(N + 1, 0);
  g->bwd_begin.assign(N + 1, 0);
  for (int32_t s = 0; s < N; s++) {
    g->fwd_begin[s + 1] = g->fwd_begin[s] + outdeg[s];
    g->bwd_begin[s + 1] = g->bwd_begin[s] + indeg[s];
  }
  g->fwd.resize(A);
  g->bwd.resize(A);
  std::vector<int32_t> fpos(g->fwd_begin.begin(), g->fwd_begin.end() - 1),
      bpos(g->bwd_begin.begin(), g->bwd_begin.end() - 1);
  for (int32_t a = 0; a < A; a++) {  // arcs visited in FST order (by source, then arc order)
    Transition t;
    t.transition_prob = std::exp(-cost[a]);  // float exp as in Kaldi (BaseFloat)
    t.pdf_id = ilabel[a] - 1;
    t.hmm_state = dst[a];
    g->fwd[fpos[src[a]]++] = t;
    t.hmm_state = src[a];
    g->bwd[bpos[dst[a]]++] = t;
  }
  // SetInitialProbs: 100 iterations of HMM propagation from the start state,
  // averaged; normalising each state's out-probs (incl. final-prob) to one.
  const int num_iters = 100;
  std::vector<double> norm(N), cur(N, 0.0), next(N, 0.0), avg(N, 0.0);
  for (int32_t s = 0; s < N; s++) {
    norm[s] = std::exp(-(double)final_cost[s]);  // exp(-inf) = 0 for non-final states
  }
  // exp(-arc.weight.Value()) is taken in double on the float cost, as Kaldi does
  for (int32_t a = 0; a < A; a++) norm[src[a]] += std::exp(-(double)cost[a]);
  for (int32_t s = 0; s < N; s++) {
    if (!(norm[s] > 0.0 && norm[s] < 100.0)) {
      g_err = "den graph state total prob out of (0,100)";
      return false;
    }
    norm[s] = 1.0 / norm[s];
  }
  // arcs grouped by source for the propagation
  std::vector<int32_t> order(A);
  {
    std::vector<int32_t> pos(g->fwd_begin.begin(), g->fwd_begin.end() - 1);
    for (int32_t a = 0; a < A; a++) order[pos[src[a]]++] = a;
  }
  cur[start] = 1.0;
  for (int iter = 0; iter < num_iters; iter++) {
    for (int32_t s = 0; s < N; s++) avg[s] += cur[s] * (1.0 / num_iters);
    for (int32_t s = 0; s < N; s++) {
      double prob = cur[s] * norm[s];
      for (int32_t k = g->fwd_begin[s]; k < g->fwd_begin[s + 1]; k++) {
        int32_t a = order[k];
        next[dst[a]] += prob * std::exp(-(double)cost[a]);
      }
    }
    cur.swap(next);
    std::fill(next.begin(), next.end(), 0.0);
    double sum = 0.0;
    for (int32_t s = 0; s < N; s++) sum += cur[s];
    for (int32_t s = 0; s < N; s++) cur[s] *= 1.0 / sum;
  }
  g->initial_probs.resize(N);
  for (int32_t s = 0; s < N; s++) g->initial_probs[s] = (float)avg[s];
  return true;
}

// ---------------------------------------------------------------------------
// Supervision (merged, epsilon-free acceptor over B*T frames).
// kaldi@22fbdd:src/chain/chain-supervision.h; created by the reference at
// /root/reference/src/my_lib_example.cpp:71-76.
// ---------------------------------------------------------------------------
struct Supervision {
  float weight = 1.0f;
  int32_t num_sequences = 0, frames_per_sequence = 0, label_dim = 0;
  int32_t num_states = 0;
  std::vector<int32_t> arc_begin;  // size num_states+1 (arcs grouped by source state, state order)
  std::vector<int32_t> arc_dst, arc_label;
  std::vector<float> arc_cost;
  std::vector<float> final_cost;  // +inf = not final
};

bool BuildSupervision(float weight, int32_t B, int32_t T, int32_t label_dim, int32_t S,
                      const int32_t* arc_begin, const int32_t* arc_dst, const int32_t* arc_label,
                      const float* arc_cost, const float* final_cost, Supervision* sup) {
  if (B <= 0 || T <= 0 || S <= 0 || label_dim <= 0) {
    g_err = "bad supervision sizes";
    return false;
  }
  sup->weight = weight;
  sup->num_sequences = B;
  sup->frames_per_sequence = T;
  sup->label_dim = label_dim;
  sup->num_states = S;
  sup->arc_begin.assign(arc_begin, arc_begin + S + 1);
  int32_t A = arc_begin[S];
  sup->arc_dst.assign(arc_dst, arc_dst + A);
  sup->arc_label.assign(arc_label, arc_label + A);
  sup->arc_cost.assign(arc_cost, arc_cost + A);
  sup->final_cost.assign(final_cost, final_cost + S);
  for (int32_t a = 0; a < A; a++) {
    if (arc_dst[a] < 0 || arc_dst[a] >= S || arc_label[a] < 1 || arc_label[a] > label_dim) {
      g_err = "supervision arc out of range";
      return false;
    }
  }
  return true;
}

// kaldi@22fbdd:src/chain/chain-supervision.cc ComputeFstStateTimes: state 0 has
// time 0, every arc advances time by one, each state must have a unique time,
// final states must sit at the total length.  Returns the total length or -1.
int32_t ComputeFstStateTimes(const Supervision& sup, std::vector<int32_t>* times) {
  int32_t S = sup.num_states;
  times->assign(S, -1);
  (*times)[0] = 0;
  int32_t total = -1;
  for (int32_t s = 0; s < S; s++) {
    int32_t t = (*times)[s];
    if (t < 0) {
      g_err = "supervision FST: state unreachable or not topologically sorted";
      return -1;
    }
    for (int32_t a = sup.arc_begin[s]; a < sup.arc_begin[s + 1]; a++) {
      int32_t d = sup.arc_dst[a];
      if ((*times)[d] == -1) {
        (*times)[d] = t + 1;
      } else if ((*times)[d] != t + 1) {
        g_err = "supervision FST: state has inconsistent times";
        return -1;
      }
    }
    if (sup.final_cost[s] != std::numeric_limits<float>::infinity()) {
      if (total == -1) total = t;
      else if (total != t) {
        g_err = "supervision FST: final states at different times";
        return -1;
      }
    }
  }
  return total;
}

// ---------------------------------------------------------------------------
// NumeratorComputation.  kaldi@22fbdd:src/chain/chain-numerator.{h,cc}
// ---------------------------------------------------------------------------
template <typename Real>
struct Numerator {
  const Supervision& sup;
  const Real* y;  // (T*B) x P, row stride ystride
  int64_t ystride;
  std::vector<int32_t> state_times;
  std::vector<int32_t> fst_output_indexes;              // per arc -> index into nnet_output_indexes
  std::vector<std::pair<int32_t, int32_t>> nnet_output_indexes;  // distinct (row, col)
  std::vector<Real> nnet_logprobs;
  std::vector<double> log_alpha, log_beta;
  double tot_log_prob = kLogZero;

  Numerator(const Supervision& s, const Real* y_, int64_t ys) : sup(s), y(y_), ystride(ys) {}

  // ComputeLookupIndexes (chain-numerator.cc): de-duplicate (row, pdf) pairs
  // per time step; row = sequence + B * frame  (ComputeRowIndex).
  bool ComputeLookupIndexes() {
    int32_t B = sup.num_sequences, T = sup.frames_per_sequence;
    int32_t total = ComputeFstStateTimes(sup, &state_times);
    if (total < 0) return false;
    if (total != B * T) {
      g_err = "supervision FST length != num_sequences * frames_per_sequence";
      return false;
    }
    int32_t S = sup.num_states, A = sup.arc_begin[S];
    fst_output_indexes.resize(A);
    // Kaldi keeps, per time step, a map pdf -> index.  Arcs are visited in
    // state order, states are in time order, so indexes are assigned in
    // first-seen order.
    std::vector<int32_t> map_pdf(sup.label_dim, -1);
    std::vector<int32_t> touched;
    int32_t cur_time = 0;
    for (int32_t s = 0; s < S; s++) {
      int32_t t = state_times[s];
      if (t != cur_time) {
        if (t < cur_time) {
          g_err = "supervision FST: states not in time order";
          return false;
        }
        for (int32_t p : touched) map_pdf[p] = -1;
        touched.clear();
        cur_time = t;
      }
      for (int32_t a = sup.arc_begin[s]; a < sup.arc_begin[s + 1]; a++) {
        int32_t pdf = sup.arc_label[a] - 1;
        if (map_pdf[pdf] == -1) {
          int32_t seq = t / T, frame = t % T;
          int32_t row = seq + B * frame;
          map_pdf[pdf] = (int32_t)nnet_output_indexes.size();
          touched.push_back(pdf);
          nnet_output_indexes.emplace_back(row, pdf);
        }
        fst_output_indexes[a] = map_pdf[pdf];
      }
    }
    return true;
  }

  bool Forward(double* weighted_logprob) {
    if (!ComputeLookupIndexes()) return false;
    nnet_logprobs.resize(nnet_output_indexes.size());
    for (size_t i = 0; i < nnet_output_indexes.size(); i++)  // CuMatrixBase::Lookup
      nnet_logprobs[i] = y[(int64_t)nnet_output_indexes[i].first * ystride + nnet_output_indexes[i].second];
    int32_t S = sup.num_states;
    log_alpha.assign(S, kLogZero);
    tot_log_prob = kLogZero;
    log_alpha[0] = 0.0;
    for (int32_t s = 0; s < S; s++) {
      double this_log_alpha = log_alpha[s];
      for (int32_t a = sup.arc_begin[s]; a < sup.arc_begin[s + 1]; a++) {
        Real transition_logprob = -(Real)sup.arc_cost[a];
        Real pseudo_loglike = nnet_logprobs[fst_output_indexes[a]];
        double& next = log_alpha[sup.arc_dst[a]];
        next = LogAdd(next, (double)(pseudo_loglike + transition_logprob) + this_log_alpha);
      }
      if (sup.final_cost[s] != std::numeric_limits<float>::infinity())  // != TropicalWeight::Zero()
        tot_log_prob = LogAdd(tot_log_prob, this_log_alpha - (double)sup.final_cost[s]);
    }
    *weighted_logprob = tot_log_prob * sup.weight;
    return true;
  }

  // Backward: deriv(row, col) += weight * occupancy.
  void Backward(Real* deriv, int64_t dstride) {
    int32_t S = sup.num_states;
    log_beta.assign(S, kLogZero);
    std::vector<Real> logprob_derivs(nnet_output_indexes.size(), (Real)0);
    for (int32_t s = S - 1; s >= 0; s--) {
      double this_log_alpha = log_alpha[s];
      double this_log_beta = -(double)sup.final_cost[s];  // -inf if not final
      for (int32_t a = sup.arc_begin[s]; a < sup.arc_begin[s + 1]; a++) {
        double next_log_beta = log_beta[sup.arc_dst[a]];
        Real transition_logprob = -(Real)sup.arc_cost[a];
        int32_t index = fst_output_indexes[a];
        Real pseudo_loglike = nnet_logprobs[index];
        this_log_beta = LogAdd(this_log_beta, (double)(pseudo_loglike + transition_logprob) + next_log_beta);
        Real occupation_logprob =
            (Real)((double)(transition_logprob + pseudo_loglike) + this_log_alpha + next_log_beta - tot_log_prob);
        Real occupation_prob = std::exp(occupation_logprob);
        logprob_derivs[index] += occupation_prob;
      }
      log_beta[s] = this_log_beta;
    }
    for (size_t i = 0; i < nnet_output_indexes.size(); i++)  // CuMatrixBase::AddElements(weight, ...)
      deriv[(int64_t)nnet_output_indexes[i].first * dstride + nnet_output_indexes[i].second] +=
          (Real)sup.weight * logprob_derivs[i];
  }
};

// ---------------------------------------------------------------------------
// DenominatorComputation.  kaldi@22fbdd:src/chain/chain-denominator.{h,cc},
// CPU (#else) branches of AlphaGeneralFrame / BetaDashGeneralFrame.
// ---------------------------------------------------------------------------
template <typename Real>
struct Denominator {
  const DenGraph& g;
  int32_t B, T, N, P;
  Real leaky;
  int threads;
  bool clamp_exp;
  std::vector<Real> alpha;     // (T+1) x ((N+1)*B); last B of each row = alpha-sum
  std::vector<Real> beta;      // 2 x ((N+1)*B)
  std::vector<Real> tot_prob;  // B
  std::vector<Real> expyT;     // P x (T*B): element [pdf][t*B+s]
  std::vector<Real> derivT;    // P x (T*B) (Kaldi keeps an 8-frame chunk; layout-equivalent)
  bool ok = true;

  Denominator(const DenGraph& g_, int32_t B_, int32_t T_, Real leaky_, const Real* y, int64_t ystride,
              int threads_, bool clamp)
      : g(g_), B(B_), T(T_), N(g_.num_states), P(g_.num_pdfs), leaky(leaky_), threads(threads_), clamp_exp(clamp) {
    alpha.assign((size_t)(T + 1) * (N + 1) * B, (Real)0);
    beta.assign((size_t)2 * (N + 1) * B, (Real)0);
    tot_prob.assign(B, (Real)0);
    expyT.resize((size_t)P * T * B);
    derivT.assign((size_t)P * T * B, (Real)0);
    // exp_nnet_output_transposed_ = exp(y^T)   (CopyFromMat(kTrans) + ApplyExp)
#pragma omp parallel for num_threads(threads) if (threads > 1)
    for (int32_t p = 0; p < P; p++)
      for (int64_t r = 0; r < (int64_t)T * B; r++) {
        Real v = y[r * ystride + p];
        if (clamp_exp) v = std::min<Real>(std::max<Real>(v, -30), 30);
        expyT[(size_t)p * T * B + r] = std::exp(v);
      }
  }

  Real* AlphaRow(int32_t t) { return alpha.data() + (size_t)t * (N + 1) * B; }

  void AlphaFirstFrame() {
    Real* a = AlphaRow(0);
    for (int32_t h = 0; h < N; h++)
      for (int32_t s = 0; s < B; s++) a[(size_t)h * B + s] = g.initial_probs[h];
  }

  // AlphaDash(t): alpha-sum = sum_h alpha; alpha += leaky * init(h) * alpha-sum.
  void AlphaDash(int32_t t) {
    Real* a = AlphaRow(t);
    Real* sum = a + (size_t)N * B;
    for (int32_t s = 0; s < B; s++) {  // AddRowSumMat (double accumulation in Kaldi's CPU path)
      double acc = 0.0;
      for (int32_t h = 0; h < N; h++) acc += a[(size_t)h * B + s];
      sum[s] = (Real)acc;
    }
    for (int32_t h = 0; h < N; h++) {  // AddVecVec
      Real c = leaky * (Real)g.initial_probs[h];
      for (int32_t s = 0; s < B; s++) a[(size_t)h * B + s] += c * sum[s];
    }
  }

  void AlphaGeneralFrame(int32_t t) {
    const Real* prev = AlphaRow(t - 1);
    Real* cur = AlphaRow(t);
    const Real* prob = expyT.data() + (size_t)(t - 1) * B;  // + pdf * T*B + s
    const size_t pstride = (size_t)T * B;
    if (threads <= 1) {  // Kaldi's loop order: for h, for s, for arcs
      for (int32_t h = 0; h < N; h++) {
        for (int32_t s = 0; s < B; s++) {
          double this_tot_alpha = 0.0;
          for (int32_t k = g.bwd_begin[h]; k < g.bwd_begin[h + 1]; k++) {
            const Transition& tr = g.bwd[k];
            Real p = prob[(size_t)tr.pdf_id * pstride + s];
            Real pa = prev[(size_t)tr.hmm_state * B + s];
            this_tot_alpha += pa * (Real)tr.transition_prob * p;
          }
          Real arbitrary_scale = (Real)1.0 / prev[(size_t)N * B + s];
          cur[(size_t)h * B + s] = (Real)(this_tot_alpha * arbitrary_scale);
        }
      }
    } else {  // same per-(h,s) summation order, parallel over states
#pragma omp parallel for num_threads(threads) schedule(static)
      for (int32_t h = 0; h < N; h++) {
        for (int32_t s = 0; s < B; s++) {
          double this_tot_alpha = 0.0;
          for (int32_t k = g.bwd_begin[h]; k < g.bwd_begin[h + 1]; k++) {
            const Transition& tr = g.bwd[k];
            Real p = prob[(size_t)tr.pdf_id * pstride + s];
            Real pa = prev[(size_t)tr.hmm_state * B + s];
            this_tot_alpha += pa * (Real)tr.transition_prob * p;
          }
          Real arbitrary_scale = (Real)1.0 / prev[(size_t)N * B + s];
          cur[(size_t)h * B + s] = (Real)(this_tot_alpha * arbitrary_scale);
        }
      }
    }
  }

  double ComputeTotLogLike() {
    const Real* last = AlphaRow(T);
    double tot_log_prob = 0.0;
    for (int32_t s = 0; s < B; s++) {
      double acc = 0.0;
      for (int32_t h = 0; h < N; h++) acc += last[(size_t)h * B + s];
      tot_prob[s] = (Real)acc;
      tot_log_prob += std::log((Real)acc);
    }
    double log_inv_scales = 0.0;
    for (int32_t t = 0; t < T; t++) {
      const Real* sum = AlphaRow(t) + (size_t)N * B;
      for (int32_t s = 0; s < B; s++) log_inv_scales += std::log(sum[s]);
    }
    return tot_log_prob + log_inv_scales;
  }

  double Forward() {
    AlphaFirstFrame();
    AlphaDash(0);
    for (int32_t t = 1; t <= T; t++) {
      AlphaGeneralFrame(t);
      AlphaDash(t);
    }
    return ComputeTotLogLike();
  }

  Real* BetaRow(int32_t t) { return beta.data() + (size_t)(t % 2) * (N + 1) * B; }

  void BetaDashLastFrame() {
    Real* b = BetaRow(T);
    for (int32_t h = 0; h < N; h++)
      for (int32_t s = 0; s < B; s++) b[(size_t)h * B + s] = (Real)1.0 / tot_prob[s];
  }

  // Beta(t): beta-dash-sum(s) = leaky * sum_h init(h) beta-dash(h,s); beta = beta-dash + that.
  void Beta(int32_t t) {
    Real* b = BetaRow(t);
    Real* sum = b + (size_t)N * B;
    for (int32_t s = 0; s < B; s++) {
      double acc = 0.0;
      for (int32_t h = 0; h < N; h++) acc += (double)b[(size_t)h * B + s] * g.initial_probs[h];
      sum[s] = (Real)(leaky * acc);
    }
    for (int32_t h = 0; h < N; h++)
      for (int32_t s = 0; s < B; s++) b[(size_t)h * B + s] += sum[s];
  }

  void BetaDashGeneralFrame(int32_t t) {
    const Real* this_alpha_dash = AlphaRow(t);
    const Real* next_beta = BetaRow(t + 1);
    Real* this_beta_dash = BetaRow(t);
    const Real* prob = expyT.data() + (size_t)t * B;
    Real* dprob = derivT.data() + (size_t)t * B;
    const size_t pstride = (size_t)T * B;
    auto body = [&](int32_t h, int32_t s) {
      Real this_alpha_dash_prob = this_alpha_dash[(size_t)h * B + s],
           inv_arbitrary_scale = this_alpha_dash[(size_t)N * B + s];
      double tot_variable_factor = 0.0;
      Real occupation_factor = this_alpha_dash_prob / inv_arbitrary_scale;
      for (int32_t k = g.fwd_begin[h]; k < g.fwd_begin[h + 1]; k++) {
        const Transition& tr = g.fwd[k];
        Real variable_factor =
            (Real)tr.transition_prob * next_beta[(size_t)tr.hmm_state * B + s] * prob[(size_t)tr.pdf_id * pstride + s];
        tot_variable_factor += variable_factor;
        Real occupation_prob = variable_factor * occupation_factor;
        dprob[(size_t)tr.pdf_id * pstride + s] += occupation_prob;
      }
      this_beta_dash[(size_t)h * B + s] = (Real)(tot_variable_factor / inv_arbitrary_scale);
    };
    if (threads <= 1) {
      for (int32_t h = 0; h < N; h++)
        for (int32_t s = 0; s < B; s++) body(h, s);
    } else {
      // parallel over sequences: every (pdf, s) accumulator is touched by one
      // thread only and in the same (h, arc) order as the serial loop.
#pragma omp parallel for num_threads(threads) schedule(static)
      for (int32_t s = 0; s < B; s++)
        for (int32_t h = 0; h < N; h++) body(h, s);
    }
  }

  // BetaGeneralFrameDebug(t): alpha-dash . beta-dash ~= B and sum of this
  // frame's occupancies ~= B; off by more than 2 => ok = false.
  void BetaGeneralFrameDebug(int32_t t) {
    const Real* a = AlphaRow(t);
    const Real* b = BetaRow(t);
    double ab = 0.0;
    for (size_t i = 0; i < (size_t)N * B; i++) ab += (double)a[i] * b[i];
    double dsum = 0.0;
    for (int32_t p = 0; p < P; p++)
      for (int32_t s = 0; s < B; s++) dsum += derivT[(size_t)p * T * B + (size_t)t * B + s];
    if (!(std::fabs(ab - B) <= 2.0)) ok = false;
    if (!(std::fabs(dsum - B) <= 2.0)) ok = false;
  }

  // Backward: deriv += deriv_weight * occupancies.  (Kaldi commits every
  // kMaxDerivTimeSteps=8 frames from a transposed chunk; same arithmetic.)
  bool Backward(Real deriv_weight, Real* deriv, int64_t dstride) {
    BetaDashLastFrame();
    Beta(T);
    for (int32_t t = T - 1; t >= 0; t--) {
      BetaDashGeneralFrame(t);
      if (t == 0) BetaGeneralFrameDebug(t);
      Beta(t);
    }
#pragma omp parallel for num_threads(threads) if (threads > 1)
    for (int64_t r = 0; r < (int64_t)T * B; r++)
      for (int32_t p = 0; p < P; p++) deriv[r * dstride + p] += deriv_weight * derivT[(size_t)p * T * B + r];
    return ok;
  }
};

// ---------------------------------------------------------------------------
// ComputeChainObjfAndDeriv.  kaldi@22fbdd:src/chain/chain-training.cc; called
// by the reference at /root/reference/src/my_lib_chain.cpp:129-131.
// ---------------------------------------------------------------------------
template <typename Real>
bool ComputeChainObjfAndDeriv(const DenGraph& g, const Supervision& sup, const Real* y, int64_t rows, int64_t cols,
                              int64_t ystride, double l2_regularize, double leaky, int threads, bool clamp_exp,
                              double* results /* objf, l2_term, weight, num_logprob_weighted, den_logprob, ok */,
                              Real* deriv, int64_t dstride, Real* xent_deriv, int64_t xstride) {
  int32_t B = sup.num_sequences, T = sup.frames_per_sequence;
  if (rows != (int64_t)B * T || cols != g.num_pdfs || sup.label_dim != g.num_pdfs) {
    g_err = "nnet_output shape does not match supervision / den graph";
    return false;
  }
  if (!(leaky > 0.0 && leaky < 1.0)) {  // KALDI_ASSERT in DenominatorComputation ctor
    g_err = "leaky_hmm_coefficient must be in (0,1)";
    return false;
  }
  if (deriv)
    for (int64_t r = 0; r < rows; r++) std::memset(deriv + r * dstride, 0, sizeof(Real) * cols);
  double num_logprob_weighted = 0.0;
  {
    Numerator<Real> num(sup, y, ystride);
    if (!num.Forward(&num_logprob_weighted)) return false;
    if (deriv) {
      num.Backward(deriv, dstride);
      if (xent_deriv)
        for (int64_t r = 0; r < rows; r++)
          std::memcpy(xent_deriv + r * xstride, deriv + r * dstride, sizeof(Real) * cols);
    } else if (xent_deriv) {
      for (int64_t r = 0; r < rows; r++) std::memset(xent_deriv + r * xstride, 0, sizeof(Real) * cols);
      num.Backward(xent_deriv, xstride);
    }
  }
  Denominator<Real> den(g, B, T, (Real)leaky, y, ystride, threads, clamp_exp);
  double den_logprob = den.Forward();
  bool ok = true;
  if (deriv) ok = den.Backward(-(Real)sup.weight, deriv, dstride);
  // Kaldi holds these as BaseFloat
  Real objf = (Real)(num_logprob_weighted - (double)sup.weight * den_logprob);
  Real weight = (Real)((double)sup.weight * B * T);
  if (!(objf - objf == 0) || !ok) {
    if (deriv)
      for (int64_t r = 0; r < rows; r++) std::memset(deriv + r * dstride, 0, sizeof(Real) * cols);
    if (xent_deriv)
      for (int64_t r = 0; r < rows; r++) std::memset(xent_deriv + r * xstride, 0, sizeof(Real) * cols);
    objf = (Real)-10.0 * weight;
  }
  Real l2_term = 0;
  if (l2_regularize != 0.0) {
    Real scale = (Real)((double)sup.weight * l2_regularize);
    double tr = 0.0;  // TraceMatMat(y, y, kTrans)
    for (int64_t r = 0; r < rows; r++)
      for (int64_t c = 0; c < cols; c++) tr += (double)y[r * ystride + c] * y[r * ystride + c];
    l2_term = (Real)(-0.5 * scale * tr);
    if (deriv)
      for (int64_t r = 0; r < rows; r++)
        for (int64_t c = 0; c < cols; c++) deriv[r * dstride + c] += -scale * y[r * ystride + c];
  }
  results[0] = objf;
  results[1] = l2_term;
  results[2] = weight;
  results[3] = num_logprob_weighted;
  results[4] = den_logprob;
  results[5] = ok ? 1.0 : 0.0;
  return true;
}

}  // namespace

// ---------------------------------------------------------------------------
// C ABI (ctypes from tests/ and bench.py only)
// ---------------------------------------------------------------------------
extern "C" {

const char* oracle_last_error() { return g_err.c_str(); }

int oracle_max_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void* oracle_den_graph_new(int32_t N, int32_t A, const int32_t* src, const int32_t* dst, const int32_t* ilabel,
                           const float* cost, int32_t start, const float* final_cost, int32_t num_pdfs) {
  auto* g = new DenGraph();
  if (!BuildDenGraph(N, A, src, dst, ilabel, cost, start, final_cost, num_pdfs, g)) {
    delete g;
    return nullptr;
  }
  return g;
}
void oracle_den_graph_free(void* g) { delete static_cast<DenGraph*>(g); }
void oracle_den_graph_initial_probs(void* g, float* out) {
  auto* d = static_cast<DenGraph*>(g);
  std::memcpy(out, d->initial_probs.data(), sizeof(float) * d->num_states);
}

void* oracle_supervision_new(float weight, int32_t B, int32_t T, int32_t label_dim, int32_t S,
                             const int32_t* arc_begin, const int32_t* arc_dst, const int32_t* arc_label,
                             const float* arc_cost, const float* final_cost) {
  auto* s = new Supervision();
  if (!BuildSupervision(weight, B, T, label_dim, S, arc_begin, arc_dst, arc_label, arc_cost, final_cost, s)) {
    delete s;
    return nullptr;
  }
  return s;
}
void oracle_supervision_free(void* s) { delete static_cast<Supervision*>(s); }

// mode 0: kaldi_fp32 (y, deriv are float);  mode 1: fp64 (y, deriv are double).
// results: double[6] = objf, l2_term, weight, num_logprob_weighted, den_logprob, ok.
int oracle_chain_f32(void* g, void* sup, const float* y, int64_t rows, int64_t cols, int64_t ystride,
                     double l2, double leaky, int threads, int clamp_exp, double* results, float* deriv,
                     int64_t dstride, float* xent_deriv, int64_t xstride) {
  return ComputeChainObjfAndDeriv<float>(*static_cast<DenGraph*>(g), *static_cast<Supervision*>(sup), y, rows, cols,
                                         ystride, l2, leaky, threads, clamp_exp != 0, results, deriv, dstride,
                                         xent_deriv, xstride)
             ? 1
             : 0;
}
int oracle_chain_f64(void* g, void* sup, const double* y, int64_t rows, int64_t cols, int64_t ystride,
                     double l2, double leaky, int threads, int clamp_exp, double* results, double* deriv,
                     int64_t dstride, double* xent_deriv, int64_t xstride) {
  return ComputeChainObjfAndDeriv<double>(*static_cast<DenGraph*>(g), *static_cast<Supervision*>(sup), y, rows, cols,
                                          ystride, l2, leaky, threads, clamp_exp != 0, results, deriv, dstride,
                                          xent_deriv, xstride)
             ? 1
             : 0;
}

// Denominator only (ChainDenominatorTest, /root/reference/src/chain-supervision-test.hpp:388-463):
// forward log-prob and d(logprob)/dy with deriv_weight = +1.
int oracle_den_only_f64(void* gp, int32_t B, int32_t T, const double* y, int64_t ystride, double leaky,
                        double* logprob, double* deriv, int64_t dstride) {
  const DenGraph& g = *static_cast<DenGraph*>(gp);
  Denominator<double> den(g, B, T, leaky, y, ystride, 1, false);
  *logprob = den.Forward();
  if (deriv) {
    for (int64_t r = 0; r < (int64_t)B * T; r++) std::memset(deriv + r * dstride, 0, sizeof(double) * g.num_pdfs);
    return den.Backward(1.0, deriv, dstride) ? 1 : 0;
  }
  return 1;
}

}  // extern "C"

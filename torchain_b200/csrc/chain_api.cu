// chain_api.cu -- the C ABI of libtorchain_b200.so (include/torchain_b200.h):
// handle types, host-side graph preparation, workspace management and the
// ComputeChainObjfAndDeriv entry points.
//
// Mirrors, for the chain-loss path only, the reference's cffi surface
// /root/reference/src/my_lib.h:19-48 and its glue
// /root/reference/src/my_lib_chain.cpp:104-136, /root/reference/src/my_lib_example.cpp:71-95,129-138.
// All arithmetic that Kaldi did underneath is in den_kernel.cu / num_kernel.cu;
// the only host arithmetic here is the one-off DenominatorGraph preparation
// (transition probabilities and initial-state probabilities).
// There is no CPU implementation of the loss in this library.

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <limits>
#include <map>
#include <memory>
#include <mutex>
#include <queue>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/torchain_b200.h"
#include "../../include/torchain_b200_profile.h"
#include "chain_kernels.h"

namespace {

thread_local std::string g_error;
std::atomic<int> g_profile{0};
std::mutex g_prof_mu;
double g_prof_den_ms = 0.0;
int64_t g_prof_den_launches = 0;
int fail(const std::string& msg) {
  g_error = msg;
  return 0;
}
void* fail_ptr(const std::string& msg) {
  g_error = msg;
  return nullptr;
}
#define CUDA_OK(expr)                                                                         \
  do {                                                                                        \
    cudaError_t e__ = (expr);                                                                 \
    if (e__ != cudaSuccess) {                                                                 \
      return fail(std::string(#expr) + ": " + cudaGetErrorString(e__));                       \
    }                                                                                         \
  } while (0)

const float kInf = std::numeric_limits<float>::infinity();

// ---------------------------------------------------------------------------
// device buffer helpers
// ---------------------------------------------------------------------------
struct DevBuf {
  void* ptr = nullptr;
  size_t bytes = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    bytes = 0;
  }
  void abandon() {  // forget the allocation without touching the runtime (process / thread teardown)
    ptr = nullptr;
    bytes = 0;
  }
  // Grows the buffer (contents are not kept).  The old allocation goes first: workspaces are hundreds of MB and
  // holding both would double the peak; a failed grow therefore leaves the buffer EMPTY (ptr == nullptr, bytes == 0),
  // consistently, and the next call simply allocates again.  cudaErrorMemoryAllocation reaches the caller as
  // "out of memory" in torchain_last_error(); the Python op then frees torch's cached blocks and retries once.
  cudaError_t reserve(size_t n) {
    if (n <= bytes) return cudaSuccess;
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    bytes = 0;
    cudaError_t e = cudaMalloc(&ptr, n);
    if (e == cudaSuccess) bytes = n;
    else ptr = nullptr;
    return e;
  }
  template <typename T>
  T* as() const {
    return static_cast<T*>(ptr);
  }
};

template <typename T>
cudaError_t upload(DevBuf& b, const std::vector<T>& v) {
  cudaError_t e = b.reserve(std::max<size_t>(sizeof(T) * v.size(), 16));
  if (e != cudaSuccess) return e;
  if (v.empty()) return cudaSuccess;
  return cudaMemcpy(b.ptr, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice);
}

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) ok = false;
    if (ok && prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// ---------------------------------------------------------------------------
// Denominator graph handle
// ---------------------------------------------------------------------------
struct DirHost {  // one direction, partitioned
  std::vector<int4> arcs, tasks;
  std::vector<int> cta_arc_off, cta_task_off, warp_task_off, warp_arc_off;
  int max_arcs_cta = 0, max_tasks_cta = 0;
  bool has_split = false;
};

struct DenPartition {  // one uploaded partition of both directions
  DevBuf arcs[2], tasks[2], cta_arc_off[2], cta_task_off[2], warp_task_off[2], warp_arc_off[2];
  int max_arcs_cta[2] = {0, 0}, max_tasks_cta[2] = {0, 0};
  bool has_split = false;
  std::vector<double> work[2];  // [grid*kDenMainWarps] cost units assigned to each warp
};

struct TiedPartition {  // the run-based kernels' streams (uploaded once per device; not autotuned)
  DevBuf arcs[2], tasks[2], cta_arc_off[2], cta_task_off[2], warp_task_off[2], warp_arc_off[2];
  int max_arcs_cta[2] = {0, 0}, max_tasks_cta[2] = {0, 0};
  DevBuf runs, cta_run_off, warp_run_off;
  int max_runs_cta = 0;
};

struct DenDevice {  // per (device) state
  int grid = 0;
  std::unique_ptr<TiedPartition> tied;  // built on first use if the graph qualifies
  // Partitions are never freed or overwritten while the handle lives: launches already
  // enqueued on other streams keep reading the generation they were given.
  std::vector<std::unique_ptr<DenPartition>> generations;
  DenPartition* cur = nullptr;              // subs = 1 (batches of 128+ sequences; autotuned)
  DenPartition* cur_sub[2] = {nullptr, nullptr};  // subs = 2 / 4 (BP = 64 / 32), built on first use
  // measured relative speed of every (CTA, main warp) bin per direction; 1.0 until calibrated
  std::vector<double> speed[2];
  int calib_rounds_done = 0;
  DevBuf warp_busy;  // [2][grid][kDenWarps] cycles, written by the last launch pair
  // A call that does not synchronise (no host results) cannot read the counters back itself: it
  // records this event behind its launch pair and a later call harvests the round once it completed.
  cudaEvent_t calib_ev = nullptr;
  bool calib_pending = false;
  const DenPartition* calib_part = nullptr;  // the partition the pending counters were measured on
  ~DenDevice() {
    if (calib_ev) cudaEventDestroy(calib_ev);
  }
};
constexpr int kCalibRounds = 3;

struct Workspace {
  DevBuf alpha, beta, tiles, zeroed, num_scratch, scratch_deriv, cta_trace, arun, wbuf;
  int trace_frames = 0, trace_grid = 0;
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  float* pinned = nullptr;  // 4 floats
  char* timeline_dev = nullptr;
  int timeline_len = 0;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof;  // around each denominator launch (profiling only)
  // A workspace (alpha store, zeroed scalars and barrier counters, numerator scratch, the pinned result
  // words, the fork / join events) serves ONE call at a time: two host threads that call on the same
  // (device, stream) -- the legacy default stream, say -- must not interleave their memsets and launches.
  // Held from a call's first enqueue to its last enqueue (or, for a call that delivers results to the
  // host, its stream synchronisation); work already enqueued is ordered by the stream itself.
  std::mutex call_mu;
  int in_use = 0;                // calls currently between get_workspace and their return (guarded by DenGraph::mu)
  unsigned long long last_use = 0;  // DenGraph::ws_clock at the last get_workspace
  ~Workspace() {
    for (auto& pr : prof) {
      cudaEventDestroy(pr.first);
      cudaEventDestroy(pr.second);
    }
    if (side) cudaStreamDestroy(side);
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_join) cudaEventDestroy(ev_join);
    if (pinned) cudaFreeHost(pinned);
  }
};

// Test hooks, read from the environment ONCE per handle (when the graph is created), never per call:
//   TORCHAIN_SPLIT_CAP=n        states with more than n arcs in a direction are split into sub-tasks (default 192)
//   TORCHAIN_FORCE_UNSTAGED=1   arc slices stay in global memory even when they would fit shared memory
//   TORCHAIN_NO_AUTOTUNE=1      the partition is not re-dealt from measured per-warp busy cycles
//   TORCHAIN_TIED=0|1           never / always use the run-based kernels (default: when the graph has few runs per arc)
//   TORCHAIN_ALIGN_STREAMS=0    small batches: the lane groups of a warp get independently dealt streams (the A/B switch
//                               of the aligned-stream partition, see build_dir)
struct GraphHooks {
  int split_cap = 192;
  int tied = -1;
  bool force_unstaged = false, no_autotune = false, align_streams = true;
  static GraphHooks from_env() {
    GraphHooks h;
    if (const char* e = getenv("TORCHAIN_SPLIT_CAP")) h.split_cap = std::max(2, atoi(e));
    h.force_unstaged = getenv("TORCHAIN_FORCE_UNSTAGED") != nullptr;
    h.no_autotune = getenv("TORCHAIN_NO_AUTOTUNE") != nullptr;
    if (const char* e = getenv("TORCHAIN_TIED")) h.tied = atoi(e) != 0;
    if (const char* e = getenv("TORCHAIN_ALIGN_STREAMS")) h.align_streams = atoi(e) != 0;
    return h;
  }
};

struct DenGraph {
  GraphHooks hooks;
  int N = 0, A = 0, P = 0, start = 0;
  std::vector<int> src, dst, pdf;
  std::vector<float> prob;
  std::vector<float> pi;
  double sum_pi = 0.0;
  // runs = distinct (destination state, pdf) pairs, numbered in (state, pdf) order; run_of[a] = run of arc a.
  // A graph whose pdfs are tied to the state being entered (a real den.fst) has ~2 runs per state; the
  // run-based kernels (den_kernel.cu, "tied") are used when runs are few against arcs.
  int R = 0;
  std::vector<int> run_of, run_dst, run_pdf;
  bool tied = false;
  std::mutex mu;
  std::map<int, std::unique_ptr<DenDevice>> dev;
  std::map<std::pair<int, void*>, std::unique_ptr<Workspace>> ws;
  unsigned long long ws_clock = 0;
};
// A workspace holds the alpha store (261 MB at config 2).  Callers that hop between streams (torch hands
// out 32 pooled streams per device) must not accumulate one per stream: beyond this many per device the
// least recently used idle one is dropped (its buffers are freed, which waits for the device).
constexpr int kMaxWorkspacesPerDevice = 4;

// Longest-processing-time-first assignment of weighted items to bins of different speeds:
// each item (heaviest first) goes to the bin that would finish it earliest.
void lpt(const std::vector<int>& items, const std::vector<int>& cost, const std::vector<double>& speed,
         std::vector<std::vector<int>>* out) {
  const int bins = (int)speed.size();
  out->assign(bins, {});
  std::vector<int> order(items);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost[a] > cost[b]; });
  std::vector<double> load(bins, 0.0);
  for (int it : order) {
    int best = 0;
    double best_t = 1e300;
    for (int b = 0; b < bins; b++) {
      const double t = (load[b] + cost[it]) / speed[b];
      if (t < best_t) {
        best_t = t;
        best = b;
      }
    }
    (*out)[best].push_back(it);
    load[best] += cost[it];
  }
}

// Partition one direction over grid CTAs x kDenMainWarps warps.
//   dir 0: tasks = destination states with their in-arcs (other = source)
//   dir 1: tasks = source states with their out-arcs     (other = destination)
void build_dir(const DenGraph& g, int dir, int grid, int bins_per_cta, const std::vector<double>& speed /* [grid*bins_per_cta] */,
               unsigned rowbytes /* bytes of one [state|pdf] row of this geometry: 4 * BP */, DirHost* out,
               std::vector<double>* work, const std::vector<int>* x_index = nullptr /* per arc: row index stored in
               the record's first field instead of the other state (tied phase B: the arc's run); never split */,
               int align = 1 /* lane groups per warp (small batches): see "aligned streams" below */) {
  const int N = g.N, A = g.A;
  std::vector<int> deg(N, 0), begin(N + 1, 0);
  const std::vector<int>& key = dir == 0 ? g.dst : g.src;
  const std::vector<int>& other = dir == 0 ? g.src : g.dst;
  for (int a = 0; a < A; a++) deg[key[a]]++;
  for (int s = 0; s < N; s++) begin[s + 1] = begin[s] + deg[s];
  std::vector<int> arc_of(A), pos(begin.begin(), begin.end() - 1);
  for (int a = 0; a < A; a++) arc_of[pos[key[a]]++] = a;  // stable: FST arc order within a state

  const int kTaskOverhead = 3;  // fixed per-state cost in arc units (prologue, store, partial sums)
  // States with very many arcs (real den graphs are heavy-tailed) would be one serial warp task;
  // they are split into sub-tasks of at most `cap` arcs that accumulate with REDs.
  const int cap = x_index ? std::numeric_limits<int>::max() : g.hooks.split_cap;
  struct Item {
    int state, first, count, flag;  // arcs [first, first+count) of the state's CSR slice
  };
  std::vector<Item> items;
  out->has_split = false;
  for (int s = 0; s < N; s++) {
    if (deg[s] <= cap) {
      items.push_back({s, 0, deg[s], 0});
    } else {
      out->has_split = true;
      const int k = (deg[s] + cap - 1) / cap, per = (deg[s] + k - 1) / k;
      for (int j = 0, first = 0; j < k; j++, first += per)
        items.push_back({s, first, std::min(per, deg[s] - first), j == 0 ? 2 : 1});
    }
  }
  const int NI = (int)items.size();
  std::vector<int> cost(NI), all(NI);
  for (int i = 0; i < NI; i++) {
    cost[i] = items[i].count + kTaskOverhead;
    all[i] = i;
  }
  std::vector<std::vector<int>> by_cta;
  std::vector<double> cta_speed(grid, 0.0);
  for (int c = 0; c < grid; c++)
    for (int w = 0; w < bins_per_cta; w++) cta_speed[c] += speed[c * bins_per_cta + w];
  lpt(all, cost, cta_speed, &by_cta);
  work->assign((size_t)grid * bins_per_cta, 0.0);

  out->arcs.clear();
  out->tasks.clear();
  out->cta_arc_off.assign(1, 0);
  out->cta_task_off.assign(1, 0);
  out->warp_task_off.assign(1, 0);
  out->warp_arc_off.assign(1, 0);
  out->max_arcs_cta = out->max_tasks_cta = 0;
  // Aligned streams (align = lane groups per warp > 1): the groups of a warp walk their own streams in lockstep, so
  // an end of state in one group sends the whole warp through the end-of-state block and keeps it off the
  // straight-line path of the arc loop.  States of (nearly) equal arc count are therefore put side by side in the
  // `align` streams of one warp and padded with zero-probability arcs to the same length: every end of state is then
  // warp-uniform, as it is with one stream per warp.  (Arc counts are Poisson-like, so neighbours in sorted order
  // differ by 0-1 arcs: the padding is ~1 % of the arcs.)
  struct Slot {
    int item, pad_to;  // pad_to = arcs this item occupies in its stream (>= its own count)
  };
  for (int c = 0; c < grid; c++) {
    std::vector<std::vector<Slot>> by_bin(bins_per_cta);
    if (align <= 1) {
      std::vector<std::vector<int>> by_warp;
      lpt(by_cta[c], cost,
          std::vector<double>(speed.begin() + c * bins_per_cta, speed.begin() + (c + 1) * bins_per_cta), &by_warp);
      for (int w = 0; w < bins_per_cta; w++)
        for (int it : by_warp[w]) by_bin[w].push_back({it, 0});
    } else {
      std::vector<int> order(by_cta[c]);
      std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost[a] > cost[b]; });
      const int ngroups = ((int)order.size() + align - 1) / align, warps = bins_per_cta / align;
      std::vector<int> gcost(ngroups), gid(ngroups);
      for (int gi = 0; gi < ngroups; gi++) {
        gcost[gi] = cost[order[(size_t)gi * align]];  // the first member is the largest
        gid[gi] = gi;
      }
      std::vector<double> wspeed(warps, 0.0);
      for (int w = 0; w < warps; w++)
        for (int j = 0; j < align; j++) wspeed[w] += speed[c * bins_per_cta + w * align + j];
      std::vector<std::vector<int>> by_warp;
      lpt(gid, gcost, wspeed, &by_warp);
      for (int w = 0; w < warps; w++)
        for (int gi : by_warp[w]) {
          const int pad_to = std::max(1, items[order[(size_t)gi * align]].count);
          for (int j = 0; j < align && (size_t)gi * align + j < order.size(); j++)
            by_bin[w * align + j].push_back({order[(size_t)gi * align + j], pad_to});
        }
    }
    const int arc0 = (int)out->arcs.size(), task0 = (int)out->tasks.size();
    for (int w = 0; w < bins_per_cta; w++) {
      for (const Slot& slot : by_bin[w]) {
        const int it = slot.item;
        const Item& item = items[it];
        const int s = item.state;
        (*work)[c * bins_per_cta + w] += cost[it];
        const size_t item_arc0 = out->arcs.size();
        int4 tk;
        tk.x = s;
        tk.y = (int)out->arcs.size() - arc0;
        tk.z = item.flag;
        float pis = g.pi[s];
        std::memcpy(&tk.w, &pis, 4);
        out->tasks.push_back(tk);
        for (int k = begin[s] + item.first; k < begin[s] + item.first + item.count; k++) {
          const int a = arc_of[k];
          int4 ar;  // state / pdf stored as row BYTE offsets: the gather address is base + offset
          ar.x = (int)((unsigned)(x_index ? (*x_index)[a] : other[a]) * rowbytes);
          ar.y = (int)((unsigned)g.pdf[a] * rowbytes);
          float pr = g.prob[a], pio = g.pi[other[a]];
          std::memcpy(&ar.z, &pr, 4);
          std::memcpy(&ar.w, &pio, 4);
          out->arcs.push_back(ar);
        }
        if (item.count == 0) {  // a state without arcs still needs its end-of-state marker
          int4 ar;
          ar.x = x_index ? 0 : (int)((unsigned)s * rowbytes);
          ar.y = 0;
          ar.z = 0;  // probability 0.0f
          ar.w = 0;
          out->arcs.push_back(ar);
        }
        while ((int)(out->arcs.size() - item_arc0) < slot.pad_to) {  // aligned streams: same length as the neighbours
          int4 ar = out->arcs.back();
          ar.z = 0;
          ar.w = 0;
          out->arcs.push_back(ar);
        }
        out->arcs.back().z |= (int)0x80000000u;  // end-of-state flag = sign bit of the (non-negative) probability
      }
      // pad the stream to a multiple of kStreamPad arcs: zero-probability arcs without an end flag,
      // pointing at rows the stream touches anyway (they add 0 to an accumulator nobody reads)
      const int stream_begin = out->warp_arc_off.back();
      while (((int)out->arcs.size() - stream_begin) % tb::kStreamPad != 0) {
        int4 ar = out->arcs.size() > (size_t)stream_begin ? out->arcs.back() : make_int4(0, 0, 0, 0);
        ar.z = 0;  // +0.0f: no probability, no end-of-state flag
        ar.w = 0;
        out->arcs.push_back(ar);
      }
      out->warp_task_off.push_back((int)out->tasks.size());
      out->warp_arc_off.push_back((int)out->arcs.size());
    }
    out->cta_arc_off.push_back((int)out->arcs.size());
    out->cta_task_off.push_back((int)out->tasks.size());
    out->max_arcs_cta = std::max(out->max_arcs_cta, (int)out->arcs.size() - arc0);
    out->max_tasks_cta = std::max(out->max_tasks_cta, (int)out->tasks.size() - task0);
  }
}

// Kaldi DenominatorGraph ctor: SetTransitions + SetInitialProbs.
DenGraph* make_den_graph(int32_t N, int32_t A, const int32_t* src, const int32_t* dst, const int32_t* ilabel,
                         const float* cost, int32_t start, const float* final_cost, int32_t num_pdfs) {
  if (N <= 0 || A < 0 || num_pdfs <= 0 || start < 0 || start >= N) {
    g_error = "denominator graph: bad sizes / start state";
    return nullptr;
  }
  auto g = std::make_unique<DenGraph>();
  g->hooks = GraphHooks::from_env();
  g->N = N;
  g->A = A;
  g->P = num_pdfs;
  g->start = start;
  g->src.assign(src, src + A);
  g->dst.assign(dst, dst + A);
  g->pdf.resize(A);
  g->prob.resize(A);
  std::vector<double> tot(N);
  for (int s = 0; s < N; s++) tot[s] = std::exp(-(double)final_cost[s]);
  for (int a = 0; a < A; a++) {
    if (src[a] < 0 || src[a] >= N || dst[a] < 0 || dst[a] >= N) {
      g_error = "denominator graph: arc endpoint out of range";
      return nullptr;
    }
    if (ilabel[a] < 1 || ilabel[a] > num_pdfs) {
      g_error = "denominator graph: arc label not in [1, num_pdfs]";
      return nullptr;
    }
    g->pdf[a] = ilabel[a] - 1;
    g->prob[a] = std::exp(-cost[a]);
    tot[src[a]] += std::exp(-(double)cost[a]);
  }
  for (int s = 0; s < N; s++) {
    if (!(tot[s] > 0.0 && tot[s] < 100.0)) {
      g_error = "denominator graph: a state's total outgoing probability is not in (0, 100)";
      return nullptr;
    }
  }
  // initial probabilities: average of 100 propagation steps from the start state
  std::vector<int> begin(N + 1, 0);
  for (int a = 0; a < A; a++) begin[src[a] + 1]++;
  for (int s = 0; s < N; s++) begin[s + 1] += begin[s];
  std::vector<int> by_src(A), pos(begin.begin(), begin.end() - 1);
  for (int a = 0; a < A; a++) by_src[pos[src[a]]++] = a;
  std::vector<double> w(A);
  for (int a = 0; a < A; a++) w[a] = std::exp(-(double)cost[a]) / tot[src[a]];
  std::vector<double> cur(N, 0.0), nxt(N, 0.0), avg(N, 0.0);
  cur[start] = 1.0;
  const int iters = 100;
  for (int it = 0; it < iters; it++) {
    for (int s = 0; s < N; s++) avg[s] += cur[s] / iters;
    for (int s = 0; s < N; s++) {
      if (cur[s] == 0.0) continue;
      for (int k = begin[s]; k < begin[s + 1]; k++) nxt[dst[by_src[k]]] += cur[s] * w[by_src[k]];
    }
    double sum = 0.0;
    for (int s = 0; s < N; s++) sum += nxt[s];
    if (!(sum > 0.0) || !std::isfinite(sum)) {  // Kaldi would go on with NaN initial probabilities
      g_error = "denominator graph: the states reachable from the start state have no outgoing arcs";
      return nullptr;
    }
    for (int s = 0; s < N; s++) {
      cur[s] = nxt[s] / sum;
      nxt[s] = 0.0;
    }
  }
  g->pi.resize(N);
  g->sum_pi = 0.0;
  for (int s = 0; s < N; s++) {
    g->pi[s] = (float)avg[s];
    g->sum_pi += (double)g->pi[s];
  }
  // runs: arcs sorted by (destination, pdf)
  {
    std::vector<int> order(A);
    for (int a = 0; a < A; a++) order[a] = a;
    std::sort(order.begin(), order.end(), [&](int x, int y) {
      if (g->dst[x] != g->dst[y]) return g->dst[x] < g->dst[y];
      if (g->pdf[x] != g->pdf[y]) return g->pdf[x] < g->pdf[y];
      return x < y;
    });
    g->run_of.assign(A, 0);
    for (int k = 0; k < A; k++) {
      const int a = order[k];
      if (k == 0 || g->dst[a] != g->dst[order[k - 1]] || g->pdf[a] != g->pdf[order[k - 1]]) {
        g->run_dst.push_back(g->dst[a]);
        g->run_pdf.push_back(g->pdf[a]);
      }
      g->run_of[a] = (int)g->run_dst.size() - 1;
    }
    g->R = (int)g->run_dst.size();
    // worth it when an arc's exp(y) gather and occupancy RED are shared by >= 3 arcs on average; split states
    // (more than split_cap arcs in a direction) are only handled by the per-arc kernels
    std::vector<int> din(N, 0), dout(N, 0);
    for (int a = 0; a < A; a++) {
      din[g->dst[a]]++;
      dout[g->src[a]]++;
    }
    int maxdeg = 0;
    for (int s = 0; s < N; s++) maxdeg = std::max(maxdeg, std::max(din[s], dout[s]));
    g->tied = A > 0 && (int64_t)g->R * 3 <= (int64_t)A && maxdeg <= g->hooks.split_cap;
    if (g->hooks.tied == 0) g->tied = false;
    if (g->hooks.tied == 1 && A > 0 && maxdeg <= 4 * g->hooks.split_cap) g->tied = true;
  }
  return g.release();
}

// ---- FST file readers (OpenFst binary "vector"/"standard", or AT&T text) ----
bool read_fst_file(const char* path, std::vector<int32_t>* src, std::vector<int32_t>* dst,
                   std::vector<int32_t>* ilabel, std::vector<float>* cost, int32_t* start,
                   std::vector<float>* final_cost) {
  std::ifstream is(path, std::ios::binary);
  if (!is) {
    g_error = std::string("cannot open ") + path;
    return false;
  }
  int32_t magic = 0;
  is.read(reinterpret_cast<char*>(&magic), 4);
  if (is && magic == 2125659606) {
    auto read_str = [&](std::string* s) {
      int32_t n = 0;
      is.read(reinterpret_cast<char*>(&n), 4);
      if (!is || n < 0 || n > 4096) return false;
      s->resize(n);
      is.read(&(*s)[0], n);
      return (bool)is;
    };
    std::string fsttype, arctype;
    int32_t version = 0, flags = 0;
    uint64_t props = 0;
    int64_t st = 0, nstates = 0, narcs = 0;
    if (!read_str(&fsttype) || !read_str(&arctype)) {
      g_error = "FST header truncated";
      return false;
    }
    is.read(reinterpret_cast<char*>(&version), 4);
    is.read(reinterpret_cast<char*>(&flags), 4);
    is.read(reinterpret_cast<char*>(&props), 8);
    is.read(reinterpret_cast<char*>(&st), 8);
    is.read(reinterpret_cast<char*>(&nstates), 8);
    is.read(reinterpret_cast<char*>(&narcs), 8);
    if (!is || fsttype != "vector" || arctype != "standard") {
      g_error = "FST must be a binary VectorFst<StdArc> (got fsttype='" + fsttype + "', arctype='" + arctype + "')";
      return false;
    }
    if (flags & 3) {
      g_error = "FST files with embedded symbol tables are not supported";
      return false;
    }
    if (nstates <= 0 || nstates > (1 << 30)) {
      g_error = "FST has an implausible number of states";
      return false;
    }
    *start = (int32_t)st;
    final_cost->resize(nstates);
    for (int64_t s = 0; s < nstates; s++) {
      float fw;
      int64_t na;
      is.read(reinterpret_cast<char*>(&fw), 4);
      is.read(reinterpret_cast<char*>(&na), 8);
      if (!is || na < 0) {
        g_error = "FST body truncated";
        return false;
      }
      (*final_cost)[s] = fw;
      for (int64_t k = 0; k < na; k++) {
        int32_t il, ol, ns;
        float wt;
        is.read(reinterpret_cast<char*>(&il), 4);
        is.read(reinterpret_cast<char*>(&ol), 4);
        is.read(reinterpret_cast<char*>(&wt), 4);
        is.read(reinterpret_cast<char*>(&ns), 4);
        if (!is) {
          g_error = "FST body truncated";
          return false;
        }
        src->push_back((int32_t)s);
        dst->push_back(ns);
        ilabel->push_back(il);
        cost->push_back(wt);
      }
    }
    return true;
  }
  // text: "src dst ilabel olabel [weight]" | "state [weight]"
  is.clear();
  is.seekg(0);
  std::string line;
  bool first = true;
  int32_t max_state = -1;
  std::vector<std::pair<int32_t, float>> finals;
  while (std::getline(is, line)) {
    std::istringstream ls(line);
    std::vector<std::string> tok;
    std::string t;
    while (ls >> t) tok.push_back(t);
    if (tok.empty()) continue;
    try {
      if (tok.size() >= 4) {
        int32_t s = std::stoi(tok[0]), d = std::stoi(tok[1]), il = std::stoi(tok[2]);
        float w = tok.size() >= 5 ? std::stof(tok[4]) : 0.f;
        if (first) *start = s;
        src->push_back(s);
        dst->push_back(d);
        ilabel->push_back(il);
        cost->push_back(w);
        max_state = std::max(max_state, std::max(s, d));
      } else {
        int32_t s = std::stoi(tok[0]);
        float w = tok.size() >= 2 ? std::stof(tok[1]) : 0.f;
        if (s < 0) {
          g_error = std::string("negative state id in text FST: ") + path;
          return false;
        }
        if (first) *start = s;
        finals.emplace_back(s, w);
        max_state = std::max(max_state, s);
      }
    } catch (...) {
      g_error = std::string("cannot parse FST (neither OpenFst binary nor text): ") + path;
      return false;
    }
    first = false;
  }
  if (max_state < 0) {
    g_error = std::string("empty FST: ") + path;
    return false;
  }
  if (max_state >= (1 << 27)) {
    g_error = std::string("text FST has an implausible state id: ") + path;
    return false;
  }
  final_cost->assign(max_state + 1, kInf);
  for (auto& f : finals) (*final_cost)[f.first] = f.second;
  return true;
}

// ---------------------------------------------------------------------------
// Supervision handle
// ---------------------------------------------------------------------------
size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// The device copy of a supervision is ONE block filled by ONE asynchronous copy out of pinned staging
// memory.  A training loop builds a new Supervision for every minibatch (the reference does,
// /root/reference/example/chime5/train_faster.py:116-141), so nothing on this path may synchronise the
// device: blocks (device memory + pinned staging of the same size) come from a process-wide pool and go back
// to it when the handle dies -- no cudaMalloc / cudaFree per minibatch, no synchronous cudaMemcpy.  A block
// remembers, per stream, an event behind the last work that read it; whoever gets the block next makes its
// upload stream wait for those events first.
struct SupBlock {
  int device = -1;
  char* dev = nullptr;
  char* pinned = nullptr;
  size_t cap = 0;
  std::map<cudaStream_t, cudaEvent_t> last_use;  // behind the last kernels that read `dev`, per stream
  ~SupBlock() {  // only when the pool sheds a block (the device is idle-synchronised by cudaFree itself)
    for (auto& kv : last_use) cudaEventDestroy(kv.second);
    if (dev) cudaFree(dev);
    if (pinned) cudaFreeHost(pinned);
  }
};

class SupBlockPool {
 public:
  // the longest-idle free block of `device` that holds `bytes` without being more than 4x too large (its
  // last readers are the most likely to have finished), else a new one
  cudaError_t acquire(int device, size_t bytes, std::unique_ptr<SupBlock>* out) {
    {
      std::lock_guard<std::mutex> lock(mu_);
      int best = -1;
      for (int i = 0; i < (int)free_.size() && best < 0; i++) {
        const SupBlock& b = *free_[i];
        if (b.device != device || b.cap < bytes || b.cap > 4 * std::max<size_t>(bytes, 1 << 16)) continue;
        best = i;
      }
      if (best >= 0) {
        *out = std::move(free_[best]);
        free_.erase(free_.begin() + best);
        return cudaSuccess;
      }
    }
    auto b = std::make_unique<SupBlock>();
    b->device = device;
    b->cap = align_up(bytes + bytes / 4, 1 << 16);  // headroom: the next minibatch is about this size
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&b->dev), b->cap);
    if (e != cudaSuccess) return e;
    e = cudaHostAlloc(reinterpret_cast<void**>(&b->pinned), b->cap, cudaHostAllocDefault);
    if (e != cudaSuccess) return e;
    *out = std::move(b);
    return cudaSuccess;
  }
  void release(std::unique_ptr<SupBlock> b) {
    std::unique_ptr<SupBlock> shed;  // destroyed outside the lock
    std::lock_guard<std::mutex> lock(mu_);
    free_.push_back(std::move(b));
    if (free_.size() > kMaxFree) {  // oldest first
      shed = std::move(free_.front());
      free_.erase(free_.begin());
    }
  }
  void trim() {  // frees every idle block (cudaFree waits for the device)
    std::vector<std::unique_ptr<SupBlock>> all;
    {
      std::lock_guard<std::mutex> lock(mu_);
      all.swap(free_);
    }
  }

 private:
  static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
  static constexpr size_t kMaxFree = 8;
  std::mutex mu_;
  std::vector<std::unique_ptr<SupBlock>> free_;
};
// Leaked on purpose: handles may be freed by a host language's finalisers during process teardown, after
// static destructors (and the CUDA runtime) are gone.
SupBlockPool& sup_pool() {
  static SupBlockPool* pool = new SupBlockPool();
  return *pool;
}

struct DevView {  // a typed window into SupDevice::block
  void* ptr = nullptr;
  template <typename T>
  T* as() const {
    return static_cast<T*>(ptr);
  }
};
struct SupDevice {
  std::unique_ptr<SupBlock> block;
  cudaEvent_t uploaded = nullptr;  // behind the upload copy, on the stream that performed it
  cudaStream_t upload_stream = nullptr;
  DevView chain_level_off, level_state_off, chain_arc_off, in_off, in_arcs, out_off, out_arcs, final_cost, arc_row,
      arc_pdf, arc_cost, arc_seq, frame_off, frame_arcs;
  ~SupDevice() {
    if (uploaded) cudaEventDestroy(uploaded);
    if (block) sup_pool().release(std::move(block));
  }
};

struct Supervision {
  float weight = 1.f;
  int B = 0, T = 0, P = 0;
  // split representation (host)
  int num_chains = 0, S = 0, A = 0;
  std::vector<int> chain_level_off, level_state_off, chain_arc_off, in_off, out_off;
  std::vector<int2> in_arcs, out_arcs;
  std::vector<float> final_cost, arc_cost;
  std::vector<int> arc_row, arc_pdf, arc_seq, frame_off, frame_arcs;  // frame_*: arc ids ordered by frame
  std::mutex mu;
  std::map<int, std::unique_ptr<SupDevice>> dev;
};

struct RawArc {
  int src, dst, label;
  float cost;
};

// Build the per-chain level structure from a list of chains; each chain is a
// list of states (already in time order, first = its start) and arcs.
Supervision* make_supervision(float weight, int32_t B, int32_t T, int32_t label_dim, int32_t S,
                              const int32_t* arc_begin, const int32_t* arc_dst, const int32_t* arc_label,
                              const float* arc_cost, const float* final_cost) {
  if (B <= 0 || T <= 0 || label_dim <= 0 || S <= 0 || (int64_t)B * T > (int64_t)(1 << 30)) {
    g_error = "supervision: bad sizes";
    return nullptr;
  }
  // the CSR offsets are checked as a whole before anything is indexed with them
  if (arc_begin[0] != 0) {
    g_error = "supervision: arc_begin[0] != 0";
    return nullptr;
  }
  for (int s = 0; s < S; s++)
    if (arc_begin[s + 1] < arc_begin[s]) {
      g_error = "supervision: arc_begin not monotone";
      return nullptr;
    }
  const int A = arc_begin[S];
  // --- state times (Kaldi ComputeFstStateTimes): unique time per state, arcs advance by one
  std::vector<int> time(S, -1);
  time[0] = 0;
  int total = -1;
  for (int s = 0; s < S; s++) {
    if (time[s] < 0) {
      g_error = "supervision: FST not connected / not topologically sorted";
      return nullptr;
    }
    if (s > 0 && time[s] < time[s - 1]) {
      g_error = "supervision: states are not in time order";
      return nullptr;
    }
    if (arc_begin[s + 1] < arc_begin[s]) {
      g_error = "supervision: arc_begin not monotone";
      return nullptr;
    }
    for (int a = arc_begin[s]; a < arc_begin[s + 1]; a++) {
      const int d = arc_dst[a];
      if (d <= s || d >= S) {
        g_error = "supervision: arc destination out of range or not forward";
        return nullptr;
      }
      if (arc_label[a] < 1 || arc_label[a] > label_dim) {
        g_error = "supervision: arc label not in [1, label_dim]";
        return nullptr;
      }
      if (time[d] == -1) time[d] = time[s] + 1;
      else if (time[d] != time[s] + 1) {
        g_error = "supervision: a state is reachable at two different times";
        return nullptr;
      }
    }
    if (final_cost[s] != kInf) {
      if (total == -1) total = time[s];
      else if (total != time[s]) {
        g_error = "supervision: final states at different times";
        return nullptr;
      }
    }
  }
  if (total != B * T) {
    g_error = "supervision: FST length != num_sequences * frames_per_sequence";
    return nullptr;
  }
  for (int s = 0; s < S; s++) {
    if (time[s] == total && arc_begin[s + 1] != arc_begin[s]) {
      g_error = "supervision: arcs leave a state at the final time";
      return nullptr;
    }
    if (time[s] < total && final_cost[s] != kInf) {
      g_error = "supervision: final state before the final time";
      return nullptr;
    }
  }

  // --- try to factorise per sequence: at each boundary time k*T the
  // boundary states must carry identical out-arc sets up to a per-state constant.
  std::vector<int> first_at(total + 2, S);  // first state index with time >= t
  for (int s = S - 1; s >= 0; s--) first_at[time[s]] = s;
  for (int t = total; t >= 0; t--)
    if (first_at[t] == S) first_at[t] = (t == total) ? S : first_at[t + 1];
  first_at[total + 1] = S;
  auto states_at = [&](int t, int* lo, int* hi) {
    *lo = first_at[t];
    *hi = first_at[t + 1];
  };

  bool split = true;
  std::vector<float> boundary_final(S, kInf);  // per-sequence final cost assigned to boundary states
  if (B > 1) {
    struct Key {
      int dst, label;
      float cost;
    };
    for (int k = 1; k < B && split; k++) {
      int lo, hi;
      states_at(k * T, &lo, &hi);
      if (lo >= hi) {
        split = false;
        break;
      }
      auto arcs_of = [&](int s) {
        std::vector<Key> v;
        for (int a = arc_begin[s]; a < arc_begin[s + 1]; a++) v.push_back({arc_dst[a], arc_label[a], arc_cost[a]});
        std::sort(v.begin(), v.end(), [](const Key& x, const Key& y) {
          if (x.dst != y.dst) return x.dst < y.dst;
          if (x.label != y.label) return x.label < y.label;
          return x.cost < y.cost;
        });
        return v;
      };
      const std::vector<Key> ref = arcs_of(lo);
      if (ref.empty()) {
        split = false;
        break;
      }
      boundary_final[lo] = 0.f;
      for (int s = lo + 1; s < hi && split; s++) {
        const std::vector<Key> v = arcs_of(s);
        if (v.size() != ref.size()) {
          split = false;
          break;
        }
        const float delta = v[0].cost - ref[0].cost;
        for (size_t i = 0; i < v.size(); i++) {
          const float d = v[i].cost - ref[i].cost;
          if (v[i].dst != ref[i].dst || v[i].label != ref[i].label ||
              std::fabs(d - delta) > 1e-4f * std::max(1.f, std::fabs(delta))) {
            split = false;
            break;
          }
        }
        boundary_final[s] = delta;
      }
    }
  }

  auto sup = std::make_unique<Supervision>();
  sup->weight = weight;
  sup->B = B;
  sup->T = T;
  sup->P = label_dim;

  // New state numbering: chain by chain, level by level.
  // split:   chain k = {virtual start (k>0) or state 0} + original states with time in (k*T, (k+1)*T]
  // unsplit: one chain = all states.
  std::vector<RawArc> arcs;      // in new state ids, grouped by chain
  std::vector<int> arc_time;     // FST time of each new arc's source
  std::vector<float> fin;        // per new state
  sup->chain_level_off.assign(1, 0);
  sup->chain_arc_off.assign(1, 0);
  const int nchains = split ? B : 1;
  const int levels_per_chain = split ? T : total;
  std::vector<int> newid(S, -1);
  int next_state = 0;
  for (int c = 0; c < nchains; c++) {
    const int t0 = c * levels_per_chain;
    // level 0
    int start_new = next_state++;
    fin.push_back(kInf);
    sup->level_state_off.push_back(start_new);
    int start_orig;
    if (c == 0) {
      start_orig = 0;
      newid[0] = start_new;
    } else {
      int lo, hi;
      states_at(t0, &lo, &hi);
      start_orig = lo;  // the reference boundary state f*: its out-arcs are this chain's start arcs
    }
    // levels 1..levels_per_chain
    for (int l = 1; l <= levels_per_chain; l++) {
      int lo, hi;
      states_at(t0 + l, &lo, &hi);
      sup->level_state_off.push_back(next_state);
      for (int s = lo; s < hi; s++) {
        newid[s] = next_state++;
        float f = kInf;
        if (l == levels_per_chain) f = (split && c + 1 < nchains) ? boundary_final[s] : final_cost[s];
        fin.push_back(f);
      }
    }
    sup->level_state_off.push_back(next_state);  // end of last level
    sup->chain_level_off.push_back((int)sup->level_state_off.size());
    // arcs: start arcs, then arcs of states at levels 1..levels_per_chain-1
    for (int a = arc_begin[start_orig]; a < arc_begin[start_orig + 1]; a++) {
      arcs.push_back({start_new, newid[arc_dst[a]], arc_label[a], arc_cost[a]});
      arc_time.push_back(t0);
    }
    for (int l = 1; l < levels_per_chain; l++) {
      int lo, hi;
      states_at(t0 + l, &lo, &hi);
      for (int s = lo; s < hi; s++)
        for (int a = arc_begin[s]; a < arc_begin[s + 1]; a++) {
          arcs.push_back({newid[s], newid[arc_dst[a]], arc_label[a], arc_cost[a]});
          arc_time.push_back(t0 + l);
        }
    }
    sup->chain_arc_off.push_back((int)arcs.size());
  }
  // level_state_off layout per chain: levels_per_chain+1 level starts + 1 end => nlevels+1 entries
  const int NS = next_state, NA = (int)arcs.size();
  (void)A;
  sup->num_chains = nchains;
  sup->S = NS;
  sup->A = NA;
  sup->final_cost = fin;
  sup->arc_cost.resize(NA);
  sup->arc_row.resize(NA);
  sup->arc_pdf.resize(NA);
  sup->in_off.assign(NS + 1, 0);
  sup->out_off.assign(NS + 1, 0);
  for (int a = 0; a < NA; a++) {
    sup->arc_cost[a] = arcs[a].cost;
    sup->arc_pdf[a] = arcs[a].label - 1;
    const int seq = arc_time[a] / T, frame = arc_time[a] % T;
    sup->arc_row[a] = frame * B + seq;  // Kaldi ComputeRowIndex
    sup->in_off[arcs[a].dst + 1]++;
    sup->out_off[arcs[a].src + 1]++;
  }
  for (int s = 0; s < NS; s++) {
    sup->in_off[s + 1] += sup->in_off[s];
    sup->out_off[s + 1] += sup->out_off[s];
  }
  sup->in_arcs.resize(NA);
  sup->out_arcs.resize(NA);
  {
    std::vector<int> ip(sup->in_off.begin(), sup->in_off.end() - 1), op(sup->out_off.begin(), sup->out_off.end() - 1);
    for (int a = 0; a < NA; a++) {
      sup->in_arcs[ip[arcs[a].dst]++] = make_int2(arcs[a].src, a);
      sup->out_arcs[op[arcs[a].src]++] = make_int2(arcs[a].dst, a);
    }
  }
  // arcs ordered by frame, for the backward launch's per-frame numerator scatter
  {
    sup->arc_seq.resize(NA);
    sup->frame_off.assign(T + 1, 0);
    for (int a = 0; a < NA; a++) {
      sup->arc_seq[a] = arc_time[a] / T;
      sup->frame_off[arc_time[a] % T + 1]++;
    }
    for (int t = 0; t < T; t++) sup->frame_off[t + 1] += sup->frame_off[t];
    sup->frame_arcs.resize(NA);
    std::vector<int> pos(sup->frame_off.begin(), sup->frame_off.end() - 1);
    for (int a = 0; a < NA; a++) sup->frame_arcs[pos[arc_time[a] % T]++] = a;
  }
  return sup.release();
}

// ---------------------------------------------------------------------------
// per-device preparation
// ---------------------------------------------------------------------------
// Build + upload one partition generation: subs = 1 from the device's current speed estimates,
// subs = 2 / 4 (one bin per lane group, small batches) with uniform speeds.
int build_partition(DenGraph* g, DenDevice* d, int subs = 1) {
  auto part = std::make_unique<DenPartition>();
  const int bins = tb::kDenMainWarps * subs;
  const std::vector<double> uniform((size_t)d->grid * bins, 1.0);
  for (int dir = 0; dir < 2; dir++) {
    DirHost h;
    build_dir(*g, dir, d->grid, bins, subs == 1 ? d->speed[dir] : uniform, 4u * (unsigned)(tb::kCtxBP / subs), &h,
              &part->work[dir], nullptr, g->hooks.align_streams ? subs : 1);
    CUDA_OK(upload(part->arcs[dir], h.arcs));
    CUDA_OK(upload(part->tasks[dir], h.tasks));
    CUDA_OK(upload(part->cta_arc_off[dir], h.cta_arc_off));
    CUDA_OK(upload(part->cta_task_off[dir], h.cta_task_off));
    CUDA_OK(upload(part->warp_task_off[dir], h.warp_task_off));
    CUDA_OK(upload(part->warp_arc_off[dir], h.warp_arc_off));
    part->max_arcs_cta[dir] = h.max_arcs_cta;
    part->max_tasks_cta[dir] = h.max_tasks_cta;
    part->has_split = part->has_split || h.has_split;
  }
  if (subs == 1) d->cur = part.get();
  else d->cur_sub[subs == 2 ? 0 : 1] = part.get();
  d->generations.push_back(std::move(part));
  return 1;
}


// Streams of the run-based kernels (den_kernel.cu "tied"): forward = destination states with their in-arcs
// grouped by run, every run closed by a run record; phase A = the runs dealt evenly; phase B = source states
// with their out-arcs, each pointing at the W row of its run.
int build_tied_partition(DenGraph* g, DenDevice* d) {
  const int N = g->N, A = g->A, R = g->R, grid = d->grid, bins = tb::kDenMainWarps;
  const unsigned rowbytes = 4u * (unsigned)tb::kCtxBP;
  auto part = std::make_unique<TiedPartition>();
  // ---- forward
  {
    std::vector<int> deg(N, 0), begin(N + 1, 0);
    for (int a = 0; a < A; a++) deg[g->dst[a]]++;
    for (int s = 0; s < N; s++) begin[s + 1] = begin[s] + deg[s];
    std::vector<int> arc_of(A), pos(begin.begin(), begin.end() - 1);
    for (int a = 0; a < A; a++) arc_of[pos[g->dst[a]]++] = a;
    for (int s = 0; s < N; s++)  // runs are numbered in (state, pdf) order: sorting by run groups a state's arcs
      std::stable_sort(arc_of.begin() + begin[s], arc_of.begin() + begin[s + 1],
                       [&](int x, int y) { return g->run_of[x] < g->run_of[y]; });
    std::vector<int> cost(N), all(N);
    for (int s = 0; s < N; s++) {
      int runs = 0;
      for (int k = begin[s]; k < begin[s + 1]; k++)
        if (k == begin[s] || g->run_of[arc_of[k]] != g->run_of[arc_of[k - 1]]) runs++;
      cost[s] = deg[s] + std::max(runs, 1) + 2;
      all[s] = s;
    }
    std::vector<std::vector<int>> by_cta;
    lpt(all, cost, std::vector<double>(grid, 1.0), &by_cta);
    std::vector<int4> recs, tasks;
    std::vector<int> cta_arc_off(1, 0), cta_task_off(1, 0), warp_task_off(1, 0), warp_arc_off(1, 0);
    for (int c = 0; c < grid; c++) {
      std::vector<std::vector<int>> by_warp;
      lpt(by_cta[c], cost, std::vector<double>(bins, 1.0), &by_warp);
      const int arc0 = (int)recs.size(), task0 = (int)tasks.size();
      for (int w = 0; w < bins; w++) {
        for (int s : by_warp[w]) {
          int4 tk = make_int4(s, (int)recs.size() - arc0, 0, 0);
          float pis = g->pi[s];
          std::memcpy(&tk.w, &pis, 4);
          tasks.push_back(tk);
          auto run_record = [&](int pdf, int run, bool last) {
            recs.push_back(make_int4((int)((unsigned)pdf * rowbytes), (int)(0x80000000u | (last ? 1u : 0u)),
                                     (int)((unsigned)run * rowbytes), 0));
          };
          for (int k = begin[s]; k < begin[s + 1]; k++) {
            const int a = arc_of[k];
            int4 ar = make_int4((int)((unsigned)g->src[a] * rowbytes), 0, 0, 0);
            float pr = g->prob[a], pio = g->pi[g->src[a]];
            std::memcpy(&ar.y, &pr, 4);
            std::memcpy(&ar.z, &pio, 4);
            recs.push_back(ar);
            const bool run_ends = k + 1 == begin[s + 1] || g->run_of[arc_of[k + 1]] != g->run_of[a];
            if (run_ends) run_record(g->pdf[a], g->run_of[a], k + 1 == begin[s + 1]);
          }
          if (deg[s] == 0) run_record(0, R, true);  // no in-arcs: alpha = 0, written through the scratch run row R
        }
        const int stream_begin = warp_arc_off.back();
        while (((int)recs.size() - stream_begin) % tb::kStreamPad != 0)
          recs.push_back(make_int4(0, 0, 0, 0));  // zero-probability arc behind the last run record: feeds nothing
        warp_task_off.push_back((int)tasks.size());
        warp_arc_off.push_back((int)recs.size());
      }
      cta_arc_off.push_back((int)recs.size());
      cta_task_off.push_back((int)tasks.size());
      part->max_arcs_cta[0] = std::max(part->max_arcs_cta[0], (int)recs.size() - arc0);
      part->max_tasks_cta[0] = std::max(part->max_tasks_cta[0], (int)tasks.size() - task0);
    }
    CUDA_OK(upload(part->arcs[0], recs));
    CUDA_OK(upload(part->tasks[0], tasks));
    CUDA_OK(upload(part->cta_arc_off[0], cta_arc_off));
    CUDA_OK(upload(part->cta_task_off[0], cta_task_off));
    CUDA_OK(upload(part->warp_task_off[0], warp_task_off));
    CUDA_OK(upload(part->warp_arc_off[0], warp_arc_off));
  }
  // ---- phase B: the per-arc builder with the arc's run as the gathered row, never split
  {
    DirHost h;
    std::vector<double> work;
    build_dir(*g, 1, grid, bins, std::vector<double>((size_t)grid * bins, 1.0), rowbytes, &h, &work, &g->run_of);
    CUDA_OK(upload(part->arcs[1], h.arcs));
    CUDA_OK(upload(part->tasks[1], h.tasks));
    CUDA_OK(upload(part->cta_arc_off[1], h.cta_arc_off));
    CUDA_OK(upload(part->cta_task_off[1], h.cta_task_off));
    CUDA_OK(upload(part->warp_task_off[1], h.warp_task_off));
    CUDA_OK(upload(part->warp_arc_off[1], h.warp_arc_off));
    part->max_arcs_cta[1] = h.max_arcs_cta;
    part->max_tasks_cta[1] = h.max_tasks_cta;
  }
  // ---- phase A: runs, contiguous equal shares (run order = destination-state order)
  {
    std::vector<int4> runs(R);
    for (int r = 0; r < R; r++)
      runs[r] = make_int4((int)((unsigned)g->run_dst[r] * rowbytes), (int)((unsigned)g->run_pdf[r] * rowbytes),
                          (int)((unsigned)r * rowbytes), 0);
    const int nb = grid * bins;
    std::vector<int> cta_run_off(grid + 1), warp_run_off(nb + 1);
    for (int b = 0; b <= nb; b++) warp_run_off[b] = (int)((int64_t)R * b / nb);
    for (int c = 0; c <= grid; c++) cta_run_off[c] = warp_run_off[c * bins];
    for (int c = 0; c < grid; c++) part->max_runs_cta = std::max(part->max_runs_cta, cta_run_off[c + 1] - cta_run_off[c]);
    CUDA_OK(upload(part->runs, runs));
    CUDA_OK(upload(part->cta_run_off, cta_run_off));
    CUDA_OK(upload(part->warp_run_off, warp_run_off));
  }
  d->tied = std::move(part);
  return 1;
}

int get_den_device(DenGraph* g, int device, DenDevice** out) {
  std::lock_guard<std::mutex> lock(g->mu);
  auto it = g->dev.find(device);
  if (it != g->dev.end()) {
    *out = it->second.get();
    return 1;
  }
  auto d = std::make_unique<DenDevice>();
  d->grid = tb::den_max_grid(device);
  if (d->grid <= 0) return fail("cannot query SM count");
  for (int dir = 0; dir < 2; dir++) d->speed[dir].assign((size_t)d->grid * tb::kDenMainWarps, 1.0);
  CUDA_OK(d->warp_busy.reserve(sizeof(unsigned long long) * 2 * (size_t)d->grid * tb::kDenWarps));
  if (g->hooks.no_autotune) d->calib_rounds_done = kCalibRounds;
  if (!build_partition(g, d.get())) return 0;
  *out = d.get();
  g->dev[device] = std::move(d);
  return 1;
}

// Partition autotuner.  The grid barrier makes every frame as slow as its slowest warp, and
// warps do not run at the same speed: SMs differ in their distance to the L2 slices and the
// warps of an SM share schedulers unevenly with the aux warps.  These differences are stable
// from frame to frame and call to call (profiles/r01_summary.md), so after each of the first
// kCalibRounds synchronised calls the measured per-warp busy cycles of the last launch pair
// turn into speed estimates (assigned work / busy time) and the states are re-dealt.
int autotune_partition(DenGraph* g, DenDevice* d, cudaStream_t copy_stream = nullptr) {
  std::lock_guard<std::mutex> lock(g->mu);
  if (d->calib_rounds_done >= kCalibRounds) return 1;
  const size_t n = 2 * (size_t)d->grid * tb::kDenWarps;
  std::vector<unsigned long long> busy(n);
  if (copy_stream) {  // harvesting behind an asynchronous call: do not touch the caller's streams
    CUDA_OK(cudaMemcpyAsync(busy.data(), d->warp_busy.ptr, sizeof(unsigned long long) * n, cudaMemcpyDeviceToHost,
                            copy_stream));
    CUDA_OK(cudaStreamSynchronize(copy_stream));
  } else {
    CUDA_OK(cudaMemcpy(busy.data(), d->warp_busy.ptr, sizeof(unsigned long long) * n, cudaMemcpyDeviceToHost));
  }
  for (int dir = 0; dir < 2; dir++) {
    const size_t bins = (size_t)d->grid * tb::kDenMainWarps;
    std::vector<double> est(bins, 0.0);
    double sum = 0.0, total_busy = 0.0;
    int valid = 0;
    for (int c = 0; c < d->grid; c++)
      for (int w = 0; w < tb::kDenMainWarps; w++) {
        const double b = (double)busy[((size_t)dir * d->grid + c) * tb::kDenWarps + w];
        const double wk = d->cur->work[dir][c * tb::kDenMainWarps + w];
        total_busy += b;
        if (b > 0.0 && wk > 0.0) {
          est[c * tb::kDenMainWarps + w] = wk / b;
          sum += wk / b;
          valid++;
        }
      }
    // too little work per warp to measure anything (tiny graphs): leave the partition alone
    if (valid < (int)bins || total_busy / bins < 2.0e5) {
      d->calib_rounds_done = kCalibRounds;
      return 1;
    }
    const double mean = sum / valid;
    for (size_t i = 0; i < bins; i++) {
      double rel = est[i] / mean;
      rel = std::min(1.5, std::max(0.67, rel));
      d->speed[dir][i] = rel;
    }
  }
  d->calib_rounds_done++;
  return build_partition(g, d);
}

// The device image of a supervision on `device`, uploading it on `stream` (asynchronously) if this is its
// first use there.  Every later user must order itself behind SupDevice::uploaded (wait_sup_uploaded).
int get_sup_device(Supervision* s, int device, cudaStream_t stream, SupDevice** out) {
  std::lock_guard<std::mutex> lock(s->mu);
  auto it = s->dev.find(device);
  if (it != s->dev.end()) {
    *out = it->second.get();
    return 1;
  }
  auto d = std::make_unique<SupDevice>();
  // lay the arrays out back to back (256-byte slots) in ONE pinned staging block, upload once
  struct Slot {
    DevView* view;
    const void* src;
    size_t nbytes, off;
  };
  std::vector<Slot> slots;
  size_t total = 0;
  auto add = [&](DevView& view, const void* src, size_t nbytes) {
    slots.push_back({&view, src, nbytes, total});
    total += align_up(std::max<size_t>(nbytes, 16), 256);
  };
  auto add_vec = [&](DevView& view, const auto& v) { add(view, v.data(), sizeof(v[0]) * v.size()); };
  add_vec(d->chain_level_off, s->chain_level_off);
  add_vec(d->level_state_off, s->level_state_off);
  add_vec(d->chain_arc_off, s->chain_arc_off);
  add_vec(d->in_off, s->in_off);
  add_vec(d->in_arcs, s->in_arcs);
  add_vec(d->out_off, s->out_off);
  add_vec(d->out_arcs, s->out_arcs);
  add_vec(d->final_cost, s->final_cost);
  add_vec(d->arc_row, s->arc_row);
  add_vec(d->arc_pdf, s->arc_pdf);
  add_vec(d->arc_cost, s->arc_cost);
  add_vec(d->arc_seq, s->arc_seq);
  add_vec(d->frame_off, s->frame_off);
  add_vec(d->frame_arcs, s->frame_arcs);
  CUDA_OK(sup_pool().acquire(device, total, &d->block));
  SupBlock& b = *d->block;
  // whoever read this block before (an earlier minibatch's kernels) must be done before it is overwritten;
  // the staging half was last read by that minibatch's upload, which those kernels were ordered behind
  for (auto& kv : b.last_use) {
    if (kv.first == stream) continue;  // same stream: ordered anyway
    CUDA_OK(cudaStreamWaitEvent(stream, kv.second, 0));
  }
  {
    // the staging memory is reused too: the previous upload out of it must have completed.  It was ordered
    // before kernels whose events we just made `stream` wait for -- but the HOST is about to write it now.
    for (auto& kv : b.last_use) {
      const cudaError_t q = cudaEventQuery(kv.second);
      if (q == cudaErrorNotReady) CUDA_OK(cudaEventSynchronize(kv.second));
      else if (q != cudaSuccess) CUDA_OK(q);
    }
  }
  for (const Slot& sl : slots) {
    if (sl.nbytes) std::memcpy(b.pinned + sl.off, sl.src, sl.nbytes);
    sl.view->ptr = b.dev + sl.off;
  }
  CUDA_OK(cudaMemcpyAsync(b.dev, b.pinned, total, cudaMemcpyHostToDevice, stream));
  CUDA_OK(cudaEventCreateWithFlags(&d->uploaded, cudaEventDisableTiming));
  CUDA_OK(cudaEventRecord(d->uploaded, stream));
  d->upload_stream = stream;
  {  // the upload itself reads the staging half: a handle freed before any kernel used it is covered too
    cudaEvent_t& ev = b.last_use[stream];
    if (!ev) CUDA_OK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    CUDA_OK(cudaEventRecord(ev, stream));
  }
  *out = d.get();
  s->dev[device] = std::move(d);
  return 1;
}

// Orders `stream` behind the upload of `sd` (no-op on the stream that uploaded it).
int wait_sup_uploaded(Supervision* s, SupDevice* sd, cudaStream_t stream) {
  if (sd->upload_stream == stream) return 1;
  std::lock_guard<std::mutex> lock(s->mu);
  CUDA_OK(cudaStreamWaitEvent(stream, sd->uploaded, 0));
  return 1;
}

// Marks `stream`'s work enqueued so far as a reader of the block (see SupBlock::last_use).
int mark_sup_used(Supervision* s, SupDevice* sd, cudaStream_t stream) {
  std::lock_guard<std::mutex> lock(s->mu);
  cudaEvent_t& ev = sd->block->last_use[stream];
  if (!ev) CUDA_OK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  CUDA_OK(cudaEventRecord(ev, stream));
  return 1;
}

int get_workspace(DenGraph* g, int device, void* stream, Workspace** out) {
  std::lock_guard<std::mutex> lock(g->mu);
  auto key = std::make_pair(device, stream);
  auto it = g->ws.find(key);
  if (it != g->ws.end()) {
    it->second->in_use++;
    it->second->last_use = ++g->ws_clock;
    *out = it->second.get();
    return 1;
  }
  auto w = std::make_unique<Workspace>();
  CUDA_OK(cudaStreamCreateWithFlags(&w->side, cudaStreamNonBlocking));
  CUDA_OK(cudaEventCreateWithFlags(&w->ev_fork, cudaEventDisableTiming));
  CUDA_OK(cudaEventCreateWithFlags(&w->ev_join, cudaEventDisableTiming));
  CUDA_OK(cudaMallocHost(&w->pinned, 4 * sizeof(float)));
  // make room first: drop idle workspaces of this device, least recently used first
  for (;;) {
    int count = 0;
    auto lru = g->ws.end();
    for (auto jt = g->ws.begin(); jt != g->ws.end(); ++jt) {
      if (jt->first.first != device) continue;
      count++;
      if (jt->second->in_use == 0 && (lru == g->ws.end() || jt->second->last_use < lru->second->last_use)) lru = jt;
    }
    if (count < kMaxWorkspacesPerDevice || lru == g->ws.end()) break;
    g->ws.erase(lru);
  }
  w->in_use = 1;
  w->last_use = ++g->ws_clock;
  *out = w.get();
  g->ws[key] = std::move(w);
  return 1;
}

// Marks the end of a call's use of its workspace (every return path of compute()).
struct WorkspaceLease {
  DenGraph* g;
  Workspace* w;
  ~WorkspaceLease() {
    std::lock_guard<std::mutex> lock(g->mu);
    w->in_use--;
  }
};

// Resolve and clear the profiling events of every workspace of this graph.
void harvest_profile(DenGraph* g) {
  std::lock_guard<std::mutex> lock(g->mu);
  for (auto& kv : g->ws) {
    for (auto& pr : kv.second->prof) {
      float ms = 0.f;
      if (cudaEventSynchronize(pr.second) == cudaSuccess && cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) {
        std::lock_guard<std::mutex> l2(g_prof_mu);
        g_prof_den_ms += ms;
        g_prof_den_launches++;
      }
      cudaEventDestroy(pr.first);
      cudaEventDestroy(pr.second);
    }
    kv.second->prof.clear();
  }
}


int pad_batch(int b) {
  if (b <= 32) return 32;
  if (b <= 64) return 64;
  return (b + 127) / 128 * 128;
}

// Optional chunked arrival of nnet_output (host-buffer entry): flags[k] != 0 once the rows of
// frames [k*chunk_frames, (k+1)*chunk_frames) have landed in device memory.
struct ChunkedInput {
  const unsigned* flags = nullptr;
  int chunk_frames = 0;
  cudaEvent_t all_landed = nullptr;  // recorded on the copy stream after the last chunk
  cudaEvent_t all_landed2 = nullptr; // same for the second copy stream (experiment), or null
  unsigned* deriv_progress = nullptr;  // device pointer of a mapped pinned counter (see DenParams)
  unsigned long long progress_mask[8] = {~0ull, ~0ull, ~0ull, ~0ull, ~0ull, ~0ull, ~0ull, ~0ull};  // values worth publishing
  float* results4_dev = nullptr;       // device float[4]: objf, l2_term, weight, ok flag of the numeric guard
};

// Frame boundaries of the chunks the host-buffer entry moves nnet_output / the derivative in:
// about 16 chunks of `base` whole frames, with shorter ones (1, 1, 2, 4, ... frames) at BOTH ends.
// What the host waits for with nothing left to overlap is the work behind the last chunk in (the
// frames it holds still have to go through the forward recurrence) and the copy of the last chunk out
// (frame 0, which is complete only when the backward launch ends), so those chunks are single frames.
std::vector<int> chunk_bounds(int T) {
  const int base = std::max(1, (T + 15) / 16);
  std::vector<int> head, tail;
  int lo = 0, hi = T;
  for (int s = 1, rep = 0; s < base && hi - lo >= 2 * s + base;) {
    head.push_back(s);
    tail.push_back(s);
    lo += s;
    hi -= s;
    if (s > 1 || ++rep == 2) s *= 2;
  }
  std::vector<int> b(1, 0);
  for (int s : head) b.push_back(b.back() + s);
  while (b.back() + base < hi) b.push_back(b.back() + base);
  if (b.back() < hi) b.push_back(hi);
  for (auto it = tail.rbegin(); it != tail.rend(); ++it) b.push_back(b.back() + *it);
  return b;  // b.front() == 0, b.back() == T, strictly increasing
}

// The whole loss, enqueued on `stream`: forward launch -> numerator -> backward launch ->
// final scalars.  results_host may be null (then nothing synchronises).
// Matrices are described by tb::Layout3 (element strides over frame / sequence / pdf): the
// reference's (T*B, P) form and a network output left in (B, P, T) go through the same kernels.
tb::Layout3 row_layout(int B, int64_t row_stride) { return tb::Layout3{(int64_t)B * row_stride, row_stride, 1}; }

int compute(DenGraph* g, Supervision* sup, const float* y, int64_t rows, int64_t cols, tb::Layout3 yl,
            float* results_host, float* results_dev_user, float* deriv, tb::Layout3 dl, float* xent,
            tb::Layout3 xl, float l2, float leaky, float xent_regularize, float deriv_scale, int device,
            cudaStream_t stream, ChunkedInput chunked = ChunkedInput()) {
  if (!g || !sup) return fail("null den_graph / supervision handle");
  if (!y) return fail("null nnet_output");
  const int B = sup->B, T = sup->T, P = g->P, N = g->N;
  if (rows != (int64_t)B * T) return fail("nnet_output rows != num_sequences * frames_per_sequence");
  if (cols != P) return fail("nnet_output cols != den_graph.NumPdfs()");
  if (sup->P != P) return fail("supervision.label_dim != den_graph.NumPdfs()");
  if (yl.t < 0 || yl.s < 0 || yl.p < 0) return fail("negative nnet_output stride");
  if ((deriv && (dl.t <= 0 || dl.s <= 0 || dl.p <= 0)) || (xent && (xl.t <= 0 || xl.s <= 0 || xl.p <= 0)))
    return fail("derivative strides must be positive");
  if (!(leaky > 0.f && leaky < 1.f)) return fail("leaky_hmm_coefficient must be in (0, 1)");
  if ((int64_t)N * tb::kCtxBP * 4 >= (1ll << 32) || (int64_t)P * tb::kCtxBP * 4 >= (1ll << 32))
    return fail("graph too large for 32-bit row byte offsets");
  DeviceGuard guard(device);
  if (!guard.ok) return fail("cannot select CUDA device");

  DenDevice* dd;
  SupDevice* sd;
  Workspace* ws;
  if (!get_den_device(g, device, &dd)) return 0;
  if (!get_sup_device(sup, device, stream, &sd)) return 0;
  if (!wait_sup_uploaded(sup, sd, stream)) return 0;
  if (!get_workspace(g, device, (void*)stream, &ws)) return 0;
  WorkspaceLease lease{g, ws};
  std::lock_guard<std::mutex> one_call(ws->call_mu);  // released before `lease` (reverse declaration order)
  const bool want_xent = xent != nullptr && xent_regularize != 0.f;

  if (!deriv) {  // objective only: derivatives go to scratch
    CUDA_OK(ws->scratch_deriv.reserve(sizeof(float) * (size_t)rows * cols));
    deriv = ws->scratch_deriv.as<float>();
    dl = row_layout(B, cols);
  }

  // ---- numerator scratch
  const size_t nA = (size_t)sup->A, nS = (size_t)sup->S;
  size_t off_arcval = 0, off_la = align_up(off_arcval + 8 * nA, 256), off_lb = align_up(off_la + 8 * nS, 256),
         off_post = align_up(off_lb + 8 * nS, 256), off_clp = align_up(off_post + 4 * nA, 256),
         num_total = align_up(off_clp + 8 * (size_t)sup->num_chains, 256);
  CUDA_OK(ws->num_scratch.reserve(num_total));
  char* nb = ws->num_scratch.as<char>();
  tb::NumParams np;
  np.num_chains = sup->num_chains;
  np.chain_level_off = sd->chain_level_off.as<int>();
  np.level_state_off = sd->level_state_off.as<int>();
  np.chain_arc_off = sd->chain_arc_off.as<int>();
  np.in_off = sd->in_off.as<int>();
  np.in_arcs = sd->in_arcs.as<int2>();
  np.out_off = sd->out_off.as<int>();
  np.out_arcs = sd->out_arcs.as<int2>();
  np.final_cost = sd->final_cost.as<float>();
  np.arc_row = sd->arc_row.as<int>();
  np.arc_pdf = sd->arc_pdf.as<int>();
  np.arc_cost = sd->arc_cost.as<float>();
  np.B = B;
  np.y = y;
  np.yl = yl;
  np.arc_val = reinterpret_cast<double*>(nb + off_arcval);
  np.log_alpha = reinterpret_cast<double*>(nb + off_la);
  np.log_beta = reinterpret_cast<double*>(nb + off_lb);
  np.arc_post = reinterpret_cast<float*>(nb + off_post);
  np.chain_logprob = reinterpret_cast<double*>(nb + off_clp);

  tb::DenParams dp = {};
  // ---- run-based ("tied") kernels: graphs with few runs per arc, passes of one context of 128 sequences
  TiedPartition* tp = nullptr;
  if (g->tied && B > tb::kCtxBP / 2) {
    std::lock_guard<std::mutex> lock(g->mu);
    if (!dd->tied && !build_tied_partition(g, dd)) return 0;
    tp = dd->tied.get();
    // the run-based kernels keep a CTA's streams in shared memory; a graph too large for that stays per-arc
    tb::DenParams probe = {};
    probe.BP = tb::kCtxBP;
    for (int dir = 0; dir < 2; dir++) {
      probe.dir[dir].max_arcs_cta = tp->max_arcs_cta[dir];
      probe.dir[dir].max_tasks_cta = tp->max_tasks_cta[dir];
    }
    probe.rdir.max_arcs_cta = tp->max_runs_cta;
    if (tb::den_tied_smem_bytes(probe) > 200 * 1024) tp = nullptr;
  }
  const int pass_cap = tp ? tb::kCtxBP : tb::kMaxBP;  // sequences per launch pair
  const int rem = B % tb::kCtxBP;
  const bool tied_rest_general = tp && rem != 0 && pad_batch(rem) != tb::kCtxBP;  // a short last pass runs per-arc

  // ---- denominator workspace.  A per-arc launch pair covers up to kMaxBP = 256 sequences as one or two
  // contexts of kCtxBP = 128 (own buffers each); larger minibatches run as consecutive passes.
  const int nctx_max = (!tp && B > tb::kCtxBP) ? 2 : 1;
  const int BPmax = B > tb::kCtxBP ? tb::kCtxBP : pad_batch(B);
  const size_t alpha_ctx = (size_t)(T + 1) * N * BPmax, beta_ctx = (size_t)3 * N * BPmax,
               tile_ctx = (size_t)2 * P * BPmax;
  if (tp) {  // alpha is a two-row ping-pong; the forward contributions of the runs take the alpha store's place
    const size_t general = tied_rest_general ? (size_t)(T + 1) * N * pad_batch(rem) : 0;
    CUDA_OK(ws->alpha.reserve(sizeof(float) * std::max((size_t)2 * N * tb::kCtxBP, general)));
    CUDA_OK(ws->arun.reserve(sizeof(float) * (size_t)T * (g->R + 1) * tb::kCtxBP));
    CUDA_OK(ws->wbuf.reserve(sizeof(float) * (size_t)(g->R + 1) * tb::kCtxBP));
  } else {
    CUDA_OK(ws->alpha.reserve(sizeof(float) * alpha_ctx * nctx_max));
  }
  CUDA_OK(ws->beta.reserve(sizeof(float) * beta_ctx * nctx_max));
  CUDA_OK(ws->tiles.reserve(sizeof(float) * tile_ctx * nctx_max));  // etile
  // zeroed block: per context {gtile | Ssum | Bsum | barriers}, then the timeline; scalars + results after it
  const size_t zc_S = align_up(4 * tile_ctx, 256), zc_B = align_up(zc_S + 8 * (size_t)(T + 1) * BPmax, 256),
               zc_bar = align_up(zc_B + 8 * (size_t)(T + 1) * BPmax, 256), zc_size = zc_bar + 256,
               z_tl = zc_size * nctx_max, z_end = align_up(z_tl + 8 * (size_t)(2 * T + 8), 256), z_scal = z_end,
               z_res = z_scal + 64, z_total = z_res + 64;
  CUDA_OK(ws->zeroed.reserve(z_total));
  char* zb = ws->zeroed.as<char>();
  CUDA_OK(cudaMemsetAsync(zb + z_scal, 0, 128, stream));

  // the partition depends on the lane-group geometry of a pass (subs = 1, 2 or 4)
  auto partition_for = [&](int subs, DenPartition** out) -> int {
    std::lock_guard<std::mutex> lock(g->mu);
    if (subs > 1 && dd->cur_sub[subs == 2 ? 0 : 1] == nullptr && !build_partition(g, dd, subs)) return 0;
    *out = subs == 1 ? dd->cur : dd->cur_sub[subs == 2 ? 0 : 1];
    return 1;
  };
  auto bind = [&](const DenPartition* part) {
    for (int dir = 0; dir < 2; dir++) {
      dp.dir[dir].arcs = part->arcs[dir].as<int4>();
      dp.dir[dir].tasks = part->tasks[dir].as<int4>();
      dp.dir[dir].cta_arc_off = part->cta_arc_off[dir].as<int>();
      dp.dir[dir].cta_task_off = part->cta_task_off[dir].as<int>();
      dp.dir[dir].warp_task_off = part->warp_task_off[dir].as<int>();
      dp.dir[dir].warp_arc_off = part->warp_arc_off[dir].as<int>();
      dp.dir[dir].max_arcs_cta = part->max_arcs_cta[dir];
      dp.dir[dir].max_tasks_cta = part->max_tasks_cta[dir];
    }
    dp.has_split = part->has_split ? 1 : 0;
  };
  // a calibration round left behind by an earlier asynchronous call, if its launches have finished
  {
    bool harvest = false;
    {
      std::lock_guard<std::mutex> lock(g->mu);
      if (dd->calib_pending) {
        const cudaError_t q = cudaEventQuery(dd->calib_ev);
        if (q == cudaSuccess) {
          dd->calib_pending = false;
          harvest = dd->calib_part == dd->cur;
        } else {
          (void)cudaGetLastError();  // cudaErrorNotReady is not an error
        }
      }
    }
    if (harvest && !autotune_partition(g, dd, ws->side)) return 0;
  }
  // only a single-pass call on the full-width partition can feed the autotuner; a synchronising call
  // reads its counters back at the end, any other leaves them for a later call (calib_ev)
  bool calibrating;
  {
    std::lock_guard<std::mutex> lock(g->mu);
    calibrating = !tp && B <= tb::kMaxBP && tb::den_subs(BPmax) == 1 && dd->calib_rounds_done < kCalibRounds &&
                  !dd->calib_pending;
  }
  const bool calib_async = calibrating && results_host == nullptr;
  dp.warp_busy = calibrating ? dd->warp_busy.as<unsigned long long>() : nullptr;
  dp.N = N;
  dp.P = P;
  dp.Btot = B;
  dp.T = T;
  dp.y = y;
  dp.yl = yl;
  dp.deriv = deriv;
  dp.dl = dl;
  dp.scalars = reinterpret_cast<double*>(zb + z_scal);
  dp.leaky = leaky;
  dp.sum_pi = (float)g->sum_pi;
  dp.scale_gamma = deriv_scale * (-sup->weight);
  dp.scale_y = deriv_scale * (-sup->weight * l2);
  dp.num_frame_off = sd->frame_off.as<int>();
  dp.num_frame_arcs = sd->frame_arcs.as<int>();
  dp.num_arc_pdf = sd->arc_pdf.as<int>();
  dp.num_arc_seq = sd->arc_seq.as<int>();
  dp.num_arc_post = np.arc_post;
  dp.y_ready = chunked.flags;
  dp.y_chunk_frames = chunked.chunk_frames;
  dp.deriv_progress = (B <= pass_cap) ? chunked.deriv_progress : nullptr;  // single-pass calls only
  for (int i = 0; i < 8; i++) dp.progress_mask[i] = chunked.progress_mask[i];
  dp.cta_trace = nullptr;
  if (g_profile.load() >= 2) {
    const size_t tb_bytes = sizeof(unsigned long long) * 2 * (size_t)(T + 1) * dd->grid * (4 + tb::kDenWarps);
    CUDA_OK(ws->cta_trace.reserve(tb_bytes));
    CUDA_OK(cudaMemsetAsync(ws->cta_trace.ptr, 0, tb_bytes, stream));
    dp.cta_trace = ws->cta_trace.as<unsigned long long>();
    ws->trace_frames = T + 1;
    ws->trace_grid = dd->grid;
  }
  dp.timeline = g_profile.load() ? reinterpret_cast<unsigned long long*>(zb + z_tl) : nullptr;
  ws->timeline_dev = zb + z_tl;
  ws->timeline_len = 2 * T + 6;
  bool numerator_done = false;
  for (int seq0 = 0; seq0 < B; seq0 += pass_cap) {
    const int pass_B = std::min(B - seq0, pass_cap);
    tb::DenBatch batch;
    batch.nctx = pass_B > tb::kCtxBP ? 2 : 1;
    dp.BP = batch.nctx == 2 ? tb::kCtxBP : pad_batch(pass_B);
    dp.subs = tb::den_subs(dp.BP);
    const bool tied_pass = tp && dp.BP == tb::kCtxBP;
    DenPartition* part = nullptr;
    if (tied_pass) {
      for (int dir = 0; dir < 2; dir++) {
        dp.dir[dir].arcs = tp->arcs[dir].as<int4>();
        dp.dir[dir].tasks = tp->tasks[dir].as<int4>();
        dp.dir[dir].cta_arc_off = tp->cta_arc_off[dir].as<int>();
        dp.dir[dir].cta_task_off = tp->cta_task_off[dir].as<int>();
        dp.dir[dir].warp_task_off = tp->warp_task_off[dir].as<int>();
        dp.dir[dir].warp_arc_off = tp->warp_arc_off[dir].as<int>();
        dp.dir[dir].max_arcs_cta = tp->max_arcs_cta[dir];
        dp.dir[dir].max_tasks_cta = tp->max_tasks_cta[dir];
      }
      dp.rdir.arcs = tp->runs.as<int4>();
      dp.rdir.tasks = nullptr;
      dp.rdir.cta_arc_off = tp->cta_run_off.as<int>();
      dp.rdir.cta_task_off = nullptr;
      dp.rdir.warp_task_off = nullptr;
      dp.rdir.warp_arc_off = tp->warp_run_off.as<int>();
      dp.rdir.max_arcs_cta = tp->max_runs_cta;
      dp.rdir.max_tasks_cta = 0;
      dp.R = g->R;
      dp.arun = ws->arun.as<float>();
      dp.wbuf = ws->wbuf.as<float>();
      dp.has_split = 0;
      dp.arcs_in_smem = 1;
    } else {
      if (!partition_for(dp.subs, &part)) return 0;
      bind(part);
      // arc / task slices are staged in shared memory unless the graph is too large for it
      dp.arcs_in_smem = 1;
      if (tb::den_smem_bytes(dp) > 200 * 1024 || g->hooks.force_unstaged) dp.arcs_in_smem = 0;
    }
    for (int ci = 0; ci < 2; ci++) {
      tb::DenParams& c = batch.c[ci];
      c = dp;
      const int cc = std::min(ci, batch.nctx - 1);  // an unused second slot mirrors the first
      c.seq0 = seq0 + cc * tb::kCtxBP;
      c.B = batch.nctx == 2 ? std::min(pass_B - cc * tb::kCtxBP, tb::kCtxBP) : pass_B;
      c.alpha = ws->alpha.as<float>() + alpha_ctx * cc;
      c.beta = ws->beta.as<float>() + beta_ctx * cc;
      c.etile = ws->tiles.as<float>() + tile_ctx * cc;
      char* zc = zb + zc_size * cc;
      c.gtile = reinterpret_cast<float*>(zc);
      c.Ssum = reinterpret_cast<double*>(zc + zc_S);
      c.Bsum = reinterpret_cast<double*>(zc + zc_B);
      c.barrier = reinterpret_cast<unsigned*>(zc + zc_bar);
    }
    CUDA_OK(cudaMemsetAsync(zb, 0, z_end, stream));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (g_profile.load()) {
      CUDA_OK(cudaEventCreate(&e0));
      CUDA_OK(cudaEventCreate(&e1));
      CUDA_OK(cudaEventRecord(e0, stream));
    }
    if (!numerator_done) {
      // The numerator runs on a side stream, on the registers the forward launch leaves free.
      // It needs all of nnet_output: resident from the start in the device entry, after the
      // last chunk has landed in the host-buffer entry (event recorded on the copy stream).
      CUDA_OK(cudaEventRecord(ws->ev_fork, stream));  // behind the supervision upload and the memsets
      CUDA_OK(cudaStreamWaitEvent(ws->side, ws->ev_fork, 0));
      if (chunked.all_landed) {
        CUDA_OK(cudaStreamWaitEvent(ws->side, chunked.all_landed, 0));
        if (chunked.all_landed2) CUDA_OK(cudaStreamWaitEvent(ws->side, chunked.all_landed2, 0));
      }
      CUDA_OK(tb::launch_numerator(np, ws->side));
      CUDA_OK(cudaEventRecord(ws->ev_join, ws->side));
    }
    CUDA_OK(tied_pass ? tb::launch_den_tied(batch, 0, dd->grid, stream) : tb::launch_den(batch, 0, dd->grid, stream));
    if (!numerator_done) {
      CUDA_OK(cudaStreamWaitEvent(stream, ws->ev_join, 0));
      numerator_done = true;
    }
    CUDA_OK(tied_pass ? tb::launch_den_tied(batch, 1, dd->grid, stream) : tb::launch_den(batch, 1, dd->grid, stream));
    if (calib_async) {
      std::lock_guard<std::mutex> lock(g->mu);
      if (!dd->calib_pending) {
        if (!dd->calib_ev) CUDA_OK(cudaEventCreateWithFlags(&dd->calib_ev, cudaEventDisableTiming));
        CUDA_OK(cudaEventRecord(dd->calib_ev, stream));
        dd->calib_part = part;
        dd->calib_pending = true;
      }
    }
    if (e0) {
      CUDA_OK(cudaEventRecord(e1, stream));
      std::lock_guard<std::mutex> lock(g->mu);
      ws->prof.emplace_back(e0, e1);
    }
  }

  // ---- final pass
  float* res_dev = reinterpret_cast<float*>(zb + z_res);
  if (want_xent) CUDA_OK(tb::launch_zero(xent, xl, T, B, P, stream));
  tb::FinalParams fp;
  fp.den_scalars = dp.scalars;
  fp.chain_logprob = np.chain_logprob;
  fp.num_chains = sup->num_chains;
  fp.B = B;
  fp.T = T;
  fp.sup_weight = sup->weight;
  fp.l2 = l2;
  fp.deriv_scale = deriv_scale;
  fp.results_dev = res_dev;
  fp.deriv = deriv;
  fp.dl = dl;
  fp.xent = want_xent ? xent : nullptr;
  fp.xl = xl;
  fp.y = y;
  fp.yl = yl;
  fp.rows = rows;
  fp.cols = cols;
  fp.num_arcs = sup->A;
  fp.arc_row = sd->arc_row.as<int>();
  fp.arc_pdf = sd->arc_pdf.as<int>();
  fp.arc_post = np.arc_post;
  CUDA_OK(tb::launch_final(fp, stream));
  if (!mark_sup_used(sup, sd, stream)) return 0;  // the numerator's side stream joined `stream` before the backward launch
  if (results_dev_user)
    CUDA_OK(cudaMemcpyAsync(results_dev_user, res_dev, 3 * sizeof(float), cudaMemcpyDeviceToDevice, stream));
  if (chunked.results4_dev)
    CUDA_OK(cudaMemcpyAsync(chunked.results4_dev, res_dev, 4 * sizeof(float), cudaMemcpyDeviceToDevice, stream));
  if (results_host) {
    CUDA_OK(cudaMemcpyAsync(ws->pinned, res_dev, 4 * sizeof(float), cudaMemcpyDeviceToHost, stream));
    CUDA_OK(cudaStreamSynchronize(stream));
    results_host[0] = ws->pinned[0];
    results_host[1] = ws->pinned[1];
    results_host[2] = ws->pinned[2];
    if (calibrating && !calib_async && !autotune_partition(g, dd)) return 0;
  }
  return 1;
}

// Rows of `width` bytes between pitched buffers; densely packed rows go as ONE linear copy (the copy
// engine then moves a single contiguous range instead of `rows` lines).
cudaError_t copy_rows(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t rows,
                      cudaMemcpyKind kind, cudaStream_t st) {
  if (rows == 0 || width == 0) return cudaSuccess;
  if (dpitch == width && spitch == width) return cudaMemcpyAsync(dst, src, width * rows, kind, st);
  return cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, rows, kind, st);
}

// Per-thread resources of the host-buffer entry.
struct HostPath {
  int device = -1;
  cudaStream_t compute = nullptr, copy = nullptr, copy2 = nullptr;
  cudaEvent_t ready = nullptr, landed = nullptr, landed2 = nullptr;
  DevBuf y, deriv, xent, flags, results4;
  unsigned* ones = nullptr;      // pinned constant 1s, one per frame flag of a chunk
  int ones_cap = 0;
  unsigned* progress = nullptr;  // pinned, device-mapped: frames of the derivative already complete
  HostPath() = default;
  HostPath(const HostPath&) = delete;
  HostPath& operator=(const HostPath&) = delete;
  // A thread_local of the main thread dies during process teardown, when the CUDA runtime may already be
  // gone: streams, events and pinned memory of an ending thread are left to the driver, and the device
  // buffers are let go WITHOUT cudaFree (DevBuf::abandon) for the same reason.
  ~HostPath() {
    y.abandon();
    deriv.abandon();
    xent.abandon();
    flags.abandon();
    results4.abandon();
  }
  void release() {  // everything belongs to `device`; called when the thread moves to another device
    if (compute) cudaStreamDestroy(compute);
    if (copy) cudaStreamDestroy(copy);
    if (copy2) cudaStreamDestroy(copy2);
    if (ready) cudaEventDestroy(ready);
    if (landed) cudaEventDestroy(landed);
    if (landed2) cudaEventDestroy(landed2);
    if (ones) cudaFreeHost(ones);
    if (progress) cudaFreeHost(progress);
    compute = copy = copy2 = nullptr;
    ready = landed = landed2 = nullptr;
    ones = nullptr;
    ones_cap = 0;
    progress = nullptr;
    y.release();
    deriv.release();
    xent.release();
    flags.release();
    results4.release();
    device = -1;
  }
};
thread_local HostPath g_host;


// ---------------------------------------------------------------------------
// Kaldi binary `Supervision` object (kaldi@22fbdd:src/chain/chain-supervision.cc
// Supervision::Write/Read, binary mode): tokens <Supervision> <Weight> float <NumSequences> int32
// <FramesPerSeq> int32 <LabelDim> int32 [<End2End> F], then the FST written with
// fst::StdCompactAcceptorFst::WriteFst (OpenFst CompactFst<StdArc, AcceptorCompactor>: FstHeader
// "compact_acceptor"/"standard", uint32 state offsets [numstates+1], 12-byte elements
// {int32 label, float weight, int32 nextstate}, a final weight being an element with label -1),
// then </Supervision>.  Layout restated from the published Kaldi / OpenFst sources; no
// Kaldi-written file is available offline to check it against.
// ---------------------------------------------------------------------------
struct ByteReader {
  const unsigned char* p;
  const unsigned char* end;
  bool ok = true;
  bool get(void* dst, size_t n) {
    if (!ok || (size_t)(end - p) < n) {
      ok = false;
      return false;
    }
    std::memcpy(dst, p, n);
    p += n;
    return true;
  }
  bool token(const char* want) {  // "<Token> " (token followed by one space)
    const size_t n = std::strlen(want);
    if (!ok || (size_t)(end - p) < n + 1 || std::memcmp(p, want, n) != 0 || p[n] != ' ') return false;
    p += n + 1;
    return true;
  }
  template <typename T>
  bool basic(T* v) {  // Kaldi WriteBasicType: one size byte, then the value
    unsigned char sz = 0;
    if (!get(&sz, 1) || sz != sizeof(T)) {
      ok = false;
      return false;
    }
    return get(v, sizeof(T));
  }
  bool str(std::string* s) {  // OpenFst string: int32 length + bytes
    int32_t n = 0;
    if (!get(&n, 4) || n < 0 || n > 4096) {
      ok = false;
      return false;
    }
    s->resize(n);
    return n == 0 || get(&(*s)[0], n);
  }
};

Supervision* parse_kaldi_supervision(const void* data, int64_t nbytes, int64_t* consumed) {
  ByteReader r{static_cast<const unsigned char*>(data), static_cast<const unsigned char*>(data) + nbytes};
  // A Kaldi binary stream announces itself with "\0B" (e.g. right after the key of an ark entry): skip it.  A text-mode
  // object ("<Supervision> <Weight> 1 ...") fails the size-byte checks below and is reported as such.
  if (nbytes >= 2 && r.p[0] == 0 && r.p[1] == 'B') r.p += 2;
  float weight = 0.f;
  int32_t B = 0, T = 0, P = 0;
  if (!r.token("<Supervision>") || !r.token("<Weight>") || !r.basic(&weight) || !r.token("<NumSequences>") ||
      !r.basic(&B) || !r.token("<FramesPerSeq>") || !r.basic(&T) || !r.token("<LabelDim>") || !r.basic(&P)) {
    g_error = "not a binary Kaldi Supervision object (header tokens)";
    return nullptr;
  }
  if (r.token("<End2End>")) {  // written by later Kaldi revisions
    unsigned char e2e = 0;
    if (!r.get(&e2e, 1) || e2e != 'F') {
      g_error = "end-to-end supervisions are not supported";
      return nullptr;
    }
  }
  int32_t magic = 0, version = 0, flags = 0;
  uint64_t props = 0;
  int64_t start = 0, nstates = 0, narcs = 0;
  std::string fsttype, arctype;
  if (!r.get(&magic, 4) || magic != 2125659606 || !r.str(&fsttype) || !r.str(&arctype) || !r.get(&version, 4) ||
      !r.get(&flags, 4) || !r.get(&props, 8) || !r.get(&start, 8) || !r.get(&nstates, 8) || !r.get(&narcs, 8)) {
    g_error = "Supervision: truncated or foreign FST header";
    return nullptr;
  }
  if (fsttype != "compact_acceptor" || arctype != "standard" || (flags & 7) != 0) {
    g_error = "Supervision: expected an unaligned compact_acceptor/standard FST without symbol tables (got '" +
              fsttype + "'/'" + arctype + "')";
    return nullptr;
  }
  if (nstates <= 0 || nstates > (1 << 28) || start != 0) {
    g_error = "Supervision: implausible FST header (states / start)";
    return nullptr;
  }
  std::vector<uint32_t> offs(nstates + 1);
  if (!r.get(offs.data(), sizeof(uint32_t) * (nstates + 1))) {
    g_error = "Supervision: truncated state table";
    return nullptr;
  }
  const uint32_t ncompacts = offs[nstates];
  struct Element {
    int32_t label;
    float weight;
    int32_t nextstate;
  };
  std::vector<Element> el(ncompacts);
  if (ncompacts > 0 && !r.get(el.data(), sizeof(Element) * (size_t)ncompacts)) {
    g_error = "Supervision: truncated arc table";
    return nullptr;
  }
  if (!r.token("</Supervision>")) {
    g_error = "Supervision: missing </Supervision>";
    return nullptr;
  }
  std::vector<int32_t> arc_begin(nstates + 1, 0), arc_dst, arc_label;
  std::vector<float> arc_cost, final_cost(nstates, kInf);
  for (int64_t s = 0; s < nstates; s++) {
    if (offs[s + 1] < offs[s] || offs[s + 1] > ncompacts) {
      g_error = "Supervision: corrupt state table";
      return nullptr;
    }
    for (uint32_t k = offs[s]; k < offs[s + 1]; k++) {
      if (el[k].label == -1) {
        final_cost[s] = el[k].weight;
      } else {
        arc_dst.push_back(el[k].nextstate);
        arc_label.push_back(el[k].label);
        arc_cost.push_back(el[k].weight);
      }
    }
    arc_begin[s + 1] = (int32_t)arc_dst.size();
  }
  if (consumed) *consumed = (int64_t)(r.p - static_cast<const unsigned char*>(data));
  return make_supervision(weight, B, T, P, (int32_t)nstates, arc_begin.data(), arc_dst.data(), arc_label.data(),
                          arc_cost.data(), final_cost.data());
}

// Synchronises the three streams of the host-buffer entry when an error return leaves its scope, so that no copy
// still reads or writes the caller's host buffers (and no kernel still waits for chunks) after the call returned.
struct HostDrain {
  HostPath& h;
  bool armed = true;
  ~HostDrain() {
    if (!armed) return;
    cudaStreamSynchronize(h.copy);
    cudaStreamSynchronize(h.copy2);
    cudaStreamSynchronize(h.compute);
  }
};

int host_entry(DenGraph* g, Supervision* sp, const float* y_host, int64_t rows, int64_t cols, int64_t ystride,
               float* results, float* deriv_host, int64_t dstride, float* xent_host, int64_t xstride, float l2,
               float leaky, float xent_regularize, int device) {
  HostPath& h = g_host;
  if (h.device != device) {
    h.release();
    CUDA_OK(cudaStreamCreateWithFlags(&h.compute, cudaStreamNonBlocking));
    CUDA_OK(cudaStreamCreateWithFlags(&h.copy, cudaStreamNonBlocking));
    CUDA_OK(cudaStreamCreateWithFlags(&h.copy2, cudaStreamNonBlocking));
    CUDA_OK(cudaEventCreateWithFlags(&h.ready, cudaEventDisableTiming));
    CUDA_OK(cudaEventCreateWithFlags(&h.landed, cudaEventDisableTiming));
    CUDA_OK(cudaEventCreateWithFlags(&h.landed2, cudaEventDisableTiming));
    CUDA_OK(cudaHostAlloc(&h.progress, sizeof(unsigned), cudaHostAllocMapped));
    CUDA_OK(h.results4.reserve(4 * sizeof(float)));
    h.device = device;
  }
  const size_t bytes = sizeof(float) * (size_t)rows * cols;
  CUDA_OK(h.y.reserve(bytes));
  const bool want_xent = xent_host != nullptr && xent_regularize != 0.f;
  if (want_xent) CUDA_OK(h.xent.reserve(bytes));

  // The derivative is produced in device memory and copied out chunk by chunk (whole frames)
  // by the copy engine WHILE the backward recurrence is still running: the kernel publishes how
  // many trailing frames are complete in a mapped pinned counter (system-scope release), this
  // (blocked anyway) host thread polls it and enqueues the copies.  Only the chunk holding frame 0
  // waits for the end.
  float* deriv_dev = nullptr;
  const int64_t deriv_dev_stride = cols;
  if (deriv_host) {
    CUDA_OK(h.deriv.reserve(bytes));
    deriv_dev = h.deriv.as<float>();
  }
  // nnet_output goes up in chunks of whole frames (chunk_bounds), alternating between TWO copy streams
  // (measured, profiles/r02_host_trace_variants.txt: 3.78 ms per call against 3.92 ms with one stream --
  // the 4-byte flag copy behind every chunk no longer stalls the next chunk's transfer; setting the flags
  // by memset or by stream memory operations instead made no difference and was dropped).  The forward
  // launch starts as soon as the first chunk has landed and waits, per frame, on the flag of the frame it
  // stages next.
  const int T = sp->T, B = sp->B;
  // TORCHAIN_HOST_TRACE=1: one line per call on stderr with the host-side times (us since entry) at which
  // everything was enqueued, the first / last early chunk left, the compute stream and the copies ended
  static const bool trace = getenv("TORCHAIN_HOST_TRACE") != nullptr;
  auto now_us = [] {
    return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
  };
  const double tr0 = trace ? now_us() : 0.0;
  double tr_enq = 0.0, tr_first = -1.0, tr_last = -1.0, tr_sync = 0.0;
  const std::vector<int> cb = chunk_bounds(T);
  const int nchunks = (int)cb.size() - 1;
  if (h.ones_cap < T) {  // no copy of an earlier call is in flight: every call ends with a sync of both copy streams
    if (h.ones) cudaFreeHost(h.ones);
    h.ones = nullptr;
    h.ones_cap = 0;
    const int cap = std::max(T, 64);
    CUDA_OK(cudaMallocHost(&h.ones, sizeof(unsigned) * (size_t)cap));
    for (int i = 0; i < cap; i++) h.ones[i] = 1u;
    h.ones_cap = cap;
  }
  CUDA_OK(h.flags.reserve(sizeof(unsigned) * (size_t)T));
  HostDrain drain{h};  // from here on, an error return waits for whatever has been enqueued
  CUDA_OK(cudaMemsetAsync(h.flags.ptr, 0, sizeof(unsigned) * (size_t)T, h.compute));
  CUDA_OK(cudaEventRecord(h.ready, h.compute));
  CUDA_OK(cudaStreamWaitEvent(h.copy, h.ready, 0));
  CUDA_OK(cudaStreamWaitEvent(h.copy2, h.ready, 0));
  for (int k = 0; k < nchunks; k++) {
    const int64_t r0 = (int64_t)cb[k] * B, nr = (int64_t)(cb[k + 1] - cb[k]) * B;
    const cudaStream_t st = (k & 1) ? h.copy2 : h.copy;
    CUDA_OK(copy_rows(h.y.as<float>() + r0 * cols, sizeof(float) * cols, y_host + r0 * ystride,
                      sizeof(float) * ystride, sizeof(float) * cols, nr, cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaMemcpyAsync(h.flags.as<unsigned>() + cb[k], h.ones, sizeof(unsigned) * (size_t)(cb[k + 1] - cb[k]),
                            cudaMemcpyHostToDevice, st));
  }
  CUDA_OK(cudaEventRecord(h.landed, h.copy));
  CUDA_OK(cudaEventRecord(h.landed2, h.copy2));
  ChunkedInput ci;
  ci.flags = h.flags.as<unsigned>();
  ci.chunk_frames = 1;  // one flag per frame
  ci.all_landed = h.landed;
  ci.all_landed2 = h.landed2;
  const bool pipelined_out = deriv_host != nullptr && B <= tb::kMaxBP;
  if (pipelined_out) {
    *reinterpret_cast<volatile unsigned*>(h.progress) = 0u;
    unsigned* dptr = nullptr;
    CUDA_OK(cudaHostGetDevicePointer(reinterpret_cast<void**>(&dptr), h.progress, 0));
    ci.deriv_progress = dptr;
    if (T <= 512) {  // chunk k (frames cb[k] ...) may leave once T - cb[k] trailing frames are complete
      for (int i = 0; i < 8; i++) ci.progress_mask[i] = 0ull;
      for (int k = 1; k < nchunks; k++) {
        const int v = T - cb[k];
        ci.progress_mask[(v >> 6) & 7] |= 1ull << (v & 63);
      }
    }
  }
  ci.results4_dev = h.results4.as<float>();
  // enqueue everything without synchronising (results are fetched below)
  if (!compute(g, sp, h.y.as<float>(), rows, cols, row_layout(B, cols), nullptr, nullptr, deriv_dev,
               row_layout(B, deriv_dev_stride), want_xent ? h.xent.as<float>() : nullptr, row_layout(B, cols), l2,
               leaky, xent_regularize, 1.0f, device, h.compute, ci))
    return 0;
  auto copy_chunk = [&](int k, cudaStream_t st) -> cudaError_t {
    const int64_t r0 = (int64_t)cb[k] * B, nr = (int64_t)(cb[k + 1] - cb[k]) * B;
    return copy_rows(deriv_host + r0 * dstride, sizeof(float) * dstride, deriv_dev + r0 * cols,
                     sizeof(float) * cols, sizeof(float) * cols, nr, cudaMemcpyDeviceToHost, st);
  };
  int next = nchunks - 1;
  if (trace) tr_enq = now_us() - tr0;
  if (pipelined_out) {
    // frames >= T - progress are complete; chunk k covers frames [cb[k], cb[k+1])
    while (next >= 1) {
      const unsigned need = (unsigned)(T - cb[next]);
      bool done = false;
      // acquire: pairs with the kernel's st.release.sys, so the copy enqueued below is ordered behind the rows
      while (__atomic_load_n(h.progress, __ATOMIC_ACQUIRE) < need) {
        if (cudaStreamQuery(h.compute) != cudaErrorNotReady) {
          done = true;
          break;
        }
      }
      if (done) break;
      CUDA_OK(copy_chunk(next, h.copy));
      if (trace) {
        tr_last = now_us() - tr0;
        if (tr_first < 0.0) tr_first = tr_last;
      }
      next--;
    }
  }
  const int first_early = next + 1;  // chunks [first_early, nchunks) left before the final pass had run
  CUDA_OK(cudaStreamSynchronize(h.compute));
  CUDA_OK(cudaStreamSynchronize(h.copy2));  // long done: the forward launch consumed its chunks
  if (trace) tr_sync = now_us() - tr0;
  if (deriv_host)
    for (; next >= 0; next--) CUDA_OK(copy_chunk(next, h.copy));
  if (want_xent)
    CUDA_OK(copy_rows(xent_host, sizeof(float) * xstride, h.xent.ptr, sizeof(float) * cols,
                      sizeof(float) * cols, rows, cudaMemcpyDeviceToHost, h.copy));
  float res_host[4];
  CUDA_OK(cudaMemcpyAsync(res_host, ci.results4_dev, sizeof(res_host), cudaMemcpyDeviceToHost, h.copy));
  CUDA_OK(cudaStreamSynchronize(h.copy));
  if (trace)
    fprintf(stderr, "host_trace T=%d chunks=%d early=%d enqueued=%.0f first_early=%.0f last_early=%.0f compute_done=%.0f end=%.0f us\n",
            T, nchunks, nchunks - first_early, tr_enq, tr_first, tr_last, tr_sync, now_us() - tr0);
  if (deriv_host && first_early < nchunks && res_host[3] == 0.f) {
    // The numeric guard tripped (Kaldi: objf = -10 * weight, derivative reset to the l2 part): the final
    // pass rewrote the derivative AFTER the early chunks had left, so they go out again.
    for (int k = first_early; k < nchunks; k++) CUDA_OK(copy_chunk(k, h.copy));
    CUDA_OK(cudaStreamSynchronize(h.copy));
  }
  drain.armed = false;  // everything has been waited for
  results[0] = res_host[0];
  results[1] = res_host[1];
  results[2] = res_host[2];
  return 1;
}

// nnet3 chain examples (kaldi@22fbdd:src/nnet3/nnet-chain-example.cc, NnetChainExample::Write, binary):
//   [key ' ' '\0' 'B'] <Nnet3ChainEg> <NumInputs> ... NnetIo ... <NumOutputs> n
//   <NnetChainSup> name <I1V> indexes <Supervision> ... </Supervision> <DW>|<DW2> deriv weights </NnetChainSup> ...
//   </Nnet3ChainEg>
// The reference pulls the Supervision out of such an example with Kaldi itself
// (/root/reference/src/my_lib_example.cpp:71-76: eg.outputs[0].supervision) and finds the examples in an
// ark by scanning for the <Nnet3ChainEg> token (/root/reference/example/chime5/ark2scp.py:24-40).  Same
// scan here: the first <Supervision> object after the next <Nnet3ChainEg> token is parsed, everything
// else in the example (features, i-vectors, derivative weights) is skipped.
const unsigned char* find_bytes(const unsigned char* p, const unsigned char* end, const char* token) {
  const size_t n = std::strlen(token);
  if ((size_t)(end - p) < n) return nullptr;
  const unsigned char* last = end - n;
  for (; p <= last; p++) {
    p = static_cast<const unsigned char*>(std::memchr(p, token[0], (size_t)(last - p) + 1));
    if (!p) return nullptr;
    if (std::memcmp(p, token, n) == 0) return p;
  }
  return nullptr;
}

Supervision* parse_chain_eg(const void* data, int64_t nbytes, int64_t* consumed) {
  const unsigned char* base = static_cast<const unsigned char*>(data);
  const unsigned char* end = base + nbytes;
  const unsigned char* eg = find_bytes(base, end, "<Nnet3ChainEg>");
  if (!eg) {
    g_error = "no <Nnet3ChainEg> token in the buffer";
    return nullptr;
  }
  // binary examples are announced by "\0B" right before the token; text-mode egs cannot be parsed here
  if (eg - base >= 2 && !(eg[-2] == 0 && eg[-1] == 'B')) {
    g_error = "text-mode nnet3 chain examples are not supported (expected the binary marker before <Nnet3ChainEg>)";
    return nullptr;
  }
  const unsigned char* next_eg = find_bytes(eg + 14, end, "<Nnet3ChainEg>");
  const unsigned char* sup = find_bytes(eg, next_eg ? next_eg : end, "<Supervision> ");
  if (!sup) {
    g_error = "<Nnet3ChainEg> without a <Supervision> object";
    return nullptr;
  }
  int64_t used = 0;
  Supervision* out = parse_kaldi_supervision(sup, (int64_t)((next_eg ? next_eg : end) - sup), &used);
  if (!out) return nullptr;
  const unsigned char* tail = find_bytes(sup + used, next_eg ? next_eg : end, "</Nnet3ChainEg>");
  const unsigned char* stop = tail ? tail + 15 : sup + used;
  if (tail && stop < end && *stop == ' ') stop++;
  if (consumed) *consumed = (int64_t)(stop - base);
  return out;
}

// Nothing may propagate through the C ABI: a std::bad_alloc / std::length_error from a corrupt size field
// or an exhausted host ends as an error return with torchain_last_error() set.
int abi_exception_int() {
  try {
    throw;
  } catch (const std::exception& e) {
    g_error = std::string("internal error: ") + e.what();
  } catch (...) {
    g_error = "internal error: unknown exception";
  }
  return 0;
}
void* abi_exception_ptr() {
  abi_exception_int();
  return nullptr;
}

}  // namespace

// ===========================================================================
// C ABI
// ===========================================================================
extern "C" {

const char* torchain_last_error(void) { return g_error.c_str(); }
int torchain_abi_version(void) { return 1; }
int64_t torchain_kernel_launch_count(void) { return tb::launch_count(); }

int my_lib_set_kaldi_device(int device) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return fail("no CUDA driver / device");
  if (device < 0 || device >= n) return fail("CUDA device index out of range");
  return 1;
}

void* torchain_denominator_graph_from_arrays(int32_t num_states, int32_t num_arcs, const int32_t* src,
                                             const int32_t* dst, const int32_t* ilabel, const float* cost,
                                             int32_t start, const float* final_cost, int32_t num_pdfs) try {
  if (num_arcs > 0 && (!src || !dst || !ilabel || !cost)) return fail_ptr("denominator graph: null arc arrays");
  if (!final_cost) return fail_ptr("denominator graph: null final_cost");
  return make_den_graph(num_states, num_arcs, src, dst, ilabel, cost, start, final_cost, num_pdfs);
} catch (...) {
  return abi_exception_ptr();
}

void* my_lib_denominator_graph_new(const char* rxfilename, int num_pdf) try {
  if (!rxfilename) return fail_ptr("null rxfilename");
  std::vector<int32_t> src, dst, il;
  std::vector<float> cost, fin;
  int32_t start = 0;
  if (!read_fst_file(rxfilename, &src, &dst, &il, &cost, &start, &fin)) return nullptr;
  return make_den_graph((int32_t)fin.size(), (int32_t)src.size(), src.data(), dst.data(), il.data(), cost.data(),
                        start, fin.data(), num_pdf);
} catch (...) {
  return abi_exception_ptr();
}

void my_lib_denominator_graph_free(void* p) { delete static_cast<DenGraph*>(p); }
int torchain_denominator_graph_num_states(void* p) { return p ? static_cast<DenGraph*>(p)->N : 0; }
int torchain_denominator_graph_num_arcs(void* p) { return p ? static_cast<DenGraph*>(p)->A : 0; }
int torchain_denominator_graph_num_pdfs(void* p) { return p ? static_cast<DenGraph*>(p)->P : 0; }
int torchain_denominator_graph_num_runs(void* p) { return p ? static_cast<DenGraph*>(p)->R : 0; }
int torchain_denominator_graph_is_tied(void* p) { return p && static_cast<DenGraph*>(p)->tied ? 1 : 0; }
int torchain_denominator_graph_initial_probs(void* p, float* out) try {
  if (!p || !out) return fail("null argument");
  auto* g = static_cast<DenGraph*>(p);
  std::memcpy(out, g->pi.data(), sizeof(float) * g->N);
  return 1;
} catch (...) {
  return abi_exception_int();
}

void* torchain_supervision_from_arrays(float weight, int32_t num_sequences, int32_t frames_per_sequence,
                                       int32_t label_dim, int32_t num_states, const int32_t* arc_begin,
                                       const int32_t* arc_dst, const int32_t* arc_label, const float* arc_cost,
                                       const float* final_cost) try {
  if (!arc_begin || !final_cost) return fail_ptr("supervision: null arrays");
  if (arc_begin[num_states > 0 ? num_states : 0] > 0 && (!arc_dst || !arc_label || !arc_cost))
    return fail_ptr("supervision: null arc arrays");
  return make_supervision(weight, num_sequences, frames_per_sequence, label_dim, num_states, arc_begin, arc_dst,
                          arc_label, arc_cost, final_cost);
} catch (...) {
  return abi_exception_ptr();
}

void* torchain_supervision_from_kaldi_bytes(const void* data, int64_t nbytes, int64_t* consumed) try {
  if (!data || nbytes <= 0) return fail_ptr("supervision: null / empty buffer");
  return parse_kaldi_supervision(data, nbytes, consumed);
} catch (...) {
  return abi_exception_ptr();
}

void* torchain_supervision_from_chain_egs(const void* data, int64_t nbytes, int64_t* consumed) try {
  if (!data || nbytes <= 0) return fail_ptr("supervision: null / empty buffer");
  return parse_chain_eg(data, nbytes, consumed);
} catch (...) {
  return abi_exception_ptr();
}

void* my_lib_supervision_new(void*) {
  return fail_ptr("Kaldi egs readers are outside the chain-loss hot path; use torchain_supervision_from_arrays");
}
void* my_lib_supervision_rand_new(void*) {
  return fail_ptr("Kaldi egs readers are outside the chain-loss hot path; use torchain_supervision_from_arrays");
}
// Minibatch assembly (the hot-path part of what the reference gets from Kaldi's MergeChainExamples /
// AppendSupervision, /root/reference/src/my_lib_example_rand.cpp:160): examples with equal weight,
// frames_per_sequence and label_dim are appended sequence-major; Kaldi concatenates the FSTs and removes
// the epsilons, which leaves the per-sequence chains independent -- exactly the split representation
// kept here, so merging is a concatenation of chain lists with shifted state / arc / sequence indices.
void* torchain_supervision_merge(void* const* sups, int32_t n) try {
  if (!sups || n <= 0) return fail_ptr("supervision merge: empty list");
  for (int i = 0; i < n; i++)
    if (!sups[i]) return fail_ptr("supervision merge: null handle");
  const Supervision* first = static_cast<const Supervision*>(sups[0]);
  int64_t Btot = 0;
  for (int i = 0; i < n; i++) {
    const Supervision* in = static_cast<const Supervision*>(sups[i]);
    // Kaldi merges only equal weights / lengths and fails the minibatch otherwise
    // ("Failed to merge 'chain' examples-- inconsistent lengths or weights?")
    if (in->weight != first->weight || in->T != first->T || in->P != first->P)
      return fail_ptr("supervision merge: inconsistent weight, frames_per_sequence or label_dim");
    Btot += in->B;
  }
  if (Btot * first->T >= (int64_t)1 << 31) return fail_ptr("supervision merge: too many frames");
  auto out = std::make_unique<Supervision>();
  out->weight = first->weight;
  out->T = first->T;
  out->P = first->P;
  out->B = (int)Btot;
  out->chain_level_off.assign(1, 0);
  out->chain_arc_off.assign(1, 0);
  out->in_off.assign(1, 0);
  out->out_off.assign(1, 0);
  int seq_off = 0;
  for (int i = 0; i < n; i++) {
    const Supervision* in = static_cast<const Supervision*>(sups[i]);
    const int state_off = out->S, arc_off = out->A, lvl_off = (int)out->level_state_off.size();
    for (int v : in->level_state_off) out->level_state_off.push_back(v + state_off);
    for (size_t k = 1; k < in->chain_level_off.size(); k++) out->chain_level_off.push_back(in->chain_level_off[k] + lvl_off);
    for (size_t k = 1; k < in->chain_arc_off.size(); k++) out->chain_arc_off.push_back(in->chain_arc_off[k] + arc_off);
    for (size_t k = 1; k < in->in_off.size(); k++) out->in_off.push_back(in->in_off[k] + arc_off);
    for (size_t k = 1; k < in->out_off.size(); k++) out->out_off.push_back(in->out_off[k] + arc_off);
    for (const int2& a : in->in_arcs) out->in_arcs.push_back(make_int2(a.x + state_off, a.y + arc_off));
    for (const int2& a : in->out_arcs) out->out_arcs.push_back(make_int2(a.x + state_off, a.y + arc_off));
    out->final_cost.insert(out->final_cost.end(), in->final_cost.begin(), in->final_cost.end());
    out->arc_cost.insert(out->arc_cost.end(), in->arc_cost.begin(), in->arc_cost.end());
    out->arc_pdf.insert(out->arc_pdf.end(), in->arc_pdf.begin(), in->arc_pdf.end());
    for (int a = 0; a < in->A; a++) {
      const int seq = in->arc_seq[a], frame = (in->arc_row[a] - seq) / in->B;  // row = frame * B + seq
      out->arc_seq.push_back(seq + seq_off);
      out->arc_row.push_back(frame * out->B + seq + seq_off);
    }
    out->num_chains += in->num_chains;
    out->S += in->S;
    out->A += in->A;
    seq_off += in->B;
  }
  // arcs ordered by frame, for the backward launch's per-frame numerator scatter
  const int T = out->T, NA = out->A;
  out->frame_off.assign(T + 1, 0);
  for (int a = 0; a < NA; a++) out->frame_off[out->arc_row[a] / out->B + 1]++;
  for (int t = 0; t < T; t++) out->frame_off[t + 1] += out->frame_off[t];
  out->frame_arcs.resize(NA);
  std::vector<int> pos(out->frame_off.begin(), out->frame_off.end() - 1);
  for (int a = 0; a < NA; a++) out->frame_arcs[pos[out->arc_row[a] / out->B]++] = a;
  return out.release();
} catch (...) {
  return abi_exception_ptr();
}

int torchain_supervision_prefetch(void* sup, int device, void* stream) try {
  if (!sup) return fail("null supervision handle");
  DeviceGuard guard(device);
  if (!guard.ok) return fail("cannot select CUDA device");
  SupDevice* sd = nullptr;
  return get_sup_device(static_cast<Supervision*>(sup), device, static_cast<cudaStream_t>(stream), &sd);
} catch (...) {
  return abi_exception_int();
}

void torchain_supervision_pool_trim(void) { sup_pool().trim(); }

void my_lib_supervision_free(void* p) { delete static_cast<Supervision*>(p); }
int my_lib_supervision_num_pdf(void* p) { return p ? static_cast<Supervision*>(p)->P : 0; }
int my_lib_supervision_num_sequence(void* p) { return p ? static_cast<Supervision*>(p)->B : 0; }
int my_lib_supervision_num_frame(void* p) { return p ? static_cast<Supervision*>(p)->T : 0; }
int torchain_supervision_num_chains(void* p) { return p ? static_cast<Supervision*>(p)->num_chains : 0; }

int torchain_ComputeChainObjfAndDeriv_ex(void* den, void* sup, const float* y, int64_t rows, int64_t cols,
                                         int64_t ystride, float* results, float* results_dev, float* deriv,
                                         int64_t dstride, float* xent, int64_t xstride, float l2, float leaky,
                                         float xent_regularize, float deriv_scale, int device, void* stream) try {
  if (!sup) return fail("null den_graph / supervision handle");
  if (ystride < cols || (deriv && dstride < cols) || (xent && xstride < cols)) return fail("row stride < cols");
  const int B = static_cast<Supervision*>(sup)->B;
  return compute(static_cast<DenGraph*>(den), static_cast<Supervision*>(sup), y, rows, cols, row_layout(B, ystride),
                 results, results_dev, deriv, row_layout(B, dstride), xent, row_layout(B, xstride), l2, leaky,
                 xent_regularize, deriv_scale, device, static_cast<cudaStream_t>(stream));
} catch (...) {
  return abi_exception_int();
}

int torchain_ComputeChainObjfAndDeriv_strided(void* den, void* sup, const float* y, int64_t num_frames,
                                              int64_t num_sequences, int64_t num_pdfs, const int64_t* y_strides,
                                              float* results, float* results_dev, float* deriv,
                                              const int64_t* deriv_strides, float* xent, const int64_t* xent_strides,
                                              float l2, float leaky, float xent_regularize, float deriv_scale,
                                              int device, void* stream) try {
  if (!sup) return fail("null den_graph / supervision handle");
  auto* sp = static_cast<Supervision*>(sup);
  if (num_frames != sp->T || num_sequences != sp->B)
    return fail("nnet_output frames / sequences != supervision frames_per_sequence / num_sequences");
  if (!y_strides || (deriv && !deriv_strides) || (xent && !xent_strides)) return fail("null stride triple");
  auto lay = [](const int64_t* st) { return st ? tb::Layout3{st[0], st[1], st[2]} : tb::Layout3{0, 0, 0}; };
  return compute(static_cast<DenGraph*>(den), sp, y, num_frames * num_sequences, num_pdfs, lay(y_strides), results,
                 results_dev, deriv, lay(deriv_strides), xent, lay(xent_strides), l2, leaky, xent_regularize,
                 deriv_scale, device, static_cast<cudaStream_t>(stream));
} catch (...) {
  return abi_exception_int();
}

int my_lib_ComputeChainObjfAndDeriv(void* den, void* sup, const float* y, int64_t rows, int64_t cols,
                                    int64_t ystride, float* results, float* deriv, int64_t dstride, float* xent,
                                    int64_t xstride, float l2, float leaky, float xent_regularize, int device,
                                    void* stream) try {
  if (!results) return fail("null results");
  if (!sup) return fail("null den_graph / supervision handle");
  if (ystride < cols || (deriv && dstride < cols) || (xent && xstride < cols)) return fail("row stride < cols");
  const int B = static_cast<Supervision*>(sup)->B;
  return compute(static_cast<DenGraph*>(den), static_cast<Supervision*>(sup), y, rows, cols, row_layout(B, ystride),
                 results, nullptr, deriv, row_layout(B, dstride), xent, row_layout(B, xstride), l2, leaky,
                 xent_regularize, 1.0f, device, static_cast<cudaStream_t>(stream));
} catch (...) {
  return abi_exception_int();
}

int torchain_to2d(const float* x_bpt, int64_t B, int64_t P, int64_t T, float* out_rows, int64_t out_stride, int device,
                  void* stream) try {
  if (!x_bpt || !out_rows) return fail("null tensor");
  if (B < 0 || P < 0 || T < 0 || out_stride < P) return fail("bad shape / row stride < cols");
  DeviceGuard guard(device);
  if (!guard.ok) return fail("cannot select CUDA device");
  cudaError_t e = tb::launch_to2d(x_bpt, out_rows, (int)B, (int)P, (int)T, out_stride, static_cast<cudaStream_t>(stream));
  if (e == cudaErrorInvalidValue) return fail("to2d: frames_per_sequence or num_sequences too large for the tiled kernel");
  CUDA_OK(e);
  return 1;
} catch (...) {
  return abi_exception_int();
}

int torchain_from2d(const float* rows, int64_t in_stride, int64_t B, int64_t P, int64_t T, float* out_bpt, float scale,
                    int device, void* stream) try {
  if (!rows || !out_bpt) return fail("null tensor");
  if (B < 0 || P < 0 || T < 0 || in_stride < P) return fail("bad shape / row stride < cols");
  DeviceGuard guard(device);
  if (!guard.ok) return fail("cannot select CUDA device");
  cudaError_t e = tb::launch_from2d(rows, out_bpt, (int)B, (int)P, (int)T, in_stride, scale, static_cast<cudaStream_t>(stream));
  if (e == cudaErrorInvalidValue) return fail("from2d: frames_per_sequence or num_sequences too large for the tiled kernel");
  CUDA_OK(e);
  return 1;
} catch (...) {
  return abi_exception_int();
}

int torchain_ComputeChainObjfAndDeriv_host(void* den, void* sup, const float* y_host, int64_t rows, int64_t cols,
                                           int64_t ystride, float* results, float* deriv_host, int64_t dstride,
                                           float* xent_host, int64_t xstride, float l2, float leaky,
                                           float xent_regularize, int device) try {
  if (!den || !sup) return fail("null den_graph / supervision handle");
  if (!y_host || !results) return fail("null nnet_output / results");
  DeviceGuard guard(device);
  if (!guard.ok) return fail("cannot select CUDA device");
  auto* g = static_cast<DenGraph*>(den);
  auto* sp = static_cast<Supervision*>(sup);
  if (rows != (int64_t)sp->B * sp->T) return fail("nnet_output rows != num_sequences * frames_per_sequence");
  if (cols != g->P) return fail("nnet_output cols != den_graph.NumPdfs()");
  if (ystride < cols || (deriv_host && dstride < cols) || (xent_host && xstride < cols))
    return fail("row stride < cols");
  return host_entry(g, sp, y_host, rows, cols, ystride, results, deriv_host, dstride, xent_host, xstride, l2, leaky,
                    xent_regularize, device);
} catch (...) {
  return abi_exception_int();
}

int torchain_host_chunk_bounds(int num_frames, int* bounds, int max_entries) try {
  if (num_frames < 1) return 0;
  const std::vector<int> b = chunk_bounds(num_frames);
  for (int i = 0; i < (int)b.size() && i < max_entries && bounds; i++) bounds[i] = b[i];
  return (int)b.size();
} catch (...) {
  return abi_exception_int();
}

void torchain_profile_enable(int on) { g_profile.store(on); }
int torchain_profile_read(void* den_graph_ptr, double* den_kernel_ms, int64_t* den_kernel_launches, int reset) {
  if (!den_graph_ptr) return fail("null den_graph handle");
  harvest_profile(static_cast<DenGraph*>(den_graph_ptr));
  std::lock_guard<std::mutex> lock(g_prof_mu);
  if (den_kernel_ms) *den_kernel_ms = g_prof_den_ms;
  if (den_kernel_launches) *den_kernel_launches = g_prof_den_launches;
  if (reset) {
    g_prof_den_ms = 0.0;
    g_prof_den_launches = 0;
  }
  return 1;
}

int torchain_profile_timeline(void* den_graph_ptr, int64_t* out_ns, int max_entries) {
  if (!den_graph_ptr || !out_ns) return 0;
  auto* g = static_cast<DenGraph*>(den_graph_ptr);
  std::lock_guard<std::mutex> lock(g->mu);
  for (auto& kv : g->ws) {
    Workspace* w = kv.second.get();
    if (!w->timeline_dev || w->timeline_len <= 0) continue;
    const int n = std::min(max_entries, w->timeline_len);
    cudaSetDevice(kv.first.first);
    if (cudaMemcpy(out_ns, w->timeline_dev, sizeof(int64_t) * n, cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
    return n;
  }
  return 0;
}

// level-2 profiling (torchain_profile_enable(2)): per-CTA, per-frame trace of the last launch pair.
// out: [2][frames][grid][4] int64 ns; returns the number of int64 written, dims through the out params.
int64_t torchain_profile_cta_trace(void* den_graph_ptr, int64_t* out_ns, int64_t max_entries, int* frames, int* grid) {
  if (!den_graph_ptr) return 0;
  auto* g = static_cast<DenGraph*>(den_graph_ptr);
  std::lock_guard<std::mutex> lock(g->mu);
  for (auto& kv : g->ws) {
    Workspace* w = kv.second.get();
    if (!w->cta_trace.ptr || w->trace_frames <= 0) continue;
    const int64_t n = (int64_t)2 * w->trace_frames * w->trace_grid * (4 + tb::kDenWarps);
    if (frames) *frames = w->trace_frames;
    if (grid) *grid = w->trace_grid;
    if (!out_ns || max_entries < n) return n;
    cudaSetDevice(kv.first.first);
    if (cudaMemcpy(out_ns, w->cta_trace.ptr, sizeof(int64_t) * n, cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
    return n;
  }
  return 0;
}

}  // extern "C"

/* torchain_b200_profile.h -- instrumentation interface of libtorchain_b200.so.
 *
 * NOT part of the drop-in boundary (include/torchain_b200.h): nothing here has a counterpart in
 * the reference's src/my_lib.h.  bench.py and tools/ use it to time the denominator launch pair
 * inside a call and to read the per-frame timeline the kernels can stamp.  The micro-benchmarks
 * that measure the roofline denominators live in a separate bench-only library
 * (tools/microbench/libtorchain_microbench.so).
 */
#ifndef TORCHAIN_B200_PROFILE_H_
#define TORCHAIN_B200_PROFILE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Per-kernel timing for the roofline line of bench.py: when enabled, every launch
 * of the persistent denominator kernel is bracketed by CUDA events on the
 * caller's stream; torchain_profile_read waits for them and returns the
 * accumulated device time (ms) and launch count since the last reset. */
void torchain_profile_enable(int on);
int torchain_profile_read(void* den_graph_ptr, double* den_kernel_ms, int64_t* den_kernel_launches, int reset);

/* With profiling enabled the kernel also stamps %globaltimer (ns) at every phase
 * boundary of its last launch: [start, prologue, T forward frames, backward
 * prologue, T backward frames, end].  Copies up to max_entries; returns the count. */
int torchain_profile_timeline(void* den_graph_ptr, int64_t* out_ns, int max_entries);

/* torchain_profile_enable(2) additionally records, for every CTA and frame of the last
 * forward/backward launch pair, %globaltimer when main warp 0 finished its arcs, when aux
 * warp 0 finished, when the CTA passed its block reduction, and when it left the grid barrier.
 * Layout [2][frames][grid][4] int64 ns; returns the entry count (call with out_ns = NULL to size). */
int64_t torchain_profile_cta_trace(void* den_graph_ptr, int64_t* out_ns, int64_t max_entries, int* frames, int* grid);

#ifdef __cplusplus
}
#endif
#endif /* TORCHAIN_B200_PROFILE_H_ */

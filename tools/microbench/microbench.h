// microbench.h -- interface of libtorchain_microbench.so (bench-only; see microbench.cu).
#pragma once
#include <stdint.h>

namespace tb {
double microbench_l2_gather(int device, int64_t table_bytes, int row_bytes, int iters);   // GB/s
double microbench_l2_red(int device, int64_t table_bytes, int row_bytes, int iters);      // GB/s of operand bytes
double microbench_grid_barrier(int device, int iters);                                    // us per barrier
double microbench_pattern(int device, int warps, int depth, int iters);                   // GB/s
double microbench_pattern_async(int device, int warps, int depth, int mode, int iters);   // GB/s
double microbench_dsmem_pattern(int device, int cluster_size, int warps, int depth, int mode, int iters,
                                int* ctas_out);                                           // GB/s per 148 SMs
double microbench_sliced(int device, int S, int depth, int mode, int frames);             // us per frame
}  // namespace tb

extern "C" {
double tb_microbench_l2_gather(int device, int64_t table_bytes, int row_bytes, int iters);
double tb_microbench_l2_red(int device, int64_t table_bytes, int row_bytes, int iters);
double tb_microbench_grid_barrier(int device, int iters);
double tb_microbench_pattern(int device, int warps, int depth, int iters);
double tb_microbench_pattern_async(int device, int warps, int depth, int mode, int iters);
double tb_microbench_dsmem_pattern(int device, int cluster_size, int warps, int depth, int mode, int iters,
                                   int* ctas_out);
double tb_microbench_sliced(int device, int seqs_per_cta, int depth, int mode, int frames);
long long tb_microbench_launch_count(void);
}

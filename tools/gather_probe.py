"""How should the recurrences fetch their 512-byte rows?  GB/s of the forward pass's gather pattern (a 5 MB and a
1.5 MB table, both L2-resident, one CTA per SM) for
  * register loads (ld.global.cg.v4 per lane; what den_kernel.cu does today),
  * an LDGSTS ring in shared memory (cp.async.cg, 16 B per lane),
  * a TMA bulk-copy ring in shared memory (cp.async.bulk 512 B per row, mbarrier per slot),
  * the exp(y) table in the distributed shared memory of a cluster,
over warps per SM and rows in flight per warp."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "microbench"))
import mb as _mb  # noqa: E402

mb = _mb.load()
IT = 3000
print("L2 random-row gather, 32 MB table, 512 B rows: %.0f GB/s" % mb.tb_microbench_l2_gather(0, 32 << 20, 512, 400))
print("red.global.add.v4.f32, 32 MB table, 512 B rows: %.0f GB/s" % mb.tb_microbench_l2_red(0, 32 << 20, 512, 400))
print("\nregister loads (ld.global.cg.v4): warps x depth (rows in flight per warp)")
for w in (14, 16, 20, 24, 28, 32):
    print("%2d" % w, ["%6.0f" % mb.tb_microbench_pattern(0, w, d, IT) for d in (2, 4, 8, 12)])
for mode, name in ((0, "LDGSTS ring (cp.async.cg 16 B/lane)"), (1, "TMA bulk ring (cp.async.bulk 512 B rows)")):
    print("\n%s: warps x depth (arcs = row pairs in flight per warp)" % name)
    for w in (4, 8, 12, 14, 16, 20, 24, 32):
        row = []
        for d in (2, 4, 6, 8, 12, 16):
            v = mb.tb_microbench_pattern_async(0, w, d, mode, IT) if w * d * 1032 <= 227 * 1024 else float("nan")
            row.append("%6.0f" % v)
        print("%2d" % w, row)
print("\nexp(y) table in distributed shared memory: GB/s scaled to 148 SMs; cluster x warps, depth 4 / 8")
for cs in (8, 16):
    for w in (14, 16, 28, 32):
        for mode, name in ((0, "L2 + DSMEM"), (1, "DSMEM only"), (2, "L2 only")):
            n = ctypes.c_int(0)
            v4 = mb.tb_microbench_dsmem_pattern(0, cs, w, 4, mode, IT, ctypes.byref(n))
            v8 = mb.tb_microbench_dsmem_pattern(0, cs, w, 8, mode, IT, ctypes.byref(n))
            print("cluster %2d  warps %2d  %-11s  %6.0f / %6.0f  (%d CTAs ran)" % (cs, w, name, v4, v8, n.value))

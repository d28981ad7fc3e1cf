"""A/B probe for library variants (tools/build_variants.sh): for every .so given on the command line (loaded in a fresh
process via TORCHAIN_B200_LIB) times BASELINE.json config 2 (or --config/--batch) and prints ms per step, the mean forward
/ backward frame time from the in-kernel timeline, and the parity of the step against the fp64 oracle."""
import argparse
import ctypes
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one(args):
    import torch
    import bench
    from torchain_b200 import _lib
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    c = bench.make_workload(args.config, args.batch, args.frames, 0)
    R = bench.Runner(torch, lib, c, dev, 0, 1, 0)
    ms = R.time_steps(args.steps, 4)
    T = c["T"]
    lib.torchain_profile_enable(1)
    R.step()
    torch.cuda.synchronize()
    n = 2 * T + 6
    tl = np.zeros(n, dtype=np.int64)
    got = lib.torchain_profile_timeline(R.g.ptr, tl.ctypes.data_as(ctypes.c_void_p), n)
    lib.torchain_profile_enable(0)
    f, b = tl[:T + 3], tl[T + 3:]
    fwd, bwd = (np.diff(f) / 1e3)[1:1 + T], (np.diff(b) / 1e3)[1:1 + T]
    out = {"lib": os.path.basename(_lib.LIB_PATH), "tied": R.g.tied, "runs": R.g.n_run, "config": args.config, "B": c["B"], "T": T, "ms_per_step": round(ms, 4),
           "fwd_frame_us": round(float(fwd.mean()), 2), "bwd_frame_us": round(float(bwd.mean()), 2),
           "fwd_launch_us": round(float(f[-1] - f[0]) / 1e3, 1), "bwd_launch_us": round(float(b[-1] - b[0]) / 1e3, 1)}
    if not args.no_parity:
        R.step(host_results=True)
        torch.cuda.synchronize()
        p = bench.parity_check(c, c["sup"], R.y.cpu().numpy(), R.res_host.copy(), R.deriv.cpu().numpy(), bench.host_cores())
        out.update(parity_ok=p["ok"], objf_rel=float("%.2e" % p["objf_rel_err"]), deriv_rel=float("%.2e" % p["deriv_rel_err"]))
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("libs", nargs="*")
    ap.add_argument("--config", default="chime5")
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--frames", type=int, default=0)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--child", action="store_true")
    args = ap.parse_args()
    if args.child:
        return one(args)
    for lib in args.libs or [""]:
        env = dict(os.environ)
        if lib:
            env["TORCHAIN_B200_LIB"] = os.path.abspath(lib)
        cmd = [sys.executable, os.path.abspath(__file__), "--child", "--config", args.config, "--batch", str(args.batch),
               "--frames", str(args.frames), "--steps", str(args.steps)] + (["--no-parity"] if args.no_parity else [])
        r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
        sys.stdout.write(r.stdout if r.returncode == 0 else "%s FAILED rc=%d: %s\n" % (lib, r.returncode, r.stderr[-600:]))
        sys.stdout.flush()


if __name__ == "__main__":
    main()

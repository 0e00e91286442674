#!/usr/bin/env python
"""Pipeline trace of k_phi_mma from a -DDIBS_PHI_TRACE build (tools/build_variant.sh trace -DDIBS_PHI_TRACE).

    python tools/phi_trace.py _exp/libdibs_b200_trace.so [workload]
Runs a few steps of the bench workload through the variant library (loaded INSTEAD of the product library, debugging
only) and prints, per pipeline role, the median clock deltas of the K steps over all CTAs of the last phi launch.
"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dibs_b200 import _native as nat  # noqa: E402

nat.LIB_PATH = os.path.join(ROOT, sys.argv[1])
import torch  # noqa: E402
import bench  # noqa: E402

MHZ = 1965.0


def main(wl="t_lin"):
    dev = torch.device("cuda:0")
    x = torch.from_numpy(bench.workload_data(wl)).to(dev)
    model = bench.build_model(wl, x, dev)
    m = bench.WORKLOADS[wl][3]
    model.sample(key=np.array([0, 1], np.uint32), n_particles=m, steps=6)
    torch.cuda.synchronize()
    buf = np.zeros(1024 * 64, np.uint64)
    fn = nat.lib().dibs_debug_phi_trace
    fn.restype, fn.argtypes = ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]
    nat.check(fn(buf.ctypes.data, buf.size))
    tr = buf.reshape(1024, 64).astype(np.int64)
    live = tr[:, 0] > 0
    tr = tr[live]
    us = lambda cyc: cyc / MHZ
    t0g = tr[:, 0].min()
    print(f"{wl}: {len(tr)} CTAs traced; kernel span {(tr[:, 2].max() - t0g) / 1e3:.1f} us (globaltimer), "
          f"finishing CTAs end at {(tr[:, 4].max() - t0g) / 1e3:.1f} us")
    start, end = tr[:, 50], tr[:, 51]
    print(f"CTA duration (to the arrival counter): median {us(np.median(end - start)):.2f} us, "
          f"min {us((end - start).min()):.2f}, max {us((end - start).max()):.2f}")
    fin = tr[tr[:, 3] == 1]
    if len(fin):
        print(f"finish tile (last j slice, {len(fin)} CTAs): median {us(np.median(fin[:, 52] - fin[:, 51])):.2f} us")
    starts_us = (tr[:, 0] - t0g) / 1e3
    print("CTA start times (us) percentiles 0/25/50/75/100:", np.percentile(starts_us, [0, 25, 50, 75, 100]).round(1))
    print("first TMA issue after CTA start: median %.2f us" % us(np.median(tr[:, 8] - start)))
    n_it = 8
    rows = []
    for it in range(n_it):
        issue, full, split, mwait, mdone = tr[:, 8 + it], tr[:, 16 + it], tr[:, 24 + it], tr[:, 32 + it], tr[:, 40 + it]
        ok = (issue > 0) & (full > 0)
        rows.append((it, us(np.median((issue - start)[ok])), us(np.median((full - issue)[ok])), us(np.median((split - full)[ok])),
                     us(np.median((mwait - split)[ok])), us(np.median((mdone - mwait)[ok]))))
    print("it  issue@  tma_land  split  ->mma_wakeup  mma_issue+commit   (us, medians)")
    for r in rows:
        print("%2d  %6.2f  %7.2f  %6.2f  %8.2f  %8.2f" % r)
    print("accumulators complete @ %.2f us, epilogue %.2f us" % (us(np.median(tr[:, 48] - start)), us(np.median(tr[:, 49] - tr[:, 48]))))
    k_step = np.median((tr[:, 16 + 7] - tr[:, 16 + 1]) / 6.0)
    print("steady-state K step (full[7]-full[1])/6: %.2f us" % us(k_step))


if __name__ == "__main__":
    main(*sys.argv[2:])

#!/usr/bin/env python
"""Where does a sample() call spend its time?  Three consecutive calls per workload through the public API, with the
plan look-up (`_plan`) and the native step call (`dibs_svgd_steps`) timed separately and the plan-cache size printed.

    python tools/e2e_probe.py [workloads...]     (diagnostics for the e2e / value gap of the BGe lines, RESULTS.md)
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from dibs_b200.inference import dibs as D  # noqa: E402


STEPS = [int(v) for v in os.environ.get("PROBE_STEPS", "1,1,1").split(",")]


def main(*wls):
    dev = torch.device("cuda:0")
    for wl in wls or ("t_lin", "t_bge", "c3"):
        x_host = torch.from_numpy(bench.workload_data(wl)).pin_memory()
        m = bench.WORKLOADS[wl][3]
        for call, n_steps in enumerate(STEPS):
            acc = {"plan": 0.0, "steps": 0.0, "n_plan": 0}
            mdl = bench.build_model(wl, x_host, dev)
            plan0, call0 = mdl._plan, mdl._call

            def plan_t(*a, **k):
                torch.cuda.synchronize(); t0 = time.perf_counter()
                r = plan0(*a, **k)
                torch.cuda.synchronize(); acc["plan"] += time.perf_counter() - t0; acc["n_plan"] += 1
                return r

            def call_t(fn, *a):
                if fn != "dibs_svgd_steps":
                    return call0(fn, *a)
                torch.cuda.synchronize(); t0 = time.perf_counter()
                r = call0(fn, *a)
                torch.cuda.synchronize(); acc["steps"] += time.perf_counter() - t0
                return r

            mdl._plan, mdl._call = plan_t, call_t
            torch.cuda.synchronize(); t0 = time.perf_counter()
            mdl.sample(key=np.array([0, 1], np.uint32), n_particles=m, steps=n_steps)
            torch.cuda.synchronize(); tot = time.perf_counter() - t0
            print(f"{wl:6s} call {call}: total {tot * 1e3:9.2f} ms | _plan x{acc['n_plan']} {acc['plan'] * 1e3:9.2f} ms | "
                  f"dibs_svgd_steps({n_steps:3d} steps) {acc['steps'] * 1e3:9.2f} ms | rest {(tot - acc['plan'] - acc['steps']) * 1e3:8.2f} ms | "
                  f"plans cached {len(D._PLAN_CACHE)}", flush=True)


if __name__ == "__main__":
    main(*sys.argv[1:])

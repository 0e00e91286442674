#!/usr/bin/env python
"""Which kernel carries the dependence of a step's cost on the step index t?  (profiles/r02/RESULTS.md, BGe finding)

    python tools/t_dependence.py [workloads...]
Per workload: fresh particles, per-kernel device times (`dibs_svgd_steps_timed`, per_kernel = 1, mean of 3 steps) of the
steps starting at t = 0, at t = 100 on the same fresh particles (bench.py's `value` operating point) and at t = 100 after
100 real steps; with the expected number of parents per node, d * mean(P_ij), of the particles at each point.
"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from dibs_b200 import _native as nat  # noqa: E402
from dibs_b200.inference.dibs import PRNGKey, split, keys_to_device  # noqa: E402


def main(*wls):
    dev = torch.device("cuda:0")
    lib = nat.lib()
    for wl in wls or ("t_lin", "t_bge", "c3"):
        _, lik, d, m, s, a, h = bench.WORKLOADS[wl]
        joint = lik != "bge"
        x = torch.from_numpy(bench.workload_data(wl)).to(dev)
        model = bench.build_model(wl, x, dev)
        plan = model._plan(m, d, sharded=True)
        sptr = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

        def fresh():
            key = PRNGKey(0)
            key, subk = split(key, 2)
            init = model._sample_initial_random_particles(key=subk, n_particles=m, n_dim=d, plan=plan)
            z, th = init if joint else (init, None)
            z = z.contiguous(); th = th.contiguous() if joint else None
            return dict(z=z, th=th, vz=torch.zeros_like(z), vt=torch.zeros_like(th) if joint else None,
                        sf=torch.zeros(m, dtype=torch.float32, device=dev), key=keys_to_device(key, dev))

        def run(st, t0, n, per_kernel):
            step_ms = np.zeros(n, np.float32); ph = np.zeros(len(nat.PHASES), np.float32)
            nat.check(lib.dibs_svgd_steps_timed(plan.handle, t0, n, nat.ptr(st["z"]), nat.ptr(st["th"]), nat.ptr(st["vz"]), nat.ptr(st["vt"]),
                                                nat.ptr(st["key"]), nat.ptr(st["sf"]), sptr, int(per_kernel), None, 0,
                                                nat.ptr(step_ms), nat.ptr(ph) if per_kernel else None))
            return step_ms, ph

        def parents(st, t):
            p = torch.empty((m, d, d), dtype=torch.float32, device=dev)
            nat.check(lib.dibs_edge_probs(plan.handle, nat.ptr(st["z"]), m, int(t), nat.ptr(p), sptr))
            torch.cuda.synchronize()
            return float(p.mean()) * d

        run(fresh(), 100, 2, 0)                      # warm (graphs, allocator)
        for label, t0, pre in (("t=0, fresh particles", 0, 0), ("t=100, fresh particles (bench `value`)", 100, 0),
                               ("t=100 after 100 real steps", 100, 100)):
            st = fresh()
            if pre:
                run(st, 0, pre, 0)
            par = parents(st, t0)
            _, ph = run(st, t0, 3, 1)
            us = {p: round(float(ph[i]) / 3 * 1e3, 1) for i, p in enumerate(nat.PHASES) if ph[i] > 0}
            print(f"{wl:6s} {label:40s} E[parents/node] {par:5.2f} of {d - 1} | sum {sum(us.values()):9.1f} us | {us}", flush=True)


if __name__ == "__main__":
    main(*sys.argv[1:])

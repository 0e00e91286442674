#!/usr/bin/env python
"""Multi-GPU consistency check: run the same sample() on 1 rank and on N ranks (torchrun) and compare particles.

    python tools/mgpu_check.py                      # single process  -> gpurun_out/mgpu_w1_<case>.npz
    torchrun --nproc-per-node 2 tools/mgpu_check.py # sharded         -> gpurun_out/mgpu_w2_<case>.npz
    python tools/mgpu_check.py --compare 2          # asserts bit-identical particles for every case
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.environ.get("MGPU_OUT") or os.path.join(ROOT, "gpurun_out")
CASES = {"lin": dict(d=20, m=64, s=32, steps=7), "bge": dict(d=12, m=32, s=16, steps=5), "nn": dict(d=10, m=16, s=8, steps=4),
         # >= 512 particles: the tensor-core phi kernel (its choice depends on the GLOBAL particle count only)
         "linL": dict(d=12, m=512, s=8, steps=3)}


def run():
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get("MGPU_WATCHDOG", "90")), exit=True)
    import torch
    import torch.distributed as dist
    from dibs_b200.inference import JointDiBS, MarginalDiBS, PRNGKey
    from dibs_b200.models import BGe, LinearGaussian, DenseNonlinearGaussian, ErdosReniDAGDistribution
    from dibs_b200.synthetic import make_linear_gaussian_data

    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # MGPU_SAME_GPU=1: every rank on cuda:0 (time-sliced contexts) with a gloo process group -- NCCL refuses two ranks on
    # one device, the peer-memory exchange (CUDA IPC between processes) does not care; lets a ONE-GPU box run the
    # multi-rank path (kernels_peer.cuh) end to end
    same = os.environ.get("MGPU_SAME_GPU") == "1"
    if same:
        local = 0
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if same:
            dist.init_process_group("gloo")
        else:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    os.makedirs(OUT, exist_ok=True)
    for name in os.environ.get("MGPU_CASES", "bge,lin,nn,linL").split(","):
        c = CASES[name]
        d = c["d"]
        x = make_linear_gaussian_data(seed=1, n_vars=d, n_observations=50)["x"]
        gm = ErdosReniDAGDistribution(n_vars=d)
        if name in ("lin", "linL"):
            model = JointDiBS(x=x, graph_model=gm, likelihood_model=LinearGaussian(n_vars=d), n_grad_mc_samples=c["s"])
        elif name == "nn":
            model = JointDiBS(x=x, graph_model=gm, likelihood_model=DenseNonlinearGaussian(n_vars=d, hidden_layers=(5,)),
                              n_grad_mc_samples=c["s"])
        else:
            model = MarginalDiBS(x=x, graph_model=gm, likelihood_model=BGe(n_vars=d), n_grad_mc_samples=c["s"])
        model.sample(key=PRNGKey(3), n_particles=c["m"], steps=c["steps"], callback_every=None)
        st = model._last_state
        if world > 1 and os.environ.get("MGPU_EXPECT"):
            from dibs_b200.inference.dibs import _PLAN_CACHE
            kinds = {p.exchange for p in _PLAN_CACHE.values() if p.cfg.world_size > 1}
            assert kinds == {os.environ["MGPU_EXPECT"]}, kinds
        if int(os.environ.get("RANK", "0")) == 0:
            np.savez(os.path.join(OUT, f"mgpu_w{world}_{name}.npz"), z=st["z"].cpu().numpy(),
                     theta=st["theta"].cpu().numpy() if st["theta"] is not None else np.zeros(0))
            print(f"[mgpu_check] world={world} case={name} done, |z|={float(st['z'].abs().mean()):.6f}", flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def compare(w):
    ok = True
    for name in os.environ.get("MGPU_CASES", "bge,lin,nn,linL").split(","):
        a = np.load(os.path.join(OUT, f"mgpu_w1_{name}.npz"))
        b = np.load(os.path.join(OUT, f"mgpu_w{w}_{name}.npz"))
        for k in ("z", "theta"):
            same = np.array_equal(a[k], b[k])
            diff = float(np.abs(a[k] - b[k]).max()) if a[k].size else 0.0
            print(f"[mgpu_check] {name}.{k}: world 1 vs {w}: bit-identical={same} max|diff|={diff:.3e}")
            ok &= same or diff < 1e-6
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    if "--compare" in sys.argv:
        compare(int(sys.argv[sys.argv.index("--compare") + 1]))
    else:
        run()

#!/usr/bin/env python
"""bench.py -- SVGD steps/sec of the DiBS particle-update hot path on B200 (driver contract: see DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload t_lin] [--impl native|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one full ``_svgd_step`` (reference: dibs/inference/svgd.py:226-267 / 673-721) over ALL particles of the
workload; the metric is SVGD steps/sec (BASELINE.json).  Default workload: the north-star target shape, JointDiBS
LinearGaussian n_vars=20 n_particles=1024 n_mc=128 (`t_lin`); the other single-GPU configs of BASELINE.json ride in the
same JSON line as short runs (`also`).  N > 1 shards the particles of the SAME workload over the ranks (strong scaling).
Rank 0 prints ONE JSON line.

native arm
  value      K steps, state resident in HBM, every step bracketed by CUDA events on the launching stream with a
             256 MiB memset (> 126 MB L2) before each step outside the bracket (cold L2); max over ranks.
  e2e        the same metric through the public API with HOST buffers: ``JointDiBS(x=<pinned host array>, ...)
             .sample(key=, n_particles=, steps=K)`` and the returned particles copied back to host memory,
             wall clock around the whole call (upload of x, particle init, K steps, final Z -> G, download; the native
             plan is cached per process, keyed on configuration + data digest, and was created by a warm-up call).
  roofline   dominant kernel of the step: per-kernel device time from events after every launch (eager replay of
             the same kernels behind a delay kernel, so launch gaps are not charged), algorithmic flops/bytes from
             DESIGN.md section 4; `passes` = the kernel-matrix and edge-probability passes the north star names,
             against SURVEY 8(d)'s algorithmic bytes and flops; `traffic` = ncu DRAM bytes (profiles/traffic.json).
  cpu_baseline  the NumPy oracle (restated reference) on this box's host cores on a bounded sample of particles.
reference arm (--impl reference): the oracle port of the reference's CPU path timed on the host cores (JAX is not
  installable in this image and the reference has no native sources, so there is no oracle/_ref build).
DIBS_BENCH_TIMELINE=1 adds `timeline_end_us`: the step GRAPH replayed with an event node behind every kernel.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

# ---------------------------------------------------------------------------------------------------------------
# workloads (BASELINE.json configs)
# ---------------------------------------------------------------------------------------------------------------
WORKLOADS = {
    # name: (description, likelihood, d, M, S, A, hidden)
    "c1": ("MarginalDiBS BGe n_vars=5 n_particles=4 (reference CPU plumbing case)", "bge", 5, 4, 128, 32, 0),
    "c2": ("JointDiBS LinearGaussian n_vars=20 n_particles=256 n_mc=64", "lingauss", 20, 256, 64, 32, 0),
    "c3": ("MarginalDiBS BGe n_vars=50 n_particles=1024 n_mc=128", "bge", 50, 1024, 128, 32, 0),
    "c4": ("JointDiBS DenseNonlinearGaussian(5,) n_vars=20 n_particles=1024 n_mc=128", "densenn", 20, 1024, 128, 32, 5),
    "c5": ("JointDiBS LinearGaussian n_vars=100 n_particles=4096 n_mc=128", "lingauss", 100, 4096, 128, 32, 0),
    "t_bge": ("MarginalDiBS BGe n_vars=20 n_particles=1024 n_mc=128 (north-star target shape)", "bge", 20, 1024, 128, 32, 0),
    "t_lin": ("JointDiBS LinearGaussian n_vars=20 n_particles=1024 n_mc=128 (north-star target shape)", "lingauss", 20, 1024, 128, 32, 0),
    # diagnostics: per-rank gradient work of t_lin on 8 GPUs (128 particles per rank) reproduced on 2
    "t_lin_q": ("JointDiBS LinearGaussian n_vars=20 n_particles=256 n_mc=128 (quarter of the target shape)", "lingauss", 20, 256, 128, 32, 0),
}
N_OBS = 100
# `value` is timed from step T_MID on freshly initialised particles, so that alpha(t), beta(t) are non-degenerate.
# The work of a step is NOT t-independent for every model: LinearGaussian is flat, but a BGe step at t = 0 costs ~6x
# (n_vars = 20) to ~20x (n_vars = 50) a step at t >= 100 (profiles/r02/RESULTS.md) -- which is why the line also carries
# `value_from_t0` (the same device timing over steps 0..K-1 of re-initialised particles, what sample() actually runs) and
# why `e2e`, which starts at t = 0 by construction, is the figure comparable to a real sample() call.  The CPU arm
# (`cpu_baseline`, `--impl reference`) is timed at T_MID too: `value` and the CPU figure cover the same step range.
T_MID = 100
L2_FLUSH_BYTES = 256 << 20
FP32_SIMT_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12     # theoretical; MEASURED_PEAKS.json has no fp32 SIMT figure


def workload_data(name):
    from dibs_b200.synthetic import make_linear_gaussian_data, make_nonlinear_gaussian_data
    _, lik, d, _, _, _, hidden = WORKLOADS[name]
    if lik == "densenn":
        return make_nonlinear_gaussian_data(seed=0, n_vars=d, n_observations=N_OBS, hidden=hidden)["x"]
    return make_linear_gaussian_data(seed=0, n_vars=d, n_observations=N_OBS)["x"]


def er_edges(d):
    # ER prior with 2 edges/node has p >= 1 for d <= 5: use 1 edge/node there
    return 2 if d > 5 else 1


def simt_peaks():
    """fp32 / fp64 SIMT FMA peaks in TFLOP/s: measured on this pool's B200 with tools/ubench.cu
    (profiles/r01/ubench_pipes.json: FFMA2 74.2, DFMA 37.1), else the theoretical 148 SM x 128 lanes x 2 x 1.965 GHz."""
    p = os.path.join(ROOT, "profiles", "r01", "ubench_pipes.json")
    try:
        j = json.load(open(p))
        return max(float(j["ffma_rrr_tflops"]), float(j.get("ffma2_tflops", 0.0))), float(j["dfma_tflops"]), \
            "measured (tools/ubench.cu -> profiles/r01/ubench_pipes.json)"
    except Exception:
        return FP32_SIMT_TFLOPS, FP32_SIMT_TFLOPS / 2, "theoretical SIMT FMA peak (148 SM x 128 lanes x 2 x 1.965 GHz)"


def step_bound(work, kernels, hbm_gbs, fp32_tf, fp64_tf):
    """t_bound = sum over the kernels of the step of max(bytes / BW_HBM, flops / peak_pipe); returns the
    bound in microseconds and, per kernel, which side binds.  Only kernels that actually ran (``kernels``) count."""
    total, per = 0.0, {}
    for name in kernels:
        w = work.get(name)
        if not w or w.get("bound") in ("nvlink", "latency"):
            continue
        peak = fp64_tf if w["bound"] == "fp64" else fp32_tf
        t_mem = w["bytes"] / (hbm_gbs * 1e9) * 1e6
        t_alu = w["flops"] / (peak * 1e12) * 1e6
        per[name] = {"us": round(max(t_mem, t_alu), 3), "side": "hbm" if t_mem >= t_alu else "alu"}
        total += max(t_mem, t_alu)
    return total, per


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            j = json.load(open(p))
            return float(j["hbm_gbs"]), float(j.get("bf16_tflops_sustained", j["bf16_tflops"])), "measured"
        except Exception:
            pass
    return 6650.0, 1590.0, "fallback"


# ---------------------------------------------------------------------------------------------------------------
# algorithmic work per launch of each kernel of the step (fp32)
# ---------------------------------------------------------------------------------------------------------------
def kernel_work(name, m_loc, m_all):
    """-> {phase: dict(flops=, bytes=, bound=)} for one step on a rank that owns m_loc of m_all particles."""
    _, lik, d, _, s, a, h = WORKLOADS[name]
    k, n = d, N_OBS
    dz = 2 * d * k
    dth = {"bge": 0, "lingauss": d * d, "densenn": d * (d * h + 2 * h + 1)}[lik]
    dd = dz + dth
    w = {}
    edge = 2 * d * d * k                                   # U V^T per particle
    if lik == "lingauss":
        fwd_bwd = 2 * (2 * n * d * d)                      # x @ (G*Theta) and x^T R, per graph
        w["mc_theta"] = dict(flops=m_loc * (s * fwd_bwd + edge), bytes=m_loc * 4 * (dz + 2 * dth), bound="fp32")
        w["mc_z"] = dict(flops=m_loc * (s * fwd_bwd + edge), bytes=m_loc * 4 * (dz + dth + d * d), bound="fp32")
        if d <= 32:
            # what the QR-form kernel EXECUTES (kernels_mc_lin_qr.cuh): two triangular d x d mat-vecs per (graph,
            # node) instead of the two N x d products above -- d (d + 1) FMAs; the reference's algorithm is charged
            # in `flops`, this is the figure pipe utilisation is judged by
            qr = 2 * d * d * (d + 1)
            w["mc_theta"]["flops_executed"] = m_loc * (s * qr + edge)
            w["mc_z"]["flops_executed"] = m_loc * (s * qr + edge)
    elif lik == "densenn":
        fwd_bwd = 2 * d * (2 * n * d * h + 2 * n * h)
        w["mc_theta"] = dict(flops=m_loc * (s * fwd_bwd + edge), bytes=m_loc * 4 * (dz + 2 * dth), bound="fp32")
        w["mc_z"] = dict(flops=m_loc * (s * fwd_bwd + edge), bytes=m_loc * 4 * (dz + dth + d * d), bound="fp32")
    else:
        # one Cholesky of the (l+1)x(l+1) parent block per (graph, node); l ~ d/2 at P = 0.5 -> (d/2)^3/3 flops, fp64.
        # NOTE: this charges the work of the HIGH-entropy regime (t = 0).  At `value`'s operating point (t >= 100) the
        # kernel does far less (common parents eliminated once per particle and node, kernels_mc_bge.cuh:13-22), so the
        # frac reported for mc_z there is not pipe utilisation; at t = 0 it is 8-17x lower (profiles/r02/RESULTS.md)
        chol = d * ((d / 2.0) ** 3) / 3.0
        w["mc_z"] = dict(flops=m_loc * (s * chol + edge), bytes=m_loc * 4 * (dz + d * d), bound="fp64")
    nmat = (int(np.floor(np.log2(max(d - 1, 1)))) + bin(max(d - 1, 1)).count("1") - 1)
    # acyclicity pass; in the step loop a particle's last gradient CTA also runs the assemble step (chain rule through
    # S = U V^T, priors, next keys) -- its flops / bytes are charged here
    w["acyclic"] = dict(flops=m_loc * (a * nmat * 2 * d ** 3 + edge + 4 * d * d * k),
                        bytes=m_loc * 4 * (dz + d * d + 2 * dz + 2 * d * d + 2 * dth), bound="fp32")
    w["allgather"] = dict(flops=0, bytes=(m_all - m_loc) * 4 * 2 * dd, bound="nvlink")
    w["pair_dist"] = dict(flops=3 * m_loc * m_all * dd, bytes=4 * (m_all * dd + m_loc * m_all), bound="fp32")
    w["pair_kernel"] = dict(flops=4 * m_loc * m_all, bytes=4 * m_loc * m_all * (3 if dth else 2), bound="hbm")
    # phi + optimizer step (one kernel: the tile's last j-slice CTA finishes the sum and updates x, v)
    w["phi_update"] = dict(flops=2 * 2 * m_loc * m_all * dd + 8 * m_loc * dd,
                           bytes=4 * (2 * m_all * dd + 2 * m_loc * m_all + 6 * m_loc * dd), bound="fp32")
    # next step's raw scores U V^T from the updated latent rows = the edge-probability pass
    w["scores"] = dict(flops=m_loc * edge, bytes=m_loc * 4 * (dz + d * d), bound="hbm")
    return w


# ---------------------------------------------------------------------------------------------------------------
# clocks during the timed region (nvidia-smi recipe of B200_PROFILING.md, NVML bindings as the fallback)
# ---------------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index, period=0.05):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.h = None

    def run(self):
        while not self._stop_evt.is_set():
            if self.h is not None:
                try:
                    self.samples.append(int(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                    bits = int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)) if hasattr(
                        self.nv, "nvmlDeviceGetCurrentClocksEventReasons") else int(
                        self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                    for b, nm in self.REASONS.items():
                        if bits & b:
                            self.reasons.add(nm)
                except Exception:
                    pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        if not self.samples:
            return self._smi_once()
        return {"sm_mhz": int(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples), "source": "nvml"}

    def _smi_once(self):
        import subprocess
        try:
            out = subprocess.run(["nvidia-smi", f"--id={self.index}", "--query-gpu=clocks.sm,clocks.max.sm",
                                  "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10).stdout
            a, b = [int(v) for v in out.strip().split(",")]
            return {"sm_mhz": a, "sm_max_mhz": b, "reasons": [], "samples": 1, "source": "nvidia-smi (after the run)"}
        except Exception as e:  # noqa: BLE001
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": f"unavailable: {e}"}


# ---------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle (restated reference) on the host cores
# ---------------------------------------------------------------------------------------------------------------
_CPU = {}


def _cpu_worker(args):
    cfg_kw, particles, t = args
    from oracle import dibs_oracle as orc
    st, x, mask, cfg = _CPU["st"], _CPU["x"], _CPU["mask"], _CPU["cfg"]
    t0 = time.perf_counter()
    dz, dth, _, _, _ = orc.particle_grads(cfg, st, t, x, mask, np.float32, particles=particles)
    return particles, dz[particles], None if dth is None else dth[particles], time.perf_counter() - t0


def oracle_config(name):
    from oracle import dibs_oracle as orc
    _, lik, d, _, s, a, h = WORKLOADS[name]
    joint = lik != "bge"
    return orc.Config(lik=orc.Likelihood(kind=lik, n_vars=d, hidden=h or 5), prior=orc.GraphPrior("er", d, er_edges(d)),
                      joint=joint, alpha_linear=0.05 if joint else 1.0, grad_estimator_z="reparam" if joint else "score",
                      n_grad_mc_samples=s, n_acyclicity_mc_samples=a)


def cpu_steps(name, n_steps, warmup, budget_s, procs=None):
    """Time the oracle's ``svgd_step`` on a bounded sample: the gradient phase (O(M), >99% of the CPU time) runs on a
    subset of ``m_s`` particles spread over all host cores and is scaled by M/m_s; the pairwise phase (kernel matrix,
    phi, optimizer) runs on all M particles.  Returns (steps_per_sec, seconds_per_full_step, sample description, cores)."""
    import multiprocessing as mp
    from oracle import dibs_oracle as orc
    _, lik, d, m, s, a, h = WORKLOADS[name]
    cores = procs or os.cpu_count() or 1
    x = workload_data(name)
    mask = np.zeros(x.shape, np.int32)
    cfg = oracle_config(name)
    st = orc.init_particles(cfg, np.array([0, 0], np.uint32), m, None, np.float32)
    _CPU.update(st=st, x=x, mask=mask, cfg=cfg)
    # calibrate: one particle on this process
    t0 = time.perf_counter()
    orc.particle_grads(cfg, st, T_MID, x, mask, np.float32, particles=[0])
    per_particle = time.perf_counter() - t0
    total = max(1, n_steps + warmup)
    m_s = int(budget_s / total / per_particle * cores * 0.7)
    m_s = max(cores, min(m, (m_s // cores) * cores))
    m_s = min(m, m_s)
    chunks = [list(range(i, m_s, cores)) for i in range(cores) if i < m_s]
    ctx = mp.get_context("fork")
    step_times = []
    with ctx.Pool(len(chunks)) as pool:
        for it in range(total):
            t0 = time.perf_counter()
            res = pool.map(_cpu_worker, [(None, c, T_MID + it) for c in chunks])
            t_grad = time.perf_counter() - t0
            dz = np.zeros(st.z.shape, np.float32)
            dth = None if st.theta is None else np.zeros(st.theta.shape, np.float32)
            for parts, gz, gth, _ in res:
                dz[parts] = gz
                if dth is not None:
                    dth[parts] = gth
            t1 = time.perf_counter()
            k_full, k_z, k_t = orc.kernel_matrix(cfg, st.z, st.theta if cfg.joint else None, np.float32)
            phi_z = orc.phi_update(k_full, k_z, cfg.h_latent, st.z, dz, np.float32)
            z_new, vz_new = orc.opt_update(cfg, st.z, st.v_z, phi_z, np.float32)
            if cfg.joint:
                phi_t = orc.phi_update(k_full, k_t, cfg.h_theta, st.theta, dth, np.float32)
                th_new, vt_new = orc.opt_update(cfg, st.theta, st.v_theta, phi_t, np.float32)
            t_pair = time.perf_counter() - t1
            if it >= warmup:
                step_times.append(t_grad * (m / m_s) + t_pair)
    sec = float(np.mean(step_times))
    sample = (f"{len(step_times)} steps; gradient phase on {m_s} of {m} particles over {len(chunks)} processes "
              f"(scaled x{m / m_s:.2f}), pairwise phase on all {m}")
    return 1.0 / sec, sec, sample, len(chunks)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = args.workload
    desc, lik, d, m, s, a, h = WORKLOADS[name]
    budget = float(os.environ.get("DIBS_BENCH_CPU_BUDGET_S", "150"))
    sps, sec, sample, cores = cpu_steps(name, args.steps, args.warmup, budget)
    line = {
        "impl": "reference", "metric": "svgd_steps_per_sec", "value": sps, "unit": "steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(name, args.gpus),
        "cpu_baseline": {"value": sps, "unit": "steps/s", "cores": cores, "kind": "port", "sample": sample,
                         "note": "NumPy oracle = restated reference (JAX not installable here; reference has no native sources)"},
        "e2e": {"value": sps, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def config_dict(name, n_gpus):
    desc, lik, d, m, s, a, h = WORKLOADS[name]
    return {"workload": f"{name}: {desc}", "n_vars": d, "n_particles": m, "n_grad_mc_samples": s,
            "n_acyclicity_mc_samples": a, "n_observations": N_OBS, "graph_prior": f"er(n_edges_per_node={er_edges(d)})",
            "optimizer": "rmsprop(0.005)", "t_start": T_MID, "particles_per_gpu": m // n_gpus,
            "parallelism": (f"particle-sharded x{n_gpus}; rows exchanged by peer-memory pushes over NVLink (particles under the "
                            "gradient phase, gradients on the critical path), NCCL all-gathers as fallback") if n_gpus > 1 else "single GPU",
            "l2": f"flushed before every timed step ({L2_FLUSH_BYTES >> 20} MiB memset outside the event bracket)"}


# ---------------------------------------------------------------------------------------------------------------
# native arm
# ---------------------------------------------------------------------------------------------------------------
def build_model(name, x, device):
    from dibs_b200.inference import JointDiBS, MarginalDiBS
    from dibs_b200.models import BGe, LinearGaussian, DenseNonlinearGaussian, ErdosReniDAGDistribution
    _, lik, d, m, s, a, h = WORKLOADS[name]
    gm = ErdosReniDAGDistribution(n_vars=d, n_edges_per_node=er_edges(d))
    kw = dict(x=x, graph_model=gm, n_grad_mc_samples=s, n_acyclicity_mc_samples=a, device=device)
    if lik == "bge":
        return MarginalDiBS(likelihood_model=BGe(n_vars=d), **kw)
    if lik == "lingauss":
        return JointDiBS(likelihood_model=LinearGaussian(n_vars=d), **kw)
    return JointDiBS(likelihood_model=DenseNonlinearGaussian(n_vars=d, hidden_layers=(h,)), **kw)


class Ctx:
    """Process-wide state of the native arm: device, ranks, helpers."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.device = torch.device("cuda", self.local)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.device)
        self.flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=self.device)

    def barrier(self):
        self.torch.cuda.synchronize(self.device)
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize(self.device)

    def max_over_ranks(self, v):
        if self.world == 1:
            return float(v)
        t = self.torch.tensor([float(v)], dtype=self.torch.float64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


def load_traffic(name, world):
    """ncu DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of every step kernel, captured on ONE
    GPU with tools/ncu_traffic.sh -> profiles/traffic.json {workload: {phase: bytes}}; multi-rank runs are never
    profiled, so N > 1 has no traffic figure."""
    if world != 1:
        return {}
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(name, {})
    except Exception:
        return {}


def measure_workload(ctx, name, K, W, n_prof, want_hot=True, want_t0=None):
    """Device-timed steps of one workload with the state resident in HBM.  Returns a dict with value (steps/s, cold L2,
    max over ranks), ms_per_step, launches, per-kernel table and the roofline entry of the dominant kernel."""
    import ctypes
    torch = ctx.torch
    from dibs_b200 import _native as nat
    from dibs_b200.inference.dibs import PRNGKey, split, keys_to_device
    desc, lik, d, m, s, a, h = WORKLOADS[name]
    if m % ctx.world:
        raise SystemExit(f"n_particles={m} not divisible by {ctx.world} ranks")
    joint = lik != "bge"
    lib = nat.lib()
    device = ctx.device
    x_dev = torch.from_numpy(workload_data(name)).to(device)
    model = build_model(name, x_dev, device)
    plan = model._plan(m, d, sharded=True)
    key = PRNGKey(0)
    key, subk = split(key, 2)
    init = model._sample_initial_random_particles(key=subk, n_particles=m, n_dim=d, plan=plan)
    z_all, th_all = init if joint else (init, None)
    lo, hi = plan.row0, plan.row0 + plan.n_local
    z = z_all[lo:hi].contiguous()
    theta = th_all[lo:hi].contiguous() if joint else None
    v_z = torch.zeros_like(z)
    v_th = torch.zeros_like(theta) if joint else None
    sf = torch.zeros(plan.n_local, dtype=torch.float32, device=device)
    key_dev = keys_to_device(key, device)
    stream = torch.cuda.current_stream(device)
    sptr = ctypes.c_void_p(stream.cuda_stream)
    flush = ctx.flush

    def steps(t0, n):
        nat.check(lib.dibs_svgd_steps(plan.handle, t0, n, nat.ptr(z), nat.ptr(theta), nat.ptr(v_z), nat.ptr(v_th),
                                      nat.ptr(key_dev), nat.ptr(sf), sptr))

    def steps_timed(t0, n, per_kernel, use_flush=True):
        step_ms = np.zeros(n, np.float32)
        phase_ms = np.zeros(len(nat.PHASES), np.float32)
        nat.check(lib.dibs_svgd_steps_timed(plan.handle, t0, n, nat.ptr(z), nat.ptr(theta), nat.ptr(v_z), nat.ptr(v_th),
                                            nat.ptr(key_dev), nat.ptr(sf), sptr, int(per_kernel),
                                            nat.ptr(flush) if use_flush else None, L2_FLUSH_BYTES,
                                            nat.ptr(step_ms), nat.ptr(phase_ms) if per_kernel else None))
        return step_ms, phase_ms

    t = T_MID
    steps(t, W); t += W                      # warm-up (also captures the CUDA graphs)
    ctx.barrier()
    launches0 = lib.dibs_launch_count()
    step_ms, _ = steps_timed(t, K, per_kernel=False); t += K
    launches = lib.dibs_launch_count() - launches0
    ctx.barrier()
    total_ms = ctx.max_over_ranks(float(step_ms.sum()))
    res = {"value": K / (total_ms / 1e3), "ms_per_step": total_ms / K, "launches": int(launches), "steps": K,
           "joint": joint, "plan": plan, "exchange": plan.exchange}
    if want_hot:
        # the same K steps back to back with the state L2-resident (no flush), one event pair around the whole chunk
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ctx.barrier()
        e0.record(stream); steps(t, K); e1.record(stream); t += K
        ctx.barrier()
        res["value_l2_resident"] = K / (ctx.max_over_ranks(e0.elapsed_time(e1)) / 1e3)

    # ---- per-kernel times (eager launches behind a delay kernel, event after every kernel, cold L2 at step start) ---
    _, phase_ms = steps_timed(t, n_prof, per_kernel=True); t += n_prof
    work = kernel_work(name, plan.n_local, m)
    hbm_peak, bf16_peak, peak_src = measured_peaks()
    fp32_tf, fp64_tf, simt_src = simt_peaks()
    traffic = load_traffic(name, ctx.world)
    kernels = {}
    for i, ph in enumerate(nat.PHASES):
        if phase_ms[i] <= 0 or ph not in work:
            continue
        us = float(phase_ms[i]) / n_prof * 1e3
        wk = work[ph]
        kernels[ph] = {"us": round(us, 2), "share": None, "gflop_per_launch": round(wk["flops"] / 1e9, 4),
                       "mb_per_launch": round(wk["bytes"] / 1e6, 4), "tflops": round(wk["flops"] / us / 1e6, 3),
                       "gbs": round(wk["bytes"] / us / 1e3, 2), "bound": wk["bound"], "traffic": traffic.get(ph)}
        if "flops_executed" in wk:
            kernels[ph]["tflops_executed"] = round(wk["flops_executed"] / us / 1e6, 3)
    tot_us = sum(e["us"] for e in kernels.values())
    for e in kernels.values():
        e["share"] = round(e["us"] / tot_us, 4)
    dom = max(kernels, key=lambda k_: kernels[k_]["us"]) if kernels else None
    roofline = None
    if dom:
        e = kernels[dom]
        if e["bound"] == "hbm":
            roofline = {"bound": "hbm", "achieved": e["gbs"], "peak": hbm_peak, "unit": "GB/s",
                        "frac": round(e["gbs"] / hbm_peak, 4), "traffic": e["traffic"],
                        "peak_source": f"{peak_src} (MEASURED_PEAKS.json)"}
        else:
            # the dominant kernels are fp32 (fp64 for BGe) SIMT arithmetic: neither the HBM nor the tensor-pipe roofline
            # bounds them, so the denominator is the fp32 FMA pipe, which MEASURED_PEAKS.json
            # does not carry -> measured with tools/ubench.cu (theoretical 148 SM x 128 lanes x 2 x 1.965 GHz)
            peak = fp64_tf if e["bound"] == "fp64" else fp32_tf
            roofline = {"bound": e["bound"] + "_simt", "achieved": e["tflops"], "peak": round(peak, 2), "unit": "TFLOP/s",
                        "frac": round(e["tflops"] / peak, 4), "traffic": e["traffic"],
                        "peak_source": simt_src + "; MEASURED_PEAKS.json carries no SIMT figure",
                        "hbm_frac_of_measured": round(e["gbs"] / hbm_peak, 5)}
            if "tflops_executed" in e:
                # `achieved` charges the reference's algorithm (N x d products); the kernel executes ~N/d times fewer
                # FMAs (QR form), so frac can exceed 1 -- the pipe-utilisation figure is this one
                roofline["achieved_executed"] = e["tflops_executed"]
                roofline["frac_executed"] = round(e["tflops_executed"] / peak, 4)
        roofline["kernel"] = dom
        roofline["us_per_launch"] = e["us"]
        tb, per = step_bound(work, kernels, hbm_peak, fp32_tf, fp64_tf)
        roofline["step_bound_us"] = round(tb, 2)
        roofline["step_frac"] = round(tb / (total_ms / K * 1e3), 4)
        roofline["step_bound_by_kernel"] = per
    if os.environ.get("DIBS_BENCH_TIMELINE") == "1":
        # diagnostics: the step GRAPH replayed with an event node behind every kernel -> END time of each kernel
        # relative to the step's start, concurrency and peer waits included (us, mean over 20 steps, this rank)
        n_tl = 20
        step_ms_tl, tl_ms = steps_timed(t, n_tl, per_kernel=2); t += n_tl
        res["timeline_end_us"] = {ph: round(float(tl_ms[i]) / n_tl * 1e3, 1) for i, ph in enumerate(nat.PHASES) if tl_ms[i] > 0}
        res["timeline_end_us"]["step"] = round(float(step_ms_tl.mean()) * 1e3, 1)
    # LAST use of this state: everything above (per-kernel profile, roofline, timeline) saw the particles of `value`
    if want_hot if want_t0 is None else want_t0:
        # the same K device-timed steps (cold L2, events, max over ranks) but from t = 0 on re-initialised particles and
        # optimizer state: the steps a sample() call really executes
        key0 = PRNGKey(0)
        key0, subk0 = split(key0, 2)
        init0 = model._sample_initial_random_particles(key=subk0, n_particles=m, n_dim=d, plan=plan)
        z0_all, th0_all = init0 if joint else (init0, None)
        z.copy_(z0_all[lo:hi]); v_z.zero_(); sf.zero_()
        if joint:
            theta.copy_(th0_all[lo:hi]); v_th.zero_()
        key_dev.copy_(keys_to_device(key0, device))
        ctx.barrier()
        ms0, _ = steps_timed(0, K, per_kernel=False)
        res["value_from_t0"] = K / (ctx.max_over_ranks(float(ms0.sum())) / 1e3)
        ctx.barrier()
    res.update(kernels=kernels, roofline=roofline, model=model, t_next=t)
    return res


def pass_rooflines(ctx, name, res):
    """The two passes BASELINE.json's north star names, each against SURVEY 8(d)'s ALGORITHMIC bytes and flops:
    kernel-matrix pass = the pairwise kernels of the step (per-kernel events of the step itself); edge-prob pass =
    `dibs_edge_probs` timed standalone with CUDA events -- at the workload's M (a ~1 us pass: launch-latency floor) and
    at a particle count whose traffic exceeds L2 (what the kernel sustains when the pass is large enough to measure)."""
    import ctypes
    torch = ctx.torch
    from dibs_b200 import _native as nat
    desc, lik, d, m, s, a, h = WORKLOADS[name]
    k = d
    dz = 2 * d * k
    dth = {"bge": 0, "lingauss": d * d, "densenn": d * (d * h + 2 * h + 1)}[lik]
    D = dz + dth
    hbm_peak, _, peak_src = measured_peaks()
    fp32_tf, _, _ = simt_peaks()
    m_loc = m // ctx.world
    out = {}
    kn = res["kernels"]
    km_us = sum(kn[p]["us"] for p in ("pair_dist", "pair_kernel") if p in kn)
    if km_us > 0:
        by = 4 * m * D + 4 * m_loc * m                      # read particles once, write K once
        fl = 3 * m_loc * m * D                              # difference form
        out["kernel_matrix"] = {"us": round(km_us, 2), "alg_bytes": by, "alg_flops": fl,
                                "gbs": round(by / km_us / 1e3, 2), "hbm_frac": round(by / km_us / 1e3 / hbm_peak, 4),
                                "tflops": round(fl / km_us / 1e6, 3), "fp32_frac": round(fl / km_us / 1e6 / fp32_tf, 4),
                                "bound": "fp32_simt (AI = %.0f flop/B >> ridge %.1f)" % (fl / by, fp32_tf * 1e3 / hbm_peak),
                                "traffic": (kn.get("pair_dist", {}).get("traffic") or 0) + (kn.get("pair_kernel", {}).get("traffic") or 0) or None}
    model = res["model"]
    stream = torch.cuda.current_stream(ctx.device)

    def time_edge(n, reps):
        z = torch.randn((n, d, k, 2), dtype=torch.float32, device=ctx.device)
        o = torch.empty((n, d, d), dtype=torch.float32, device=ctx.device)
        plan = model._plan(n, k)
        sp = ctypes.c_void_p(stream.cuda_stream)
        for _ in range(3):
            nat.check(nat.lib().dibs_edge_probs(plan.handle, nat.ptr(z), n, T_MID, nat.ptr(o), sp))
        ts = []
        for _ in range(reps):
            ctx.flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            nat.check(nat.lib().dibs_edge_probs(plan.handle, nat.ptr(z), n, T_MID, nat.ptr(o), sp))
            e1.record(stream)
            e1.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        return float(np.median(ts))
    by1 = 8 * d * k + 4 * d * d                             # per particle: read Z once, write P once
    n_big = max(m_loc, int(3 * 126e6 / by1))
    us_m, us_big = time_edge(m_loc, 20), time_edge(n_big, 10)
    out["edge_prob"] = {"us_at_workload_m": round(us_m, 2), "alg_bytes": m_loc * by1, "alg_flops": m_loc * 2 * d * d * k,
                        "gbs_at_workload_m": round(m_loc * by1 / us_m / 1e3, 2),
                        "hbm_frac_at_workload_m": round(m_loc * by1 / us_m / 1e3 / hbm_peak, 4),
                        "n_large": n_big, "us_large": round(us_big, 2), "gbs_large": round(n_big * by1 / us_big / 1e3, 2),
                        "hbm_frac_large": round(n_big * by1 / us_big / 1e3 / hbm_peak, 4),
                        "note": "inside the step the pass is fused (raw scores U V^T are produced where the new Z row is on chip; "
                                "P never reaches HBM); standalone it is launch-latency-bound at the workload's M, so the kernel is "
                                "also timed at a particle count whose traffic exceeds L2"}
    out["peak_hbm_gbs"] = hbm_peak
    out["peak_source"] = f"{peak_src} (MEASURED_PEAKS.json)"
    return out


def run_native(args):
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- dibs_b200 has no CPU path (use --impl reference for the CPU arm)")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus != world and world == 1 and args.gpus > 1:
        raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    ctx = Ctx()
    dist, device, rank = ctx.dist, ctx.device, ctx.rank
    from dibs_b200.inference.dibs import PRNGKey

    name = args.workload
    desc, lik, d, m, s, a, h = WORKLOADS[name]
    joint = lik != "bge"
    x_host = torch.from_numpy(workload_data(name)).pin_memory()
    K, W = args.steps, max(args.warmup, 3)

    sampler = ClockSampler(ctx.local)
    sampler.start()
    res = measure_workload(ctx, name, K, W, min(K, 50))
    clocks = sampler.stop()
    passes_rl = pass_rooflines(ctx, name, res)

    # ---- end to end through the public API with host buffers -------------------------------------------------
    def e2e_once(n_steps):
        ctx.barrier()
        t0 = time.perf_counter()
        mdl = build_model(name, x_host, device)                     # H2D of x from pinned host memory
        out = mdl.sample(key=PRNGKey(0), n_particles=m, steps=n_steps)
        g = out[0] if joint else out
        g_host = g.cpu()                                            # D2H of the result
        th_host = None
        if joint:
            th = out[1]
            flat = mdl.likelihood_model.flatten(th) if not isinstance(th, torch.Tensor) else th
            th_host = flat.cpu()
        torch.cuda.synchronize(device)
        dt_ = time.perf_counter() - t0
        d2h = g_host.numel() * g_host.element_size() + (th_host.numel() * th_host.element_size() if th_host is not None else 0)
        return ctx.max_over_ranks(dt_), d2h

    e2e_once(min(K, 20))                                            # warm the public path once (library, allocator)
    e2e_s, d2h_bytes = e2e_once(K)
    h2d_bytes = x_host.numel() * x_host.element_size() + 8

    # ---- the other single-GPU configs of BASELINE.json, short runs in the same line --------------------------------
    also = {}
    for other in ([] if args.no_also else [w for w in ("c2", "t_bge", "c3") if w != name]):
        o_steps = {"c3": 20}.get(other, 50)
        r = measure_workload(ctx, other, o_steps, 3, min(o_steps, 10), want_hot=False, want_t0=True)
        rl = r["roofline"] or {}
        also[other] = {"workload": WORKLOADS[other][0], "value": r["value"], "value_from_t0": r.get("value_from_t0"), "t_start": T_MID,
                       "unit": "steps/s", "ms_per_step": r["ms_per_step"],
                       "steps": o_steps, "dominant_kernel": rl.get("kernel"), "frac": rl.get("frac"), "bound": rl.get("bound"),
                       "us_per_launch": rl.get("us_per_launch"),
                       "kernels_us": {k_: v["us"] for k_, v in r["kernels"].items()}}
        del r

    if rank != 0:
        if ctx.world > 1:
            dist.destroy_process_group()
        return

    # ---- CPU baseline (rank 0, N = 1 only) -------------------------------------------------------------------
    cpu_baseline = None
    if ctx.world == 1 and not args.no_cpu_baseline:
        budget = float(os.environ.get("DIBS_BENCH_CPU_BUDGET_S", "20"))
        sps, sec, sample, cores = cpu_steps(name, 2, 1, budget)
        cpu_baseline = {"value": sps, "unit": "steps/s", "cores": cores, "kind": "port", "sample": sample}

    n_passes = 2 if joint else 1
    value = res["value"]
    roofline = res["roofline"]
    if roofline is not None:
        roofline["passes"] = passes_rl
    cfg = config_dict(name, ctx.world)
    cfg["exchange"] = res["exchange"]
    line = {
        "metric": "svgd_steps_per_sec", "value": value, "unit": "steps/s", "n_gpus": ctx.world, "steps": K, "warmup": W,
        "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32" if lik != "bge" else "f32 (BGe Cholesky in f64)", "data": "synthetic",
        "config": cfg,
        "graphs_scored_per_sec": value * m * s * n_passes,
        "value_l2_resident": res.get("value_l2_resident"),
        "value_from_t0": res.get("value_from_t0"),
        "gpu_launches": res["launches"],
        "clocks": clocks,
        "e2e": {"value": K / e2e_s, "unit": "steps/s", "h2d_bytes_per_step": h2d_bytes / K, "d2h_bytes_per_step": d2h_bytes / K,
                "what": "JointDiBS/MarginalDiBS(x=<pinned host>).sample(steps=K) + results to host, wall clock incl. upload of x, "
                        "particle init and the final Z -> G (the native plan -- workspace, data factorisation, CUDA graphs, "
                        "peer-memory handshake -- is cached per process and was created by the warm-up call)",
                "h2d_bytes_total": h2d_bytes, "d2h_bytes_total": d2h_bytes, "seconds": e2e_s},
        "roofline": roofline,
        "kernels": res["kernels"],
        "also": also,
        "timeline_end_us": res.get("timeline_end_us"),
        "cpu_baseline": cpu_baseline,
    }
    print(json.dumps(line), flush=True)
    if ctx.world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--workload", default="t_lin", choices=sorted(WORKLOADS),
                    help="default: the north-star target shape (JointDiBS LinearGaussian n_vars=20 n_particles=1024 n_mc=128)")
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the short runs of the other BASELINE configs")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()

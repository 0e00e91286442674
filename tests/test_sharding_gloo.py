"""World-size-2 check of the multi-GPU decomposition on CPU (gloo): the step is sharded by particle exactly as
dibs_b200/csrc/dibs_abi.cu::enqueue_step shards it -- every rank derives the SAME per-particle sub-keys from the
replicated loop key, computes the gradients of its own rows, one all-gather brings the particle rows and one the
gradient rows to every rank, then each rank updates its own rows from its slab of the kernel matrix -- and the result
must equal the single-process oracle step bit for bit.  Also covers the host-side pieces the sharded plan uses
(rank / world discovery, row partition, the unique-id broadcast)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import dibs_oracle as orc
from oracle import threefry as tf


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _case():
    rng = np.random.default_rng(0)
    d, n = 5, 20
    x = rng.normal(size=(n, d)).astype(np.float32)
    cfg = orc.Config(lik=orc.Likelihood(kind="lingauss", n_vars=d), prior=orc.GraphPrior("er", d, 1), joint=True,
                     n_grad_mc_samples=4, n_acyclicity_mc_samples=4)
    st = orc.init_particles(cfg, tf.prng_key(7), 6, None, np.float32)
    return cfg, st, x, np.zeros((n, d), np.int32)


def _gather(local):
    out = [torch.empty_like(local) for _ in range(dist.get_world_size())]
    dist.all_gather(out, local.contiguous())
    return torch.cat(out, 0)


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cfg, st, x, mask = _case()
        m = st.z.shape[0]
        # host-side discovery + partition used by DiBS._plan(sharded=True) / _Plan.__init__
        from dibs_b200.inference.dibs import DiBS
        w, r = DiBS._dist(None)
        assert (w, r) == (world, rank)
        n_local = m // world
        row0 = rank * n_local
        rows = range(row0, row0 + n_local)
        # the unique-id broadcast of _Plan.attach_nccl (CPU tensors under gloo)
        ident = torch.arange(128, dtype=torch.uint8) if rank == 0 else torch.zeros(128, dtype=torch.uint8)
        dist.broadcast(ident, src=0)
        assert ident.tolist() == list(range(128))

        t = 3
        # gradient phase on the rank's rows only; keys come from the replicated loop key
        dz, dth, key, base, _ = orc.particle_grads(cfg, st, t, x, mask, np.float32, particles=rows)
        # exchange 1 (side branch of the step graph): particle rows; exchange 2 (critical path): gradient rows
        z_all = _gather(torch.from_numpy(st.z[row0:row0 + n_local])).numpy()
        th_all = _gather(torch.from_numpy(st.theta[row0:row0 + n_local])).numpy()
        dz_all = _gather(torch.from_numpy(dz[row0:row0 + n_local])).numpy()
        dth_all = _gather(torch.from_numpy(dth[row0:row0 + n_local])).numpy()
        assert np.array_equal(z_all, st.z) and np.array_equal(th_all, st.theta)
        # pairwise phase: every rank needs all particles and gradients, and updates its own rows
        k_full, k_z, k_t = orc.kernel_matrix(cfg, z_all, th_all, np.float32)
        phi_z = orc.phi_update(k_full, k_z, cfg.h_latent, z_all, dz_all, np.float32)[row0:row0 + n_local]
        phi_t = orc.phi_update(k_full, k_t, cfg.h_theta, th_all, dth_all, np.float32)[row0:row0 + n_local]
        z_new, _ = orc.opt_update(cfg, st.z[row0:row0 + n_local], st.v_z[row0:row0 + n_local], phi_z, np.float32)
        th_new, _ = orc.opt_update(cfg, st.theta[row0:row0 + n_local], st.v_theta[row0:row0 + n_local], phi_t, np.float32)
        z_fin = _gather(torch.from_numpy(z_new)).numpy()
        th_fin = _gather(torch.from_numpy(th_new)).numpy()
        base_fin = _gather(torch.from_numpy(base[row0:row0 + n_local])).numpy()
        if rank == 0:
            ref = orc.svgd_step(cfg, st, t, x, mask, np.float32)
            ret["z"] = bool(np.array_equal(z_fin, ref.z))
            ret["theta"] = bool(np.array_equal(th_fin, ref.theta))
            ret["key"] = bool(np.array_equal(key, ref.key))
            ret["base"] = bool(np.array_equal(base_fin, ref.sf_baseline))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_step_equals_single_process():
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
        assert dict(ret) == {"z": True, "theta": True, "key": True, "base": True}


def test_sharded_plan_rejects_uneven_split():
    """n_particles must divide evenly over the ranks (checked in dibs_plan_create before any device work)."""
    import ctypes
    from dibs_b200 import _native as nat
    if not os.path.exists(nat.LIB_PATH):
        pytest.skip("native library not built")
    c = nat.DibsConfig()
    for name, _ in nat.DibsConfig._fields_:
        pass
    c.n_vars, c.n_dim, c.n_particles = 5, 5, 7
    c.n_grad_mc_samples, c.n_acyclicity_mc_samples = 4, 4
    c.world_size, c.rank = 2, 0
    out = ctypes.c_void_p()
    assert nat.lib().dibs_plan_create(ctypes.byref(c), ctypes.byref(out)) == -1
    assert b"divisible" in nat.lib().dibs_last_error()

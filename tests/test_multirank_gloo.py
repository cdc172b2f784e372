"""The N > 1 path on CPU: world_size-2 `gloo` run of the sharded Gauss-Newton loop.  Every rank holds a
replica of the map and a contiguous shard of the queries (sage_shard_range, the product's own host arithmetic); per
iteration each rank reduces its shard to the normal-equation sums, one all-reduce(sum) makes them global, and every rank
takes the same step.  On the GPU box the per-shard reduction is nn_search_kernel and the exchange is NCCL; here the
oracle stands in for the kernel so that the sharding, the exchange and the lock-step property are covered without a GPU."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASIC_LABELS = [40, 44, 48, 49, 50, 70, 72]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import sage_icp_b200 as sg
    from sage_icp_b200 import synthetic as syn
    from oracle import oracle_py as orc
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        m = orc.OracleMap(0.8, 100.0, 20, 20, BASIC_LABELS, evict_faithful=False)
        m.add_points(syn.sample_street_map(60_000, 3, -30.0, 30.0))
        scan = syn.make_scan(5, (0.0, 0.0, 0.0), n_beams=16, n_az=400)
        r = np.linalg.norm(scan[:, :3], axis=1)
        scan = scan[(r > 4) & (r < 35)]
        guess = syn.pose7_from_xyyaw((0.2, -0.1, 0.006))
        b, e = sg.shard_range(len(scan), rank, world)
        src_full = np.c_[np.array([orc.se3_act(guess, p[:3]) for p in scan]), scan[:, 3]]
        src = src_full[b:e].copy()
        T = np.array([0, 0, 0, 0, 0, 0, 1.0])
        steps = []
        for _ in range(4):
            s, t, _ = m.get_correspondences(src, 3.0, 0.4)
            JTJ, JTr, _, _ = orc.align_clouds(s, t, 1.0 / 3.0) if len(s) else (np.zeros((6, 6)), np.zeros(6), None, None)
            buf = torch.from_numpy(np.r_[JTJ.ravel(), JTr, float(len(s))].copy())
            dist.all_reduce(buf, op=dist.ReduceOp.SUM)  # the only exchange of the path: 43 doubles here, 17 on the GPU
            A, g = buf[:36].numpy().reshape(6, 6), buf[36:42].numpy()
            est = orc.se3_exp(orc.ldlt6_solve(A, -g))
            src[:, :3] = np.array([orc.se3_act(est, p[:3]) for p in src])
            T = orc.se3_mul(est, T)
            steps.append(np.r_[est, buf[42].item()])
        # unsharded run of the same iterations on rank 0 for comparison
        ref = None
        if rank == 0:
            full, Tr, ref = src_full.copy(), np.array([0, 0, 0, 0, 0, 0, 1.0]), []
            for _ in range(4):
                s, t, _ = m.get_correspondences(full, 3.0, 0.4)
                _, _, _, est = orc.align_clouds(s, t, 1.0 / 3.0)
                full[:, :3] = np.array([orc.se3_act(est, p[:3]) for p in full])
                Tr = orc.se3_mul(est, Tr)
                ref.append(np.r_[est, float(len(s))])
        q.put((rank, np.array(steps), None if ref is None else np.array(ref), (b, e), len(scan)))
    finally:
        dist.destroy_process_group()


def test_sharded_gauss_newton_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = {}
    for _ in range(2):
        rank, steps, ref, shard, n = q.get(timeout=240)
        out[rank] = (steps, ref, shard, n)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (s0, ref, sh0, n), (s1, _, sh1, _) = out[0], out[1]
    assert sh0[0] == 0 and sh0[1] == sh1[0] and sh1[1] == n and abs((sh0[1] - sh0[0]) - (sh1[1] - sh1[0])) <= 1
    assert np.array_equal(s0, s1)  # lock-step: both ranks took bit-identical steps (all-reduce gives every rank the same sums)
    assert np.array_equal(s0[:, 7], ref[:, 7])  # same number of correspondences as the unsharded run, every iteration
    assert np.allclose(s0[:, :7], ref[:, :7], atol=1e-11)  # and the same step up to summation order
    assert np.linalg.norm(s0[-1, :3]) < np.linalg.norm(s0[0, :3])  # converging

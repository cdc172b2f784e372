"""N > 1 on real GPUs (skipped on a single-GPU box; `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multirank.py -m gpu`):
two ranks, replicated map, sharded queries, all-reduce of the normal equations either fused into the search kernel over
NVLink peer memory (sage_map_comm_peer_*) or by NCCL.  Every rank must return the bit-identical pose, equal (to rounding) to
the unsharded single-GPU registration and to the oracle within the pose tolerance."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASIC_LABELS = [40, 44, 48, 49, 50, 70, 72]


def _worker(rank, world, port, comm, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import sage_icp_b200 as sg
    from sage_icp_b200 import synthetic as syn
    import bench
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        m = sg.SageMap(0.8, 1e9, 20, 20, BASIC_LABELS, device=rank)
        m.add_points(syn.sample_street_map(400_000, 3, -60.0, 60.0))
        scan = syn.make_scan(5, (0.0, 0.0, 0.0), n_beams=64, n_az=1000)
        guess = syn.pose7_from_xyyaw((0.2, -0.1, 0.006))
        full_pose, full_it = m.register_frame(scan, guess, 3.0, 1.0 / 3.0, 0.4)  # unsharded, before any communicator exists
        if comm == "nccl":
            uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if rank == 0:
                uid.copy_(torch.frombuffer(bytearray(sg.nccl_unique_id()), dtype=torch.uint8))
            dist.broadcast(uid, 0)
            m.comm_init(rank, world, bytes(uid.cpu().numpy().tobytes()))
        else:
            mine = torch.frombuffer(bytearray(m.comm_peer_handle()), dtype=torch.uint8).cuda()
            allh = [torch.zeros(64, dtype=torch.uint8, device="cuda") for _ in range(world)]
            dist.all_gather(allh, mine)
            m.comm_peer_attach(rank, world, b"".join(bytes(h.cpu().numpy().tobytes()) for h in allh))
        dist.barrier()
        shard = bench.shard_of(scan, rank, world)
        poses = []
        for _ in range(3):  # repeated registrations: tags / parity slots keep working
            pose, it = m.register_frame(shard, guess, 3.0, 1.0 / 3.0, 0.4)
            poses.append(np.r_[pose, it])
        q.put((rank, np.array(poses), np.r_[full_pose, full_it], len(shard)))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("comm", ["peer", "nccl"])
def test_sharded_registration_two_gpus(orc, comm):
    import sage_icp_b200 as sg
    if sg.device_count() < 2:
        pytest.skip("needs two B200s")
    import torch.multiprocessing as mp
    from sage_icp_b200 import synthetic as syn
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29800 + (os.getpid() % 150) + (0 if comm == "peer" else 1)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, comm, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = {}
    for _ in range(2):
        rank, poses, full, n = q.get(timeout=300)
        out[rank] = (poses, full, n)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    (p0, full0, n0), (p1, full1, n1) = out[0], out[1]
    assert n0 + n1 == 64000 and abs(n0 - n1) <= 32
    assert np.array_equal(p0, p1)  # lock-step: bit-identical pose and iteration count on both ranks, every time
    assert np.array_equal(p0[0], p0[1]) and np.array_equal(p0[1], p0[2])  # and reproducible
    assert np.array_equal(full0, full1)
    assert p0[0][7] == full0[7]  # same iteration count as the unsharded registration
    assert np.abs(p0[0][:7] - full0[:7]).max() < 1e-10  # same pose up to summation order
    # and against the oracle
    o = orc.OracleMap(0.8, 1e9, 20, 20, BASIC_LABELS, evict_faithful=False)
    o.add_points(syn.sample_street_map(400_000, 3, -60.0, 60.0))
    pose_o, it_o = o.register_frame_core(syn.make_scan(5, (0.0, 0.0, 0.0), n_beams=64, n_az=1000), syn.pose7_from_xyyaw((0.2, -0.1, 0.006)),
                                         3.0, 1.0 / 3.0, 0.4, threads=orc.max_threads())
    from conftest import POSE_TOL_M, POSE_TOL_RAD, pose_delta
    dt, da = pose_delta(p0[0][:7], pose_o)
    assert it_o == int(p0[0][7]) and dt <= POSE_TOL_M and da <= POSE_TOL_RAD

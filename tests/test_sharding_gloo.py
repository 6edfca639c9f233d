"""Multi-GPU path of bench.py on CPU: two gloo ranks shard the batch (weak scaling, no data-path
collective), time with max-over-ranks, and rank 0 aggregates.  The per-rank solve itself is replaced
by the oracle on a tiny batch, so the host logic (sharding, seeds, aggregation) is what is tested."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    import oracle_py as O
    B = 6
    p = O.benchmark_problem()
    q0, v0 = bench.initial_states(rank * B, B, list(p.q_min), list(p.q_max))
    batch = O.Batch(p, B)
    for b, s in enumerate(batch.solvers):
        s.set_solution("q", q0[b])
        s.set_solution("v", v0[b])
    batch.update_solution(0.0, q0, v0, False, 1)
    kkt = batch.kkt_error(0.0, q0, v0, 1)
    # max-over-ranks timing reduction, as bench.py does with NCCL
    t = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    gathered = [torch.zeros(B, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(kkt))
    if rank == 0:
        out.put((float(t.item()), [g.numpy() for g in gathered], q0))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import bench
    import oracle_py as O
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    tmax, gathered, q0_rank0 = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert tmax == 11.0                              # max over ranks
    # the union of the two shards is the single-process batch of 12 instances, in order
    B = 6
    prob = O.benchmark_problem()
    q0, v0 = bench.initial_states(0, 2 * B, list(prob.q_min), list(prob.q_max))
    assert np.array_equal(q0[:B], q0_rank0)
    batch = O.Batch(prob, 2 * B)
    for b, s in enumerate(batch.solvers):
        s.set_solution("q", q0[b])
        s.set_solution("v", v0[b])
    batch.update_solution(0.0, q0, v0, False, 1)
    ref = batch.kkt_error(0.0, q0, v0, 1)
    assert np.array_equal(np.concatenate(gathered), ref)


def test_initial_states_are_inside_the_joint_limits():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import bench
    import oracle_py as O
    p = O.benchmark_problem()
    q0, v0 = bench.initial_states(0, 4096, list(p.q_min), list(p.q_max))
    assert np.all(q0 > np.array(list(p.q_min))) and np.all(q0 < np.array(list(p.q_max)))
    assert np.all(np.abs(v0) <= 0.5)
    # counter-based: shard k of size n == rows [k n, (k+1) n) of the big batch
    q1, _ = bench.initial_states(1024, 16, list(p.q_min), list(p.q_max))
    assert np.array_equal(q1, q0[1024:1040])


def _anymal_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import anymal_problems as tp
    import fb_py
    import oracle_py
    from idocp_b200 import problems as P
    oracle_py.build()
    fb_py.lib()
    B = 2
    pr = tp.TrottingProblem()
    q0, v0 = P.anymal_initial_states(rank * B, B, q_nominal=pr.q0)
    kkt = np.zeros(B)
    for b in range(B):
        o = pr.make_oracle(fb_py, q0=q0[b], v0=v0[b])
        o.update_solution(0.0, q0[b], v0[b])
        o.compute_kkt_residual(0.0, q0[b], v0[b])
        kkt[b] = o.kkt_error()
    gathered = [torch.zeros(B, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(kkt))
    if rank == 0:
        out.put([g.numpy() for g in gathered])
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_of_the_anymal_batch():
    """bench.py --workload anymal_*: rank r owns instances [r B, (r+1) B) of the counter-based splitmix64 stream; the
    union of two shards is the single-process batch (ANYmal OCPSolver, no data-path collective)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import anymal_problems as tp
    import fb_py
    import oracle_py
    from idocp_b200 import problems as P
    oracle_py.build()
    fb_py.lib()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 30100 + (os.getpid() % 500)
    procs = [ctx.Process(target=_anymal_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    gathered = out.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    pr = tp.TrottingProblem()
    q0, v0 = P.anymal_initial_states(0, 4, q_nominal=pr.q0)
    assert np.allclose(np.linalg.norm(q0[:, 3:7], axis=1), 1.0, atol=1e-15)
    assert np.all(np.abs(q0[:, :3] - pr.q0[:3]) <= 0.01) and np.all(np.abs(q0[:, 7:] - pr.q0[7:]) <= 0.02) and np.all(np.abs(v0) <= 0.1)
    ref = []
    for b in range(4):
        o = pr.make_oracle(fb_py, q0=q0[b], v0=v0[b])
        o.update_solution(0.0, q0[b], v0[b])
        o.compute_kkt_residual(0.0, q0[b], v0[b])
        ref.append(o.kkt_error())
    assert np.array_equal(np.concatenate(gathered), np.array(ref))

"""Multi-rank path on CPU: world_size 2, gloo.  The same batch of registrations solved on 1 rank and sharded over 2 ranks gives the
identical result table on every rank (the GPU solver is replaced here by the CPU oracle — tests may use it; shard.py itself never does)."""
import os
import socket

import numpy as np
import pytest

from randt_slam_b200 import shard


def test_partition_blocks_cover_and_balance():
    assert shard.partition(10, 3) == [(0, 3), (3, 6), (6, 10)]
    assert shard.partition(2, 4) == [(0, 0), (0, 1), (1, 1), (1, 2)]
    w = [10, 1, 1, 1, 1, 10, 1, 1]
    blocks = shard.partition(8, 2, w)
    assert blocks[0][0] == 0 and blocks[-1][1] == 8 and blocks[0][1] == blocks[1][0]
    tot = [sum(w[a:b]) for a, b in blocks]
    assert abs(tot[0] - tot[1]) <= 10
    for world in (1, 2, 3, 8):
        b = shard.partition(256, world, np.random.default_rng(0).integers(100, 300, 256))
        assert b[0][0] == 0 and b[-1][1] == 256 and all(b[i][1] == b[i + 1][0] for i in range(world - 1))
        sizes = [e - a for a, e in b]
        assert max(sizes) - min(sizes) <= 0.25 * 256 / world + 2


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _solve_factory():
    from oracle import oracle_py as O
    from randt_slam_b200 import params as P
    from tests import helpers as H
    p = P.OXFORD
    cases = [H.make_registration_case(O, p, seed=30 + i, n_fixed_scans=2, true_pose=(0.6, -0.4, 0.03), guess=(0.4 + 0.03 * i, -0.3, 0.02)) for i in range(5)]

    def solve_block(begin, end):
        rows = np.zeros((end - begin, shard.ROW))
        for j, c in enumerate(cases[begin:end]):
            f = c["fixed"]
            o = O.loop_constraint(f["cells"], f["slot"], p.size_x, p.size_y, p.resolution, p.max_neighbor_linf_distance, c["moving"]["cells"], c["pose0"],
                                  p.n_results_nn_lookup, matcher_loss_scale=p.loss_function_scale, loop_scale=p.loop_closure_scale,
                                  alpha=p.loss_function_convexity, divisor=p.gnc_control_parameter_divisor, max_gnc_steps=2, pairs=(c["im"], c["jf"]))
            rows[j, :4] = o["pose"]; rows[j, 4] = o["score"]; rows[j, 5] = o["iterations"]; rows[j, 6] = o["status"]
        return rows
    return solve_block, len(cases), [len(c["im"]) for c in cases]


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        solve, n, weights = _solve_factory()
        table = shard.register_sharded(solve, n, rank, world, weights)
        q.put((rank, table))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_ranks_gloo_match_single_rank():
    import torch.multiprocessing as mp
    solve, n, weights = _solve_factory()
    single = shard.register_sharded(solve, n, 0, 1, weights)
    assert single.shape == (n, shard.ROW) and np.array_equal(single[:, 7], np.arange(n))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    got = dict(q.get(timeout=240) for _ in range(2))
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    for r in range(2):
        assert np.array_equal(got[r], single), "rank %d table differs from the single-rank result" % r

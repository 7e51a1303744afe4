"""Multi-GPU sharding of independent registrations (BASELINE configs[3]).

Each (moving scan, fixed submap) registration is a closed problem — Matcher::estimateLoopConstraint builds a fresh ceres::Problem per
call (R/src/ndt_registration/ndt_matcher.cpp:427) — so a batch shards with no data-path collective: rank r owns one contiguous block
of problems (balanced by pair count when the counts are known), runs its own K3/K4 loop on its own GPU, and ONE all_gather of the
per-problem result rows returns every pose to every rank.  One process per GPU; torch.distributed is only the plumbing
(NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
import numpy as np

ROW = 8   # [cos, sin, tx, ty, score, iterations, status, problem index]


def partition(n_problems, world, weights=None):
    """Contiguous blocks [begin, end) per rank.  weights (e.g. pairs per problem) balances the blocks by total weight."""
    if world <= 0:
        raise ValueError("world must be positive")
    if weights is None:
        cuts = [(n_problems * r) // world for r in range(world + 1)]
    else:
        w = np.asarray(weights, np.float64)
        if len(w) != n_problems:
            raise ValueError("one weight per problem")
        c = np.concatenate([[0.0], np.cumsum(w)])
        total = c[-1]
        cuts = [0]
        for r in range(1, world):
            target = total * r / world
            k = int(np.searchsorted(c, target, side="left"))
            k = min(max(k, cuts[-1]), n_problems)
            # pick the nearer of the two neighbouring cut points
            if k > cuts[-1] and k <= n_problems and abs(c[k - 1] - target) <= abs(c[min(k, n_problems)] - target):
                k -= 1
            cuts.append(max(k, cuts[-1]))
        cuts.append(n_problems)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def gather_results(local_rows, begin, n_problems, rank, world, device=None):
    """all_gather of the per-problem rows: every rank returns the full [n_problems, ROW] table, in problem order.

    local_rows: [n_local, ROW] float64 of this rank's block starting at problem `begin` (column 7 is overwritten with the global index).
    """
    import torch
    import torch.distributed as dist

    local = np.ascontiguousarray(local_rows, np.float64).reshape(-1, ROW).copy()
    local[:, 7] = np.arange(begin, begin + len(local))
    if world == 1:
        return local
    counts = torch.zeros(world, dtype=torch.int64, device=device)
    counts[rank] = len(local)
    dist.all_reduce(counts)
    cap = int(counts.max().item())
    buf = torch.full((cap, ROW), float("nan"), dtype=torch.float64, device=device)
    if len(local):
        buf[: len(local)] = torch.from_numpy(local).to(buf.device)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)                      # the job's only data exchange: 64 B per problem
    table = np.full((n_problems, ROW), np.nan)
    for r in range(world):
        rows = out[r][: int(counts[r].item())].cpu().numpy()
        if len(rows):
            table[rows[:, 7].astype(np.int64)] = rows
    return table


def register_sharded(solve_block, n_problems, rank, world, weights=None, device=None):
    """solve_block(begin, end) -> [end - begin, ROW] rows for this rank's block; returns the gathered table on every rank."""
    begin, end = partition(n_problems, world, weights)[rank]
    rows = solve_block(begin, end) if end > begin else np.zeros((0, ROW))
    return gather_results(rows, begin, n_problems, rank, world, device)

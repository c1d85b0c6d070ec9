"""Multi-GPU sharding of the pair lists (one process per GPU, torch.distributed).

The path shards naturally: every ordered pair (and every refined child of it) is independent
(/root/reference has no multi-GPU code at all).  Each class's task list is cut into `world` contiguous shards of
equal predicted cost; all refined descendants of a task stay on the task's rank, so the per-task sums and the Runge
decisions are local.  The only exchange step is the gather of the per-pair results (Point3, 24 B/pair) to the
exporting rank, done with NCCL point-to-point transfers over NVLink (variable-length shards, no padding).
"""
from __future__ import annotations

# relative cost of one task of a class at refinement level 0 (FP64 instruction counts of the three kernels,
# profiles/ round 1): the adjacent classes evaluate thetaPsi in the reference's operation order plus the singular part.
CLASS_COST = (2.6, 2.4, 1.0)


def shard_bounds(n: int, world: int, weights=None):
    """Contiguous [lo, hi) ranges that split n tasks (optionally with per-task integer costs) into `world` parts of
    equal cost.  Without weights the parts differ by at most one task."""
    if weights is None:
        base, rem = divmod(n, world)
        out, lo = [], 0
        for r in range(world):
            hi = lo + base + (1 if r < rem else 0)
            out.append((lo, hi))
            lo = hi
        return out
    import numpy as np
    c = np.concatenate([[0], np.cumsum(np.asarray(weights, dtype=np.float64))])
    total = c[-1]
    cuts = [int(np.searchsorted(c, total * r / world, side="left")) for r in range(world + 1)]
    cuts[0], cuts[-1] = 0, n
    for r in range(1, world + 1):
        cuts[r] = max(cuts[r], cuts[r - 1])
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def predicted_task_cost(cls: int, level: int, adaptive_depth=None):
    """Cost model used for balancing: class weight x 4^level (fixed) or x sum_{l<=depth} 4^l (adaptive, per task)."""
    if adaptive_depth is None:
        return CLASS_COST[cls] * (4 ** max(level, 0))
    return CLASS_COST[cls] * sum(4 ** l for l in range(int(adaptive_depth) + 1))


def gather_results(local, full, bounds, rank: int, world: int, dst: int = 0):
    """Gather contiguous row shards `local` (rows bounds[rank]) into `full` on rank dst (NCCL/gloo point-to-point)."""
    import torch.distributed as dist
    if world == 1:
        if full is not None and full.data_ptr() != local.data_ptr():
            full.copy_(local)
        return
    ops = []
    if rank == dst:
        lo, hi = bounds[dst]
        full[lo:hi].copy_(local)
        for r in range(world):
            if r == dst:
                continue
            lo, hi = bounds[r]
            if hi > lo:
                ops.append(dist.P2POp(dist.irecv, full[lo:hi], r))
    else:
        lo, hi = bounds[rank]
        if hi > lo:
            ops.append(dist.P2POp(dist.isend, local, dst))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


def integrate_and_gather(ctx, cls, tasks_local, level, out_local, full, bounds, rank, world, side_stream, chunks=8, dst=0):
    """Integrate this rank's shard of a class chunk by chunk and ship every finished chunk of per-pair results to rank
    `dst` while the next chunk is being integrated (NCCL point-to-point on a side stream, overlapped with compute).

    tasks_local: int32[n_local,3] device tensor (rows bounds[rank] of the class), out_local = (integrals, results) device
    tensors for the shard, full: float64[n_total,3] on rank dst (None elsewhere).  Returns the list of pending NCCL works;
    call wait_all() on it before reading `full`."""
    import torch
    import torch.distributed as dist
    n_local = int(tasks_local.shape[0])
    integrals, results = out_local
    works = []
    n_max = max(hi - lo for lo, hi in bounds)
    step = max(1, -(-n_max // max(1, chunks)))           # same chunk grid on every rank
    main = torch.cuda.current_stream()
    for c0 in range(0, n_max, step):
        lo_c, hi_c = min(c0, n_local), min(c0 + step, n_local)
        if hi_c > lo_c:
            ctx.integrate_class(cls, tasks_local[lo_c:hi_c], level, want_stats=False,
                                out=(integrals[lo_c:hi_c], results[lo_c:hi_c]))
        if world == 1:
            continue
        ev = torch.cuda.Event()
        ev.record(main)
        with torch.cuda.stream(side_stream):
            side_stream.wait_event(ev)
            ops = []
            if rank == dst:
                glo, _ = bounds[dst]
                if hi_c > lo_c:
                    full[glo + lo_c:glo + hi_c].copy_(results[lo_c:hi_c], non_blocking=True)
                for r in range(world):
                    if r == dst:
                        continue
                    rlo, rhi = bounds[r]
                    a, b = min(c0, rhi - rlo), min(c0 + step, rhi - rlo)
                    if b > a:
                        ops.append(dist.P2POp(dist.irecv, full[rlo + a:rlo + b], r))
            elif hi_c > lo_c:
                ops.append(dist.P2POp(dist.isend, results[lo_c:hi_c], dst))
            if ops:
                works += dist.batch_isend_irecv(ops)
    return works


def wait_all(works, side_stream=None):
    import torch
    for w in works:
        w.wait()
    if side_stream is not None:
        torch.cuda.current_stream().wait_stream(side_stream)

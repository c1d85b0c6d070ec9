"""Multi-GPU helpers on the Python side (one process per GPU, torch.distributed).

The product's multi-GPU layer is native: i2_mgpu_* in include/i2_abi.h (integrator2_b200/abi.py: MultiGpu) shards the
prepare, integrates, and runs NCCL inside the library.  What remains here: the host-side shard arithmetic (mirrored by tests),
the cost model, PeerExport (CUDA IPC plumbing for the peer-store export) and `exchange_rounds`, the "bring your own
communicator" variant of what i2_mgpu_run does between its two halves.

The path shards naturally: every ordered pair (and every refined child of it) is independent
(/root/reference has no multi-GPU code at all).  Each class's task list is cut into `world` contiguous shards of
equal predicted cost; all refined descendants of a task stay on the task's rank, so the per-task sums and the Runge
decisions are local.  The only exchange step is the gather of the per-pair results (Point3, 24 B/pair) to the
exporting rank.  Two implementations: `PeerExport` (default for the export use case) maps the exporting rank's array into
every process and lets the integrate kernels store their results into it directly over NVLink (compute + gather fused in
one kernel); `integrate_and_gather` is the NCCL point-to-point baseline (variable-length shards, chunks overlapped).
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
    if level < 0:
        # error control: the value a converged task ends with depends on the parity of the LAST round of the list it was integrated
        # with (the reference's result ping-pong) — a chunk would get its own last round.  One call per shard; a
        # multi-GPU run that must reproduce the single-GPU values uses the C ABI's i2_mgpu_run (MultiGpu.run), which agrees on the
        # last rounds across the GPUs before the final assembly.
        chunks = 1
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


class PeerExport:
    """Export array of one class (Point3[n_total]) living on rank `dst`, mapped into every other rank's address space
    with CUDA IPC (C ABI: i2_peer_alloc / i2_peer_open).  A rank passes `results_arg()` as the results pointer of
    integrate_class / integrate_all: the kernels' final-assembly stores then land in rank dst's HBM directly over
    NVLink — compute and gather are ONE kernel, there is no staging buffer, no NCCL call and no second pass over the data.
    After the step: torch.cuda.synchronize() on every rank, then a host barrier; `full` (rank dst) is complete."""

    def __init__(self, ctx, n_total, bounds, rank, world, dst=0, handle_exchange=None):
        self.ctx, self.bounds, self.rank, self.dst = ctx, bounds, rank, dst
        self.full, self.addr, self.mapped = None, 0, False
        handle = None
        if rank == dst:
            self.full, handle, self.addr = ctx.peer_alloc(n_total)
        if world > 1:
            if handle_exchange is None:
                import torch.distributed as dist
                box = [handle]
                dist.broadcast_object_list(box, src=dst)
                handle = box[0]
            else:
                handle = handle_exchange(handle)
            if rank != dst:
                self.addr = ctx.peer_open(handle)
                self.mapped = True

    def results_arg(self):
        """raw device address of this rank's row block inside the export array (24 B per Point3 row)"""
        return self.addr + 24 * self.bounds[self.rank][0]

    def close(self):
        if self.mapped:
            self.ctx.peer_close(self.addr)
        elif self.addr:
            self.full = None
            self.ctx.peer_free(self.addr)
        self.addr, self.mapped = 0, False


def wait_all(works, side_stream=None):
    import torch
    for w in works:
        w.wait()
    if side_stream is not None:
        torch.cuda.current_stream().wait_stream(side_stream)


# Probability that a regular task is still unconverged after the first refinement round, as a function of
# rho = |c_i - c_j| / sqrt(max(S_i, S_j)) (centroid distance over panel size), measured with the CPU oracle on the
# reference's s5m.dat airplane (scale 0.0005; tools' calibration in DESIGN.md section 7).  Far pairs never refine twice.
_RHO_EDGES = (1.0, 1.5, 2.0, 3.0, 4.0, 6.0, 10.0)
_P_UNCONVERGED = (0.69, 0.42, 0.13, 0.031, 0.0066, 0.0020, 0.0006, 0.0)
# expected child integrations of a task that survives round 1: 16 + 0.33 (64 + 0.23 (256 + 0.13 * 1024))
_TAIL_COST = 66.7
_BASE_COST = 5.0     # rounds 0 and 1 are unconditional: 1 + 4 child integrations


def adaptive_task_cost(vertices, cells, tasks, cls=2):
    """Predicted cost (in level-0 pair integrations) of every task under adaptive error control: class weight x
    (5 + 66.7 * P(rho)).  Works on numpy arrays or on torch tensors (any device); returns the same kind."""
    is_torch = type(tasks).__module__.startswith("torch")     # (numpy >= 2 arrays also have a .device attribute)
    if is_torch:
        import torch
        v = torch.as_tensor(vertices, device=tasks.device, dtype=torch.float64)
        c = torch.as_tensor(cells, device=tasks.device).long()
        tri = v[c]                                              # [nc, 3, 3]
        cent = tri.mean(1)
        area = 0.5 * torch.linalg.norm(torch.linalg.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]), dim=1)
        i, j = tasks[:, 0].long(), tasks[:, 1].long()
        rho = torch.linalg.norm(cent[i] - cent[j], dim=1) / torch.sqrt(torch.maximum(area[i], area[j]))
        idx = torch.bucketize(rho, torch.tensor(_RHO_EDGES, device=tasks.device, dtype=torch.float64), right=True)
        p = torch.tensor(_P_UNCONVERGED, device=tasks.device, dtype=torch.float64)[idx]
        return CLASS_COST[cls] * (_BASE_COST + _TAIL_COST * p)
    import numpy as np
    tri = np.asarray(vertices)[np.asarray(cells)]
    cent = tri.mean(1)
    area = 0.5 * np.linalg.norm(np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]), axis=1)
    i, j = tasks[:, 0], tasks[:, 1]
    rho = np.linalg.norm(cent[i] - cent[j], axis=1) / np.sqrt(np.maximum(area[i], area[j]))
    p = np.asarray(_P_UNCONVERGED)[np.searchsorted(np.asarray(_RHO_EDGES), rho, side="right")]
    return CLASS_COST[cls] * (_BASE_COST + _TAIL_COST * p)


def cost_balanced_bounds(cost, world):
    """shard_bounds for a per-task cost vector that may live on the GPU (torch) or on the host (numpy)."""
    if type(cost).__module__.startswith("torch"):
        import torch
        n = int(cost.shape[0])
        c = torch.cumsum(cost.double(), 0)
        total = float(c[-1]) if n else 0.0
        targets = torch.tensor([total * r / world for r in range(1, world)], device=cost.device, dtype=torch.float64)
        cuts = [0] + [int(x) for x in torch.searchsorted(c, targets).tolist()] + [n]
        for r in range(1, world + 1):
            cuts[r] = max(cuts[r], cuts[r - 1])
        return [(cuts[r], cuts[r + 1]) for r in range(world)]
    return shard_bounds(len(cost), world, cost)


def forward_ranges(pairs: int, world: int):
    """Forward-slot ranges [lo, hi) of the `world` shards of a class with `pairs` unordered pairs: equal counts, bounds at
    multiples of 32 (the C side: i2_host_prepare / i2_mgpu_prepare).  Rank r owns those pairs in BOTH orders: its task list is
    [pairs lo..hi ; their reversed pairs], 2 (hi - lo) tasks."""
    cuts = [0] + [(pairs * r // world) & ~31 for r in range(1, world)] + [pairs]
    for r in range(1, world + 1):
        cuts[r] = max(cuts[r], cuts[r - 1])
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def exchange_rounds(ctx, level, check=False, group=None):
    """What i2_mgpu_run does between i2_host_run_rounds and i2_host_run_finalize, with torch.distributed as the communicator
    (any backend): under error control the shards agree on each class's last round and on the per-cell refinement counters
    (element-wise maximum) before the final assembly.  `ctx` has run host_prepare with its shard set."""
    import numpy as np
    import torch
    import torch.distributed as dist
    ctx.host_run_rounds(level)
    if level < 0 and dist.is_initialized() and dist.get_world_size(group) > 1:
        last = torch.tensor(ctx.host_last_rounds(), dtype=torch.int32)
        ref = torch.from_numpy(np.ascontiguousarray(ctx.host_refinements()))
        if dist.get_backend(group) == "nccl":
            dev = torch.device("cuda", torch.cuda.current_device())
            last, ref = last.to(dev), ref.to(dev)
        dist.all_reduce(last, op=dist.ReduceOp.MAX, group=group)
        dist.all_reduce(ref, op=dist.ReduceOp.MAX, group=group)
        ctx.host_last_rounds([int(x) for x in last.cpu().tolist()])
        ctx.host_refinements(ref.cpu().numpy())
    ctx.host_run_finalize(level, check)

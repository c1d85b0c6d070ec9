#!/usr/bin/env python
"""bench.py — headline benchmark of the integrator2 hot path on B200.

Metric (BASELINE.json): ordered triangle-pair integrals per second (each = one J(K_i,K_j): Theta + Psi -> Point3),
all three neighbour classes, on the reference's example mesh Vint16k.dat (configs[2]: 16 930 triangles,
286 607 970 ordered pairs, `-r 0`).  One "step" = one full pass of EvaluatorJ3DK::integrateOver{Simple,Attached,Not}
Neighbors over the mesh's task lists.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mesh NAME] [--level L]

  value     device-resident pass (task lists and mesh already in HBM), CUDA events, max over ranks
  e2e       the same pass through the host-buffer C ABI (i2_host_prepare + i2_host_run): host mesh in, pinned host
            results + task keys out, copies inside the timed region
  roofline  regular-pair kernel against the FP64-pipe peak measured on this device by i2_peak_rates
  cpu_baseline  the OpenMP CPU oracle (oracle/, kind "port") on a bounded sample of the same task list
  --impl reference  the reference's own CUDA build (oracle/_ref/integrator2test3D, unmodified sources) on the same
            GPU — the reference has no CPU implementation of this path; falls back to the CPU oracle
            port when that binary is absent.

N > 1 (torchrun): every class's task list is split into N equal-cost contiguous shards (tasks of one class cost the
same at a fixed level) and each rank integrates its shard; tasks are independent, so there is no data-path collective:
the per-pair results stay resident on the rank that computed them (a checksum all-reduce after the timed region validates them).
The export variant (all per-pair results gathered to rank 0 with NCCL, overlapped with compute) is timed separately and
reported as `with_gather_to_rank0`.  `--workload matrixfree` runs BASELINE.json configs[4] (108 544-triangle sphere).
"""
from __future__ import annotations

import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_REGULAR_PAIR = 6.1e3   # 13 points x ~450 FP64 flop + ~250 per pair
METRIC = "triangle-pair integrals/s (potential+gradient)"
UNIT = "pairs/s"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, reasons, mx = [], set(), None
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx = float(f[2])
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:  # noqa: BLE001
            pass
        if sm:
            sm.sort()
            # median over the samples taken under load (upper half: idle samples before/after are lower)
            out["sm_mhz"] = sm[len(sm) // 2]
        out["sm_max_mhz"] = mx
        out["reasons"] = sorted(reasons)
        return out


def mesh_dat_path(name, mesh):
    """A .dat file of the mesh for the reference CLI: the staged reference example if present, else re-written from the fixture."""
    p = os.path.join(ROOT, "oracle", "_ref", "examples", name + ".dat")
    if os.path.exists(p):
        return p
    from integrator2_b200.meshio import write_dat
    p = os.path.join(tempfile.gettempdir(), f"i2_{name}.dat")
    write_dat(p, mesh)
    return p


def run_reference(args, mesh, n_pairs_total):
    """Reference arm: the UNMODIFIED reference CUDA build on this GPU, timed by its own GpuTimer lines."""
    binary = os.path.join(ROOT, "oracle", "_ref", "integrator2test3D")
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = {"workload": workload_name(args.mesh, args.scale, args.level, class_pair_counts(mesh)), "triangles": mesh.n_cells,
           "quadrature": "Cowper 13-point (order 7)", "command": f"integrator2test3D -f {args.mesh}.dat -s {args.scale} -r {args.level}" if args.level >= 0
           else f"integrator2test3D -f {args.mesh}.dat -s {args.scale}", "l2": "inputs+outputs per step are far larger than L2; no flush needed"}
    if os.path.exists(binary):
        path = mesh_dat_path(args.mesh, mesh)
        times = []
        cmd = [binary, "-f", path, "-r", str(args.level)]
        if args.scale != 1.0:
            cmd += ["-s", repr(args.scale)]
        ok = True
        for it in range(args.warmup + args.steps):
            t0 = time.time()
            try:
                out = subprocess.run(cmd, capture_output=True, text=True, timeout=1200, cwd=tempfile.gettempdir()).stdout
            except Exception as e:  # noqa: BLE001
                out, ok = "", False
            ms = [float(x) for x in re.findall(r"Time for .* integration:\s*([0-9.]+) ms", out)]
            if len(ms) != 3:
                ok = False
                break
            if it >= args.warmup:
                times.append(sum(ms))
        if ok and times:
            ms_step = sum(times) / len(times)
            value = n_pairs_total / (ms_step * 1e-3)
            line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
                    "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                    "data": f"reference example mesh (tests/golden/meshes.npz, parsed from {args.mesh}.dat); no random data", "config": cfg,
                    "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "reference",
                                     "sample": "whole workload; the reference is CUDA-only (no CPU path): its unmodified sources compiled for sm_100, "
                                               "1 host thread + this B200, time window = its own three 'Time for ... integration' lines"},
                    "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
            print(json.dumps(line), flush=True)
            return
    # fallback: CPU oracle port on all host cores, bounded sample
    from oracle import oracle_py as O
    om = O.OracleMesh(mesh.vertices, mesh.cells)
    v, c, sample = cpu_oracle_rate_from_mesh(O, om, mesh, args.level, budget_s=15.0)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": n_pairs_total / v * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": f"reference example mesh (tests/golden/meshes.npz, parsed from {args.mesh}.dat); no random data", "config": cfg,
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": c, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mesh", default="Vint16k")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--level", type=int, default=0, help="fixed refinement level (-1 = adaptive error control)")
    ap.add_argument("--workload", default="lists", choices=["lists", "matrixfree"],
                    help="lists: the reference's task-list path (headline); matrixfree: list-free row sums on a refined sphere (configs[4])")
    ap.add_argument("--sphere-level", type=int, default=5, help="matrixfree: the base mesh is refined this many times by midpoint subdivision (G1: 5 -> 108 544 triangles)")
    ap.add_argument("--mf-mesh", default="G1", help="matrixfree: base mesh (G1 -> configs[4]; s5m2 with --scale 0.0005 --sphere-level 2 --level -1 -> configs[3], 125 280 triangles)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    import numpy as np
    from integrator2_b200.meshio import load_fixture
    if args.workload == "matrixfree":
        return run_matrix_free(args)
    mesh = load_fixture(args.mesh, args.scale)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        # pair counts without a GPU: from the mesh size and the two small classes
        run_reference(args, mesh, sum(class_pair_counts(mesh)))   # pair counts without a GPU: vertex / edge incidence
        return

    import torch
    import torch.distributed as dist
    from integrator2_b200 import abi

    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    dev = torch.device(f"cuda:{local}")
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)          # every torch op and every library kernel of this process runs on this stream
    ctx = abi.Context(local)

    # ---- resident inputs: mesh SoA + the three ordered task lists (built on the device) ----------------------
    ctx.set_mesh(mesh.vertices, mesh.cells)
    lists = ctx.classify()
    tasks_full = [ctx.tasks_from_pairs(p) for p in lists]
    del lists
    counts = [int(t.shape[0]) for t in tasks_full]
    total_pairs = sum(counts)
    # equal-cost contiguous shard of every class for this rank
    from integrator2_b200.multigpu import adaptive_task_cost, cost_balanced_bounds, shard_bounds
    if args.level < 0 and world > 1:
        # adaptive error control: shards of equal PREDICTED cost (class weight x expected refinement depth from the
        # centroid-distance / panel-size ratio), still contiguous so that all descendants of a task stay on its rank
        all_bounds = [cost_balanced_bounds(adaptive_task_cost(mesh.vertices, mesh.cells, t, cls), world) for cls, t in enumerate(tasks_full)]
    else:
        all_bounds = [shard_bounds(n, world) for n in counts]
    bounds = [b[rank] for b in all_bounds]
    tasks = [t[lo:hi].contiguous() for t, (lo, hi) in zip(tasks_full, bounds)]
    if world > 1:
        del tasks_full
    my_counts = [int(t.shape[0]) for t in tasks]
    outs = [(torch.empty((n, 4), dtype=torch.float64, device=dev), torch.empty((n, 3), dtype=torch.float64, device=dev)) for n in my_counts]
    gathered = None
    if world > 1 and rank == 0:
        gathered = [torch.empty((n, 3), dtype=torch.float64, device=dev) for n in counts]
    refin = torch.zeros((mesh.n_cells,), dtype=torch.uint8, device=dev) if args.level < 0 else None
    refins = [torch.zeros((mesh.n_cells,), dtype=torch.uint8, device=dev) for _ in range(3)] if args.level < 0 else None

    side = torch.cuda.Stream(device=dev) if world > 1 else None
    chk = torch.zeros((3, 4), dtype=torch.float64, device=dev)

    def step():
        # every rank integrates its shard of every class; the per-pair results stay resident on the rank that computed them.
        # Tasks are independent, so the step has no data-path collective; ranks meet at the barrier that brackets the timing.
        # One i2_integrate_all call = the three integrateOver* virtuals; the two adjacent classes overlap the regular one.
        if refins is not None:
            for r in refins:
                r.zero_()
        ctx.integrate_all(tasks, args.level, want_stats=False, refinements=refins, out=outs)

    def global_checksum():
        # validation outside the timed region: per-class sum |J|_1 over all ranks (NCCL all-reduce of 3 doubles)
        for cls in range(3):
            chk[cls, 3] = outs[cls][1].abs().sum()
        if world > 1:
            dist.all_reduce(chk)
        return [float(x) for x in chk[:, 3].tolist()]

    def step_with_gather():
        # export variant: finished chunks of per-pair results travel to rank 0 over NCCL while the next chunk computes
        from integrator2_b200.multigpu import integrate_and_gather, wait_all
        works = []
        for cls in (2, 0, 1):
            works += integrate_and_gather(ctx, cls, tasks[cls], args.level, outs[cls], gathered[cls] if rank == 0 else None,
                                          all_bounds[cls], rank, world, side, chunks=8 if cls == 2 else 1)
        wait_all(works, side)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    launches0 = abi.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(args.steps):
            step()
        e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = abi.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
        lt = torch.tensor([launches], dtype=torch.int64, device=dev)
        dist.all_reduce(lt)
        launches = int(lt.item())
    ms_step = ms_total / args.steps
    value = total_pairs / (ms_step * 1e-3)
    checksum = global_checksum()

    # ---- N > 1: the export variant (all per-pair results gathered to rank 0 over NVLink), timed separately ------------
    gather_info = None
    if world > 1:
        for _ in range(2):
            step_with_gather()
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record(stream)
        for _ in range(args.steps):
            step_with_gather()
        g1.record(stream)
        barrier()
        tg = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device=dev)
        dist.all_reduce(tg, op=dist.ReduceOp.MAX)
        ms_g = float(tg.item()) / args.steps
        gather_info = {"value": total_pairs / (ms_g * 1e-3), "unit": UNIT, "ms_per_step": ms_g,
                       "bytes_into_rank0_per_step": int(sum(counts) * 24 * (world - 1) / world),
                       "what": "same step, but every per-pair Point3 result is also gathered to rank 0 (NCCL point-to-point, 8 chunks "
                               "overlapped with compute); bound by rank 0's NVLink ingest, reported for the export use case"}

    # ---- N > 1: export variant with the gather FUSED into the integrate kernels: rank 0's export arrays are mapped into every
    # process (CUDA IPC) and each rank's kernels store their per-pair results straight into them over NVLink/NVSwitch
    peer_info = None
    if world > 1:
        from integrator2_b200.multigpu import PeerExport
        gathered = None
        torch.cuda.empty_cache()
        exports = [PeerExport(ctx, counts[k], all_bounds[k], rank, world) for k in range(3)]
        peer_out = [(outs[k][0], exports[k].results_arg()) for k in range(3)]

        def step_peer():
            if refins is not None:
                for r in refins:
                    r.zero_()
            ctx.integrate_all(tasks, args.level, want_stats=False, refinements=refins, out=peer_out)

        for _ in range(2):
            step_peer()
        barrier()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record(stream)
        for _ in range(args.steps):
            step_peer()
        p1.record(stream)
        barrier()
        tp = torch.tensor([p0.elapsed_time(p1)], dtype=torch.float64, device=dev)
        dist.all_reduce(tp, op=dist.ReduceOp.MAX)
        ms_p = float(tp.item()) / args.steps
        chk_peer = [float(e.full.abs().sum()) for e in exports] if rank == 0 else None
        peer_info = {"value": total_pairs / (ms_p * 1e-3), "unit": UNIT, "ms_per_step": ms_p,
                     "bytes_into_rank0_per_step": int(sum(counts) * 24 * (world - 1) / world),
                     "checksum_sum_abs_J_on_rank0": chk_peer,
                     "what": "same step with compute and gather fused: every rank's kernels store their per-pair Point3 results directly into "
                             "rank 0's export arrays over NVLink peer mappings (i2_peer_alloc / i2_peer_open, multigpu.PeerExport); no NCCL "
                             "call, no staging copy"}
        barrier()
        for e in exports:
            e.close()
        barrier()

    # ---- roofline of the dominant kernel (regular pairs), measured live with CUDA events on the launch stream ----
    roof = None
    if rank == 0:
        ctx.set_profiling(True)
        t_int = []
        with torch.cuda.stream(stream):
            for _ in range(3):
                ctx.integrate_class(2, tasks[2], args.level, want_stats=False, refinements=refin, out=outs[2])
                if args.level >= 0:
                    t_int.append(ctx.profile_last()[0])
        ctx.set_profiling(False)
        dfma_tf, mufu_g = ctx.peak_rates()
        dfma3_tf = ctx.peak_dfma_three_operand()
        if t_int:
            ms_k = sum(t_int) / len(t_int)
            flops = FLOP_PER_REGULAR_PAIR * my_counts[2] * (4 ** max(args.level, 0))
            achieved = flops / (ms_k * 1e-3) / 1e12
            # dram__bytes_read+write of this kernel from the ncu --set full capture in profiles/r01_ncu_k_regular_grouped_v7_final.txt:
            # 19.447 GB for 286 403 650 pairs = 67.90 B/pair (algorithmic: 12 B task + 32 B integrals + 24 B result = 68 B)
            traffic = 67.90 * my_counts[2] if args.level == 0 else None
            roof = {"bound": "fp64", "achieved": achieved, "peak": dfma_tf, "unit": "TFLOP/s", "frac": achieved / dfma_tf,
                    "traffic": traffic, "kernel": "k_regular_grouped (not-neighbours, level 0, fused assembly)", "kernel_ms": ms_k,
                    "algorithmic_flop_per_pair": FLOP_PER_REGULAR_PAIR, "pairs_per_launch": my_counts[2],
                    "peak_source": "measured on this device by i2_peak_rates (DFMA chains); MEASURED_PEAKS.json has no FP64 figure",
                    "mufu_peak_gops": mufu_g,
                    "peak_three_register_operands": dfma3_tf,
                    "note": "achieved uses the per-point work model of SURVEY.md 8(d) (6.1 kflop/pair); the grouped kernel executes ~1130 FP64 "
                            "instructions (~1.8 kflop) per pair, i.e. frac > 1 means work removed, not a faster pipe; peak_three_register_operands is the DFMA rate when every "
                            "instruction reads three distinct registers (the practical ceiling of real code, ~75 % of peak); FP64-pipe active 63 % (84 % of the three-operand ceiling) in "
                            "profiles/r01_ncu_k_regular_grouped_v7_final.txt",
                    "executed_fp64_inst_per_pair": 1130,
                    "fp64_pipe_active_frac": 0.629,
                    # 55 MUFU (RSQ64H / RCP64H) warp instructions per pair in the same capture; the XU pipe is not a limiter
                    "executed_mufu_per_pair": 55,
                    "mufu_frac": (55.0 * my_counts[2] / (ms_k * 1e-3) / 1e9) / mufu_g if (args.level == 0 and mufu_g > 0) else None,
                    "issue_slots_active_frac": 0.572,
                    "dispatch_bound_frac": 0.89}

    # ---- end to end through the host-buffer C ABI (N=1): host mesh in -> prepare (H2D, geometry, classification, task
    # lists) -> three classes -> results.  Two variants, both timed with the host clock around the blocking calls:
    #   e2e.value            results stay resident in HBM (exactly what Evaluator3D::runAllPairs leaves behind) and the
    #                        per-class checksums (96 B) are read back as the step's metric;
    #   e2e.full_d2h         additionally every per-pair result and its (i,j) key is copied to pinned host memory
    #                        (what outputResultsToFile does before formatting): 36 B/pair, PCIe-bound.
    e2e = None
    if not args.no_e2e:
        del outs, tasks
        if world == 1:
            del tasks_full
        torch.cuda.empty_cache()
        c2 = abi.Context(local)
        c2.host_set_shard(rank, world)     # N > 1: every rank prepares the (replicated, 0.4 MB) mesh and integrates its shard
        cnt_full = c2.host_prepare(mesh.vertices, mesh.cells)
        cnt = c2.host_shard()[1]
        reps = max(2, min(args.steps, 5))

        def timed(fn):
            # host clock around the blocking calls; N > 1: ranks start together and the slowest rank's time counts
            for _ in range(2):
                fn()
            barrier()
            t0 = time.perf_counter()
            for _ in range(reps):
                fn()
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / reps
            if world > 1:
                tt = torch.tensor([dt], dtype=torch.float64, device=dev)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                dt = float(tt.item())
            return dt

        sums = []

        def resident():
            c2.host_prepare(mesh.vertices, mesh.cells)
            c2.host_run(args.level, None, None)
            sums.append(c2.host_checksums())

        dt_res = timed(resident)
        if world > 1:      # per-class checksums of the shards -> whole-job checksums (validation, outside the timed region)
            st = torch.as_tensor(sums[-1], device=dev)
            dist.all_reduce(st)
            sums.append(st.cpu().numpy())
        ht = [torch.empty((n, 3), dtype=torch.int32, pin_memory=True) for n in cnt]
        hr = [torch.empty((n, 3), dtype=torch.float64, pin_memory=True) for n in cnt]

        def full():
            c2.host_prepare(mesh.vertices, mesh.cells)
            c2.host_run(args.level, ht, hr)

        dt_full = timed(full)
        d2h = int(sum(cnt_full) * (24 + 12))
        e2e = {"value": sum(cnt_full) / dt_res, "unit": UNIT, "h2d_bytes_per_step": int(mesh.vertices.nbytes + mesh.cells.nbytes) * world,
               "d2h_bytes_per_step": 96 * world, "ms_per_step": dt_res * 1e3,
               "what": "i2_host_prepare (H2D mesh, geometry, classification, ordered task lists) + i2_host_run (3 classes) with the per-pair "
                       "results left in HBM like Evaluator3D::runAllPairs does, + D2H of the per-class checksums",
               "checksum_sum_abs_J": [float(x) for x in sums[-1][:, 3]],
               "full_d2h": {"value": sum(cnt_full) / dt_full, "unit": UNIT, "ms_per_step": dt_full * 1e3, "d2h_bytes_per_step": d2h,
                            "d2h_gb_per_s": d2h / dt_full / 1e9,
                            "what": "same, plus every per-pair result (24 B) and (i,j,k) key (12 B) copied to pinned host memory, chunks overlapped with compute"}}
        c2.close()

    cpu = None
    if rank == 0 and not args.no_cpu:
        from oracle import oracle_py as O
        om = O.OracleMesh(mesh.vertices, mesh.cells)
        om._pairs = None
        # task list for the sample comes from the device classification already validated against the oracle
        v, cores, sample = cpu_oracle_rate_from_mesh(O, om, mesh, args.level)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": f"reference example mesh (tests/golden/meshes.npz, parsed from {args.mesh}.dat); no random data",
                "config": {"workload": workload_name(args.mesh, args.scale, args.level, counts),
                           "triangles": mesh.n_cells, "quadrature": "Cowper 13-point (order 7)", "sharding": f"{world} contiguous equal-cost shards per class, results resident per rank (no data-path collective)",
                           "l2": "inputs+outputs per step (task lists 12 B/pair, results 56 B/pair) are far larger than L2; no flush needed"},
                "clocks": clocks, "gpu_launches": launches, "e2e": e2e, "roofline": roof, "cpu_baseline": cpu,
                "with_gather_to_rank0": gather_info, "with_peer_store_to_rank0": peer_info, "checksum_sum_abs_J": checksum}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def class_pair_counts(mesh):
    """(vertex-adjacent, edge-adjacent, regular) ORDERED pair counts from vertex / edge incidence, without an O(N^2) pass and
    without a GPU: a pair sharing an edge is counted twice in sum_v k_v (k_v - 1), a pair sharing one vertex once."""
    import numpy as np
    n = mesh.n_cells
    c = mesh.cells.astype(np.int64)
    kv = np.bincount(c.ravel())
    share_vertex = int((kv * (kv - 1)).sum())
    e = np.sort(np.concatenate([c[:, [0, 1]], c[:, [1, 2]], c[:, [2, 0]]]), axis=1)
    _, cnt = np.unique(e[:, 0] * (c.max() + 1) + e[:, 1], return_counts=True)
    share_edge = int((cnt * (cnt - 1)).sum())
    va = share_vertex - 2 * share_edge
    return va, share_edge, n * (n - 1) - va - share_edge


def workload_name(mesh_name, scale, level, counts):
    """config.workload: the same string for our arm and for the reference arm"""
    return (f"{mesh_name}.dat scale {scale} level {'adaptive' if level < 0 else level}: "
            f"{counts[0]} vertex-adjacent + {counts[1]} edge-adjacent + {counts[2]} regular = {sum(counts)} ordered pairs")


def regular_pair_count(mesh):
    """ordered pairs that share no vertex (closed or open surface, any valence), without an O(N^2) pass."""
    import numpy as np
    n = mesh.n_cells
    c = mesh.cells.astype(np.int64)
    kv = np.bincount(c.ravel())
    share_vertex = int((kv * (kv - 1)).sum())                      # ordered pairs counted once per shared vertex
    e = np.sort(np.concatenate([c[:, [0, 1]], c[:, [1, 2]], c[:, [2, 0]]]), axis=1)
    _, cnt = np.unique(e[:, 0] * (c.max() + 1) + e[:, 1], return_counts=True)
    share_edge = int((cnt * (cnt - 1)).sum())                      # ordered pairs sharing an edge: counted twice above
    return n * (n - 1) - (share_vertex - share_edge)


def run_matrix_free(args):
    """BASELINE.json configs[4]: G1 sphere refined 5 times (106 * 4^5 = 108 544 triangles, 1.18e10 ordered regular pairs),
    rows sharded over the GPUs, list-free kernel (i2_apply_regular), row sums gathered to rank 0 with NCCL."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from integrator2_b200 import abi
    from integrator2_b200.meshio import load_fixture, subdivide
    from integrator2_b200.multigpu import gather_results, shard_bounds
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    mesh = subdivide(load_fixture(args.mf_mesh, args.scale), args.sphere_level)
    adaptive = args.level < 0
    n = mesh.n_cells
    pairs = regular_pair_count(mesh)
    ctx = abi.Context(local)
    ctx.set_mesh(mesh.vertices, mesh.cells)
    bounds = shard_bounds(n, world)
    lo, hi = bounds[rank]
    out = torch.empty((hi - lo, 3), dtype=torch.float64, device=dev)
    full = torch.empty((n, 3), dtype=torch.float64, device=dev) if rank == 0 else None

    last = {}

    def step():
        if adaptive:
            # Runge loop per pair inside one kernel; the value of a converged pair depends on the parity of the class's LAST
            # round (the reference's ping-pong buffers), which is global: all-reduce max of the ranks' last rounds (one int)
            a = ctx.apply_regular_adaptive(lo, hi)
            L = torch.tensor([a["stats"]["last_round"]], dtype=torch.int32, device=dev)
            if world > 1:
                dist.all_reduce(L, op=dist.ReduceOp.MAX)
            same = (int(L.item()) & 1) == (a["stats"]["last_round"] & 1)
            out.copy_(a["out"] if same else a["other"])
            last.update(a["stats"], global_last_round=int(L.item()), refinements=a["refinements"])
        else:
            ctx.apply_regular(lo, hi, None, out)
        if world > 1:
            gather_results(out, full, bounds, rank, world)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 1)):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    l0 = abi.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = abi.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        lt = torch.tensor([launches], dtype=torch.int64, device=dev)
        dist.all_reduce(lt)
        launches = int(lt.item())
    ms_step = ms / args.steps
    if rank == 0:
        dfma_tf, _ = ctx.peak_rates()
        achieved = FLOP_PER_REGULAR_PAIR * pairs / (ms_step * 1e-3) / 1e12
        chk = float((full if world > 1 else out).abs().sum())
        if adaptive:
            hist = torch.bincount(last["refinements"].int()).tolist()
            print(json.dumps({"metric": METRIC, "value": pairs / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                              "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                              "data": f"synthetic: {args.mf_mesh} (scale {args.scale}) refined {args.sphere_level}x by midpoint subdivision (deterministic, no RNG)",
                              "config": {"workload": f"matrix-free regular class under automatic error control (Runge rule, <= 5 rounds per pair): {n} triangles, "
                                                     f"{pairs} ordered regular pairs, row sums sum_j J(K_i,K_j); no task list, no refined mesh, no per-pair output",
                                         "sharding": f"{world} contiguous row blocks", "l2": "mesh SoA is L2-resident by design; no per-pair HBM traffic"},
                              "clocks": clocks, "gpu_launches": launches, "e2e": None, "roofline": None, "cpu_baseline": None,
                              "rank0_rounds": {k: last[k] for k in ("last_round", "global_last_round", "integrated", "unconverged")},
                              "rank0_refinement_histogram": hist, "checksum_sum_abs": chk}), flush=True)
            if world > 1:
                dist.destroy_process_group()
            return
        print(json.dumps({"metric": METRIC, "value": pairs / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                          "data": "synthetic: G1 sphere refined by midpoint subdivision (deterministic, no RNG)",
                          "config": {"workload": f"matrix-free regular class, G1 sphere refined {args.sphere_level}x: {n} triangles, {pairs} ordered regular pairs "
                                                 "(row sums sum_j J(K_i,K_j); no task list, no per-pair output)",
                                     "sharding": f"{world} contiguous row blocks", "l2": "mesh SoA (24 MB) is L2-resident by design; no per-pair HBM traffic"},
                          "clocks": clocks, "gpu_launches": launches, "e2e": None,
                          "roofline": {"bound": "fp64", "achieved": achieved / world, "peak": dfma_tf, "unit": "TFLOP/s", "frac": achieved / world / dfma_tf,
                                       "traffic": None, "kernel": "k_apply_regular", "note": "per GPU; same work model as the list kernel"},
                          "cpu_baseline": None, "checksum_sum_abs": chk}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def reference_total_pairs(mesh):
    """ordered pairs of all classes = N(N-1) minus pairs sharing 3 vertices (none in the example meshes)."""
    n = mesh.n_cells
    return n * (n - 1)


def cpu_oracle_rate_from_mesh(O, om, mesh, level, budget_s=15.0):
    """Bounded CPU sample without the O(N^2) host classification: all regular partners j of every s-th control panel i
    (rows of the pair matrix), s chosen so that the OpenMP oracle works for about budget_s seconds."""
    import numpy as np
    n = mesh.n_cells

    def rows_tasks(rows):
        I = np.repeat(rows, n)
        J = np.tile(np.arange(n), rows.size)
        ci, cj = mesh.cells[I], mesh.cells[J]
        shared = (ci[:, :, None] == cj[:, None, :]).any(axis=(1, 2))
        keep = ~shared
        return np.ascontiguousarray(np.stack([I[keep], J[keep], np.arange(int(keep.sum()))], axis=1).astype(np.int32))

    nrows, dt, sample, rows = 16, 0.0, None, None
    for _ in range(4):   # grow the sample until it costs about budget_s (first calls also warm up the OpenMP team)
        rows = np.unique(np.linspace(0, n - 1, nrows).astype(np.int64))
        sample = rows_tasks(rows)
        t0 = time.time()
        om.run_class(2, sample, level)
        dt = max(time.time() - t0, 1e-3)
        if dt >= 0.6 * budget_s or rows.size >= n:
            break
        nrows = int(min(n, max(nrows + 1, nrows * budget_s / dt)))
    return sample.shape[0] / dt, O.num_threads(), (f"{sample.shape[0]} regular pairs of the same mesh (all partners j of {rows.size} evenly spaced "
                                                   f"control panels i), OpenMP oracle, {dt:.1f} s")


if __name__ == "__main__":
    main()

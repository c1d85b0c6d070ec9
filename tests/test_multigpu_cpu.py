"""Host-side logic of the multi-GPU path on CPU: shard bounds and the variable-length gather (gloo, world_size 2)."""
import os
import subprocess
import sys
import textwrap

import numpy as np

from conftest import ROOT
from integrator2_b200.multigpu import predicted_task_cost, shard_bounds


def test_shard_bounds_cover_everything_once():
    for n in (0, 1, 7, 898, 286403650):
        for w in (1, 2, 3, 4, 8):
            b = shard_bounds(n, w)
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[r][1] == b[r + 1][0] for r in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def test_cost_balanced_bounds():
    rng = np.random.default_rng(0)
    depth = rng.integers(1, 6, size=10000)
    cost = np.array([predicted_task_cost(2, 0, d) for d in depth])
    b = shard_bounds(cost.size, 8, cost)
    per = np.array([cost[lo:hi].sum() for lo, hi in b])
    assert per.max() / per.mean() < 1.02
    assert b[0][0] == 0 and b[-1][1] == cost.size


def test_gather_two_ranks_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys
        sys.path.insert(0, {ROOT!r})
        import torch, torch.distributed as dist
        from integrator2_b200.multigpu import shard_bounds, gather_results
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        n = 1001
        ref = torch.arange(n * 3, dtype=torch.float64).reshape(n, 3) * 0.5      # stand-in for per-pair results
        b = shard_bounds(n, world)
        lo, hi = b[rank]
        local = ref[lo:hi].clone()
        full = torch.zeros_like(ref) if rank == 0 else None
        gather_results(local, full, b, rank, world)
        if rank == 0:
            assert torch.equal(full, ref)
            print("GATHER_OK")
        dist.destroy_process_group()
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29611")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29611", str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert "GATHER_OK" in out.stdout, out.stdout + out.stderr


def test_adaptive_cost_model_orders_near_before_far():
    """Predicted adaptive cost: near pairs (small centroid distance / panel size) cost several times a far pair; numpy and
    torch paths agree; cost-balanced bounds cover the list."""
    import torch
    from integrator2_b200.meshio import load_fixture
    from integrator2_b200.multigpu import adaptive_task_cost, cost_balanced_bounds
    from oracle import oracle_py as O
    m = load_fixture("s5m", 0.0005)
    om = O.OracleMesh(m.vertices, m.cells)
    t = om.tasks(2)[::50]
    c_np = adaptive_task_cost(m.vertices, m.cells, t)
    c_t = adaptive_task_cost(m.vertices, m.cells, torch.as_tensor(t)).numpy()
    assert np.allclose(c_np, c_t)
    assert c_np.min() == 5.0 and c_np.max() > 40.0
    b = cost_balanced_bounds(torch.as_tensor(c_np), 4)
    assert b[0][0] == 0 and b[-1][1] == t.shape[0] and all(b[r][1] == b[r + 1][0] for r in range(3))
    per = np.array([c_np[lo:hi].sum() for lo, hi in b])
    assert per.max() / per.mean() < 1.02
    assert b == shard_bounds(t.shape[0], 4, c_np) or abs(b[1][0] - shard_bounds(t.shape[0], 4, c_np)[1][0]) <= 1


def test_peer_export_bookkeeping_without_a_gpu():
    """multigpu.PeerExport: who allocates, who maps, which address a rank hands to the kernels, who frees / unmaps
    (the CUDA side — i2_peer_alloc / i2_peer_open and the stores over NVLink — is covered by tests/test_gpu_peer_export.py)."""
    from integrator2_b200.multigpu import PeerExport

    class FakeCtx:
        def __init__(self):
            self.calls = []

        def peer_alloc(self, rows, cols=3):
            self.calls.append(("alloc", rows))
            return "FULL", b"h" * 64, 1 << 20

        def peer_open(self, handle):
            assert handle == b"h" * 64
            self.calls.append(("open",))
            return 1 << 30

        def peer_close(self, addr):
            self.calls.append(("close", addr))

        def peer_free(self, addr):
            self.calls.append(("free", addr))

    bounds = shard_bounds(1001, 3)
    box = {}

    def exchange(handle):          # stands in for dist.broadcast_object_list
        if handle is not None:
            box["h"] = handle
        return box["h"]

    owner, writer = FakeCtx(), FakeCtx()
    e0 = PeerExport(owner, 1001, bounds, 0, 3, handle_exchange=exchange)
    e2 = PeerExport(writer, 1001, bounds, 2, 3, handle_exchange=exchange)
    assert e0.full == "FULL" and e0.results_arg() == (1 << 20) + 24 * bounds[0][0]
    assert e2.full is None and e2.results_arg() == (1 << 30) + 24 * bounds[2][0]
    e0.close(); e2.close(); e2.close()
    assert owner.calls == [("alloc", 1001), ("free", 1 << 20)]
    assert writer.calls == [("open",), ("close", 1 << 30)]
    single = FakeCtx()
    e = PeerExport(single, 7, [(0, 7)], 0, 1)
    assert e.results_arg() == 1 << 20 and single.calls == [("alloc", 7)]


def test_forward_ranges_tile_and_align():
    from integrator2_b200.multigpu import forward_ranges
    for pairs in (0, 5, 31, 32, 449, 143201825):
        for world in (1, 2, 3, 8):
            r = forward_ranges(pairs, world)
            assert r[0][0] == 0 and r[-1][1] == pairs and len(r) == world
            for (lo, hi), (lo2, _) in zip(r, r[1:]):
                assert hi == lo2 and lo <= hi
            assert all(lo % 32 == 0 for lo, _ in r)
            if pairs >= 64 * world:      # balanced to within one warp group
                sizes = [hi - lo for lo, hi in r]
                assert max(sizes) - min(sizes) <= 64


def _exchange_worker(rank, world, port, q):
    import torch.distributed as dist
    from integrator2_b200.multigpu import exchange_rounds
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)

    class FakeShard:
        """records what exchange_rounds does to a context (no GPU here): local state in, agreed state out"""
        def __init__(self):
            self.nc = 4
            self.last = [1, 3, 2] if rank == 0 else [4, 2, 2]
            self.ref = np.array([[1, 2, 1, 0], [0, 0, 3, 1], [2, 2, 2, 2]], dtype=np.uint8) if rank == 0 else \
                np.array([[2, 1, 1, 0], [1, 0, 1, 1], [2, 5, 2, 0]], dtype=np.uint8)
            self.calls = []

        def host_run_rounds(self, level):
            self.calls.append(("rounds", level))

        def host_last_rounds(self, set_to=None):
            if set_to is not None:
                self.last = list(set_to)
            return list(self.last)

        def host_refinements(self, set_to=None):
            if set_to is not None:
                self.ref = np.array(set_to, dtype=np.uint8)
            return self.ref

        def host_run_finalize(self, level, check=False):
            self.calls.append(("finalize", level, check))

    c = FakeShard()
    exchange_rounds(c, -1, check=True)
    q.put((rank, c.last, c.ref.tolist(), c.calls))
    c2 = FakeShard()
    exchange_rounds(c2, 0)            # fixed level: nothing to agree on
    q.put((rank + 10, c2.last, c2.ref.tolist(), c2.calls))
    dist.destroy_process_group()


def test_exchange_rounds_world_size_2_gloo():
    """The exchange step of a sharded run under error control (what i2_mgpu_run does with NCCL), with gloo and two processes:
    every shard ends with the MAXIMUM of the last rounds and of the per-cell refinement counters, then finalizes."""
    import socket
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_exchange_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(4):
        k, last, ref, calls = q.get(timeout=120)
        got[k] = (last, ref, calls)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in (0, 1):
        assert got[r][0] == [4, 3, 2]
        assert got[r][1] == [[2, 2, 1, 0], [1, 0, 3, 1], [2, 5, 2, 2]]
        assert got[r][2] == [("rounds", -1), ("finalize", -1, True)]
    assert got[10][0] == [1, 3, 2] and got[11][0] == [4, 2, 2]          # untouched at a fixed level
    assert got[10][2] == [("rounds", 0), ("finalize", 0, False)]

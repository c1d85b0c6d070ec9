"""Multi-GPU layer of the C ABI (i2_mgpu_*, include/i2_abi.h) on the GPU: world = 1 always; two GPUs driven by one process
(i2_mgpu_create_local) and two processes with one GPU each (i2_mgpu_create_rank, NCCL id through a file) when the box has
two GPUs (`gpurun --gpus 2`, tools/gpu_mgpu2.sh keeps the log under profiles/).  The bar is BITWISE equality with the
single-GPU run: tasks, results, (i,j)/(j,i) defects, refinement counters, per-round counts."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from conftest import ROOT
from integrator2_b200.meshio import load_fixture

pytestmark = pytest.mark.gpu


def _whole(abi, m, level, check=True):
    """single-context reference run through the host-buffer entry points -> dict per class"""
    import torch
    c = abi.Context(0)
    counts = c.host_prepare(m.vertices, m.cells)
    ht = [torch.empty((n, 3), dtype=torch.int32, pin_memory=True) for n in counts]
    hr = [torch.empty((n, 3), dtype=torch.float64, pin_memory=True) for n in counts]
    he = [torch.empty((n,), dtype=torch.float64, pin_memory=True) for n in counts]
    href = [torch.zeros((m.n_cells,), dtype=torch.uint8, pin_memory=True) for _ in range(3)]
    stats = c.host_run(level, ht, hr, he if check else None, href if level < 0 else None)
    sums = c.host_checksums()
    c.close()
    return dict(counts=counts, tasks=[t.numpy() for t in ht], results=[t.numpy() for t in hr], errors=[t.numpy() for t in he],
                refinements=[t.numpy() for t in href], stats=stats, sums=sums)


def _compare_shards(mg, whole, level, fetch):
    """fetch(rank, cls) -> dict(tasks, results, errors) of that rank's shard"""
    counts = whole["counts"]
    for cls in range(3):
        P, nxt = counts[cls] // 2, 0
        for rank in range(mg.world):
            first, cnt = mg.shard(rank)
            assert first[cls] == nxt and first[cls] % 32 == 0
            h, lo = cnt[cls] // 2, first[cls]
            nxt += h
            f = fetch(rank, cls)
            if f is None:
                continue
            for name in ("tasks", "results", "errors"):
                assert np.array_equal(f[name][:h], whole[name][cls][lo:lo + h]), (level, rank, cls, name)
                assert np.array_equal(f[name][h:], whole[name][cls][P + lo:P + lo + h]), (level, rank, cls, name, "reversed")
        assert nxt == P


@pytest.mark.parametrize("level", [0, 1, -1])
def test_world_of_one_equals_the_single_context_run(level):
    from integrator2_b200 import abi
    m = load_fixture("s5m", 0.0005)
    whole = _whole(abi, m, level)
    mg = abi.MultiGpu(local_gpus=1)
    assert mg.prepare(m.vertices, m.cells, level) == whole["counts"]
    stats = mg.run(level, check=True, want_stats=True)
    _compare_shards(mg, whole, level, lambda rank, cls: mg.fetch(rank, cls, errors=True))
    for cls in range(3):
        assert stats[cls]["last_round"] == whole["stats"][cls]["last_round"]
        assert stats[cls]["unconverged"] == whole["stats"][cls]["unconverged"]
        assert stats[cls]["integrated"] == whole["stats"][cls]["integrated"]
        if level < 0:
            assert np.array_equal(mg.refinements(cls), whole["refinements"][cls])
    _same_sums(mg.checksums(), whole["sums"])
    mg.close()


def _same_sums(a, b):
    """per-class (sum J_x, J_y, J_z, sum |J|_1): the signed sums cancel and are accumulated in a different order (atomics):
    equal to rounding at the scale of sum |J|_1"""
    assert (np.abs(a - b) <= 1e-11 * b[:, 3:4]).all(), (a, b)


def _gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("level", [0, 2, -1])
def test_two_gpus_one_process_bitwise(level):
    """i2_mgpu_create_local(2): sharded prepare, NCCL all-reduce of the last rounds / refinement counters under error
    control, row-striped fetch, NCCL gather to GPU 0 — all equal to the single-GPU run bit for bit."""
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2; log in profiles/r02_mgpu_2gpu.txt)")
    import torch
    from integrator2_b200 import abi
    m = load_fixture("s5m", 0.0005)
    whole = _whole(abi, m, level)
    mg = abi.MultiGpu(local_gpus=2)
    assert mg.prepare(m.vertices, m.cells, level) == whole["counts"]
    stats = mg.run(level, check=True, want_stats=True)
    _compare_shards(mg, whole, level, lambda rank, cls: mg.fetch(rank, cls, errors=True))
    for cls in range(3):
        assert stats[cls]["last_round"] == whole["stats"][cls]["last_round"]
        assert stats[cls]["unconverged"] == whole["stats"][cls]["unconverged"]
        if level < 0:
            assert np.array_equal(mg.refinements(cls), whole["refinements"][cls])
    _same_sums(mg.checksums(), whole["sums"])
    # export gather: shards concatenated in rank order on GPU 0, keys and values
    for cls in range(3):
        n = whole["counts"][cls]
        dst = torch.empty((n, 3), dtype=torch.float64, device="cuda:0")
        keys = torch.empty((n, 3), dtype=torch.int32, device="cuda:0")
        mg.gather(cls, 0, 0, dst)
        mg.gather(cls, 1, 0, keys)
        mg.synchronize()
        order = np.lexsort((keys.cpu().numpy()[:, 2],))
        assert np.array_equal(keys.cpu().numpy()[order], whole["tasks"][cls][np.argsort(whole["tasks"][cls][:, 2])])
        assert np.array_equal(dst.cpu().numpy()[order], whole["results"][cls][np.argsort(whole["tasks"][cls][:, 2])])
    mg.close()


WORKER = r'''
import os, sys, time
import numpy as np
sys.path.insert(0, sys.argv[1])
rank, world, level, idfile, out = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5], sys.argv[6]
import torch
from integrator2_b200 import abi
from integrator2_b200.meshio import load_fixture
torch.cuda.set_device(rank)
if rank == 0:
    uid = abi.MultiGpu.unique_id()
    open(idfile + ".tmp", "wb").write(uid)
    os.rename(idfile + ".tmp", idfile)
else:
    for _ in range(600):
        if os.path.exists(idfile):
            break
        time.sleep(0.1)
    uid = open(idfile, "rb").read()
m = load_fixture("s5m", 0.0005)
mg = abi.MultiGpu(device=rank, rank=rank, world=world, uid=uid)
counts = mg.prepare(m.vertices, m.cells, level)
stats = mg.run(level, check=True, want_stats=True)
sums = mg.checksums()
res = {"counts": np.array(counts), "sums": sums, "first": np.array(mg.shard(rank)[0]), "cnt": np.array(mg.shard(rank)[1])}
for cls in range(3):
    f = mg.fetch(0, cls, errors=True)
    for k, v in f.items():
        res[f"{k}{cls}"] = v
    res[f"last{cls}"] = np.array(stats[cls]["last_round"])
    res[f"unconv{cls}"] = np.array(stats[cls]["unconverged"])
    if level < 0:
        res[f"ref{cls}"] = mg.refinements(cls)
# operator apply by row blocks
cuts = mg.apply_prepare(m.vertices, m.cells, level if level <= 0 else 0)
vec, st = mg.apply(level if level <= 0 else 0, want_out=True, want_stats=True)
res["apply"] = vec
res["cuts"] = np.array(cuts)
np.savez(out, **res)
mg.close()
'''


@pytest.mark.parametrize("level", [0, -1])
def test_two_processes_one_gpu_each_bitwise(level, tmp_path):
    """i2_mgpu_create_rank: two processes, NCCL unique id shipped through a file (any transport works)."""
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2; log in profiles/r02_mgpu_2gpu.txt)")
    from integrator2_b200 import abi
    m = load_fixture("s5m", 0.0005)
    whole = _whole(abi, m, level)
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    idfile = str(tmp_path / "nccl_id")
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(r), "2", str(level), idfile, str(tmp_path / f"out{r}.npz")],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    for p in procs:
        out, _ = p.communicate(timeout=600)
        assert p.returncode == 0, out[-3000:]
    outs = [np.load(tmp_path / f"out{r}.npz") for r in range(2)]
    nxt = [0, 0, 0]
    for r, o in enumerate(outs):
        assert o["counts"].tolist() == whole["counts"]
        _same_sums(o["sums"], whole["sums"])     # all-reduced: every rank holds the job's checksums
        for cls in range(3):
            P, lo, h = whole["counts"][cls] // 2, int(o["first"][cls]), int(o["cnt"][cls]) // 2
            assert lo == nxt[cls]
            nxt[cls] += h
            for name in ("tasks", "results", "errors"):
                assert np.array_equal(o[f"{name}{cls}"][:h], whole[name][cls][lo:lo + h]), (r, cls, name)
                assert np.array_equal(o[f"{name}{cls}"][h:], whole[name][cls][P + lo:P + lo + h]), (r, cls, name)
            assert int(o[f"last{cls}"]) == whole["stats"][cls]["last_round"]
            assert o[f"unconv{cls}"].tolist() == whole["stats"][cls]["unconverged"]
            if level < 0:
                assert np.array_equal(o[f"ref{cls}"], whole["refinements"][cls])
    assert nxt == [c // 2 for c in whole["counts"]]
    # operator apply: both ranks hold the same full vector, equal to the single-GPU apply bit for bit
    c = abi.Context(0)
    c.set_mesh(m.vertices, m.cells)
    c.apply_prepare(0, m.n_cells)
    single = c.apply(level)["out"].cpu().numpy()
    c.close()
    assert np.array_equal(outs[0]["apply"], outs[1]["apply"])
    assert np.array_equal(outs[0]["apply"], single)

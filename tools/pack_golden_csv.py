"""Packs the csv files written by the reference CLI on the GPU box (tools/gpu_golden_csv.sh -> gpurun_out/golden_csv/) into
tests/golden/G1_r0_csv.npz (bytes of each file)."""
import os
import sys

import numpy as np

src = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden_csv"
out = {}
for name in ("SimpleNeighbors", "AttachedNeighbors", "NotNeighbors"):
    out[name] = np.frombuffer(open(os.path.join(src, name + ".csv"), "rb").read(), dtype=np.uint8)
np.savez_compressed("tests/golden/G1_r0_csv.npz", **out)
print({k: v.size for k, v in out.items()})

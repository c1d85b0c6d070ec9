"""Summary of one kernel of an `ncu --set full` capture, for profiles/ and for bench.py (which READS the counters of the
regular-pair kernel from the newest profiles/r02_ncu_k_regular_grouped*.json instead of carrying literals).

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > /tmp/raw.csv
    python tools/ncu_summary.py /tmp/raw.csv --kernel k_regular_grouped --pairs 286403650 --out profiles/r02_ncu_k_regular_grouped_v8

writes <out>.json (machine-readable) and <out>.txt (the same numbers, readable).  `--pairs` = work units of the profiled launch
(ordered pairs for the list kernel; panel evaluations for the list-free kernels)."""
import argparse
import csv
import json
import re


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("raw_csv")
    ap.add_argument("--kernel", required=True, help="regex on the kernel name")
    ap.add_argument("--pairs", type=float, required=True)
    ap.add_argument("--out", required=True)
    ap.add_argument("--note", default="")
    ap.add_argument("--build-id", default="")
    a = ap.parse_args()
    rows = list(csv.reader(open(a.raw_csv)))
    hdr, units = rows[0], rows[1]
    data = [r for r in rows[2:] if len(r) == len(hdr) and re.search(a.kernel, r[hdr.index("Kernel Name")])]
    if not data:
        raise SystemExit(f"no launch matching {a.kernel}")
    r = data[-1]

    def get(name, default=None):
        if name not in hdr:
            return default
        v = r[hdr.index(name)].replace(",", "")
        try:
            x = float(v)
        except ValueError:
            return default
        u = units[hdr.index(name)]
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(u)
        return x * scale if scale is not None else x

    cycles = get("sm__cycles_elapsed.avg")
    fp64_thread = sum(get(f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum.per_cycle_elapsed", 0.0) for op in ("dfma", "dmul", "dadd")) * cycles
    warp_inst = get("smsp__inst_executed.sum")
    xu_pct = get("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", 0.0)
    stalls = {k.split("issue_stalled_")[1].split("_per_issue")[0]: float(r[hdr.index(k)])
              for k in hdr if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")}
    total_stall = sum(stalls.values()) or 1.0
    fp64_per_pair = fp64_thread / a.pairs
    warp_inst_per_pair = warp_inst * 32.0 / a.pairs        # warp instructions per 32 pairs = thread-level instructions per pair
    out = {
        "kernel": r[hdr.index("Kernel Name")],
        "grid": r[hdr.index("Grid Size")], "block": r[hdr.index("Block Size")],
        "duration_ms_under_ncu": get("gpu__time_duration.sum"),
        "registers": get("launch__registers_per_thread"),
        "pairs_per_launch": a.pairs,
        "fp64_inst_per_pair": fp64_per_pair,
        "all_inst_per_pair": warp_inst_per_pair,
        "other_warp_inst_per_pair": warp_inst_per_pair - fp64_per_pair,
        # the XU pipe executes a warp instruction in 2 passes of 16 lanes per sub-partition: instructions = pct * cycles * 4 smsp / 2 ... reported as the
        # pipe's own utilisation instead of a derived count
        "xu_pipe_active_frac": xu_pct / 100.0,
        "mufu_warp_inst_per_pair": None,
        "fp64_pipe_active_frac": get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", 0.0) / 100.0,
        "issue_slots_active_frac": get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0.0) / 100.0,
        "achieved_occupancy_frac": get("sm__warps_active.avg.pct_of_peak_sustained_active", 0.0) / 100.0,
        "warps_per_sm": get("sm__warps_active.avg.per_cycle_active"),
        "dram_bytes_read": get("dram__bytes_read.sum"), "dram_bytes_write": get("dram__bytes_write.sum"),
        "dram_bytes_per_pair": (get("dram__bytes_read.sum", 0.0) + get("dram__bytes_write.sum", 0.0)) / a.pairs,
        "sm_clock_ghz": get("smsp__cycles_elapsed.avg.per_second"),
        "stall_wait_frac": stalls.get("wait", 0.0) / total_stall,
        "stall_math_throttle_frac": stalls.get("math_pipe_throttle", 0.0) / total_stall,
        "stalls_per_issue": stalls,
        "note": a.note, "build_id": a.build_id,
    }
    json.dump(out, open(a.out + ".json", "w"), indent=1)
    with open(a.out + ".txt", "w") as f:
        f.write(f"ncu --set full summary ({a.raw_csv}); {a.note}\n")
        for k, v in out.items():
            if k != "stalls_per_issue":
                f.write(f"{k:32s} {v}\n")
        f.write("warp-state samples per issued instruction (smsp__average_warps_issue_stalled_*_per_issue_active):\n")
        for k, v in sorted(stalls.items(), key=lambda kv: -kv[1]):
            f.write(f"    {k:24s} {v:8.3f}  ({100 * v / total_stall:5.1f} %)\n")
    print(json.dumps({k: out[k] for k in ("duration_ms_under_ncu", "registers", "fp64_inst_per_pair", "other_warp_inst_per_pair", "fp64_pipe_active_frac",
                                          "issue_slots_active_frac", "dram_bytes_per_pair")}))


if __name__ == "__main__":
    main()

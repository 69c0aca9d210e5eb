"""profiles/traffic.json from committed ncu `--set full` raw pages (ncu -i X.ncu-rep --page raw --csv).

usage: python scripts/make_traffic_json.py BUILD_ID  workload=csv:cells[:also,also]  ...
  e.g. python scripts/make_traffic_json.py $(python -c "import mpi_parallel_multiscale_diffusion_fem_b200 as p; print(p.build_id())") \
           target=profiles/r02e_ncu_full_solve_fused_1184.csv:1184:cfg3  cfg1=profiles/r02e_ncu_full_solve_cluster_cfg1x64_1024.csv:1024:cfg1x64

Per workload: DRAM bytes per coarse cell of ONE launch of the dominant kernel (dram__bytes_read.sum + dram__bytes_write.sum
divided by the cells of the captured launch; bench.py multiplies by the cells of the running rank's launch) and the pipe
utilisations of the same page.  bench.py quotes them only when BUILD_ID equals msb_build_id() of the running library.
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def page(path):
    rows = list(csv.reader(open(path)))
    h, u, v = rows[0], rows[1], rows[-1]
    return {name: (v[i], u[i]) for i, name in enumerate(h)}


def num(p, name):
    val, unit = p[name]
    return float(val.replace(",", "")) * UNIT.get(unit, 1.0)


def main():
    build_id = sys.argv[1]
    out = {"build_id": build_id,
           "note": "per-cell DRAM bytes and pipe utilisations of the dominant kernel from the committed ncu --set full raw pages; "
                   "bench.py quotes them only when build_id equals msb_build_id() of the running library",
           "workloads": {}}
    for spec in sys.argv[2:]:
        name, rest = spec.split("=", 1)
        parts = rest.split(":")
        path, cells = parts[0], int(parts[1])
        also = parts[2].split(",") if len(parts) > 2 else []
        p = page(os.path.join(ROOT, path))
        dram = num(p, "dram__bytes_read.sum") + num(p, "dram__bytes_write.sum")
        wf = num(p, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")
        ent = {"dram_bytes_per_cell": dram / cells,
               "kernel": p["Kernel Name"][0],
               "smem_wavefront_frac": num(p, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed") / 100,
               "fp64_pipe_frac": num(p, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed") / 100,
               "issue_slot_frac": num(p, "smsp__issue_active.avg.pct_of_peak_sustained_active") / 100,
               "bank_conflict_wavefront_frac": (num(p, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum") / wf) if wf else None,
               "captured_launch_ms": num(p, "gpu__time_duration.sum") * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(p["gpu__time_duration.sum"][1], 1.0),
               "source": "%s: dram__bytes_read.sum + dram__bytes_write.sum = %.1f MB for one launch over %d coarse cells (ncu --set full), "
                         "scaled linearly to the launch of bench.py" % (path, dram / 1e6, cells)}
        for w in [name] + also:
            out["workloads"][w] = ent
    json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    print(json.dumps(out, indent=1)[:1500])


if __name__ == "__main__":
    main()

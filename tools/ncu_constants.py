"""Turn an `ncu --csv --metrics ...` launch dump into profiles/r02_kernel_constants.json, the per-launch counters
bench.py quotes in its roofline objects (so they are derived from a committed capture, not typed in).

    ncu --metrics gpu__time_duration.sum,sm__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum,\\
dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active \\
        --clock-control none --csv --log-file gpurun_out/r02_counters.csv python tools/prof_counters.py
    python tools/ncu_constants.py gpurun_out/r02_counters.csv profiles/r02_kernel_constants.json

Kernels are matched by name; the median launch is taken over the launches of the same kernel + grid size.  The
site-updates per launch are fixed by tools/prof_counters.py (C2: 4096 chains x 64^2 sites x 50 steps for
k_traj_resident, 4096 x 64^2 for one k_leapfrog_step)."""
import csv
import json
import sys
from collections import defaultdict

SU = {"k_traj_resident": 4096 * 64 * 64 * 50, "k_leapfrog_step": 4096 * 64 * 64}


def main(src, dst):
    rows = [r for r in csv.reader(open(src, errors="ignore")) if len(r) > 10]
    hdr = rows[0]
    col = {h: i for i, h in enumerate(hdr)}
    acc = defaultdict(lambda: defaultdict(list))
    for r in rows[1:]:
        try:
            name, metric, val = r[col["Kernel Name"]], r[col["Metric Name"]], float(r[col["Metric Value"]].replace(",", ""))
            unit = r[col["Metric Unit"]]
        except (ValueError, KeyError, IndexError):
            continue
        scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1e-6, "ms": 1e-3, "ns": 1e-9,
                 "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}.get(unit, 1.0)
        acc[name][metric].append(val * scale)
    out = {}
    for key, su in SU.items():
        for name, m in acc.items():
            if key not in name:
                continue
            mean = lambda k: (sorted(m[k])[len(m[k]) // 2]) if m.get(k) else None     # median launch (the thermalisation launch runs 20 trajectories)
            fp64 = mean("sm__inst_executed_pipe_fp64.sum")
            d = {"kernel": name[:120], "launches_seen": len(m.get("gpu__time_duration.sum", [])),
                 "ncu_time_s_per_launch": mean("gpu__time_duration.sum"),
                 "dram_bytes_per_launch": (mean("dram__bytes_read.sum") or 0) + (mean("dram__bytes_write.sum") or 0),
                 "fp64_warp_inst_per_launch": fp64,
                 "fp64_thread_inst_per_site_update": (fp64 * 32 / su) if fp64 else None,
                 "warp_inst_per_launch": mean("smsp__inst_executed.sum"),
                 "ncu_pipe_fp64_pct": mean("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
                 "site_updates_per_launch": su, "source": f"{src} (ncu --metrics pass of tools/prof_counters.py, this round)"}
            out[key] = d
            break
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])

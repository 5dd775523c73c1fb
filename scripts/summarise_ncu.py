"""gpurun_out/<tag>_<kernel>_ncu_raw.csv (ncu --page raw --csv of one --set full capture, see
scripts/capture_profiles.sh) -> profiles/<tag>_<kernel>_ncu_key_metrics.csv (the rows a reader needs) and, for K1,
profiles/r02_k1_ncu_summary.json (read by bench.py for roofline.pipe_fp64_pct / traffic).

    python scripts/summarise_ncu.py r02p
"""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = re.compile(r"^(Kernel Name|Block Size|Grid Size|gpu__time_duration\.sum|dram__bytes_(read|write)\.sum$|"
                  r"launch__(registers_per_thread|waves_per_multiprocessor|grid_size|block_size|occupancy_limit_.*|shared_mem_per_block.*)|"
                  r"sm__inst_executed_pipe_fp64\.avg\.pct_of_peak_sustained_active|sm__pipe_fp64_cycles_active\.avg\.pct_of_peak_sustained_(active|elapsed)|"
                  r"sm__issue_active\.avg\.pct_of_peak_sustained_elapsed|smsp__issue_active\.avg\.(pct_of_peak_sustained_active|per_cycle_active)|"
                  r"sm__warps_active\.avg\.pct_of_peak_sustained_active|smsp__warps_(active|eligible)\.avg\.per_cycle_active|"
                  r"smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio|smsp__inst_executed\.sum|sm__cycles_elapsed\.max|"
                  r"smsp__cycles_active\.avg|smsp__sass_thread_inst_executed_op_d(fma|mul|add)_pred_on\.sum|"
                  r"smsp__sass_inst_executed_op_(local|shared|global)_(ld|st)\.sum|l1tex__t_bytes.*lookup_(hit|miss)\.sum|"
                  r"lts__t_sector_hit_rate\.pct|sm__throughput\.avg\.pct_of_peak_sustained_elapsed|dram__throughput\.avg\.pct_of_peak_sustained_elapsed)")


def load(path):
    rows = list(csv.reader(open(path)))
    hdr, units, val = rows[0], rows[1], rows[2]
    return {h: (units[i], val[i]) for i, h in enumerate(hdr)}


def main():
    tag = sys.argv[1]
    src = os.path.join(ROOT, "gpurun_out")
    dst = os.path.join(ROOT, "profiles")
    for name in ("k1", "pt", "k4", "scan", "mle"):
        path = os.path.join(src, "%s_%s_ncu_raw.csv" % (tag, name))
        if not os.path.exists(path):
            continue
        m = load(path)
        with open(os.path.join(dst, "%s_%s_ncu_key_metrics.csv" % (tag, name)), "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["metric", "unit", "value"])
            for h, (u, v) in m.items():
                if KEEP.match(h):
                    w.writerow([h, u, v])
        if name == "k1":
            def fl(k):
                u, v = m[k]
                x = float(v.replace(",", ""))
                return x * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
            out = {
                "source": "profiles/%s_k1_ncu_key_metrics.csv: ncu --set full --clock-control none, one steady-state launch of "
                          "loglik_batch_kernel<5> inside `bench.py --steps 2 --warmup 3 --no-cpu` (scripts/capture_profiles.sh)" % tag,
                "sm__inst_executed_pipe_fp64_pct": fl("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
                "pipe_fp64_pct_of_elapsed": fl("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed"),
                "issue_active_pct": fl("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                "warps_active_per_scheduler": fl("smsp__warps_active.avg.per_cycle_active"),
                "scheduler_cycles_active_of_elapsed": fl("smsp__cycles_active.avg") / fl("sm__cycles_elapsed.max"),
                "registers_per_thread": fl("launch__registers_per_thread"),
                "gpu_time_us_under_ncu": fl("gpu__time_duration.sum"),
                "dram_bytes_per_launch": fl("dram__bytes_read.sum") + fl("dram__bytes_write.sum"),
                "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of the same capture (cold caches under the "
                                  "profiler: includes first-touch of theta, the series and the instruction stream)",
            }
            json.dump(out, open(os.path.join(dst, "r02_k1_ncu_summary.json"), "w"), indent=1)
            print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

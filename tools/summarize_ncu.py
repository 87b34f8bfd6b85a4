"""Turns `ncu -i REPORT --page raw --csv` into the two files kept under profiles/:
  <out>_summary.csv  selected counters per kernel launch
  kernel_counters.json   per stage of ONE step: dram bytes (read + write), executed warp instructions and the ncu
                         duration; read by bench.py for roofline.{achieved,traffic}
usage: python tools/summarize_ncu.py REPORT.ncu-rep profiles/r2_ncu_full_summary.csv profiles/kernel_counters.json"""
import csv
import json
import subprocess
import sys

KEEP = [
    ("gpu__time_duration.sum", "time_us"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs"),
    ("launch__occupancy_limit_registers", "occ_lim_regs"), ("launch__occupancy_limit_shared_mem", "occ_lim_smem"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes_per_inst"),
    ("smsp__inst_executed.sum", "warp_insts"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "pipe_fma_pct"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "pipe_xu_pct"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "pipe_lsu_pct"),
    ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
]
STAGE = [("preprocess_forward", "preprocess"), ("preprocess_backward", "preprocess_backward"),
         ("render_forward", "render_forward"), ("render_backward", "render_backward"), ("clear_gradients", "clear_gradients"),
         ("sort_", "binning"), ("scan_sorted", "binning"), ("ms_", "binning"), ("duplicate", "binning"),
         ("ranges_cull", "binning"), ("gaussian_heads_forward", "heads_forward"), ("gaussian_heads_backward", "heads_backward")]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def main(rep, out_csv, out_traffic):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    out, per_stage, counts = [], {}, {}
    for r in rows[2:]:
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("ocrf::", "")
        rec = {"kernel": name}
        for key, short in KEEP:
            if key not in ix:
                continue
            try:
                v = float(r[ix[key]].replace(",", ""))
            except ValueError:
                continue
            rec[short] = v * UNIT.get(units[ix[key]], 1.0)
        out.append(rec)
        for pat, st in STAGE:
            if pat in name:
                d = per_stage.setdefault(st, {"dram_bytes": 0.0, "warp_insts": 0.0, "time_us_ncu": 0.0})
                d["dram_bytes"] += rec.get("dram_read", 0.0) + rec.get("dram_write", 0.0)
                d["warp_insts"] += rec.get("warp_insts", 0.0)
                d["time_us_ncu"] += rec.get("time_us", 0.0)
                counts.setdefault(st, {}).setdefault(name, 0)
                counts[st][name] += 1
                break
    cols = ["kernel"] + [s for _, s in KEEP]
    with open(out_csv, "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=cols)
        w.writeheader()
        for rec in out:
            w.writerow({k: (round(v, 3) if isinstance(v, float) else v) for k, v in rec.items()})
    if out_traffic:
        # the capture may hold several steps: normalise to ONE step by the launch count of a once-per-step kernel
        steps = max(1, max((n for st in counts.values() for k, n in st.items() if "preprocess_forward" in k), default=1))
        res = {k: {m: x / steps for m, x in v.items()} for k, v in per_stage.items()}
        res["_source"] = "%s -> %s (%d step(s) captured)" % (rep, out_csv, steps)
        json.dump(res, open(out_traffic, "w"), indent=1)
    print("kernels:", len(out), "stages:", {k: round(v["dram_bytes"] / 1e6, 1) for k, v in per_stage.items()})


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)

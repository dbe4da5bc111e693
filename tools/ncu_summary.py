#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page raw --csv` output into the JSON kept under profiles/: per kernel the duration,
DRAM bytes read/written, tensor-pipe / SM / issue utilisation, L2 hit rate, registers, grid.
    ncu -i gpurun_out/x.ncu-rep --page raw --csv > raw.csv && python tools/ncu_summary.py raw.csv > profiles/x_summary.json"""
import csv
import json
import sys

WANT = {
    "gpu__time_duration.sum": "time_us",
    "dram__bytes_read.sum": "dram_read_MB",
    "dram__bytes_write.sum": "dram_write_MB",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__inst_executed.sum": "warp_inst",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active": "tensor_hmma_pct",
    "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active": "tensor_imma_pct",
    "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "smem_wavefront_pct",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__shared_mem_per_block_dynamic": "smem_dyn_bytes",
}


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ci = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in rows[2:]:
        if len(r) < len(hdr) // 2:
            continue
        rec = {"kernel": r[ci["Kernel Name"]][:110]}
        for m, name in WANT.items():
            if m in ci and r[ci[m]] not in ("", "n/a"):
                v = float(r[ci[m]].replace(",", ""))
                u = units[ci[m]]
                if name == "time_us":
                    v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
                if name.endswith("_MB"):
                    v = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0) * v
                rec[name] = round(v, 3)
        if "dram_read_MB" in rec and "time_us" in rec:
            rec["dram_GBs"] = round((rec["dram_read_MB"] + rec.get("dram_write_MB", 0.0)) / rec["time_us"] * 1e3, 1)
        out.append(rec)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1])

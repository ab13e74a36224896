#!/usr/bin/env python
"""Summarise an .ncu-rep (one kernel launch of `ncu --set full --import-source on`) as text for profiles/.

    python tools/ncu_summary.py gpurun_out/v9_count.ncu-rep [source-file-substring] > profiles/<name>.txt

Prints the headline metrics of the raw page, the warp-stall ranking, and -- when the report carries source
(-lineinfo builds) -- the per-source-line share of executed instructions and of stall samples.
"""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__waves_per_multiprocessor", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "sm__cycles_active.avg", "sm__cycles_elapsed.avg",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    only = sys.argv[2] if len(sys.argv) > 2 else None
    rows = page(rep, "raw")
    hdr, units, vals = rows[0], rows[1], rows[2]
    col = {h: i for i, h in enumerate(hdr)}
    print("# %s: %s" % (rep.split("/")[-1], vals[col["Kernel Name"]] if "Kernel Name" in col else ""))
    for w in WANT:
        if w in col:
            print(w, vals[col[w]], units[col[w]])
    st = []
    for h, i in col.items():
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            try:
                st.append((float(vals[i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
    st.sort(reverse=True)
    print("stalls (warps per issue): " + ", ".join("%s=%.2f" % (n, v) for v, n in st[:9]))

    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    fname, hdr, lines = "", None, []
    tot_i = tot_s = 0.0
    for r in csv.reader(io.StringIO(out)):
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = {n: i for i, n in enumerate(r)}
        elif hdr and r[0].isdigit():
            try:
                ni = float(r[hdr["Instructions Executed"]] or 0)
                nsmp = float(r[hdr["# Samples"]] or 0)
            except (ValueError, KeyError):
                continue
            tot_i += ni
            tot_s += nsmp
            lines.append((ni, nsmp, "%s:%s" % (fname, r[0]), r[1].strip()))
    if not lines:
        return
    print("\n# per source line (share of executed warp instructions / of stall samples), lines >= 1 %%; "
          "total inst %.3g, samples %.0f" % (tot_i, tot_s))
    for ni, nsmp, where, text in lines:
        if only and only not in where:
            continue
        if ni >= 0.01 * tot_i or (tot_s and nsmp >= 0.01 * tot_s):
            print("%-16s inst %5.1f%% smp %5.1f%% | %s" % (where, 100 * ni / max(tot_i, 1), 100 * nsmp / max(tot_s, 1), text[:100]))


if __name__ == "__main__":
    main()

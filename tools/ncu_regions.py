"""Share of executed warp instructions (and stall samples) per code REGION of place.cu / common.cuh from an .ncu-rep
captured with --import-source on -- profiling aid:  python tools/ncu_regions.py report.ncu-rep"""
import collections
import csv
import io
import subprocess
import sys

REGIONS = {
    "place.cu": [(16, 60, "draw_turn + ws_pick"), (61, 140, "trim_ends"), (141, 176, "insert_one"), (177, 216, "filter / drop_empty"),
                 (217, 268, "kernel set-up"), (269, 345, "main loop: window / trigger / append"),
                 (346, 400, "checkpoint dispatch, overshoot, tail"), (401, 600, "shift kernel"), (601, 640, "contig merge"),
                 (641, 760, "prep_units")],
    "common.cuh": [(28, 57, "philox"), (58, 110, "warp scans / shuffles"), (111, 166, "bitonic sort"),
                   (167, 240, "ws_cov / ws_overlap"), (241, 300, "merge0_sorted"), (301, 400, "bucket sort"),
                   (401, 480, "sort_merge0 / insert_merge0")],
}


def region(f, line):
    for a, b, n in REGIONS.get(f, []):
        if a <= line <= b:
            return "%s: %s" % (f, n)
    return f


def main():
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    fname, hdr = "", None
    inst, smp = collections.Counter(), collections.Counter()
    for r in csv.reader(io.StringIO(out)):
        if not r:
            continue
        if r[0] in ("File Path", "File Name"):
            fname = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = {n: i for i, n in enumerate(r)}
        elif hdr and r[0].isdigit() and "Instructions Executed" in hdr:
            try:
                inst[region(fname, int(r[0]))] += float(r[hdr["Instructions Executed"]] or 0)
                smp[region(fname, int(r[0]))] += float(r[hdr["# Samples"]] or 0)
            except (ValueError, KeyError, IndexError):
                pass
    ti, ts = sum(inst.values()) or 1, sum(smp.values()) or 1
    for k, v in inst.most_common():
        print("%-48s inst %5.1f%%  samples %5.1f%%" % (k, 100 * v / ti, 100 * smp[k] / ts))


if __name__ == "__main__":
    main()

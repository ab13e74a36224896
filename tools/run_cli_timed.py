"""Run `python -m gat_b200.cli` (under torch.distributed.run for N > 1) on BED files and print ONE JSON line:
wall seconds, the phases the CLI / gat_b200.run report (GATB_TIMING=1), and a checksum of the result table --
which must not depend on the number of GPUs (the placement stream is keyed by the global sample index).

    python tools/run_cli_timed.py --gpus 8 --label ns_1e6 -- --segments=s.bed --annotations=a.bed \
        --workspace=w.bed --num-samples=1000000 --random-seed=1 --qvalue-method=BH
"""
import argparse
import hashlib
import json
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--label", default="run")
    ap.add_argument("--port", type=int, default=29531)
    ap.add_argument("cli", nargs=argparse.REMAINDER)
    args = ap.parse_args()
    cli = [a for a in args.cli if a != "--"]
    out = tempfile.NamedTemporaryFile(prefix="gat_b200_", suffix=".tsv", delete=False).name
    cmd = [sys.executable]
    if args.gpus > 1:
        cmd += ["-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus), "--master-addr", "127.0.0.1",
                "--master-port", str(args.port)]
    cmd += ["-m", "gat_b200.cli"] + cli + ["-S", out]
    env = dict(os.environ, GATB_TIMING="1", PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    t0 = time.perf_counter()
    p = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True)
    wall = time.perf_counter() - t0
    line = {"label": args.label, "gpus": args.gpus, "wall_s": round(wall, 2), "returncode": p.returncode, "cli": " ".join(cli)}
    m = re.search(r"intervals loaded in ([0-9.]+) seconds", p.stderr)
    if m:
        line["load_and_prepare_s"] = float(m.group(1))
    m = re.search(r"sampling completed in ([0-9.]+) seconds", p.stderr)
    if m:
        line["run_s"] = float(m.group(1))
    m = re.search(r"# gat_b200.run phases: (.*)", p.stderr)
    if m:
        line["phases"] = dict((k.strip(), float(v[:-1])) for k, v in (x.rsplit(" ", 1) for x in m.group(1).split(", ")))
    m = re.search(r"output written in ([0-9.]+) seconds", p.stderr)
    if m:
        line["output_s"] = float(m.group(1))
    if p.returncode == 0 and os.path.exists(out):
        text = open(out).read()
        rows = text.strip().split("\n")
        line["rows"] = len(rows) - 1
        line["table_md5"] = hashlib.md5(text.encode()).hexdigest()
        line["first_row"] = rows[1][:140] if len(rows) > 1 else ""
    else:
        line["stderr_tail"] = p.stderr[-1500:]
    print(json.dumps(line))
    if os.path.exists(out):
        os.unlink(out)
    sys.exit(p.returncode)


if __name__ == "__main__":
    main()

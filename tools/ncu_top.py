#!/usr/bin/env python
"""Summarise `ncu --page source --csv` output: hottest SASS instructions by stall samples."""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    body = rows[2:]
    tot = sum(int(r[ix["# Samples"]] or 0) for r in body)
    texe = sum(int(r[ix["Instructions Executed"]] or 0) for r in body)
    print("kernel:", rows[0][1], "| samples", tot, "| warp instr executed", texe, "| SASS lines", len(body))
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    ranked = sorted(enumerate(body), key=lambda t: -int(t[1][ix["# Samples"]] or 0))[:top]
    for pos, r in sorted(ranked):
        s = int(r[ix["# Samples"]] or 0)
        stalls = sorted(((int(r[ix[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
        print("%5d %5.1f%% exe=%9s thr=%5s  %-60s %s" % (
            pos, 100.0 * s / max(tot, 1), r[ix["Instructions Executed"]], r[ix["Avg. Threads Executed"]],
            r[ix["Source"]].strip()[:60], " ".join("%s:%d" % (n, v) for v, n in stalls if v)))


if __name__ == "__main__":
    main()

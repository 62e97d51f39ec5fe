#!/usr/bin/env python
"""Whole-kernel totals of the warp-stall sampling reasons of an ncu report.

    python tools/ncu_stalls.py gpurun_out/prof.ncu-rep swnmf_fwd
"""
import csv
import io
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv",
                          "--kernel-name", f"regex:{kern}"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, tot = None, {}
    for r in rows:
        if not r:
            continue
        if r[0] == "Address":
            hdr = {h: i for i, h in enumerate(r)}
        elif hdr and r[0].startswith("0x"):
            for k, i in hdr.items():
                if k.startswith("stall_") and "Not Issued" not in k and i < len(r) and r[i].isdigit():
                    tot[k[6:]] = tot.get(k[6:], 0) + int(r[i])
    s = sum(tot.values()) or 1
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        if v:
            print(f"{100.0*v/s:5.1f}%  {k}  ({v})")


if __name__ == "__main__":
    main()

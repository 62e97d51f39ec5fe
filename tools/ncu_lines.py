#!/usr/bin/env python
"""Summarise an ncu report per CUDA source line: samples, instructions and dominant stall reasons.

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep swnmf_fwd [min_pct]
"""
import csv
import io
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv",
                          "--kernel-name", f"regex:{kern}"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    fname, hdr, lines = None, None, []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = {h: i for i, h in enumerate(r)}
        elif hdr and r[0].isdigit():
            s = r[hdr["# Samples"]]
            if s.isdigit():
                stalls = {}
                for k, i in hdr.items():
                    if k.startswith("stall_") and "Not Issued" not in k and r[i].isdigit() and int(r[i]) > 0:
                        stalls[k[6:]] = int(r[i])
                lines.append((fname, int(r[0]), int(s), int(r[hdr["Instructions Executed"]] or 0), r[1].strip(), stalls))
    tot = sum(l[2] for l in lines) or 1
    tot_inst = sum(l[3] for l in lines) or 1
    print(f"total samples {tot}, warp instructions {tot_inst}")
    for f, ln, s, inst, src, st in sorted(lines, key=lambda l: -l[2]):
        if 100.0 * s / tot < min_pct:
            break
        top = ", ".join(f"{k}:{v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:4])
        print(f"{100.0*s/tot:5.1f}% smp {100.0*inst/tot_inst:5.1f}% inst  {f}:{ln:<4d} {src[:70]:70s} | {top}")


if __name__ == "__main__":
    main()

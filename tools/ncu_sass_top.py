#!/usr/bin/env python
"""Top SASS instructions of a kernel in an ncu report by stall samples, with the dominant stall reasons.

    python tools/ncu_sass_top.py gpurun_out/prof.ncu-rep mlp_bwd_tc [n]
"""
import csv
import io
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv",
                          "--kernel-name", f"regex:{kern}"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, items = None, []
    for r in rows:
        if not r:
            continue
        if r[0] == "Address":
            hdr = {h: i for i, h in enumerate(r)}
        elif hdr and r[0].startswith("0x"):
            smp = r[hdr["# Samples"]]
            smp = int(smp) if smp.isdigit() else 0
            stalls = {k[6:]: int(r[i]) for k, i in hdr.items()
                      if k.startswith("stall_") and "Not Issued" not in k and i < len(r) and r[i].isdigit() and int(r[i]) > 0}
            ex = r[hdr["Instructions Executed"]] if "Instructions Executed" in hdr else "0"
            items.append((smp, r[0], r[hdr["Source"]], stalls, ex))
    tot = sum(i[0] for i in items) or 1
    print(f"total samples {tot}, {len(items)} SASS instructions")
    for idx, (smp, addr, src, st, ex) in enumerate(items):
        pass
    order = sorted(range(len(items)), key=lambda i: -items[i][0])[:n]
    for i in order:
        smp, addr, src, st, ex = items[i]
        top = ", ".join(f"{k}:{v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
        prev = items[i - 1][2][:60] if i > 0 else ""
        print(f"{100.0*smp/tot:5.1f}%  #{i:5d} ex={ex:>9s} {src[:70]:70s} | {top} | prev: {prev}")


if __name__ == "__main__":
    main()

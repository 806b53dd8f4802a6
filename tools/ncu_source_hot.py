#!/usr/bin/env python
"""Hottest source lines of an `ncu --set full --import-source on` capture (needs -lineinfo): stall samples per CUDA-C line.

    python tools/ncu_source_hot.py prof.ncu-rep [top]
"""
import csv
import io
import subprocess
import sys


def main():
    rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    cur, hdr, lines = None, None, []
    for r in csv.reader(io.StringIO(txt)):
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
        elif hdr and len(r) == len(hdr) and r[2] == "-" and r[0].isdigit():
            d = dict(zip(hdr, r))
            lines.append((int(d["# Samples"] or 0), cur, r[0], r[1].strip()[:100], int(d["Instructions Executed"] or 0), d))
    tot = sum(l[0] for l in lines) or 1
    print(f"total samples {tot}")
    for n, f, ln, src, inst, d in sorted(lines, key=lambda x: -x[0])[:top]:
        st = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "(Not" not in k and v not in ("", "0", "-")}
        best = ", ".join(f"{k} {v}" for k, v in sorted(st.items(), key=lambda x: -x[1])[:3])
        print(f"{n:7d} {100 * n / tot:5.1f}%  {f}:{ln:>4}  inst {inst:8d}  [{best}]  {src}")


if __name__ == "__main__":
    main()

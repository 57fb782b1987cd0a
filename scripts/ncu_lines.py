#!/usr/bin/env python
"""Per-source-line instruction counts and stall samples from an .ncu-rep captured with
--import-source on (kernels compiled with -lineinfo).
  python scripts/ncu_lines.py gpurun_out/x.ncu-rep [top N]"""
import csv
import io
import subprocess
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    fname, hdr, lines = None, None, []
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if len(r) > 4 and r[0] == "Line No":
            hdr = r
            continue
        if hdr and len(r) == len(hdr) and r[2] == "-":  # source-line summary rows have no address
            try:
                inst = int(r[hdr.index("Instructions Executed")])
                samp = int(r[hdr.index("# Samples")])
            except ValueError:
                continue
            lines.append((inst, samp, fname, r[0], r[1].strip()))
    ti = sum(l[0] for l in lines) or 1
    ts = sum(l[1] for l in lines) or 1
    print(f"# {path}: {ti} warp instructions, {ts} stall samples attributed to source lines")
    print("# --- by instructions")
    for inst, samp, f, ln, src in sorted(lines, reverse=True)[:top]:
        print(f"{100 * inst / ti:5.1f}% inst {100 * samp / ts:5.1f}% samp  {f}:{ln:>4s}  {src[:110]}")
    print("# --- by stall samples")
    for inst, samp, f, ln, src in sorted(lines, key=lambda l: -l[1])[:top]:
        print(f"{100 * inst / ti:5.1f}% inst {100 * samp / ts:5.1f}% samp  {f}:{ln:>4s}  {src[:110]}")


if __name__ == "__main__":
    main()

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
    phases = sampler_phases()
    if phases and any(l[2] == "sampler.cuh" for l in lines):
        print("# --- sampler kernel by phase (source-line ranges of csrc/sampler.cuh, common.cuh = Philox / rand_r)")
        agg = {}
        for inst, samp, f, ln, _ in lines:
            name = "rng (common.cuh)" if f == "common.cuh" else "other"
            if f == "sampler.cuh":
                n = int(ln)
                name = next((ph for ph, lo, hi in phases if lo <= n < hi), "other")
            a = agg.setdefault(name, [0, 0])
            a[0] += inst
            a[1] += samp
        for name, (inst, samp) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            print(f"{100 * inst / ti:5.1f}% inst {100 * samp / ts:5.1f}% samp  {name}")


def sampler_phases():
    """[(phase, first line, one past last line)] from anchor strings in csrc/sampler.cuh."""
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "surel_plus_b200", "csrc", "sampler.cuh")
    try:
        src = open(path).read().split("\n")
    except OSError:
        return []
    anchors = [("loads / row decode helpers", "// ---------------------------------------------------------------- cache-policy loads"),
               ("sort (network + merge path)", "// ---------------------------------------------------------------- warp merge sort"),
               ("LP-row interning", "// ---------------------------------------------------------------- LP-key interning"),
               ("kernel prologue / seed ticket", "// ---------------------------------------------------------------- the sampler kernel"),
               ("walk: trace load", "if (PARITY && a.rng_mode == SUBG_RNG_TRACE) {"),
               ("walk: first hop (incl. Fisher-Yates)", "// ---- first hop without replacement"),
               ("walk: hops", "// Walks are advanced in groups of kGW per lane"),
               ("keys -> registers, sort call", "if (lane == 0) keys[0] = (K)(uint32_t)u << OB;"),
               ("run heads + row allocation", "// ---- runs of equal node = one set member each"),
               ("landing counts (encode)", "// ---- landing counts per run"),
               ("first-visit ranks", "// ---- first-visit rank of every member"),
               ("emit (records, lookups, stores)", "// ---- emit the set"),
               ("tail", "if (lane == 0 && mx > 0) atomicMax(a.max_set, mx);")]
    pos = []
    for name, text in anchors:
        hit = next((i + 1 for i, line in enumerate(src) if text in line), None)
        if hit is not None:
            pos.append((hit, name))
    pos.sort()
    return [(name, lo, pos[i + 1][0] if i + 1 < len(pos) else 10 ** 9) for i, (lo, name) in enumerate(pos)]


if __name__ == "__main__":
    main()

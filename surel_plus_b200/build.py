"""Builds surel_plus_b200/_lib/libsubg_b200.so with nvcc for sm_100a (B200 only).

In-tree build: the .so is git-ignored but travels to the GPU box with the gpurun
snapshot.  `python -m surel_plus_b200.build` or `__graft_entry__.build()`.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "_lib")
LIB = os.path.join(LIBDIR, "libsubg_b200.so")
SOURCES = ["capi.cu", "spg.cu", "spjoin.cu", "ppr.cu", "walks.cu", "ingest.cu", "xchg.cu"] + [f"sampler_k{b}{p}.cu" for b in (32, 64) for p in "abcd"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _sha(paths, extra="") -> str:
    h = hashlib.sha256()
    for path in paths:
        with open(path, "rb") as f:
            h.update(os.path.basename(path).encode())
            h.update(f.read())
    h.update(extra.encode())
    return h.hexdigest()


def _headers():
    inc = os.path.join(os.path.dirname(HERE), "include")
    hs = [os.path.join(CSRC, n) for n in sorted(os.listdir(CSRC)) if n.endswith((".cuh", ".inc", ".h"))]
    return hs + [os.path.join(inc, n) for n in sorted(os.listdir(inc))]


def build(force: bool = False, verbose: bool = False, tag: str = "", defs: tuple = ()) -> str:
    """Incremental: an object is recompiled when its .cu, any header or the flags changed.
    tag / defs: an experiment variant (libsubg_b200_<tag>.so compiled with -D<def> ...), picked up at run time through
    SUBG_LIB=<path>; the product library is the untagged one."""
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(LIBDIR, "obj" + ("_" + tag if tag else ""))
    os.makedirs(objdir, exist_ok=True)
    headers = _headers()
    FLAGS = list(globals()["FLAGS"]) + [f"-D{d}" for d in defs]
    LIB = os.path.join(LIBDIR, f"libsubg_b200_{tag}.so") if tag else globals()["LIB"]
    flags = " ".join(FLAGS)

    def compile_one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        stamp = obj + ".sha256"
        dig = _sha([os.path.join(CSRC, src)] + headers, flags)
        if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read().strip() == dig:
            return obj, False
        cmd = [NVCC, *FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        with open(obj + ".ptxas.log", "w") as f:
            f.write(r.stderr)
        if verbose:
            sys.stderr.write(r.stderr)
        with open(stamp, "w") as f:
            f.write(dig)
        return obj, True

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as ex:
        res = list(ex.map(compile_one, SOURCES))
    objs = [o for o, _ in res]
    if not any(c for _, c in res) and os.path.exists(LIB) and not force:
        return LIB
    cmd = [NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


#: SASS mnemonics worth seeing at a glance: TMA bulk copies + their mbarriers, global / shared atomics, read-only-path
#: gathers, warp reductions, shuffles, votes, matrix pipes (there should be none: nothing here is a contraction)
SASS_WATCH = ["UBLKCP", "UTMALDG", "SYNCS", "ATOMG", "ATOMS", "REDG", "RED.", "LDG.E.CONSTANT", "LDG.E.64.CONSTANT", "LDG.E.128",
              "STG.E.128", "LDS", "STS", "REDUX", "SHFL", "VOTE", "MATCH", "VIMNMX", "POPC", "HMMA", "UTC", "LDTM", "BAR.SYNC",
              "WARPSYNC", "LDL", "STL"]


def sass_summary(out_path: str | None = None) -> str:
    """Opcode histogram of every object of the library (cuobjdump -sass), written to profiles/sass_summary.txt:
    the evidence that the kernels are sm_100a code and which Blackwell / memory-system instructions they use."""
    import collections
    import re
    objdir = os.path.join(LIBDIR, "obj")
    cuobjdump = os.path.join(os.path.dirname(NVCC), "cuobjdump")
    lines = ["# cuobjdump -sass of surel_plus_b200/_lib/obj/*.o (regenerate: python -m surel_plus_b200.build --sass)",
             "# per object: arch, kernels, instructions, then the watched mnemonics (prefix match) with their counts", ""]
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        if not os.path.exists(obj):
            continue
        r = subprocess.run([cuobjdump, "-sass", obj], capture_output=True, text=True)
        arch = sorted(set(re.findall(r"arch = (sm_\w+)", r.stdout)))
        kernels = re.findall(r"Function : (\S+)", r.stdout)
        ops = collections.Counter()
        for m in re.finditer(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", r.stdout, flags=re.M):
            ops[m.group(1)] += 1
        watched = []
        for w in SASS_WATCH:
            n = sum(c for o, c in ops.items() if o.startswith(w))
            if n:
                watched.append(f"{w}*={n}")
        lines.append(f"{src}: arch {','.join(arch) or '?'}; {len(kernels)} kernels; {sum(ops.values())} instructions")
        lines.append("    " + "  ".join(watched))
        per = collections.Counter()
        for k in kernels:
            per[re.sub(r"I[A-Za-z0-9_]*E+v.*$", "", k)[:60]] += 1
        lines.append("    kernels: " + ", ".join(f"{k} x{n}" if n > 1 else k for k, n in sorted(per.items())))
        lines.append("")
    text = "\n".join(lines)
    out_path = out_path or os.path.join(os.path.dirname(HERE), "profiles", "sass_summary.txt")
    with open(out_path, "w") as f:
        f.write(text)
    return out_path


if __name__ == "__main__":
    tag = sys.argv[sys.argv.index("--tag") + 1] if "--tag" in sys.argv else ""
    defs = tuple(a[2:] for a in sys.argv if a.startswith("-D"))
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, tag=tag, defs=defs))
    if "--sass" in sys.argv:
        print(sass_summary())

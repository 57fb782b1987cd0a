#!/usr/bin/env python
"""Acceptance run "link-prediction Hits@50 unchanged" with the reference itself in the comparison:
examples/link_prediction.py (PyG-free restatement of the reference's Net, mean / attention / LSTM pooling) trained on
features from (a) the reference's nthread=1 rand_r stream replayed on the GPU, (b) the Philox fast path, (c) independent
numpy walks, and -- added here -- (d) the COMPILED reference as users run it: gset_sampler with nthread = -1, all OpenMP
threads drawing from one shared rand_r word (subg_acc.c:731-732), followed by the scipy CSR build of subg_matrix.

    python tests/acceptance_hits50.py --aggr lstm [--steps 300 ...]      (needs oracle/_ref, i.e. a build where
                                                                          /root/reference existed; it travels to the GPU box)
Test infrastructure: this file may use oracle/; the example and the package never do."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "examples"))

import link_prediction as lp  # noqa: E402
from oracle import pyoracle as po, reference as ref  # noqa: E402


def reference_threaded_spg(G_obs, args, sample_seed):
    subg = ref.subg_acc()
    n = G_obs.shape[0]
    q = np.arange(n, dtype=np.int32)
    devnull, saved = os.open(os.devnull, os.O_WRONLY), os.dup(1)   # the reference prints '#SubGAcc' lines from C
    sys.stdout.flush()
    os.dup2(devnull, 1)
    try:
        nsize, remap, enc = subg.gset_sampler(G_obs.indptr.astype(np.int32), G_obs.indices.astype(np.int32), q,
                                              num_walks=args.num_walks, num_steps=args.num_steps - 1, nthread=-1, seed=sample_seed)
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(devnull)
        os.close(saved)
    remap = remap.copy()
    iso = np.repeat(np.diff(G_obs.indptr) == 0, nsize)
    remap[0][iso] = np.repeat(q, nsize)[iso]      # the reference leaves an isolated seed's id unwritten
    return po.subg_matrix_from(nsize, remap, enc, q, n, args.num_steps)


if __name__ == "__main__":
    if ref.subg_acc() is None:
        print("oracle/_ref is not built: the nthread=-1 source is skipped", file=sys.stderr)
    else:
        lp.EXTRA_SOURCES.append(("compiled reference, nthread=-1 (all OpenMP threads on one shared rand_r word: what users run)",
                                 reference_threaded_spg))
    lp.main()

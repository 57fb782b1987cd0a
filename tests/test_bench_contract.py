"""CPU: the parts of bench.py that need no GPU -- the reference arm's JSON line (the contract the driver reads) and the
pure helpers behind `roofline` (algorithmic bytes per seed) and the query batches."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _bench_module():
    import importlib
    return importlib.import_module("bench")


@pytest.mark.reference
def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the unmodified reference on the host cores) on a reduced collab shape: one JSON line,
    impl = reference, the metric / unit / config keys of our own arm, cpu_baseline describing the run, e2e repeating the
    value with zero copy bytes.  Needs oracle/_ref (this container; on the GPU box the prebuilt file travels)."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "collab", "--scale", "0.05",
                        "--steps", "1", "--warmup", "1", "--ref-seconds", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "sampled_node_sets_per_sec" and d["unit"] == "seeds/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "collab" in d["config"]["workload"] and "reference_sample" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_algorithmic_bytes_per_seed_follow_survey_8d():
    """B_seed = 24 + 4 min(d, M) + 12 M (m - 1) + 8 |S_u| for d > 0, 32 for an isolated seed."""
    b = _bench_module()
    deg = np.array([0, 3, 500, 200], dtype=np.int64)
    M, m = 200, 3
    T = 1 + 40 + 420 + 300        # set sizes incl. the isolated seed's singleton
    want = 32 + (24 + 4 * 3 + 12 * M * (m - 1)) + (24 + 4 * 200 + 12 * M * (m - 1)) + (24 + 4 * 200 + 12 * M * (m - 1)) + 8 * (T - 1)
    assert b.seed_algorithmic_bytes(deg, M, m, float(T)) == float(want)


@pytest.mark.parametrize("k,B", [(20, 1024), (10, 11264), (1000, 1001), (1000, 64064)])
def test_query_batches_have_the_requested_shape_and_valid_nodes(k, B):
    """1 positive : k negatives (main.py --k); the 1-vs-1000 MRR pattern repeats every source against its negatives."""
    b = _bench_module()
    from surel_plus_b200.graphs import synthetic_graph
    A = synthetic_graph(5000, 30000, seed=2)
    deg = np.diff(A.indptr)
    q = b.make_queries(deg, A.indptr, A.indices, A.shape[0], B, k, np.random.default_rng(0))
    assert q.shape == (2, B) and q.dtype == np.int64 and q.min() >= 0 and q.max() < A.shape[0]
    npos = max(B // (k + 1), 1)
    is_edge = np.array([v in A.indices[A.indptr[u]:A.indptr[u + 1]] for u, v in q.T])
    assert is_edge.sum() >= npos               # the positives (shuffled into the batch) are edges of the graph
    if k >= 100:                               # every source of the MRR pattern appears about 1 + k times
        assert len(np.unique(q[0])) <= npos


def test_workloads_cover_the_five_baseline_configs():
    b = _bench_module()
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert len(base["configs"]) == 5
    assert set(b.WORKLOADS) == {"collab", "ppa", "citation2-ppr", "dblp", "twitter"}
    assert b.WORKLOADS["ppa"]["M"] == 200 and b.WORKLOADS["ppa"]["m"] == 3 and b.WORKLOADS["ppa"]["k"] == 20          # configs[1]
    assert b.WORKLOADS["collab"]["M"] == 200 and b.WORKLOADS["collab"]["m"] == 2                                        # configs[0]
    assert b.WORKLOADS["citation2-ppr"]["topk"] == 100 and b.WORKLOADS["citation2-ppr"]["k"] == 1000                    # configs[2]
    assert b.WORKLOADS["dblp"]["M"] == 100 and b.WORKLOADS["dblp"]["join"] == "triplet"                                 # configs[3]
    assert b.WORKLOADS["twitter"]["M"] == 100 and b.WORKLOADS["twitter"].get("device_graph")                            # configs[4]

"""CPU: host-side logic that needs no GPU -- edge-list parsing (surel_plus_b200.io), seed partitioning and the
LP-table merge of the multi-GPU exchange under randomised inputs, the pooled host buffers' ownership rules."""
import gc

import numpy as np
from hypothesis import given, settings, strategies as st

from surel_plus_b200 import parallel as par


def test_read_edgelist_matches_loadtxt(tmp_path):
    from surel_plus_b200.io import read_edgelist
    rng = np.random.default_rng(0)
    e = rng.integers(0, 10_000, (5000, 2))
    path = tmp_path / "edges.txt"
    with open(path, "w") as f:
        f.write("# comment line\n")
        for i, (a, b) in enumerate(e):
            f.write(f"{a}\t{b}\n" if i % 2 else f"{a} {b}\n")        # tabs and blanks, as in the twitter / SNAP releases
    row, col = read_edgelist(str(path))
    ref_row, ref_col = np.loadtxt(path, dtype=int).T                   # subg_acc/test/test.py:16
    assert row.dtype == np.int64 and np.array_equal(row, ref_row) and np.array_equal(col, ref_col)


@settings(max_examples=60, deadline=None)
@given(st.lists(st.floats(min_value=0.0, max_value=1e3, allow_nan=False), min_size=0, max_size=200), st.integers(1, 9))
def test_partition_by_work_is_a_partition(weights, world):
    b = par.partition_by_work(np.asarray(weights, dtype=np.float64), world)
    assert len(b) == world + 1 and b[0] == 0 and b[-1] == len(weights)
    assert np.all(np.diff(b) >= 0)


@settings(max_examples=60, deadline=None)
@given(st.lists(st.lists(st.tuples(st.integers(0, 3), st.integers(0, 3)), min_size=0, max_size=12, unique=True), min_size=1, max_size=5))
def test_merge_lp_tables_is_the_first_occurrence_scan(shards):
    """Merging per-shard unique tables in rank order == the reference's serial first-occurrence scan over the
    concatenated stream (subg_acc.c:957-978)."""
    tables = [np.array(t, np.int16).reshape(-1, 2) for t in shards]
    merged, maps = par.merge_lp_tables(tables)
    seen, order = {}, []
    for t in tables:
        for row in map(tuple, t.tolist()):
            if row not in seen:
                seen[row] = len(order)
                order.append(row)
    assert [tuple(r) for r in merged.tolist()] == order
    for t, mp in zip(tables, maps):
        assert mp.tolist() == [seen[tuple(r)] for r in t.tolist()]


def test_pinned_blocks_return_to_the_pool_when_the_last_view_dies(monkeypatch):
    """Arrays handed to the caller own their page-locked block through their base chain; the block goes back to the
    pool only when the array AND every view of it are gone (no GPU: the allocator is replaced by a counter)."""
    import ctypes as C
    from surel_plus_b200 import spg

    class FakeLib:
        def __init__(self):
            self.bufs, self.freed = {}, []

        def subg_host_alloc(self, pptr, nbytes):
            buf = (C.c_char * int(nbytes))()
            self.bufs[C.addressof(buf)] = buf
            pptr._obj.value = C.addressof(buf)
            return 0

        def subg_host_free(self, p):
            self.freed.append(p.value)

    pool = spg._PinnedPool(keep_bytes=1 << 20)
    pool._lib = FakeLib()
    monkeypatch.setattr(spg, "_pinned", pool)
    a = spg.pinned_empty((1000,), np.int32)
    a[:] = 7
    view = a[10:20]
    addr = a.ctypes.data
    del a
    gc.collect()
    assert pool._free == [] and view.tolist() == [7] * 10           # the view keeps the block alive
    del view
    gc.collect()
    assert len(pool._free) == 1 and pool._free[0][1] == addr
    b = spg.pinned_empty((900,), np.int32)                          # fits the recycled block (12.5 % slack rule)
    assert b.ctypes.data == addr and pool._free == []
    del b
    gc.collect()
    big = spg.pinned_empty((1 << 19,), np.int32)                    # 2 MiB > keep_bytes: trimmed when it comes back
    del big
    gc.collect()
    assert sum(c for c, _ in pool._free) <= 1 << 20 and len(pool._lib.freed) >= 1


def test_integration_md_python_blocks_parse_and_name_real_symbols():
    """The stubs a reference maintainer would copy from INTEGRATION.md must at least be valid Python and bind symbols
    that include/subg_b200.h really declares."""
    import ast
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    assert len(blocks) >= 4
    for b in blocks:
        ast.parse(b)
    header = open(os.path.join(root, "include", "subg_b200.h")).read()
    declared = set(re.findall(r"\b(subg_[a-z0-9_]+)\s*\(", re.sub(r"/\*.*?\*/", "", header, flags=re.S)))
    used = set(re.findall(r"L\.(subg_[a-z0-9_]+)", text))
    assert used and used <= declared, sorted(used - declared)


def test_top_level_subg_acc_is_the_b200_module():
    """`from subg_acc import gset_sampler, walk_sampler` (sampler/random_walks.py:18) resolves to this repository."""
    import subg_acc
    from surel_plus_b200 import subg_acc as ours
    assert subg_acc.gset_sampler is ours.gset_sampler and subg_acc.walk_sampler is ours.walk_sampler
    assert subg_acc.walk_join is ours.walk_join and subg_acc.batch_sampler is ours.batch_sampler


import pytest  # noqa: E402


@pytest.mark.reference
def test_unmodified_reference_sampler_module_binds_the_shim():
    """The reference's sampler/random_walks.py, imported UNMODIFIED from /root/reference, picks up this repository's
    `subg_acc` (north star: "sampler/ runs unchanged").  Its other imports are not in this image (fastremap, and
    utils / dataloader pull in torch_geometric + ogb): they are stubbed with the few names subg_matrix uses."""
    import importlib.util
    import os
    import sys
    import types
    import scipy.sparse as sp
    ref_root = "/root/reference"
    saved = {k: sys.modules.get(k) for k in ("fastremap", "utils", "dataloader")}
    saved_path = list(sys.path)
    try:
        sys.modules["fastremap"] = types.ModuleType("fastremap")
        for name in ("utils", "dataloader"):
            mod = types.ModuleType(name)
            mod.np, mod.csr_matrix = np, sp.csr_matrix
            mod.__all__ = ["np", "csr_matrix"]
            sys.modules[name] = mod
        spec = importlib.util.spec_from_file_location("ref_random_walks", os.path.join(ref_root, "sampler", "random_walks.py"))
        rw = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(rw)      # runs `sys.path.append(/root/reference)` and `from subg_acc import ...`
        from surel_plus_b200 import subg_acc as ours
        assert rw.gset_sampler is ours.gset_sampler and rw.walk_sampler is ours.walk_sampler
        # the body of the reference's subg_matrix needs exactly these names besides the sampler
        assert callable(rw.subg_matrix) and rw.csr_matrix is sp.csr_matrix
    finally:
        sys.path[:] = saved_path
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_scipy_upload_cache_follows_the_buffers(monkeypatch):
    """train._as_spg keeps the device copy of a scipy matrix per object; a matrix whose `.data` is REPLACED
    (utils.encoding(..., 'PPR') assigns a new array, utils.py:36), or a different target device, is uploaded again;
    invalidate_uploads() covers writes into the same buffer.  (SpG.from_scipy is stubbed: no GPU here.)"""
    import scipy.sparse as sp
    from surel_plus_b200 import train
    calls = []

    class _Fake:
        def __init__(self, tag):
            self.tag = tag
    monkeypatch.setattr(train.SpG, "from_scipy", classmethod(lambda cls, x, device="cuda": calls.append(device) or _Fake(len(calls))))
    train.invalidate_uploads()
    x = sp.random(50, 50, density=0.1, format="csr", random_state=0)
    a = train._as_spg(x, "cuda:0")
    assert train._as_spg(x, "cuda:0") is a and len(calls) == 1              # cached by object + buffers
    x.data = (x.data + 0.1) / (x.data.max() + 0.1)                          # utils.py:36
    b = train._as_spg(x, "cuda:0")
    assert b is not a and len(calls) == 2
    assert train._as_spg(x, "cuda:1") is not b and len(calls) == 3          # the device is part of the key
    x.data[:] = 0.5                                                         # in place: invisible ...
    assert len(calls) == 3 and train._as_spg(x, "cuda:1") is not None and len(calls) == 3
    train.invalidate_uploads()                                              # ... until told
    train._as_spg(x, "cuda:1")
    assert len(calls) == 4
    train.invalidate_uploads()

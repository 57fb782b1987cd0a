"""Data formats either side of the hot path.

* edge2csr        subg_acc/test/test.py:15-19: whitespace-separated edge list -> CSR (here: resident in HBM)
* save_npz / load_npz   main.py:184-202 (--save_ppr / --load_ppr): the SpG in scipy's .npz layout, so files
  written by the reference load here and files written here load with scipy.sparse.load_npz.
"""
from __future__ import annotations

import numpy as np

from .spg import DeviceGraph, SpG


def read_edgelist(file) -> tuple[np.ndarray, np.ndarray]:
    """(row, col) int64 arrays of a text edge list (`np.loadtxt(file, dtype=int).T` in test.py:16; parsed
    with pandas' C reader when available: loadtxt needs minutes per 100 M edges)."""
    try:
        import pandas as pd
        df = pd.read_csv(file, sep=r"\s+", header=None, comment="#", usecols=[0, 1], dtype=np.int64, engine="c")
        return df[0].to_numpy(), df[1].to_numpy()
    except ImportError:
        row, col = np.loadtxt(file, dtype=np.int64, usecols=(0, 1), ndmin=2).T
        return np.ascontiguousarray(row), np.ascontiguousarray(col)


def edge2csr(file="twitter-2010.txt", device="cuda", symmetrize=False) -> DeviceGraph:
    """test.py:15-19 with the COO -> CSR conversion (sort, coalesce, row pointer) on the device.
    Returns the resident graph; `.to_scipy()` gives the reference's csr_matrix, `.csr()` its arrays."""
    row, col = read_edgelist(file)
    return DeviceGraph.from_edges(row, col, symmetrize=symmetrize, device=device)


def save_npz(file, z, compressed=True) -> None:
    """scipy.sparse.save_npz layout (main.py:202) from a device SpG (or any scipy sparse matrix)."""
    if isinstance(z, SpG):
        v = z.views()
        arrays = dict(format=np.array("csr".encode("ascii")), shape=np.array(z.shape, dtype=np.int64),
                      data=v["data"].cpu().numpy(), indices=v["indices"].cpu().numpy(), indptr=v["indptr"].cpu().numpy())
        if not str(file).endswith(".npz"):
            file = str(file) + ".npz"
        (np.savez_compressed if compressed else np.savez)(file, **arrays)
    else:
        import scipy.sparse as sp
        sp.save_npz(file, z, compressed=compressed)


def load_npz(file, device="cuda") -> SpG:
    """scipy.sparse.load_npz (main.py:187) straight into HBM: int data -> LP pointer SpG, float data -> value SpG."""
    import scipy.sparse as sp
    return SpG.from_scipy(sp.load_npz(file).tocsr(), device)

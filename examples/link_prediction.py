#!/usr/bin/env python
"""End-to-end link prediction on the B200 SubGAcc path, shaped like the reference's main.py / train.py loop:

    subg_matrix (set sampling + LP encoding + SpG, on the device)  ->  per batch: gather (SpJoin)  ->  Net  ->  BCE

The model is a PyG-free restatement of the reference's Net (model.py:45-90: pe_embedding MLP, sum over the two
slots, set pooling per segment, MergeLayer scorer) with mean, attention or LSTM pooling written in plain PyTorch
(torch_geometric is not in this image).  Purpose: the north star's acceptance item "link-prediction Hits@50
unchanged" -- the same model trained on features from (a) the reference's own rand_r stream replayed on the GPU
(bit-identical to the reference's arrays) and (b) the Philox fast path must reach the same Hits@50 up to
training noise.  Data: a synthetic graph (heavy-tailed background + planted communities) split as the reference's dataloader does (dataloader.py:
train_ratio): 80 % of the edges form the observed graph the sets are sampled on, 10 % are training targets and
10 % test positives -- neither is present in the observed graph.

    python examples/link_prediction.py [--nodes 20000] [--edges 150000] [--steps 300] [--aggr mean|attn|lstm]
"""
import argparse
import os
import sys
import time

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("SUBG_QUIET", "1")
from surel_plus_b200 import _capi, gather, subg_matrix  # noqa: E402
from surel_plus_b200.graphs import synthetic_graph  # noqa: E402


def segment_mean(x, ptr):
    """MeanAggregation(x, ptr=ptr) of model.py:66."""
    sizes = (ptr[1:] - ptr[:-1])
    seg = torch.repeat_interleave(torch.arange(sizes.numel(), device=x.device), sizes)
    out = torch.zeros(sizes.numel(), x.shape[1], device=x.device, dtype=x.dtype).index_add_(0, seg, x)
    return out / sizes.clamp(min=1).unsqueeze(1).to(x.dtype)


class AttnPool(nn.Module):
    """AttentionalAggregation(gate_nn, nn) of model.py:59-62: softmax(gate(x)) over the segment, sum of nn(x)."""

    def __init__(self, h):
        super().__init__()
        self.gate, self.fnn = nn.Linear(h, 1), nn.Linear(h, h)

    def forward(self, x, ptr):
        sizes = (ptr[1:] - ptr[:-1])
        seg = torch.repeat_interleave(torch.arange(sizes.numel(), device=x.device), sizes)
        g = self.gate(x).squeeze(1)
        mx = torch.full((sizes.numel(),), -1e30, device=x.device).scatter_reduce_(0, seg, g, "amax")
        w = torch.exp(g - mx[seg])
        den = torch.zeros(sizes.numel(), device=x.device).index_add_(0, seg, w)
        w = w / den[seg].clamp(min=1e-30)
        return torch.zeros(sizes.numel(), x.shape[1], device=x.device).index_add_(0, seg, self.fnn(x) * w.unsqueeze(1))


class LstmPool(nn.Module):
    """aggr.LSTMAggregation(hidden, hidden) of model.py:63-66 (torch_geometric 2.2): the rows of every segment, in
    order, are padded into a dense [segments, max_len, h] batch (to_dense_batch), run through one nn.LSTM
    (batch_first) and the output of the LAST time step of the padded batch is the segment's embedding -- exactly
    what PyG computes, padding included.  `index` is the per-row segment id (gather's ptr=False, train.py:24-30)."""

    def __init__(self, h):
        super().__init__()
        self.lstm = nn.LSTM(h, h, batch_first=True)

    def forward(self, x, index, num_segments):
        sizes = torch.bincount(index, minlength=num_segments)
        start = torch.cumsum(sizes, 0) - sizes
        pos = torch.arange(x.shape[0], device=x.device) - start[index]
        dense = x.new_zeros(num_segments, int(sizes.max()), x.shape[1])
        dense[index, pos] = x
        return self.lstm(dense)[0][:, -1]


class Net(nn.Module):
    def __init__(self, input_dim, hidden, aggr="mean", dropout=0.1):
        super().__init__()
        self.pe_embedding = nn.Sequential(nn.Linear(input_dim, hidden), nn.ReLU(), nn.Linear(hidden, hidden))
        self.aggr = aggr
        self.pool = AttnPool(hidden) if aggr == "attn" else (LstmPool(hidden) if aggr == "lstm" else None)
        self.fc1, self.fc2 = nn.Linear(2 * hidden, hidden), nn.Linear(hidden, 1)   # MergeLayer, model.py:7-33
        self.dropout = dropout

    def forward(self, x, ptr, num_segments=None):
        x = self.pe_embedding(x).sum(dim=-2)                                        # model.py:78
        if self.aggr == "lstm":                                                     # model.py:82-83: aggr(x, index=ptr)
            pooled = self.pool(x, ptr, num_segments)
        else:
            pooled = self.pool(x, ptr) if self.pool is not None else segment_mean(x, ptr)
        xl, xr = pooled.view(2, -1, x.shape[-1])                                    # model.py:81
        h = F.dropout(F.relu(self.fc1(torch.cat([xl, xr], dim=-1))), p=self.dropout, training=self.training)
        return self.fc2(h).squeeze(1)


def hits_at_k(pos, neg, k=50):
    """OGB Hits@K: share of positives scored above the k-th best negative."""
    kth = torch.topk(neg, k).values[-1]
    return float((pos > kth).float().mean())


def numpy_walks(G, M, m, seed):
    """The reference's sampling law (subg_acc.c:763-809) written independently with numpy's PCG64: first hop without
    replacement (w % d if d <= M, else M distinct neighbours), later hops uniform; walks[n, M, m]."""
    rng = np.random.default_rng(seed)
    indptr, indices = G.indptr.astype(np.int64), G.indices
    n = G.shape[0]
    deg = np.diff(indptr)
    off = np.arange(M)[None, :] % np.maximum(deg, 1)[:, None]
    for u in np.where(deg > M)[0]:
        off[u] = rng.choice(deg[u], M, replace=False)
    cur = np.where(deg[:, None] > 0, indices[np.minimum(indptr[:-1, None] + off, len(indices) - 1)], np.arange(n)[:, None])
    walks = np.empty((n, M, m), np.int32)
    walks[:, :, 0] = cur
    for s in range(1, m):
        d = deg[cur]
        pick = np.minimum((rng.random(cur.shape) * d).astype(np.int64), np.maximum(d - 1, 0))
        nxt = indices[np.minimum(indptr[cur] + pick, len(indices) - 1)]
        cur = np.where(d > 0, nxt, cur)
        walks[:, :, s] = cur
    return walks


#: further sources of (SpG, LP table) to compare, appended by callers: (name, fn(G_obs, args, sample_seed) -> (z, enc)).
#: tests/acceptance_hits50.py adds the compiled reference run with nthread = -1 (kept out of this file: the product side
#: never touches the reference or the oracle).
EXTRA_SOURCES = []


def run(G_obs, pos_tr, pos_te, neg_te, rng_mode, args, model_seed, sample_seed):
    dev = "cuda:0"
    t0 = time.perf_counter()
    if callable(rng_mode):
        z, enc = rng_mode(G_obs, args, sample_seed)                 # e.g. a scipy CSR: gather uploads it once
    elif rng_mode == _capi.SUBG_RNG_TRACE:
        from surel_plus_b200 import DeviceGraph, SpG
        g = DeviceGraph.from_scipy(G_obs, dev)
        z = SpG.sample(g, np.arange(G_obs.shape[0]), num_walks=args.num_walks, num_steps=args.num_steps - 1,
                       rng_mode=rng_mode, walks=numpy_walks(G_obs, args.num_walks, args.num_steps - 1, sample_seed),
                       first_visit_ranks=False)
        enc = z.enc_table()
        g.close()
    else:
        z, enc = subg_matrix(G_obs, np.arange(G_obs.shape[0]), num_walks=args.num_walks, num_steps=args.num_steps,
                             device=dev, seed=sample_seed, rng_mode=rng_mode)
    xpe = torch.from_numpy(enc).to(dev).float() / args.num_walks                  # main.py:174
    t_prep = time.perf_counter() - t0
    torch.manual_seed(model_seed)
    net = Net(args.num_steps, args.hidden, args.aggr).to(dev)
    opt = torch.optim.Adam(net.parameters(), lr=args.lr)
    rng = np.random.default_rng(model_seed)
    n, B = G_obs.shape[0], args.batch
    net.train()
    for step in range(args.steps):
        p = pos_tr[:, rng.integers(0, pos_tr.shape[1], B)]
        q = rng.integers(0, n, (2, B))
        edge = torch.from_numpy(np.concatenate([p, q], axis=1))
        y = torch.cat([torch.ones(B), torch.zeros(B)]).to(dev)
        xz, ptr = gather(edge, z, dev, args.aggr != "lstm", xpe)                   # train.py:119-127 (ptr=False for LSTM)
        loss = F.binary_cross_entropy_with_logits(net(xz, ptr, 2 * edge.shape[1]), y)
        opt.zero_grad()
        loss.backward()
        opt.step()
    net.eval()
    with torch.no_grad():
        def score(e):
            out = []
            for i in range(0, e.shape[1], 4096):
                ee = torch.from_numpy(e[:, i:i + 4096])
                xz, ptr = gather(ee, z, dev, args.aggr != "lstm", xpe)
                out.append(net(xz, ptr, 2 * ee.shape[1]))
            return torch.cat(out)
        h50 = hits_at_k(score(pos_te), score(neg_te), 50)
    if hasattr(z, "close"):
        z.close()
    return h50, float(loss.detach()), t_prep


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", type=int, default=20000)
    ap.add_argument("--edges", type=int, default=150000)
    ap.add_argument("--num_walks", type=int, default=100)
    ap.add_argument("--num_steps", type=int, default=4)
    ap.add_argument("--hidden", type=int, default=96)
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--lr", type=float, default=1e-3)
    ap.add_argument("--aggr", default="mean", choices=["mean", "attn", "lstm"])
    ap.add_argument("--communities", type=int, default=400)
    ap.add_argument("--model-seeds", type=int, default=3)
    ap.add_argument("--sample-seeds", type=int, default=4)
    args = ap.parse_args()
    # heavy-tailed background (30 % of the edges) + planted communities (70 %): held-out edges are then predictable
    # from the structure around their endpoints, which is what the LP features encode
    A = synthetic_graph(args.nodes, int(args.edges * 0.3), seed=21, gamma=2.0).tocoo()
    und = A.row < A.col
    grng = np.random.default_rng(22)
    csize = max(args.nodes // args.communities, 2)
    comm = grng.integers(0, args.communities, int(args.edges * 0.7))
    a = comm * csize + grng.integers(0, csize, comm.size)
    b = comm * csize + grng.integers(0, csize, comm.size)
    ok = (a != b) & (a < args.nodes) & (b < args.nodes)
    e = np.concatenate([np.stack([A.row[und], A.col[und]]).astype(np.int64),
                        np.stack([np.minimum(a, b)[ok], np.maximum(a, b)[ok]])], axis=1)
    e = np.unique(e, axis=1)
    perm = np.random.default_rng(0).permutation(e.shape[1])
    n_te = e.shape[1] // 10
    pos_te, pos_tr, obs = e[:, perm[:n_te]], e[:, perm[n_te:2 * n_te]], e[:, perm[2 * n_te:]]
    import scipy.sparse as sp
    rows = np.concatenate([obs[0], obs[1]])
    cols = np.concatenate([obs[1], obs[0]])
    G_obs = sp.csr_matrix((np.ones(rows.size, dtype=bool), (rows, cols)), shape=A.shape)
    G_obs.sort_indices()
    neg_te = np.random.default_rng(1).integers(0, A.shape[0], (2, 20000))
    print(f"graph: {A.shape[0]} nodes, {obs.shape[1]} observed edges, {pos_tr.shape[1]} training targets, {n_te} test positives; LP M={args.num_walks} "
          f"num_steps={args.num_steps}; Net hidden={args.hidden} aggr={args.aggr}; {args.steps} steps of {args.batch}+{args.batch}")
    res = {}
    sources = [("rand_r replay (the reference's nthread=1 stream, bit-identical arrays)", _capi.SUBG_RNG_RAND_R),
               ("philox (fast path)", _capi.SUBG_RNG_PHILOX),
               ("numpy PCG64 walks fed as traces (independent statement of the sampling law)", _capi.SUBG_RNG_TRACE)]
    sources += list(EXTRA_SOURCES)
    for name, mode in sources:
        hs = []
        for ss in range(args.sample_seeds):          # one sampled SpG per sampling seed ...
            for ms in range(args.model_seeds):       # ... and several model initialisations / batch orders on it
                h50, loss, t_prep = run(G_obs, pos_tr, pos_te, neg_te, mode, args, ms, 111413 + ss)
                hs.append(h50)
                print(f"  {name}: sampling seed {111413 + ss} model seed {ms}: Hits@50 = {h50:.4f}  final loss {loss:.4f}  "
                      f"prep {t_prep * 1e3:.0f} ms", flush=True)
        res[name] = hs
    names = list(res)
    print(f"Hits@50 over {len(res[names[0]])} runs each (sampling seeds x model seeds):")
    for nm in names:
        print(f"  {np.mean(res[nm]):.4f} +- {np.std(res[nm], ddof=1):.4f}  {nm}")
    for i in range(len(names)):
        for j in range(i + 1, len(names)):
            ha, hb = res[names[i]], res[names[j]]
            se = float(np.sqrt(np.var(ha, ddof=1) / len(ha) + np.var(hb, ddof=1) / len(hb)))
            print(f"  difference of the means [{j}] - [{i}]: {np.mean(hb) - np.mean(ha):+.4f} (standard error {se:.4f})")


if __name__ == "__main__":
    main()

"""Per-rank graph construction (pipegcn_b200/distgraph.py, the papers100M-shaped path): every rank's layout built
from its own filtered pair stream equals the layout `PartitionPlan` builds from the union graph."""
import torch

from pipegcn_b200.distgraph import build_rank_layout
from pipegcn_b200.partition import PartitionPlan
from pipegcn_b200.synthetic import GlobalGraph, random_partition


def test_rank_layouts_equal_global_plan():
    spec = dict(n_nodes=3000, n_edges=40_000, n_feat=12, n_class=5, train_frac=0.3)
    P = 3
    lays = [build_rank_layout(spec, r, P, "cpu", pair_chunk=7000)[0] for r in range(P)]
    # union graph in global ids
    src, dst = [], []
    for L in lays:
        gid = torch.cat([L.inner_gid, L.halo_gid])
        rows = torch.repeat_interleave(torch.arange(L.num_in), (L.indptr[1:] - L.indptr[:-1]).long())
        src.append(gid[L.indices.long()])
        dst.append(L.inner_gid[rows])
    src, dst = torch.cat(src), torch.cat(dst)
    n = spec["n_nodes"]
    # symmetric, one self loop per node, no duplicates
    key = dst * n + src
    assert key.unique().numel() == key.numel()
    assert torch.equal(torch.sort(key).values, torch.sort(src * n + dst).values)
    assert int((src == dst).sum()) == n
    feat = torch.zeros(n, 12)
    label = torch.zeros(n, dtype=torch.int64)
    tm = torch.zeros(n, dtype=torch.bool)
    for L in lays:
        feat[L.inner_gid], label[L.inner_gid], tm[L.inner_gid] = L.feat, L.label, L.train_mask
    g = GlobalGraph(n, src, dst, feat, label, tm)
    part = random_partition(n, P, seed=1)
    plan = PartitionPlan(g, part, P)
    for r in range(P):
        ref, L = plan.build(r), lays[r]
        assert (ref.num_in, ref.num_all, ref.recv_shape) == (L.num_in, L.num_all, L.recv_shape)
        for a in ("indptr", "indices", "t_indptr", "t_indices", "in_deg", "feat", "label", "train_mask", "inner_gid",
                  "halo_gid"):
            assert torch.equal(getattr(ref, a), getattr(L, a)), (r, a)
        for a, b in zip(ref.boundary, L.boundary):
            assert (a is None and b is None) or torch.equal(a, b)
        assert ref.n_train_global == L.n_train_global

"""Property tests of the index spaces on random small graphs and partitions."""
import torch
from hypothesis import given, settings, strategies as st

from oracle import dglpart
from oracle import setup as osetup
from pipegcn_b200.partition import build_layouts
from pipegcn_b200.synthetic import GlobalGraph


def random_graph(n, m, seed):
    g = torch.Generator().manual_seed(seed)
    src = torch.randint(0, n, (m,), generator=g)
    dst = torch.randint(0, n, (m,), generator=g)
    keep = src != dst
    key = torch.unique(src[keep] * n + dst[keep])
    src, dst = key // n, key % n
    loops = torch.arange(n)
    src, dst = torch.cat([src, loops]), torch.cat([dst, loops])            # one self loop per node (utils.py:94-95)
    feat = torch.randn(n, 3, generator=g)
    label = torch.randint(0, 4, (n,), generator=g)
    mask = torch.rand(n, generator=g) < 0.6
    return GlobalGraph(n, src, dst, feat, label, mask)


@settings(max_examples=25, deadline=None)
@given(n=st.integers(6, 60), m=st.integers(0, 300), p=st.integers(1, 5), seed=st.integers(0, 10_000))
def test_layout_invariants(n, m, p, seed):
    g = random_graph(n, m, seed)
    gen = torch.Generator().manual_seed(seed + 1)
    part = torch.randint(0, p, (n,), generator=gen)
    part[:p] = torch.arange(p)                     # no empty part
    layouts = build_layouts(g, part, p)
    assert sum(l.num_in for l in layouts) == n and sum(l.nnz for l in layouts) == g.n_edges
    for r, L in enumerate(layouts):
        # CSR and CSC describe the same edge set
        rows = torch.repeat_interleave(torch.arange(L.num_in), (L.indptr[1:] - L.indptr[:-1]).long())
        trows = torch.repeat_interleave(torch.arange(L.num_all), (L.t_indptr[1:] - L.t_indptr[:-1]).long())
        assert torch.equal(torch.sort(rows * L.num_all + L.indices.long()).values,
                           torch.sort(L.t_indices.long() * L.num_all + trows).values)
        # every halo row feeds at least one inner row; every inner row has its self loop
        assert bool(((L.t_indptr[1:] - L.t_indptr[:-1])[L.num_in:] > 0).all())
        assert bool((L.indices.long()[L.indptr[:-1].long()] >= 0).all())
        self_loop = torch.zeros(L.num_in, dtype=torch.bool)
        self_loop[rows[L.indices.long() == rows]] = True
        assert bool(self_loop.all())
        # train rows first (move_train_first), global in-degree carried along
        k = int(L.train_mask.sum())
        assert bool(L.train_mask[:k].all()) and not bool(L.train_mask[k:].any())
        assert torch.equal(L.in_deg, g.in_degrees()[L.inner_gid])
        # halo block of peer j == the rows j sends, in j's boundary order (sender order == receiver order)
        off = L.num_in
        for j in range(p):
            if j == r:
                assert L.boundary[j] is None and L.recv_shape[j] is None
                continue
            h = L.recv_shape[j]
            sent = layouts[j].inner_gid[layouts[j].boundary[r]]
            assert torch.equal(L.halo_gid[off - L.num_in: off - L.num_in + h], sent)
            assert layouts[j].boundary[r].unique().numel() == h          # unique inside one peer's list
            off += h
        assert off == L.num_all


@settings(max_examples=10, deadline=None)
@given(n=st.integers(8, 40), m=st.integers(10, 200), p=st.integers(2, 4), seed=st.integers(0, 1000))
def test_layout_equals_reference_procedure(n, m, p, seed):
    """One-pass builder == the restated per-process procedure of train.py / utils.py (pinned by the goldens)."""
    g = random_graph(n, m, seed)
    part = torch.randint(0, p, (n,), generator=torch.Generator().manual_seed(seed))
    part[:p] = torch.arange(p)
    layouts = build_layouts(g, part, p)
    setups = osetup.setup_world(dglpart.partition_graph(n, g.src, g.dst, part, p, g.feat, g.label, g.train_mask))
    for L, S in zip(layouts, setups):
        rows = torch.repeat_interleave(torch.arange(L.num_in), (L.indptr[1:] - L.indptr[:-1]).long())
        assert torch.equal(torch.sort(rows * L.num_all + L.indices.long()).values, torch.sort(S.v * S.num_all + S.u).values)
        assert L.recv_shape == S.recv_shape
        for a, b in zip(L.boundary, S.boundary):
            assert (a is None and b is None) or torch.equal(a, b)


def test_inductive_train_subgraph():
    """--inductive (main.py:34-35): only train nodes and the edges among them survive; every node is a train node."""
    from pipegcn_b200.synthetic import train_subgraph
    g = random_graph(50, 300, 3)
    s = train_subgraph(g)
    n = int(g.train_mask.sum())
    assert s.n_nodes == n and bool(s.train_mask.all()) and s.feat.shape[0] == n
    keep = g.train_mask
    assert s.n_edges == int((keep[g.src] & keep[g.dst]).sum())
    old = torch.nonzero(keep).flatten()
    assert torch.equal(s.feat, g.feat[old]) and torch.equal(s.label, g.label[old])
    # edges map back to original train-train edges
    back = set(zip(old[s.src].tolist(), old[s.dst].tolist()))
    orig = {(a, b) for a, b in zip(g.src.tolist(), g.dst.tolist()) if keep[a] and keep[b]}
    assert back == orig
    lays = build_layouts(s, torch.arange(n) % 2, 2)
    assert sum(l.num_in for l in lays) == n and all(bool(l.train_mask.all()) for l in lays)

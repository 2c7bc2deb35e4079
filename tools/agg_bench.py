"""Aggregate kernel benchmark on one partition of a P-way split of a named shape (what a rank of a P-GPU run executes):
forward (CSR by destination, with the division) and backward (CSC by source, accumulate) for both kernel
implementations (`agg_impl` 1 = row per lane group, 2 = chunked), CUDA-event timed; one JSON line per case.

    python tools/agg_bench.py [shape=rmat-1m] [P=1] [dtype=bf16] [d=256] [--once]
`--once`: a single launch per case after warm-up (what `ncu -k regex:agg` profiles).
"""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from pipegcn_b200 import _C, ops
from pipegcn_b200.graph import PartGraph, alloc_rows
from pipegcn_b200.partition import PartitionPlan
from pipegcn_b200.synthetic import make_graph, random_partition

argv = [a for a in sys.argv[1:] if not a.startswith("--")]
once = "--once" in sys.argv
shape = argv[0] if len(argv) > 0 else "rmat-1m"
P = int(argv[1]) if len(argv) > 1 else 1
dtype = {"bf16": torch.bfloat16, "fp32": torch.float32}[argv[2] if len(argv) > 2 else "bf16"]
g = make_graph(shape, device="cuda")
d = int(argv[3]) if len(argv) > 3 else g.n_feat
part = random_partition(g.n_nodes, P, seed=1, device="cuda")
lay = PartitionPlan(g, part, P).build(0)
del g
graph = PartGraph.from_layout(lay, device="cuda")
x = alloc_rows(lay.num_all, d, dtype, "cuda")
x.copy_(torch.randn(lay.num_all, d, device="cuda"))
gy = alloc_rows(lay.num_in, d, dtype, "cuda")
gy.copy_(torch.randn(lay.num_in, d, device="cuda"))
gx = alloc_rows(lay.num_all, d, dtype, "cuda", zero=True)
info = dict(shape=shape, P=P, dtype=str(dtype), d=d, n_in=lay.num_in, num_all=lay.num_all, nnz=lay.nnz,
            chunks_fwd=graph.fwd.n_chunks, chunks_bwd=graph.bwd.n_chunks, seg_len=graph.fwd.seg_len)


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    if once:
        n = 1
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


ref = {}
for impl in (1, 2):
    _C.check(_C.lib.pg_set_option(b"agg_impl", impl))
    of = ops.aggregate(graph.fwd, x, row_div=graph.in_deg_f).clone()
    ob = ops.aggregate(graph.bwd, gy, out=gx.clone(), acc_rows=lay.num_in).clone()
    if impl == 1:
        ref = dict(f=of, b=ob)
    same = bool(torch.equal(of, ref["f"]) and torch.equal(ob, ref["b"]))
    tf = timeit(lambda: ops.aggregate(graph.fwd, x, row_div=graph.in_deg_f))
    tb = timeit(lambda: ops.aggregate(graph.bwd, gy, out=gx, acc_rows=lay.num_in))
    bf = ops.aggregate_bytes(graph.fwd, x, True)
    bb = ops.aggregate_bytes(graph.bwd, gy, False)
    es = x.element_size()
    print(json.dumps(dict(info, impl=impl, same_as_impl1=same, fwd_ms=tf, bwd_ms=tb,
                          fwd_alg_gbs=bf / tf / 1e6, bwd_alg_gbs=bb / tb / 1e6,
                          fwd_gather_tbs=lay.nnz * d * es / tf / 1e9, bwd_gather_tbs=lay.nnz * d * es / tb / 1e9)), flush=True)

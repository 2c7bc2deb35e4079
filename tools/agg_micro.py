"""Aggregate micro-benchmark on one partition of a P-way split (what a rank of a P-GPU run executes)."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from pipegcn_b200 import _C, ops
from pipegcn_b200.graph import PartGraph, alloc_rows
from pipegcn_b200.partition import PartitionPlan
from pipegcn_b200.synthetic import make_graph, random_partition

P = int(sys.argv[1]) if len(sys.argv) > 1 else 8
g = make_graph("rmat-1m", device="cuda")
part = random_partition(g.n_nodes, P, seed=1, device="cuda")
lay = PartitionPlan(g, part, P).build(0)
graph = PartGraph.from_layout(lay, device="cuda")
d = 256
x = alloc_rows(lay.num_all, d, torch.bfloat16, "cuda"); x.copy_(torch.randn(lay.num_all, d, device="cuda"))
gy = alloc_rows(lay.num_in, d, torch.bfloat16, "cuda"); gy.copy_(torch.randn(lay.num_in, d, device="cuda"))
print(f"P={P} N_in={lay.num_in} num_all={lay.num_all} nnz={lay.nnz} fwd mean deg {lay.nnz/lay.num_in:.1f} bwd mean deg {lay.nnz/lay.num_all:.1f}")

def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

ref_f = ref_b = None
for pack in (0, 1):
    _C.lib.pg_set_option(b"agg_pack_short", pack)
    of = ops.aggregate(graph.fwd, x, row_div=graph.in_deg_f)
    ob = ops.aggregate(graph.bwd, gy)
    if pack == 0: ref_f, ref_b = of.clone(), ob.clone()
    else: print("same result:", torch.equal(of, ref_f), torch.equal(ob, ref_b))
    tf = timeit(lambda: ops.aggregate(graph.fwd, x, row_div=graph.in_deg_f))
    tb = timeit(lambda: ops.aggregate(graph.bwd, gy))
    print(f"pack_short={pack}: fwd {tf*1e3:.1f} us  bwd {tb*1e3:.1f} us", flush=True)

"""Kernel-time breakdown of ONE partition out of P on a single GPU (LocalWorld): what a rank of a P-GPU run launches.

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv \
        python tools/local_profile.py 8
All P simulated ranks run (so that halo messages exist); the CSV therefore holds P copies of every kernel.
"""
import os, sys
from pathlib import Path
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import bench
from pipegcn_b200.partition import build_layouts
from pipegcn_b200.synthetic import SHAPES, make_graph, random_partition
from pipegcn_b200.train import LocalTrainer
from pipegcn_b200.world import LocalWorld

P = int(sys.argv[1]) if len(sys.argv) > 1 else 8
w = bench.WORKLOADS["rmat-1m"]
g = make_graph(w["shape"], device="cuda")
part = random_partition(g.n_nodes, P, seed=1, device="cuda")
layouts = build_layouts(g, part, P)
eargs = bench.engine_args(w, g.n_feat, SHAPES[w["shape"]]["n_class"], int(g.train_mask.sum().item()), P, 0.5)
del g
trainer = LocalTrainer(layouts, eargs, LocalWorld(P, "cuda"))
for _ in range(3):
    trainer.run_epoch()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
trainer.run_epoch()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done", flush=True)

"""Multi-GPU parity: one process per GPU (torchrun), DistWorld + CUDA IPC + NVLink pushes, against the CPU oracle.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/dist_parity.py
Every rank builds the same seeded tiny graph, runs the oracle world on the host (threads) and its own rank of the
CUDA engine, and compares logits / losses / reduced gradients per epoch in the four exchange modes.
"""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch
import torch.distributed as dist

MODES = {
    "sync": dict(),
    "sync_corr": dict(feat_corr=True, grad_corr=True, corr_momentum=0.9),
    "pipeline": dict(enable_pipeline=True),
    "pipeline_corr": dict(enable_pipeline=True, feat_corr=True, grad_corr=True, corr_momentum=0.95),
}


def main():
    rank, size, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    dist.init_process_group("nccl", rank=rank, world_size=size, device_id=dev)
    from oracle.train import initial_state, run_world
    from pipegcn_b200.helper.feature_buffer import Buffer
    from pipegcn_b200.helper.reducer import Reducer
    from pipegcn_b200.train import RankEngine
    from pipegcn_b200.world import DistWorld
    from tests.helpers import make_args, small_world

    shape = sys.argv[1] if len(sys.argv) > 1 else "tiny"
    n_class = 5 if shape == "tiny" else 16
    g, _, layouts, setups = small_world(shape, size)
    use_graph = os.environ.get("PG_PARITY_GRAPH", "0") == "1"       # epochs >= 3 replayed from CUDA graphs
    n_epochs = 6 if use_graph else 4
    ok = True
    only = os.environ.get("PG_PARITY_MODES")
    for mode, kw in MODES.items():
        if only and mode not in only.split(","):
            continue
        oargs, eargs = make_args(g, n_class, n_epochs=n_epochs, **kw)
        eargs.cuda_graph = use_graph
        init = initial_state(oargs)
        traces = run_world(setups, oargs, init_state=init)
        world = DistWorld(device=dev)
        eng = RankEngine(layouts[rank], eargs, world, init_state=init, seg_len=32)
        eng.keep_logits = True
        worst = 0.0
        for e in range(n_epochs):
            if use_graph and e == 3:
                eng.capture()
            eng.model.load_state_dict(traces[0].states[e])      # teacher forcing, see tests/test_engine_gpu.py
            loss = eng.run_epoch()
            eng.buffer.check_status()
            eng.buffer.timer.clear()
            ref = traces[rank]
            d = (eng.last_logits.float().cpu() - ref.logits[e]).abs().max().item()
            worst = max(worst, d / max(ref.logits[e].abs().max().item(), 1e-6))
            dl = abs(float(loss.item()) - ref.losses[e]) / abs(ref.losses[e])
            for n, p in eng.model.named_parameters():
                gd = (p.grad.float().cpu() - ref.grads[e][n]).abs().max().item()
                worst = max(worst, gd / max(ref.grads[e][n].abs().max().item(), 1e-6))
            worst = max(worst, dl)
            if os.environ.get("PG_PARITY_VERBOSE") and rank == 0:
                print(f"[dist_parity]   mode={mode} epoch {e} graph={int(eng.graphs is not None)} logits err {d:.3e} loss rel {dl:.3e} worst {worst:.3e}", flush=True)
        eng.buffer.synchronize()
        torch.cuda.synchronize()
        flag = torch.tensor([worst], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        if rank == 0:
            good = flag.item() < 2e-3
            ok = ok and good
            print(f"[dist_parity] {shape} P={size} graph={int(use_graph)} mode={mode:14s} worst relative error {flag.item():.3e} "
                  f"{'OK' if good else 'FAIL'}", flush=True)
        eng.graphs = None
        del eng
        torch.cuda.synchronize()
        dist.barrier()
    if rank == 0:
        print("[dist_parity] " + ("ALL OK" if ok else "FAILED"), flush=True)
    sys.stdout.flush()
    torch.cuda.synchronize()
    os._exit(0 if (ok or rank != 0) else 1)


if __name__ == "__main__":
    main()

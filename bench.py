"""Headline benchmark: epochs/s of full-graph partitioned GraphSAGE training (BASELINE.json metric).

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the CPU path (oracle port of the reference) on the host cores

Workload (BASELINE.json configs[1]): synthetic RMAT 1 M nodes / 20 M edges, one partition per GPU
(random partition), 3-layer GraphSAGE, hidden 256, bf16 activations, --enable-pipeline.  A step is one
epoch: forward, loss, backward (with the gradient halo exchange), gradient all-reduce, Adam step.
After the warm-up a short Python-launched region is instrumented with CUDA events (roofline of the aggregate kernel,
exposed communication); the timed region proper replays whole epochs from CUDA graphs (`--no-graph`: launches every
kernel from Python).  One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

WORKLOADS = {
    "rmat-1m": dict(shape="rmat-1m", n_layers=3, n_hidden=256, dtype="bf16", enable_pipeline=True,
                    feat_corr=False, grad_corr=False,
                    desc="RMAT 1M nodes / 20M edges, F=256, C=64, 3-layer GraphSAGE hidden 256, bf16, --enable-pipeline"),
    "reddit-shaped": dict(shape="reddit-shaped", n_layers=4, n_hidden=256, dtype="fp32", enable_pipeline=True,
                          feat_corr=True, grad_corr=True,
                          desc="Reddit-shaped RMAT 233K nodes / 115M edges, F=602, C=41, 4-layer GraphSAGE hidden 256, "
                               "fp32, --enable-pipeline --feat-corr --grad-corr"),
    "small": dict(shape="small", n_layers=3, n_hidden=64, dtype="bf16", enable_pipeline=True,
                  feat_corr=False, grad_corr=False, desc="20K-node RMAT (debug)"),
}


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", choices=["ours", "reference"], default="ours")
    p.add_argument("--workload", default="rmat-1m", choices=list(WORKLOADS))
    p.add_argument("--dropout", type=float, default=0.5, help="reference default (helper/parser.py:14)")
    p.add_argument("--cpu-scale", type=int, default=16, help="CPU baseline runs on a 1/cpu-scale graph")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--use-pp", action="store_true", help="--use-pp of the reference scripts: layer-0 aggregate precomputed once")
    p.add_argument("--ncu-region", action="store_true", help="cudaProfilerStart/Stop around the timed steps")
    p.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of replaying CUDA graphs")
    return p.parse_args()


def engine_args(w, g, n_class, n_parts, dropout, cuda_graph=False, use_pp=False):
    return argparse.Namespace(
        model="graphsage", backend="nccl", dtype=w["dtype"], n_layers=w["n_layers"], n_hidden=w["n_hidden"],
        n_linear=0, n_feat=g.n_feat, n_class=n_class, n_train=int(g.train_mask.sum().item()), dropout=dropout,
        norm="layer", lr=1e-2, weight_decay=0.0, use_pp=use_pp, enable_pipeline=w["enable_pipeline"],
        feat_corr=w["feat_corr"], grad_corr=w["grad_corr"], corr_momentum=0.95, seed=0, n_epochs=0,
        log_every=10, n_partitions=n_parts, cuda_graph=cuda_graph)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) >= 7:
                    self.rows.append(parts)
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=5)
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": float(self.rows[0][1]) if self.rows else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        d = json.loads(f.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------- CPU arm
def cpu_reference_run(w, args, steps, warmup):
    """The reference's algorithm on the host cores: oracle port, one partition, bounded 1/scale sample."""
    from oracle import dglpart
    from oracle import setup as osetup
    from oracle.fabric import ThreadFabric
    from oracle.train import OracleArgs, run_rank
    from pipegcn_b200.synthetic import SHAPES, make_graph
    spec = dict(SHAPES[w["shape"]])
    scale = max(1, args.cpu_scale)
    spec["n_nodes"] //= scale
    spec["n_edges"] //= scale
    from oracle.cbuild import set_threads
    n_cpu = os.cpu_count() or 1
    g = make_graph(spec, device="cpu")
    part = torch.zeros(g.n_nodes, dtype=torch.int64)
    parts = dglpart.partition_graph(g.n_nodes, g.src, g.dst, part, 1, g.feat, g.label, g.train_mask)
    setups = osetup.setup_world(parts)
    n_ep = warmup + steps
    oargs = OracleArgs(n_layers=w["n_layers"], n_hidden=w["n_hidden"], n_feat=g.n_feat, n_class=spec["n_class"],
                       n_train=int(g.train_mask.sum()), dropout=args.dropout, n_epochs=n_ep,
                       enable_pipeline=w["enable_pipeline"], feat_corr=w["feat_corr"], grad_corr=w["grad_corr"])
    # all the host threads it can USE: on many-core hosts the fastest setting is below the core count, so one
    # epoch is timed at a few thread counts first and the fastest is kept
    cand = sorted({t for t in (8, 16, 32, 64, n_cpu) if t <= n_cpu})
    probe = OracleArgs(**{**oargs.__dict__, "n_epochs": 2})
    best = None
    for t in cand:
        set_threads(t)
        w1 = run_rank(setups[0], probe, ThreadFabric(1), keep_trace=False).wall[-1]
        if best is None or w1 < best[0]:
            best = (w1, t)
    cores = best[1]
    set_threads(cores)
    tr = run_rank(setups[0], oargs, ThreadFabric(1), keep_trace=False)
    per_epoch = sum(tr.wall[warmup:]) / max(len(tr.wall[warmup:]), 1)
    eps_sample = 1.0 / max(per_epoch, 1e-9)
    return {
        "value": eps_sample / scale, "unit": "epochs/s", "cores": cores, "kind": "port",
        "sample": (f"oracle (CPU port of the reference, torch {torch.__version__}, {cores} of {n_cpu} host threads -- the "
                   f"fastest of {cand}) on a 1/{scale}-scale "
                   f"graph of the same shape ({g.n_nodes} nodes, {g.n_edges} edges, 1 partition), {steps} epochs after "
                   f"{warmup} warm-up: {eps_sample:.3f} epochs/s on the sample; value = that / {scale} "
                   f"(linear-in-edges equivalent for the full graph)"),
        "sample_epochs_per_s": eps_sample,
    }


def main():
    args = parse()
    w = WORKLOADS[args.workload]
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return
        steps, warmup = max(1, min(args.steps, 4)), max(1, min(args.warmup, 2))
        cb = cpu_reference_run(w, args, steps, warmup)
        line = {"impl": "reference", "metric": "epochs_per_sec", "value": cb["value"], "unit": "epochs/s",
                "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": 1e3 / cb["value"],
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": f"{args.workload}: {w['desc']}", "partitions": 1, "dropout": args.dropout},
                "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": "epochs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return

    assert torch.cuda.is_available(), "bench.py needs a GPU (use gpurun)"
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    import torch.distributed as dist
    from pipegcn_b200 import _C, ops
    from pipegcn_b200.partition import PartitionPlan
    from pipegcn_b200.synthetic import SHAPES, make_graph, random_partition
    from pipegcn_b200.train import RankEngine
    from pipegcn_b200.world import DistWorld, LocalWorld

    if world_size > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world_size, device_id=dev)
        world = DistWorld(device=dev)
    else:
        world = LocalWorld(1, dev).view(0)
    assert args.gpus == world_size, f"--gpus {args.gpus} but WORLD_SIZE={world_size}"

    t_setup = time.time()
    g = make_graph(w["shape"], device=dev)
    n_class = SHAPES[w["shape"]]["n_class"]
    part = random_partition(g.n_nodes, world_size, seed=1, device=dev)
    if world_size > 1:   # every rank built the graph itself; make sure they agree
        chk = torch.stack([g.src.sum(), g.dst.sum(), part.sum(), g.train_mask.sum()]).to(torch.float64)
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert torch.equal(lo, hi), "ranks generated different synthetic graphs"
    layout = PartitionPlan(g, part, world_size).build(rank)
    eargs = engine_args(w, g, n_class, world_size, args.dropout, cuda_graph=not args.no_graph, use_pp=args.use_pp)
    engine = RankEngine(layout, eargs, world)
    n_nodes, n_edges = g.n_nodes, g.n_edges
    feat_host = layout.feat.to(engine.dtype).cpu().pin_memory()
    label_host = engine.labels.cpu().pin_memory()
    del g
    torch.cuda.empty_cache()
    setup_s = time.time() - t_setup

    def barrier():
        torch.cuda.synchronize()
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n_steps, fn):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n_steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world_size > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def step():
        engine.buffer.timer.clear()
        engine.run_epoch()

    def step_e2e():
        engine.buffer.timer.clear()
        engine.set_features(feat_host)
        engine.labels.copy_(label_host, non_blocking=True)
        loss = engine.run_epoch()
        return float(loss.item())      # device -> host read of the step's result

    for _ in range(max(args.warmup, 3)):
        step()
    engine.buffer.check_status()

    # ---- instrumented eager epochs: every aggregate launch bracketed by CUDA events (roofline), the flag waits
    #      bracketed by CUDA events (exposed communication); kernels are launched from Python here
    ops.PROFILE = []
    _C.LAUNCHES = 0
    comm_s = []
    n_eager = args.steps if args.no_graph else min(args.steps, 5)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()

    def step_prof():
        engine.buffer.timer.clear()
        engine.run_epoch()
        comm_s.append(engine.buffer.timer)      # events resolved after the region
        engine.buffer.timer = type(engine.buffer.timer)()

    if args.ncu_region:
        torch.cuda.cudart().cudaProfilerStart()
    ms_eager_total = timed(n_eager, step_prof)
    if args.ncu_region:
        torch.cuda.cudart().cudaProfilerStop()
    launches_per_step = _C.LAUNCHES / max(n_eager, 1)
    prof, ops.PROFILE = ops.PROFILE, None
    engine.buffer.check_status()
    exposed = [t.tot_time() for t in comm_s]
    exposed_s = sum(exposed) / max(len(exposed), 1)
    ms_eager = ms_eager_total / n_eager

    # ---- timed region proper: K epochs replayed from CUDA graphs (one graph per epoch parity)
    graph_info = {"enabled": False}
    if not args.no_graph:
        try:
            engine.capture()
            for _ in range(2):
                step()
            graph_info = {"enabled": True, "graphs": len(engine.graphs)}
        except Exception as e:   # noqa: BLE001  -- same kernels either way; only the launch mechanism differs
            engine.graphs = None
            engine.buffer.graph_mode = False
            ops.STEP_DEV = None
            graph_info = {"enabled": False, "error": f"{type(e).__name__}: {str(e)[:200]}"}
            print(f"[bench] CUDA graph capture failed, timing eager launches: {e}", file=sys.stderr)
    if graph_info["enabled"]:
        ms_total = timed(args.steps, step)
    elif args.no_graph:
        ms_total = ms_eager_total * args.steps / n_eager
    else:
        ms_total = timed(args.steps, step)
    engine.buffer.check_status()
    launches = int(round(launches_per_step * args.steps))
    ms_step = ms_total / args.steps
    value = 1e3 / ms_step

    # ---- roofline of the dominant kernel (the forward+backward neighbour aggregate, HBM bound)
    hbm_peak, peak_src = peaks()
    agg_ms = sum(a.elapsed_time(b) for a, b, _ in prof)
    agg_bytes = sum(nb for _, _, nb in prof)
    n_agg = len(prof)
    achieved = (agg_bytes / 1e9) / (agg_ms / 1e3) if agg_ms > 0 else 0.0
    traffic = None
    tf = ROOT / "profiles" / "agg_traffic.json"
    if tf.exists():       # per-launch DRAM bytes of the aggregate kernel from the committed `ncu --set full` capture
        traffic = json.loads(tf.read_text()).get(f"{args.workload}:{world_size}")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": traffic, "kernel": "pg::agg_kernel (+fix-up), forward and backward aggregate",
                "launches": n_agg, "avg_launch_ms": agg_ms / max(n_agg, 1),
                "algorithmic_bytes_per_launch": agg_bytes / max(n_agg, 1),
                "share_of_step": agg_ms / ms_eager_total if ms_eager_total else None, "peak_source": peak_src,
                "measured_in": f"{n_eager} eager (Python-launched) epochs before the graph-replayed timed region"}

    # ---- end to end: host buffers, H2D of the step's inputs and D2H of its loss inside the timed region
    e2e = None
    if not args.no_e2e:
        step_e2e()
        ms_e2e = timed(args.steps, step_e2e) / args.steps
        e2e = {"value": 1e3 / ms_e2e, "unit": "epochs/s",
               "h2d_bytes_per_step": int(feat_host.numel() * feat_host.element_size() + label_host.numel() * 8),
               "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e}

    clocks = sampler.stop() if sampler else None      # sampled across the eager, replayed and end-to-end regions
    cb = None
    if rank == 0 and world_size == 1 and not args.no_cpu_baseline:
        cb = cpu_reference_run(w, args, steps=3, warmup=2)

    if rank == 0:
        line = {
            "metric": "epochs_per_sec", "value": value, "unit": "epochs/s", "n_gpus": world_size,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": w["dtype"], "data": "synthetic",
            "config": {"workload": f"{args.workload}: {w['desc']}", "n_nodes": n_nodes, "n_edges": n_edges,
                       "partitions": world_size, "partition_method": "random", "dropout": args.dropout, "use_pp": args.use_pp,
                       "n_in_rank0": layout.num_in, "halo_rank0": layout.num_all - layout.num_in,
                       "nnz_rank0": layout.nnz, "linear": ops.LINEAR_IMPL,
                       "l2": "per-epoch working set (features, activations, indices) exceeds the 126 MB L2; no flush",
                       "setup_s": round(setup_s, 1)},
            "exposed_comm_s_per_epoch": exposed_s,
            "exposed_comm_frac": exposed_s / (ms_eager / 1e3) if ms_eager else None,
            "eager_ms_per_step": ms_eager, "cuda_graph": graph_info,
            "roofline": roofline, "e2e": e2e, "cpu_baseline": cb, "gpu_launches": launches, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    sys.stdout.flush()
    engine.graphs = None            # drop captured graphs (they hold NCCL kernels) before the communicator goes away
    torch.cuda.synchronize()
    if world_size > 1:
        dist.barrier()
        torch.cuda.synchronize()
    os._exit(0)                     # skip NCCL communicator teardown (can block behind captured graphs)


if __name__ == "__main__":
    main()

"""GPU parity on the BASELINE.json configurations themselves (scaled to sizes the CPU oracle finishes in seconds):
same generator, same degree skew, same feature/hidden widths, same flags, same dtype -- per-layer exchange buffers and
layer outputs, logits, loss and reduced gradients against the oracle, at the tolerances SURVEY.md §8c states:

    fp32 exchange (f_buf of layer 0)    rtol 3e-5            (pure exchange / EMA; deeper f_buf carry a GEMM + LayerNorm)
    fp32 layer outputs / logits         rtol 2e-4, atol 2e-4 (3xTF32 tensor-core product vs MKL sgemm)
    fp32 loss                           rel 1e-4
    bf16 per-layer outputs / logits     rtol 2e-2, atol 1e-2 * scale  (scale = max |ref| of the tensor: LayerNorm'd
                                        activations are O(1), logits O(1); see test for the measured margins)
    bf16 loss                           rel 1e-2
    free-running k epochs, final loss   rel 1e-4 (fp32) / 1e-2 (bf16)
    fp32 reduced gradients              rtol 2e-3, atol 3e-2 * max|ref|: a weight gradient is a sum of 3e4-1e5 signed
                                        terms with heavy cancellation, which amplifies the ~1e-4 relative error the
                                        upstream gradient has picked up through three 3xTF32 layers (measured: up to 1.4 %
                                        of the largest entry on cfg2 at 1/32 scale, 4 partitions)
"""
import argparse

import pytest
import torch

pytestmark = pytest.mark.gpu

# BASELINE.json configs[0]: Reddit 2-partition GraphSAGE n-layers=2 n-hidden=128 (reference path: fp32, synchronous)
CFG1 = dict(n_nodes=233_000 // 32, n_edges=115_000_000 // 32, n_feat=602, n_class=41, train_frac=0.66)
# configs[1]: RMAT 1M/20M, 3-layer hidden 256, --enable-pipeline (bf16 storage in the engine)
CFG2 = dict(n_nodes=1_000_000 // 32, n_edges=20_000_000 // 32, n_feat=256, n_class=64, train_frac=0.66)


def _world(spec, n_parts):
    from tests.helpers import small_world
    return small_world(spec, n_parts)


def _hooks(trainer):
    caps = [dict() for _ in trainer.engines]
    for r, eng in enumerate(trainer.engines):
        for i, layer in enumerate(eng.model.layers):
            def hook(mod, inp, out, r=r, i=i):
                # no host synchronisation inside the forward: in the synchronous exchange mode a rank's stream waits
                # for pushes the host has not launched yet (LocalWorld: one host thread for all ranks)
                caps[r][i] = ((inp[1] if len(inp) > 1 else inp[0]).detach().clone(), out.detach().clone())
            layer.register_forward_hook(hook)
    return caps


def _run(spec, n_parts, n_epochs, dtype, free_running=False, **flags):
    from oracle.train import initial_state, run_world
    from pipegcn_b200.train import LocalTrainer
    from pipegcn_b200.world import LocalWorld
    from tests.helpers import make_args
    g, _, layouts, setups = _world(spec, n_parts)
    oargs, eargs = make_args(g, spec["n_class"], n_epochs=n_epochs, **flags)
    eargs.dtype = dtype
    init = initial_state(oargs)
    traces = run_world(setups, oargs, init_state=init)
    trainer = LocalTrainer(layouts, eargs, LocalWorld(n_parts, "cuda"), init_state=init)
    caps = _hooks(trainer)
    out = []
    for e in range(n_epochs):
        if not free_running:
            for eng in trainer.engines:
                eng.model.load_state_dict(traces[0].states[e])
        losses = trainer.run_epoch(keep_logits=True)
        out.append(dict(loss=[float(l.item()) for l in losses],
                        logits=[en.last_logits.float().cpu() for en in trainer.engines],
                        layers=[{i: (a.float().cpu(), b.float().cpu()) for i, (a, b) in c.items()} for c in caps],
                        grads=[{n: p.grad.detach().float().cpu().clone() for n, p in en.model.named_parameters()}
                               for en in trainer.engines]))
    return traces, out


def _close(got, ref, rtol, atol_frac, what):
    scale = ref.abs().max().item()
    bad = (got - ref).abs() > atol_frac * scale + rtol * ref.abs()
    assert not bool(bad.any()), (f"{what}: {int(bad.sum())}/{bad.numel()} outside rtol {rtol} atol {atol_frac}*{scale:.3g}; "
                                 f"max err {(got - ref).abs().max().item():.3e}")


def test_cfg1_reddit_shaped_p2_fp32_per_layer():
    """configs[0] at 1/32 scale: 7 281 nodes, 3.6 M edges (mean in-degree ~490, hub rows of > 4 000 terms), F=602."""
    traces, out = _run(CFG1, 2, 3, "fp32", n_layers=2, n_hidden=128)
    for e, ep in enumerate(out):
        for r in range(2):
            for i, rec in traces[r].layers[e].items():
                _close(ep["layers"][r][i][0], rec["f_buf"], 3e-5 if i == 0 else 2e-4, 1e-6 if i == 0 else 2e-4,
                       f"epoch {e} rank {r} f_buf[{i}]")
                _close(ep["layers"][r][i][1], rec["layer_out"], 2e-4, 2e-4, f"epoch {e} rank {r} layer_out[{i}]")
            _close(ep["logits"][r], traces[r].logits[e], 2e-4, 2e-4, f"epoch {e} rank {r} logits")
            assert abs(ep["loss"][r] - traces[r].losses[e]) <= 1e-4 * abs(traces[r].losses[e])
            for n, gref in traces[r].grads[e].items():
                _close(ep["grads"][r][n], gref, 2e-3, 3e-2, f"epoch {e} rank {r} grad {n}")


@pytest.mark.parametrize("n_parts", [1, 4])
def test_cfg2_rmat_fp32_per_layer(n_parts):
    """configs[1] at 1/32 scale, fp32 activations: 31 250 nodes, 625 K edges, F=256, 3 layers x 256, pipelined."""
    traces, out = _run(CFG2, n_parts, 3, "fp32", n_layers=3, n_hidden=256, enable_pipeline=True)
    for e, ep in enumerate(out):
        for r in range(n_parts):
            for i, rec in traces[r].layers[e].items():
                _close(ep["layers"][r][i][0], rec["f_buf"], 3e-5 if i == 0 else 2e-4, 1e-6 if i == 0 else 2e-4,
                       f"epoch {e} rank {r} f_buf[{i}]")
                _close(ep["layers"][r][i][1], rec["layer_out"], 2e-4, 2e-4, f"epoch {e} rank {r} layer_out[{i}]")
            _close(ep["logits"][r], traces[r].logits[e], 2e-4, 2e-4, f"epoch {e} rank {r} logits")
            assert abs(ep["loss"][r] - traces[r].losses[e]) <= 1e-4 * abs(traces[r].losses[e])
            for n, gref in traces[r].grads[e].items():
                _close(ep["grads"][r][n], gref, 2e-3, 3e-2, f"epoch {e} rank {r} grad {n}")


@pytest.mark.parametrize("n_parts", [1, 4])
def test_cfg2_rmat_bf16_per_layer(n_parts):
    """The headline dtype at SURVEY §8c's bf16 tolerance (rtol 2e-2, atol 1e-2 of the tensor's scale, loss rel 1e-2),
    per-layer: exchange buffer, layer output, logits."""
    traces, out = _run(CFG2, n_parts, 3, "bf16", n_layers=3, n_hidden=256, enable_pipeline=True)
    for e, ep in enumerate(out):
        for r in range(n_parts):
            for i, rec in traces[r].layers[e].items():
                _close(ep["layers"][r][i][0], rec["f_buf"], 2e-2, 1e-2, f"epoch {e} rank {r} f_buf[{i}]")
                _close(ep["layers"][r][i][1], rec["layer_out"], 2e-2, 1e-2, f"epoch {e} rank {r} layer_out[{i}]")
            _close(ep["logits"][r], traces[r].logits[e], 2e-2, 1e-2, f"epoch {e} rank {r} logits")
            assert abs(ep["loss"][r] - traces[r].losses[e]) <= 1e-2 * abs(traces[r].losses[e])


@pytest.mark.parametrize("dtype,tol", [("fp32", 1e-4), ("bf16", 1e-2)])
def test_free_running_final_loss(dtype, tol):
    """No teacher forcing: 10 epochs of the engine's own Adam trajectory (pipeline + feat/grad correction, 2 ranks)
    against the oracle's; the north-star's 'final loss' comparison."""
    spec = dict(n_nodes=20_000, n_edges=400_000, n_feat=64, n_class=16, train_frac=0.66)
    traces, out = _run(spec, 2, 10, dtype, free_running=True, n_layers=3, n_hidden=64, enable_pipeline=True,
                       feat_corr=True, grad_corr=True)
    for r in range(2):
        ref, got = traces[r].losses[-1], out[-1]["loss"][r]
        assert abs(got - ref) <= tol * abs(ref), f"rank {r}: final loss {got} vs oracle {ref} ({dtype})"
    # and the whole trajectory stays together
    for e in range(10):
        for r in range(2):
            assert abs(out[e]["loss"][r] - traces[r].losses[e]) <= 5 * tol * abs(traces[r].losses[e])


def test_n_linear_tail():
    """`--n-linear 1` (model.py:49-51): the last layer is a plain nn.Linear without exchange."""
    spec = dict(n_nodes=3_000, n_edges=40_000, n_feat=48, n_class=7, train_frac=0.66)
    traces, out = _run(spec, 2, 3, "fp32", n_layers=3, n_hidden=32, n_linear=1, enable_pipeline=True)
    for e, ep in enumerate(out):
        for r in range(2):
            _close(ep["logits"][r], traces[r].logits[e], 2e-4, 2e-4, f"epoch {e} rank {r} logits")
            assert abs(ep["loss"][r] - traces[r].losses[e]) <= 1e-4 * abs(traces[r].losses[e])
            for n, gref in traces[r].grads[e].items():
                _close(ep["grads"][r][n], gref, 2e-3, 3e-2, f"epoch {e} rank {r} grad {n}")


def _engine_vs_oracle(g, part, n_parts, n_class, n_epochs=3, **flags):
    from oracle import dglpart
    from oracle import setup as osetup
    from oracle.train import initial_state, run_world
    from pipegcn_b200.partition import build_layouts
    from pipegcn_b200.train import LocalTrainer
    from pipegcn_b200.world import LocalWorld
    from tests.helpers import make_args
    layouts = build_layouts(g, part, n_parts)
    setups = osetup.setup_world(dglpart.partition_graph(g.n_nodes, g.src, g.dst, part, n_parts, g.feat, g.label,
                                                        g.train_mask))
    oargs, eargs = make_args(g, n_class, n_epochs=n_epochs, **flags)
    init = initial_state(oargs)
    traces = run_world(setups, oargs, init_state=init)
    trainer = LocalTrainer(layouts, eargs, LocalWorld(n_parts, "cuda"), init_state=init)
    for e in range(n_epochs):
        for eng in trainer.engines:
            eng.model.load_state_dict(traces[0].states[e])
        losses = trainer.run_epoch(keep_logits=True)
        for r, eng in enumerate(trainer.engines):
            _close(eng.last_logits.float().cpu(), traces[r].logits[e], 2e-4, 2e-4, f"epoch {e} rank {r} logits")
            assert abs(float(losses[r].item()) - traces[r].losses[e]) <= 1e-4 * abs(traces[r].losses[e])
            for n, p in eng.model.named_parameters():
                _close(p.grad.float().cpu(), traces[r].grads[e][n], 2e-3, 3e-2, f"epoch {e} rank {r} grad {n}")


def test_metis_partitioned_engine():
    """`--partition-method metis` (parser.py:39-41) end to end: unequal parts, small halos."""
    from pipegcn_b200.metis import metis_partition
    from pipegcn_b200.synthetic import make_graph
    spec = dict(n_nodes=6_000, n_edges=90_000, n_feat=32, n_class=9, train_frac=0.66)
    g = make_graph(spec)
    part = metis_partition(g, 3, objtype="vol")
    assert part.unique().numel() == 3
    _engine_vs_oracle(g, part, 3, 9, n_hidden=32, enable_pipeline=True, feat_corr=True, grad_corr=True)


def test_inductive_engine():
    """`--inductive` (main.py:34-35): the engine trains on the train-node subgraph."""
    from pipegcn_b200.synthetic import make_graph, random_partition, train_subgraph
    spec = dict(n_nodes=6_000, n_edges=90_000, n_feat=32, n_class=9, train_frac=0.5)
    sub = train_subgraph(make_graph(spec))
    assert bool(sub.train_mask.all()) and sub.n_nodes < 6_000
    part = random_partition(sub.n_nodes, 2)
    _engine_vs_oracle(sub, part, 2, 9, n_hidden=32, enable_pipeline=True)


@pytest.mark.parametrize("use_pp", [False, True])
def test_eval_branch_matches_oracle(use_pp):
    """The eval branch (layer.py:52-62) on the homogeneous full graph, degrees taken from the graph: equal to the
    oracle's forward of the un-partitioned graph (one partition: no halo, in_deg == row length), incl. --use-pp
    (`cat(feat, ah)` -> one Linear, layer.py:58-60)."""
    import torch.nn.functional as F
    from oracle.train import initial_state, run_world
    from pipegcn_b200.evaluate import full_graph
    from pipegcn_b200.module.model import GraphSAGE
    from pipegcn_b200.partition import get_layer_size
    from tests.helpers import make_args, small_world
    spec = dict(n_nodes=4_000, n_edges=60_000, n_feat=40, n_class=6, train_frac=0.66)
    g, _, layouts, setups = small_world(spec, 1)
    oargs, eargs = make_args(g, 6, n_epochs=1, n_hidden=32, use_pp=use_pp, lr=0.0)
    init = initial_state(oargs)
    ref = run_world(setups, oargs, init_state=init)[0]
    model = GraphSAGE(get_layer_size(g.n_feat, 32, 6, 3), F.relu, use_pp, norm="layer", dropout=0.5).cuda()
    model.load_state_dict(init)
    model.eval()
    with torch.no_grad():
        logits = model(full_graph(g, "cuda"), g.feat.cuda())
    # the oracle's rows are in `move_train_first` order: map back through the layout
    order = layouts[0].inner_gid
    _close(logits.float().cpu()[order], ref.logits[0], 2e-4, 2e-4, "eval logits")


def test_train_eval_checkpoint_roundtrip(tmp_path, monkeypatch):
    """train.run with --eval on a graph with learnable (planted) labels: accuracy is evaluated on the GPU every
    log_every epochs, the best state_dict is saved under the reference's key names and loads into the oracle's model
    (= the reference's module tree) strictly; validation accuracy ends above chance."""
    import torch.nn.functional as F
    from oracle.model import OracleGraphSAGE
    from pipegcn_b200 import train
    from pipegcn_b200.helper import context as ctx
    from pipegcn_b200.helper.feature_buffer import Buffer
    from pipegcn_b200.helper.reducer import Reducer
    from pipegcn_b200.partition import build_layouts, get_layer_size
    from pipegcn_b200.synthetic import make_graph
    from pipegcn_b200.world import LocalWorld
    monkeypatch.chdir(tmp_path)
    spec = dict(n_nodes=3_000, n_edges=30_000, n_feat=24, n_class=4, train_frac=0.6)
    g = make_graph(spec, planted_labels=True)
    layout = build_layouts(g, torch.zeros(g.n_nodes, dtype=torch.int64), 1)[0]
    args = argparse.Namespace(
        model="graphsage", backend="nccl", dtype="fp32", n_layers=2, n_hidden=32, n_linear=0, n_feat=24, n_class=4,
        n_train=int(g.train_mask.sum()), dropout=0.1, norm="layer", lr=1e-2, weight_decay=0.0, use_pp=False,
        enable_pipeline=False, feat_corr=False, grad_corr=False, corr_momentum=0.95, seed=0, n_epochs=60, log_every=10,
        n_partitions=1, eval=True, inductive=False, dataset="synthetic:test", graph_name="roundtrip")
    world = LocalWorld(1, "cuda").view(0)
    ctx.buffer, ctx.reducer = Buffer(world), Reducer(world)
    eng = train.run(layout, args, world, eval_graph=g)
    assert eng.checkpoint_path == "model/roundtrip_final.pth.tar"
    state = torch.load(tmp_path / eng.checkpoint_path)
    ref_model = OracleGraphSAGE(get_layer_size(24, 32, 4, 2), F.relu, False, norm="layer", dropout=0.1)
    ref_model.load_state_dict(state, strict=True)
    assert eng.best_val_acc > 0.6, eng.best_val_acc             # 4 classes: chance = 0.25 (CPU oracle reaches 0.88)
    assert (tmp_path / "results").exists() and any((tmp_path / "results").iterdir())
    text = next((tmp_path / "results").iterdir()).read_text()
    assert "Validation Accuracy" in text and "Epoch 00009" in text

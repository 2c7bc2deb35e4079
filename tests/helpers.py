"""Shared builders for the parity tests: same seeded graph for the oracle and the CUDA engine."""
import argparse

import torch

from oracle import dglpart
from oracle import setup as osetup
from oracle.train import OracleArgs
from pipegcn_b200.partition import build_layouts
from pipegcn_b200.synthetic import make_graph, random_partition


def small_world(shape="tiny", n_parts=2, seed_graph=0):
    g = make_graph(shape, seed_graph=seed_graph)
    part = random_partition(g.n_nodes, n_parts)
    layouts = build_layouts(g, part, n_parts)
    parts = dglpart.partition_graph(g.n_nodes, g.src, g.dst, part, n_parts, g.feat, g.label, g.train_mask)
    setups = osetup.setup_world(parts)
    return g, part, layouts, setups


def make_args(g, n_class, **kw):
    base = dict(n_layers=3, n_hidden=16, n_linear=0, n_feat=g.n_feat, n_class=n_class,
                n_train=int(g.train_mask.sum()), dropout=0.0, norm="layer", lr=1e-2, weight_decay=0.0,
                use_pp=False, enable_pipeline=False, feat_corr=False, grad_corr=False, corr_momentum=0.95,
                seed=0, n_epochs=4, log_every=10)
    base.update(kw)
    oargs = OracleArgs(**base)
    eargs = argparse.Namespace(model="graphsage", backend="nccl", dtype="fp32", **base)
    return oargs, eargs

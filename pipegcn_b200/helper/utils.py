"""Data / partition helpers on the path into the hot loop.

Counterparts of /root/reference/helper/utils.py: `load_data` (:74-96),
`graph_partition` (:132-144), `load_partition` (:99-129), `get_layer_size` (:147-151).
The reference's datasets need DGL, ogb and a network; this engine accepts
`synthetic:<shape>` datasets (pipegcn_b200/synthetic.py) and, for the reference's
dataset names, the synthetic graph of the same shape -- saying so loudly.
`get_boundary` (:154-188) has no per-process counterpart: boundary lists come out of
the one-pass layout builder (pipegcn_b200/partition.py).
"""
import warnings

import torch

from ..partition import PartitionPlan, get_layer_size  # noqa: F401  (re-export)
from ..synthetic import SHAPES, make_graph, random_partition, train_subgraph

_ALIAS = {'reddit': 'reddit-shaped', 'ogbn-products': 'products-shaped'}
_cache = {}


def _shape_of(dataset: str) -> str:
    if dataset.startswith('synthetic:'):
        return dataset.split(':', 1)[1]
    if dataset in _ALIAS:
        warnings.warn(f"dataset '{dataset}' needs DGL and a download; using the synthetic graph "
                      f"'{_ALIAS[dataset]}' of the same shape")
        return _ALIAS[dataset]
    if dataset in SHAPES:
        return dataset
    raise ValueError('Unknown dataset: {}'.format(dataset))


def load_data(dataset, device='cpu'):
    """-> (GlobalGraph with one self loop per node, n_feat, n_class)  (utils.py:74-96)."""
    shape = _shape_of(dataset)
    key = (shape, str(device))
    if key not in _cache:
        _cache[key] = make_graph(shape, device=device)
    g = _cache[key]
    return g, g.n_feat, SHAPES[shape]['n_class']


def graph_partition(g, args):
    """Node -> partition assignment (utils.py:132-144)."""
    if args.partition_method == 'random':
        return random_partition(g.n_nodes, args.n_partitions, seed=1, device=g.src.device)
    from ..metis import metis_partition
    return metis_partition(g, args.n_partitions, objtype=args.partition_obj)


def load_partition(args, rank, device=None):
    """This rank's `PartitionLayout` (utils.py:99-129 + the set-up half of train.run)."""
    device = device if device is not None else (f'cuda:{torch.cuda.current_device()}' if torch.cuda.is_available() else 'cpu')
    g, n_feat, n_class = load_data(args.dataset, device=device)
    args.n_feat, args.n_class = n_feat, n_class
    args.n_train = int(g.train_mask.sum().item())
    if getattr(args, 'inductive', False):          # main.py:34-35: partition the train-node subgraph
        g = train_subgraph(g)
    part = graph_partition(g, args)
    return PartitionPlan(g, part, args.n_partitions).build(rank)

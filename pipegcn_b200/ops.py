"""torch-facing wrappers of the C-ABI kernels (PyTorch tensors in, PyTorch tensors out).

Each autograd Function mirrors one step of /root/reference/module/layer.py:44-51 and
its autograd; nothing here computes on the host or falls back to torch kernels for
the aggregate / exchange path.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _C
from .graph import CsrPlan, PartGraph, alloc_rows


def _rows(t: torch.Tensor):
    if t.dim() != 2 or (t.shape[1] > 1 and t.stride(1) != 1):
        raise _C.PgError(f"expected a row-major 2-D tensor, got shape {tuple(t.shape)} strides {t.stride()}")
    if not t.is_cuda:
        raise _C.PgError("the hot path runs on CUDA tensors only (no CPU fall-back)")
    return t


def aggregate(plan: CsrPlan, x: torch.Tensor, out: torch.Tensor = None, row_div: torch.Tensor = None,
              acc_rows: int = 0) -> torch.Tensor:
    """out[r] = sum_{e in row r} x[indices[e]] (/ row_div[r]) (+ out[r] for r < acc_rows)."""
    _rows(x)
    d = x.shape[1]
    if out is None:
        out = alloc_rows(plan.n_rows, d, x.dtype, x.device)
    _rows(out)
    assert out.shape[0] == plan.n_rows and out.shape[1] == d and out.dtype == x.dtype
    scratch = plan.scratch(d)
    prof = PROFILE
    if prof is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    _C.count(2 if plan.n_long else 1)
    _C.check(_C.lib.pg_aggregate(C.byref(plan.c), x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), d,
                                 _C.dtype_code(x.dtype), row_div.data_ptr() if row_div is not None else None,
                                 int(acc_rows), scratch.data_ptr() if scratch is not None else None,
                                 _C.stream_ptr()), "pg_aggregate")
    if prof is not None:
        e1.record()
        prof.append((e0, e1, aggregate_bytes(plan, x, row_div is not None)))
    return out


# when a list, every aggregate launch appends (start event, end event, algorithmic bytes)
PROFILE = None
LINEAR_IMPL = "cublas (torch.nn.functional.linear); tcgen05 kernel pending"


def aggregate_bytes(plan: CsrPlan, x: torch.Tensor, has_div: bool) -> int:
    """ALGORITHMIC bytes of one launch: column ids + row pointers + divisor +
    every source row once + every output row once."""
    s, d = x.element_size(), x.shape[1]
    return 4 * plan.nnz + 4 * (plan.n_rows + 1) + (4 * plan.n_rows if has_div else 0) \
        + s * d * x.shape[0] + s * d * plan.n_rows


def row_div(x: torch.Tensor, div: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
    _rows(x)
    if out is None:
        out = alloc_rows(x.shape[0], x.shape[1], x.dtype, x.device)
    _C.count()
    _C.check(_C.lib.pg_row_div(x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), x.shape[0], x.shape[1],
                               _C.dtype_code(x.dtype), div.data_ptr(), _C.stream_ptr()), "pg_row_div")
    return out


class SageAggregate(torch.autograd.Function):
    """ah = (A @ feat) / in_deg   (layer.py:47-50) and its gradient  g_feat = A^T @ (g_ah / in_deg)."""

    @staticmethod
    def forward(ctx, feat, graph: PartGraph, deg_f):
        ctx.graph, ctx.deg_f = graph, deg_f
        if feat.stride(1) != 1:
            feat = feat.contiguous()
        return aggregate(graph.fwd, feat, row_div=deg_f)

    @staticmethod
    def backward(ctx, g_ah):
        graph = ctx.graph
        if g_ah.stride(1) != 1:
            g_ah = g_ah.contiguous()
        gs = row_div(g_ah, ctx.deg_f)
        return aggregate(graph.bwd, gs), None, None


def sage_aggregate(feat: torch.Tensor, graph: PartGraph, deg_f: torch.Tensor = None) -> torch.Tensor:
    return SageAggregate.apply(feat, graph, graph.in_deg_f if deg_f is None else deg_f)


# ---- dense part.  Round-1 bring-up: library GEMM (cuBLAS through torch); the tcgen05 kernel
# ---- (csrc/linear_tcgen05.cu) replaces it behind the same two functions.
def linear(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor = None) -> torch.Tensor:
    w = weight if weight.dtype == x.dtype else weight.to(x.dtype)
    b = bias if bias is None or bias.dtype == x.dtype else bias.to(x.dtype)
    return torch.nn.functional.linear(x, w, b)


def sage_linear(x, ah, w1, b1, w2, b2) -> torch.Tensor:
    """x @ W1^T + b1 + ah @ W2^T + b2   (layer.py:51)."""
    return linear(x, w1, b1) + linear(ah, w2, b2)

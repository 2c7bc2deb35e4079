"""Does tcgen05 kind::tf32 truncate or round the 13 low mantissa bits of an fp32 operand?  If it truncates, the `hi` half of
the 3xTF32 split need not be materialised (the raw fp32 tensor IS the hi operand) and only `lo` has to be written."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from pipegcn_b200 import _C, ops
from pipegcn_b200.graph import alloc_rows

torch.manual_seed(0)
m, n, k = 4096, 256, 256
a = alloc_rows(m, k, torch.float32, "cuda"); a.copy_(torch.randn(m, k, device="cuda"))
b = alloc_rows(n, k, torch.float32, "cuda"); b.copy_(torch.randn(n, k, device="cuda"))
ah, al = ops.split_tf32(a)
bh, bl = ops.split_tf32(b)


def one(a_, b_):
    srcs = (_C.pg_gemm_src * 1)(_C.pg_gemm_src(a_.data_ptr(), a_.stride(0), b_.data_ptr(), b_.stride(0), k))
    out = alloc_rows(m, n, torch.float32, "cuda")
    _C.check(_C.lib.pg_linear(_C.PG_F32, _C.PG_F32, srcs, 1, None, None, out.data_ptr(), out.stride(0), m, n, _C.stream_ptr()))
    return out


hh, rr, rh = one(ah, bh), one(a, b), one(a, bh)
ref = (ah.double() @ bh.double().t())
print("hi*hi vs fp64(hi*hi):", (hh.double() - ref).abs().max().item())
print("raw*raw == hi*hi bitwise:", bool(torch.equal(hh, rr)), " max diff", (hh - rr).abs().max().item())
print("raw*hi  == hi*hi bitwise:", bool(torch.equal(hh, rh)))

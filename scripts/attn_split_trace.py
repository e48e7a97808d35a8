#!/usr/bin/env python3
"""Phase timeline of the split-precision tcgen05 attention kernel (CTA 0, first items), in SM cycles.
Needs a library whose attention_tc_split.cu was compiled with -DSVIT_ATT_TRACE (the instrumentation costs 20 % and is
compiled out by default):  nvcc ... -DSVIT_ATT_TRACE -c attention_tc_split.cu, relink, SVIT_LIB=<that .so>."""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shapley_vit_b200 import _lib, ops
from shapley_vit_b200._lib import check
lib = _lib.load()
buf = torch.zeros(24 * 16, dtype=torch.int64, device="cuda")
n_seq, T = 1024, 197
qkv = torch.randn(n_seq, T, 2304, device="cuda") * 0.5
q3 = ops.OperandArray.from_float(qkv, _lib.FMT_X3)
out = ops.OperandArray((n_seq, T, 768), torch.float16, _lib.FMT_C8, "cuda")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
call = lambda: check(lib.svit_attention_split(C.c_void_p(q3.ptr), C.c_void_p(out.ptr), _lib.FMT_C8, n_seq, T, 12, 64, st))
call(); torch.cuda.synchronize()
lib.svit_debug_attention_split_trace.argtypes = [C.c_void_p]
assert lib.svit_debug_attention_split_trace(C.c_void_p(buf.data_ptr())) == 0
call(); torch.cuda.synchronize()
lib.svit_debug_attention_split_trace(C.c_void_p(0))
t = buf.cpu().view(24, 16)
t0 = int(t[0][t[0] > 0].min())
names = ["g0 S", "g0 P", "g0 O", "g0 st", "g1 S", "g1 P", "g1 O", "g1 st", "m qk0", "m pv1", "m qk1", "m pv0"]
print("item " + " ".join(f"{n:>7s}" for n in names))
for i in range(24):
    print(f"{i:4d} " + " ".join(f"{(int(t[i][k]) - t0) if t[i][k] > 0 else -1:7d}" for k in range(12)))

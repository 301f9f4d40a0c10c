"""Timeline of the decoupled-issue attention kernel (G3C_ATTN_IMPL=v4).  Usage: G3C_ATTN_IMPL=v4 python tools/attn_trace_v4.py"""
import sys

import torch

sys.path.insert(0, ".")
from gen3c_b200 import _lib, ops  # noqa: E402

L, H = 56320, 32
q = (torch.randn(L, H * 128, device="cuda")).to(torch.bfloat16)
k = (torch.randn(L, H * 128, device="cuda")).to(torch.bfloat16)
vt = (torch.randn(H * 128, L, device="cuda")).to(torch.bfloat16)
ops.attention(q, k, vt, H)
torch.cuda.synchronize()
buf = torch.zeros(3 * 64 * 8, dtype=torch.int64, device="cuda")
lib = _lib.load()
_lib.check(lib.g3c_attn_set_trace(buf.data_ptr()), "set_trace")
ops.attention(q, k, vt, H)
torch.cuda.synchronize()
_lib.check(lib.g3c_attn_set_trace(None), "set_trace")
t = buf.cpu().view(3, 64, 8).double()
t0 = t[0, 8, 0]
names = {0: ["mma: loop top", "s_free(B) seen", "QK_A(j+1) issued", "P_A seen", "PV_A issued", "s_free(A) seen",
             "QK_B(j+1) issued", "P_B seen"],
         1: ["smxA: before wait S", "S ready", "LDTM done", "max done", "P half0 stored", "P stored+wait_st", "arrived"],
         2: ["smxB: before wait S", "S ready", "LDTM done", "max done", "P half0 stored", "P stored+wait_st", "arrived"]}
ev = []
for r in range(3):
    for j in range(8, 11):
        for s, n in enumerate(names[r]):
            ev.append((float(t[r, j, s] - t0), f"j={j} {n}"))
print("== absolute timeline (clk), steps 8..10")
for tt, n in sorted(ev):
    print(f"{tt:9.0f}  {n}")
sl = slice(8, 56)
print("== mean durations over steps 8..55 (clk)")
print(f"step period: {(t[0, 9:57, 0] - t[0, 8:56, 0]).mean():.0f}")
d = lambda r, a, b: float((t[r, sl, b] - t[r, sl, a]).mean())  # noqa: E731
print(f"mma: wait s_free(B) {d(0,0,1):.0f} | issue QK_A {d(0,1,2):.0f} | wait P_A {d(0,2,3):.0f} | issue PV_A {d(0,3,4):.0f} | "
      f"wait s_free(A) {d(0,4,5):.0f} | issue QK_B {d(0,5,6):.0f} | wait P_B {d(0,6,7):.0f}")
for r in (1, 2):
    print(f"smx{'AB'[r-1]}: wait S {d(r,0,1):.0f} | LDTM {d(r,1,2):.0f} | max {d(r,2,3):.0f} | exp half0 {d(r,3,4):.0f} | "
          f"exp half1+wait_st {d(r,4,5):.0f} | arrive {d(r,5,6):.0f} | busy {d(r,1,6):.0f}")

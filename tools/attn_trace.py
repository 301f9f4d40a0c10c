"""Per-phase clock64 timeline of one attention CTA (profiling aid; G3C_ATTN_MODE=0|2, G3C_ATTN_TRACE_MMA_ONLY=1)."""
import sys

import torch

sys.path.insert(0, ".")
from gen3c_b200 import _lib, ops  # noqa: E402

L, H = 56320, 32
# the engine's calling convention: softmax scale * log2(e) folded into Q, scale = ln 2 (reference-free fast tiles)
q = (torch.randn(L, H * 128, device="cuda") * (128 ** -0.5 * 1.4426950408889634)).to(torch.bfloat16)
LN2 = 0.6931471805599453
k = (torch.randn(L, H * 128, device="cuda")).to(torch.bfloat16)
vt = (torch.randn(H * 128, L, device="cuda")).to(torch.bfloat16)
ops.attention(q, k, vt, H, scale=LN2)
torch.cuda.synchronize()
buf = torch.zeros(3 * 64 * 8, dtype=torch.int64, device="cuda")
lib = _lib.load()
_lib.check(lib.g3c_attn_set_trace(buf.data_ptr()), "set_trace")
ops.attention(q, k, vt, H, scale=LN2)
torch.cuda.synchronize()
_lib.check(lib.g3c_attn_set_trace(None), "set_trace")
t = buf.cpu().view(3, 64, 8).double()
t0 = t[0, 8, 0]
names = {0: ["mma: before wait P_A", "after wait P_A", "issued PV_A+S_A", "after wait P_B", "issued PV_B+S_B"],
         1: ["smxA: before wait S", "S ready", "LDTM done", "max done", "P half0 stored", "P stored+wait_st", "arrived"],
         2: ["smxB: before wait S", "S ready", "LDTM done", "max done", "P half0 stored", "P stored+wait_st", "arrived"]}
print("== absolute timeline (clk, relative to MMA step 8), steps 8..11")
ev = []
for r in range(3):
    for j in range(8, 12):
        for s, n in enumerate(names[r]):
            ev.append((float(t[r, j, s] - t0), f"j={j} {n}"))
for tt, n in sorted(ev):
    print(f"{tt:9.0f}  {n}")
print("== mean durations over steps 8..55 (clk)")
sl = slice(8, 56)
per = (t[0, 9:57, 0] - t[0, 8:56, 0]).mean()
print(f"step period (MMA loop): {per:.0f}")
d = lambda r, a, b: float((t[r, sl, b] - t[r, sl, a]).mean())  # noqa: E731
print(f"mma: wait P_A {d(0,0,1):.0f} | issue A {d(0,1,2):.0f} | wait P_B {d(0,2,3):.0f} | issue B {d(0,3,4):.0f}")
for r in (1, 2):
    print(f"smx{'AB'[r-1]}: wait S {d(r,0,1):.0f} | LDTM {d(r,1,2):.0f} | max+rescale-check {d(r,2,3):.0f} | exp half0 {d(r,3,4):.0f} | "
          f"exp half1 + wait_st {d(r,4,5):.0f} | arrive {d(r,5,6):.0f} | total busy {d(r,1,6):.0f}")
# hop latencies: softmax arrive -> MMA observes P ; MMA commit -> softmax observes S
pa = (t[0, sl, 1] - t[1, sl, 6]).mean()
pb = (t[0, sl, 3] - t[2, sl, 6]).mean()
sa = (t[1, 9:57, 1] - t[0, 8:56, 2]).mean()
sb = (t[2, 9:57, 1] - t[0, 8:56, 4]).mean()
print(f"hop P_A arrive -> MMA sees it: {pa:.0f} ; P_B: {pb:.0f}")
print(f"MMA issued(PV_A+S_A(j+1)) -> softmax A sees S(j+1): {sa:.0f} ; B: {sb:.0f}   (includes MMA execution)")

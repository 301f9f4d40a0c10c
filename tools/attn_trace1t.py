"""clock64 timeline of one CTA of k_attn_fwd1t (one query tile per CTA, three S buffers); G3C_ATTN_1T=2|4 selects the variant."""
import sys

import torch

sys.path.insert(0, ".")
from gen3c_b200 import _lib, ops  # noqa: E402

L, H = 56320, 32
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
names = {0: ["iss: top of step", "stage {V_j, K_j+3} seen", "P(j) seen", "P.V issued", "S(j+3) issued", "commits issued"],
         1: ["smx h=0: before wait S", "S ready", "LDTM done", "exp+store+release done"],
         2: ["smx h=last: before wait S", "S ready", "LDTM done", "exp+store+release done"]}
t0 = t[0, 8, 0]
print("== absolute timeline (clk, relative to issuer step 8), steps 8..11")
ev = []
for r in range(3):
    for j in range(8, 12):
        for s, n in enumerate(names[r]):
            ev.append((float(t[r, j, s] - t0), f"j={j} {n}"))
for tt, n in sorted(ev):
    print(f"{tt:9.0f}  {n}")
sl = slice(8, 56)
print("== mean over steps 8..55 (clk)")
print(f"step period (issuer): {float((t[0, 9:57, 0] - t[0, 8:56, 0]).mean()):.0f}")
d = lambda r, a, b: float((t[r, sl, b] - t[r, sl, a]).mean())  # noqa: E731
print(f"issuer: wait stage {d(0,0,1):.0f} | wait P {d(0,1,2):.0f} | issue P.V {d(0,2,3):.0f} | issue S {d(0,3,4):.0f} | commits {d(0,4,5):.0f}")
for r in (1, 2):
    print(f"{names[r][0][:10]}: period {float((t[r, 9:57, 0] - t[r, 8:56, 0]).mean()):.0f} | wait S {d(r,0,1):.0f} | LDTM {d(r,1,2):.0f} | exp+store+release {d(r,2,3):.0f}")
print(f"loader (stage of step j): waits for the free slot {d(1,4,5):.0f} | issues its TMA loads {d(1,5,6):.0f} | issue -> issuer sees the stage {float((t[0, sl, 1] - t[1, sl, 6]).mean()):.0f} | slot free -> consumed, in steps of the issuer {float((t[0, sl, 1] - t[1, sl, 5]).mean()) / float((t[0, 9:57, 0] - t[0, 8:56, 0]).mean()):.2f}")
print(f"softmax step j ends -> issuer sees P(j): {float((t[0, sl, 2] - t[1, sl, 3]).mean()):.0f} (h=0) {float((t[0, sl, 2] - t[2, sl, 3]).mean()):.0f} (h=last)")
print(f"S(j+3) issued -> softmax starts step j+3 (slack; negative = softmax waits): {float((t[1, 11:59, 1] - t[0, 8:56, 4]).mean()):.0f}")

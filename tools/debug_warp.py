import os, sys
import numpy as np, torch
sys.path.insert(0, ".")
from oracle import cases
from gen3c_b200 import warp
os.makedirs("gpurun_out", exist_ok=True)
for name in ("R3", "R4"):
    g = np.load(f"tests/golden/warp_{name}.npz"); c = cases.warp_case(name)
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    w, m, d, f = warp.forward_warp(cu(c["image"]), None if c["mask"] is None else cu(c["mask"]), None, None, cu(c["w2c_tgt"]), cu(c["K"]), cu(c["K"]),
                                   render_depth=True, world_points1=cu(g["points"]))
    np.savez_compressed(f"gpurun_out/warp_{name}_gpu.npz", w=w.cpu().numpy(), m=m.cpu().numpy(), d=d.cpu().numpy(), f=f.cpu().numpy())
    for b in range(w.shape[0]):
        e = np.abs(w[b].cpu().numpy() - g["warped"][b])
        print(name, "item", b, "frac>2e-3:", (e > 2e-3).mean(), "max", e.max(), "flow err", np.abs(f[b].cpu().numpy() - g["flow"][b]).max())

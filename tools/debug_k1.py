"""Phase timeline of the fused position-embedding kernel (CTA 0), via clock64 stamps."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from petr_b200 import _lib, ops, synth, engine
from petr_b200.configs import WORKLOADS
DEV = "cuda:0"
wl = WORKLOADS["vov-p4-800x320"]
metas = synth.make_img_metas(wl, 4, 5)
i2l = torch.from_numpy(engine.img2lidar_from_metas(metas)).to(DEV)
pr = (-61.2, -61.2, -10.0, 61.2, 61.2, 10.0)
w1 = (torch.randn(1024, 192) * 0.05).to(DEV).to(torch.bfloat16); b1 = torch.zeros(1024, device=DEV)
w2 = (torch.randn(256, 1024) * 0.03).to(DEV).to(torch.bfloat16); b2 = torch.zeros(256, device=DEV)
dbg = torch.zeros(32, dtype=torch.int64, device=DEV)
L = _lib.lib()
L.petr_debug_set_ptr.argtypes = [ctypes.c_void_p]
for it in range(3):
    L.petr_debug_set_ptr(ctypes.c_void_p(dbg.data_ptr()) if it == 2 else None)
    ops.position_mlp_fused(i2l, wl.feat_h, wl.feat_w, 64, wl.img_h, wl.img_w, True, 1.0, (pr[3] - 1.0) / (64 * 65), pr, w1, b1, w2, b2)
    torch.cuda.synchronize()
L.petr_debug_set_ptr(None)
d = dbg.cpu().tolist()
t0 = d[0]
names = {0: "start", 1: "features written", **{2 + c: f"epi chunk {c} done" for c in range(8)}, 10: "o_full seen", 11: "staged", 12: "stored",
         16: "MMA: a1 ready", **{17 + c: f"MMA: G1({c}) issued" for c in range(8)}, 25: "MMA: all issued"}
for k in sorted(names):
    print(f"{names[k]:24s} +{d[k] - t0:8d} clk")

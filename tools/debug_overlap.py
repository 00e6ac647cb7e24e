"""Does the H2D copy of the next batch overlap the head forward?  (bench.py e2e diagnosis)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import petr_b200 as pb
from petr_b200 import synth
dev = torch.device("cuda:0")
wl = pb.WORKLOADS["vov-p4-1600x640"]
B = 16
head = pb.build_head(wl.head_cfg()); head.load_state_dict(synth.head_state_dict(head, 0)); head = head.eval().to(dev)
metas = synth.make_img_metas(wl, B, 0)
xh = synth.make_features(wl, B, 0, pin=True)
xd = xh.to(dev); xd2 = torch.empty_like(xd)
cs = torch.cuda.Stream(device=dev)
for _ in range(3): head([xd], metas)
torch.cuda.synchronize()
def timed(fn, n=5):
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t) / n * 1e3
def copy_only():
    with torch.cuda.stream(cs): xd2.copy_(xh, non_blocking=True)
def fwd_only(): head([xd], metas)
def both():
    with torch.cuda.stream(cs): xd2.copy_(xh, non_blocking=True)
    head([xd], metas)
def host_only():
    t = time.perf_counter(); head([xd], metas); return (time.perf_counter() - t) * 1e3
print("copy only   %.2f ms" % timed(copy_only))
print("fwd only    %.2f ms" % timed(fwd_only))
print("copy || fwd %.2f ms" % timed(both))
torch.cuda.synchronize(); print("host enqueue time of one forward %.2f ms" % host_only()); torch.cuda.synchronize()
head.use_cuda_graph = True
for _ in range(2): head([xd], metas)
print("graph fwd only    %.2f ms" % timed(fwd_only))
print("graph copy || fwd %.2f ms" % timed(both))

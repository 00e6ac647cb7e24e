#!/usr/bin/env python
"""bench.py -- PETR head (position embedding + 6-layer decoder) samples/sec on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                  [--workload vov-p4-1600x640] [--batch B_per_gpu] [--precision bf16|fp32]

One "step" = one head forward (PETRHead.forward, petr_head.py:366-468) over a batch of B
synthetic samples of the named BASELINE config (random-init weights, synthetic 6-view features and
nuScenes-like img_metas; petr_b200.synth).  N>1: launched by torch.distributed.run, one process per
GPU, samples sharded data-parallel (weak scaling, no data-path collective; NCCL only for the timing
barrier / max over ranks).  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "PETR head 6-view samples/sec at 1600x640, 1/2/4/8 B200; attn % of bf16 peak"


FLASH_KERNEL = "cross_attention_tc4_kernel"        # fast (max-free) pass; the exact pass only re-runs flagged CTAs


def load_traffic(batch):
    """DRAM bytes per launch of the flash cross-attention kernel from the committed `ncu --set full`
    capture (dram__bytes_read.sum + dram__bytes_write.sum), if one exists for this kernel and
    batch size."""
    p = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if os.path.isfile(p):
        with open(p) as f:
            d = json.load(f)
        ent = d.get(FLASH_KERNEL, {}).get(str(batch))
        if ent:
            return float(ent["dram_bytes_per_launch"])
    return None


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            d = json.load(f)
        return dict(source="measured", hbm=d["hbm_gbs"], bf16=d["bf16_tflops"],
                    bf16_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]))
    return dict(source="fallback", hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0)


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return dict(sm_mhz=float(np.median(sm)) if sm else None,
                    sm_max_mhz=float(max(smax)) if smax else None,
                    samples=len(sm), reasons=sorted(reasons))


def bind_to_gpu_numa_node(local_rank):
    """Pin this process to the CPUs next to its GPU before any pinned host buffer is allocated, so first-touch puts
    the staging memory on the GPU's NUMA node (each rank streams ~0.4 GB per step from it).  Returns the cpulist."""
    try:
        pr = torch.cuda.get_device_properties(local_rank)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/local_cpulist") as f:
            txt = f.read().strip()
        cpus = set()
        for part in txt.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = f.read().strip()
        return f"node {node}, cpus {txt}"
    except Exception as exc:
        return f"unbound ({type(exc).__name__})"


def cpu_oracle_setup(wl, seed=0):
    import petr_b200 as pb
    from petr_b200 import synth
    head = pb.build_head(wl.head_cfg())
    head.load_state_dict(synth.head_state_dict(head, seed), strict=True)
    sd = {k: v.detach().clone() for k, v in head.state_dict().items()}
    v2 = wl.head == "PETRv2Head"
    return sd, dict(with_fpe=v2, with_time=v2, with_multi=v2)


def time_cpu_oracle(wl, nsamples, seed=0):
    """Reference's CPU implementation of the path (oracle port, fp32, all host threads):
    per-sample loop as BASELINE.md section 2 prescribes."""
    from oracle import petr_oracle as po          # the one place bench.py executes oracle/
    from petr_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd, ocfg = cpu_oracle_setup(wl, seed)
    metas = synth.make_img_metas(wl, 1, seed)
    x = synth.make_features(wl, 1, seed)
    with torch.no_grad():
        po.head_forward(sd, x, metas, ocfg)                       # warm-up
        times = []
        for _ in range(nsamples):
            t = time.perf_counter()
            po.head_forward(sd, x, metas, ocfg)
            times.append(time.perf_counter() - t)
    return times, cores


def time_gpu_eager_oracle(wl, nsamples, dev, seed=0):
    """"The reference on the same box" (BASELINE.md section 2): the oracle port's plain torch ops, fp32, moved to the
    B200 (PyTorch eager: cuBLAS / cuDNN kernels, (B*8, 900, Lk) score tensors materialised), per-sample loop.
    A reported baseline; like cpu_baseline it never touches the measured path."""
    from oracle import petr_oracle as po
    from petr_b200 import synth
    sd, ocfg = cpu_oracle_setup(wl, seed)
    sd = {k: v.to(dev) for k, v in sd.items()}
    metas = synth.make_img_metas(wl, 1, seed)
    x = synth.make_features(wl, 1, seed).to(dev)
    with torch.no_grad(), torch.device(dev):
        for _ in range(2):
            po.head_forward(sd, x, metas, ocfg)
        torch.cuda.synchronize(dev)
        t = time.perf_counter()
        for _ in range(nsamples):
            po.head_forward(sd, x, metas, ocfg)
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t
    del sd, x
    torch.cuda.empty_cache()
    return nsamples / dt


def run_reference(args, wl, rank):
    if rank != 0:
        return
    warm = max(0, args.warmup)
    times, cores = time_cpu_oracle(wl, max(1, args.steps) + warm, 0)
    times = times[warm:] if len(times) > warm else times
    total = float(sum(times))
    value = len(times) / total
    line = dict(metric=METRIC, value=value, unit="samples/s", n_gpus=args.gpus, steps=len(times),
                warmup=warm, ms_per_step=1000.0 * total / len(times), higher_is_better=True,
                scaling="weak", vs_baseline=None, dtype="f32", data="synthetic", impl="reference",
                config=dict(workload=wl.name, batch_per_step=1, num_keys=wl.num_keys,
                            note="reference head forward on host CPU, fp32, per-sample loop"),
                cpu_baseline=dict(value=value, unit="samples/s", cores=cores, kind="port",
                                  sample=f"{len(times)} single-sample head forwards of {wl.name} "
                                         f"(oracle/petr_oracle.py, torch CPU fp32, {cores} threads)"),
                e2e=dict(value=value, unit="samples/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="vov-p4-1600x640")
    ap.add_argument("--batch", type=int, default=32, help="samples per GPU per step (BASELINE configs 3 and 5: batch sweep 1-32 per GPU, global 256 on 8 GPUs)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cpu-samples", type=int, default=6)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--breakdown", action="store_true", help="print per-op CUDA-event times to stderr")
    ap.add_argument("--graph", dest="graph", action="store_true", default=None,
                    help="replay the whole head as one CUDA graph per input buffer (default on one GPU; multi-process "
                         "runs launch eagerly unless this flag is given)")
    ap.add_argument("--no-graph", dest="graph", action="store_false", help="eager launches")
    ap.add_argument("--keysplit", action="store_true",
                    help="BASELINE config 4: every rank works on the SAME samples, keys split over camera "
                         "views, one NCCL all-gather of (O, m, l) partials per decoder layer (strong scaling)")
    ap.add_argument("--p2p", action="store_true",
                    help="with --keysplit: exchange the partials by pushing them over NVLink peer memory (own kernels, "
                         "graph-capturable) instead of NCCL all-gathers")
    args = ap.parse_args()

    if args.graph is None:
        # default: whole-head CUDA-graph replay, also in data-parallel multi-process runs (the graph captures no
        # collective there; NCCL is only used for the timing barrier).  Key-split captures NCCL all-gathers: eager.
        args.graph = (not args.keysplit) or args.p2p
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    import petr_b200 as pb
    from petr_b200 import ops, synth
    wl = pb.WORKLOADS[args.workload]

    if args.impl == "reference":
        run_reference(args, wl, rank)
        return

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (impl=ours) needs a CUDA device; petr_b200 has no CPU fallback")
    args.warmup = max(3, args.warmup)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    host_numa = bind_to_gpu_numa_node(local_rank) if world > 1 else "not bound (single process)"
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    B = args.batch
    head = pb.build_head(wl.head_cfg())
    head.load_state_dict(synth.head_state_dict(head, 0), strict=True)
    head = head.eval().to(dev)
    head.precision = args.precision
    head.use_cuda_graph = bool(args.graph)
    if args.keysplit and world > 1:
        if args.graph and not args.p2p:
            raise SystemExit("--graph with --keysplit needs --p2p (NCCL collectives captured in a CUDA graph hung "
                             "process-group teardown; see profiles/README.md)")
        head.keysplit = dict(world=world, rank=rank, group=None, p2p=bool(args.p2p))
    # data parallel: each rank gets its own shard of the global batch (different seeds per rank);
    # key-split: every rank gets the same samples
    data_seed = 0 if args.keysplit else rank
    metas = synth.make_img_metas(wl, B, seed=data_seed)
    x_host = synth.make_features(wl, B, seed=data_seed, pin=True)
    x_dev = x_host.to(dev, non_blocking=True)
    torch.cuda.synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v: float) -> float:
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident timing (`value`) ---------------------------------------------------
    for _ in range(args.warmup):
        out = head([x_dev], metas)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ops.reset_launch_count()
    head.engine.graph_replayed_launches = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        out = head([x_dev], metas)
    e1.record()
    barrier()
    launches = ops.launch_count() + head.engine.graph_replayed_launches
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if rank == 0 else None
    # second, separately timed pass with per-launch CUDA events (kept out of `value`: the event
    # records add host overhead per launch)
    from petr_b200 import _lib
    head.use_cuda_graph = False                       # per-launch events need eager launches
    _lib.TIMERS = {}
    for _ in range(args.steps):
        out = head([x_dev], metas)
    torch.cuda.synchronize()
    timers = _lib.TIMERS
    _lib.TIMERS = None
    head.use_cuda_graph = bool(args.graph)
    ms_per_step = ms_total / args.steps
    nrep = 1 if args.keysplit else world          # key-split: all ranks share the same samples
    value = nrep * B * args.steps / (ms_total / 1000.0)

    # ---- end-to-end through the public API with host buffers (`e2e`) -------------------------
    out_host = [{k: torch.empty(out[k].shape, dtype=out[k].dtype).pin_memory()
                 for k in ("all_cls_scores", "all_bbox_preds")} for _ in range(2)]

    # Each step copies that step's features from pinned host memory, runs the public
    # PETRHead.forward and reads the outputs back; the H2D copy of step i+1 runs on a second stream
    # while step i computes (double-buffered device input), all inside the timed region.
    copy_stream = torch.cuda.Stream(device=dev)
    d2h_stream = torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream(dev)
    x_bufs = [torch.empty_like(x_dev) for _ in range(2)]
    ev_copied = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]
    ev_out_ready = [torch.cuda.Event() for _ in range(2)]
    ev_out_read = [torch.cuda.Event() for _ in range(2)]

    def e2e_run(nsteps):
        def prefetch(i):
            slot = i % 2
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(ev_free[slot])
                x_bufs[slot].copy_(x_host, non_blocking=True)
                ev_copied[slot].record(copy_stream)
        for slot in range(2):
            ev_free[slot].record(main_stream)
            ev_out_read[slot].record(main_stream)
        prefetch(0)
        held = [None, None]
        for i in range(nsteps):
            slot = i % 2
            if i + 1 < nsteps:
                prefetch(i + 1)
            main_stream.wait_event(ev_copied[slot])
            # the outputs of this input buffer's graph are static: its previous results must have left the device
            main_stream.wait_event(ev_out_read[slot])
            o = head([x_bufs[slot]], metas)
            ev_free[slot].record(main_stream)
            ev_out_ready[slot].record(main_stream)
            held[slot] = o
            with torch.cuda.stream(d2h_stream):         # D2H of step i runs under the compute of step i + 1
                d2h_stream.wait_event(ev_out_ready[slot])
                for k in out_host[slot]:
                    out_host[slot][k].copy_(o[k], non_blocking=True)
                    o[k].record_stream(d2h_stream)      # eager mode: outputs are allocator blocks of the main stream
                ev_out_read[slot].record(d2h_stream)
        main_stream.synchronize()
        d2h_stream.synchronize()

    # raw H2D bandwidth of one step's features (context for the PCIe-bound e2e figure)
    hs, he = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    hs.record()
    for _ in range(3):
        x_bufs[0].copy_(x_host, non_blocking=True)
    he.record()
    torch.cuda.synchronize()
    h2d_gbs = 3 * x_host.numel() * 4 / (hs.elapsed_time(he) * 1e-3) / 1e9
    h2d_min = -max_over_ranks(-h2d_gbs)              # all ranks copy at once here: the slowest rank's host feed

    e2e_run(2)
    barrier()
    t0 = time.perf_counter()
    e2e_run(args.steps)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = nrep * B * args.steps / e2e_s

    # Secondary, informational: the same loop with the features handed over as bf16 (an AMP backbone's output / the
    # head's bf16 wire format): half the host -> device bytes, bit-identical results on the bf16 path (the first
    # kernel rounds fp32 features to bf16 anyway).  NOT the headline e2e, which keeps the reference's fp32 features.
    e2e_bf16 = None
    if args.precision == "bf16":
        x_host_f32, x_bufs_f32 = x_host, x_bufs
        x_host = x_host_f32.to(torch.bfloat16).pin_memory()
        x_bufs = [torch.empty(x_dev.shape, dtype=torch.bfloat16, device=dev) for _ in range(2)]
        e2e_run(2)
        barrier()
        t0 = time.perf_counter()
        e2e_run(args.steps)
        barrier()
        e2e_bf16 = dict(value=nrep * B * args.steps / max_over_ranks(time.perf_counter() - t0), unit="samples/s",
                        h2d_bytes_per_step=int(x_host.numel() * 2 + B * wl.num_cams * (64 + 8)),
                        note="features handed over as bf16 (half the H2D bytes); informational, not the headline")
        x_host, x_bufs = x_host_f32, x_bufs_f32
    h2d = x_host.numel() * 4 + B * wl.num_cams * (64 + 8)
    d2h = sum(t.numel() * 4 for t in out_host[0].values())

    # ---- roofline of the dominant kernel (flash cross-attention) ------------------------------
    peaks = load_peaks()
    roofline = None
    breakdown = None
    if timers:
        breakdown = {k: dict(ms_per_step=float(sum(a.elapsed_time(b) for a, b in v)) / args.steps,
                             launches_per_step=len(v) / args.steps) for k, v in timers.items()}
        breakdown = dict(sorted(breakdown.items(), key=lambda kv: -kv[1]["ms_per_step"]))
    if timers and timers.get("cross_attention_bf16"):
        evs = timers["cross_attention_bf16"]
        avg_ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
        s = head.engine.spec
        flops = 4.0 * s.Lq * wl.num_keys * s.E * B                  # per launch (one layer, B samples)
        achieved = flops / (avg_ms * 1e-3) / 1e12
        # the honest secondary bound at head_dim 32: one exponential per score element, 16 MUFU ex2 per clock per SM;
        # rows are exponentiated in whole 32-row warps (900 -> 928), the kernel moves half of them to the FMA pipe
        sm_clk = (clocks or {}).get("sm_mhz") or 1965.0
        exps = 928.0 * wl.num_keys * s.heads * B
        exp_ms = exps / (16.0 * 148 * sm_clk * 1e6) * 1e3
        roofline = dict(kernel=FLASH_KERNEL, bound="tensor", achieved=achieved,
                        peak=peaks["bf16_sustained"], unit="TFLOP/s", frac=achieved / peaks["bf16_sustained"],
                        frac_of_burst_peak=achieved / peaks["bf16"],
                        traffic=load_traffic(B) if wl.name == "vov-p4-1600x640" else None,
                        peak_source=peaks["source"] + " bf16 sustained (frac) / burst (frac_of_burst_peak)",
                        launches_timed=len(evs), avg_launch_ms=avg_ms,
                        share_of_step=avg_ms * s.L / ms_per_step,
                        alg_flops_per_launch=flops,
                        exp_bound=dict(exps_per_launch=exps, mufu_only_ms=exp_ms, frac_of_mufu_only_time=exp_ms / avg_ms,
                                       note="time of the launch's exponentials at 16 ex2/clk/SM if all ran on MUFU; the "
                                            "kernel evaluates 2 of every 4 pairs with an FMA-pipe polynomial"))
    roofline_k1 = None
    if timers and timers.get("position_mlp_fused"):
        evs = timers["position_mlp_fused"]
        avg_ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
        T = B * wl.num_keys
        flops = 2.0 * T * (192 * 1024 + 1024 * 256)
        byts = T * 256 * (4 + 2) + B * wl.num_cams * 64 + (192 * 1024 + 1024 * 256) * 2     # residual in, bf16 out, matrices, weights once
        roofline_k1 = dict(kernel="position_mlp_fused_kernel", bound="tensor", achieved=flops / (avg_ms * 1e-3) / 1e12,
                           peak=peaks["bf16_sustained"], unit="TFLOP/s",
                           frac=flops / (avg_ms * 1e-3) / 1e12 / peaks["bf16_sustained"],
                           hbm_achieved_gbs=byts / (avg_ms * 1e-3) / 1e9, hbm_peak_gbs=peaks["hbm"],
                           hbm_frac=byts / (avg_ms * 1e-3) / 1e9 / peaks["hbm"], avg_launch_ms=avg_ms,
                           alg_flops_per_launch=flops, alg_bytes_per_launch=byts,
                           note="tensor pipe is the binding roof (896 flop/B); the coordinate tensor, the MLP input and the "
                                "hidden activation contribute no HBM bytes")

    if rank == 0:
        cpu = None
        gpu_eager = None
        if world == 1 and not args.no_cpu_baseline:
            times, cores = time_cpu_oracle(wl, args.cpu_samples, 0)
            v = len(times) / float(sum(times))
            cpu = dict(value=v, unit="samples/s", cores=cores, kind="port",
                       sample=f"{len(times)} single-sample head forwards of {wl.name} "
                              f"(oracle/petr_oracle.py, torch CPU fp32, {cores} threads)")
            try:
                ge = time_gpu_eager_oracle(wl, 10, dev, 0)
                gpu_eager = dict(value=ge, unit="samples/s", kind="port on cuda (PyTorch eager fp32, cuBLAS/cuDNN)",
                                 sample=f"10 single-sample head forwards of {wl.name} (oracle/petr_oracle.py on {dev})")
            except Exception as exc:                                   # reported, never fatal for the measured line
                gpu_eager = dict(unavailable=f"{type(exc).__name__}: {exc}"[:200])
        line = dict(metric=METRIC, value=value, unit="samples/s", n_gpus=world, steps=args.steps,
                    warmup=args.warmup, ms_per_step=ms_per_step, higher_is_better=True,
                    scaling="strong" if args.keysplit else "weak",
                    vs_baseline=None, dtype="bf16" if args.precision == "bf16" else "f32",
                    data="synthetic",
                    config=dict(workload=wl.name, batch_per_gpu=B, global_batch=B * nrep,
                                num_keys=wl.num_keys, num_query=900, decoder_layers=6,
                                parallelism=((f"keysplit{world}" + ("-p2p" if args.p2p else "-nccl")) if args.keysplit else f"dp{world}"),
                                precision=args.precision, cuda_graph=bool(args.graph),
                                l2="inputs larger than L2 (%.0f MB features per step)" % (x_host.numel() * 4 / 1e6)),
                    e2e=dict(value=e2e_value, unit="samples/s", h2d_bytes_per_step=int(h2d),
                             d2h_bytes_per_step=int(d2h), h2d_gbs_measured=round(h2d_gbs, 1),
                             h2d_gbs_min_over_ranks=round(h2d_min, 1), host_numa=host_numa,
                             note="H2D of step i+1 and D2H of step i-1 overlap the compute of step i (three streams)"),
                    e2e_bf16_features=e2e_bf16,
                    gpu_launches=int(launches), clocks=clocks, roofline=roofline, roofline_k1=roofline_k1,
                    cpu_baseline=cpu, gpu_eager_baseline=gpu_eager)
        if args.breakdown and breakdown:
            sys.stderr.write("per-op breakdown (ms per step, CUDA events):\n")
            for k, v in breakdown.items():
                sys.stderr.write(f"  {k:32s} {v['ms_per_step']:9.3f} ms  x{v['launches_per_step']:.0f}\n")
        print(json.dumps(line), flush=True)
    if dist is not None:
        # The JSON line is out.  Tear down properly: drop the captured graphs first, then the process group; a
        # teardown that raises or does not finish within 60 s makes the run exit NON-zero (the line stays valid,
        # the failure is visible) instead of hanging the job.
        import threading
        sys.stdout.flush()
        sys.stderr.flush()
        killer = threading.Timer(60.0, lambda: (sys.stderr.write("bench.py: teardown timed out\n"), os._exit(3)))
        killer.daemon = True
        killer.start()
        head.engine.invalidate()
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()
        killer.cancel()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- the reference's headline metric on B200: fused quantize -> decode -> mean throughput (coords/s)
of the type quantizer (BASELINE.json `metric`), on the named workloads of BASELINE.json `configs`.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload metric|cfg2|cfg3|cfg4|biased]
                    [--d D] [--n N] [--rate R] [--mode unbiased|biased]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

  --workload metric (default): d = 2^24, n = 128 clients per GPU, R = 1, unbiased -- the shape BASELINE.json's metric is quoted on
             cfg2 : d = 2^16, n = 1000 (the largest point of the NMSE sweep)
             cfg3 : rotated variant, d = 2^20, n = 128: randomized Hadamard -> type quantizer -> mean in the rotated domain -> inverse
             cfg4 : Flower hook shape, d = 122 626 (the reference's CIFAR-10 CNN, TU:40-49), n = 100 client deltas
             biased: the metric shape with the biased (Reznik) quantizer
One "step" = one pass of the hot path over one batch of synthetic client vectors resident in HBM
(N(0,1) i.i.d., torch.Generator seed 42 + rank).  Rank 0 prints ONE JSON line.
  value     : whole-job coords/s = N_gpus * n * d * K / (max-over-ranks device time), inputs in HBM
  e2e       : same metric through the public host-buffer call (pinned host rows -> H2D -> fused path -> D2H of the mean)
  roofline  : dominant kernel by per-kernel CUDA events (dme_profile_*), algorithmic bytes 4*n*d + 4*d of the step / its time,
              against MEASURED_PEAKS.json hbm_gbs; kernel_ms lists every kernel of the step
  cpu_baseline : oracle port (oracle/dme_oracle.c) on the box's host cores, bounded sample
--impl reference times that CPU port alone (the reference is pure Python/torch and cannot travel to the GPU box).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

UNIT = "coords/s"
L2_BYTES = 256 << 20          # inputs smaller than this are rotated over several copies so that no step re-reads a cached input
WORKLOADS = {
    "metric": dict(d=1 << 24, n=128, rate=1, mode="unbiased", kind="type"),
    "cfg2": dict(d=1 << 16, n=1000, rate=1, mode="unbiased", kind="type"),
    "cfg3": dict(d=1 << 20, n=128, rate=1, mode="unbiased", kind="rotated"),
    "cfg4": dict(d=122626, n=100, rate=1, mode="unbiased", kind="type"),
    "biased": dict(d=1 << 24, n=128, rate=1, mode="biased", kind="type"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="metric", choices=sorted(WORKLOADS))
    ap.add_argument("--d", type=int, default=None)
    ap.add_argument("--n", type=int, default=None, help="clients per GPU")
    ap.add_argument("--rate", type=float, default=None)
    ap.add_argument("--mode", default=None, choices=["unbiased", "biased"])
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--eager", action="store_true", help="call the eager API every step instead of replaying the captured CUDA graph")
    ap.add_argument("--reduce", default="peer", choices=["peer", "nccl"],
                    help="N > 1: sum of the ranks' partial means by our peer-memory kernel (dme_peer_sum_slice) or by ncclAllReduce")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    a = ap.parse_args()
    w = WORKLOADS[a.workload]
    a.kind = w["kind"]
    for k in ("d", "n", "rate", "mode"):
        if getattr(a, k) is None:
            setattr(a, k, w[k])
    return a


def metric_name(a):
    q = "rotated type quantizer" if a.kind == "rotated" else f"{a.mode} type quantizer"
    return f"quantize+decode+mean coords/s ({q}, R={a.rate:g})"


def workload_name(a):
    rot = "rotated (RHT -> type -> mean -> inverse) " if a.kind == "rotated" else ""
    return f"{a.workload}: {rot}type_{a.mode} R={a.rate:g} d={a.d} n={a.n}/gpu N(0,1) synthetic"


def rate_key(r):
    return int(r) if float(r) == int(r) else float(r)


# ------------------------------------------------------------------ CPU arm (oracle port; the checker, timed)
def cpu_arm(a, steps, warmup, budget_s):
    """Times the oracle's quantize->dequantize->mean loop (ND:133-147) on host cores.  Unbiased type quantizer: the C port with one
    pthread per core, `cores` clients of the full length d per step.  Rotated / biased: the port's single-threaded pieces driven
    from Python, a few clients per step."""
    from oracle import oracle as orc
    cores = os.cpu_count() or 1
    d = a.d
    rng = np.random.default_rng(42)
    R = rate_key(a.rate)
    if a.kind == "type" and a.mode == "unbiased":
        ns = max(1, min(a.n, cores))
        X = rng.standard_normal((ns, d), dtype=np.float32)
        Xs = rng.random(ns, dtype=np.float32)
        m = orc.m_for(R, d)
        fn = lambda: orc.quantize_mean_unbiased(X, m, Xs, threads=cores)
        used, what = cores, f"oracle/dme_oracle.c with {cores} pthreads"
    else:
        ns = max(1, min(a.n, 2 if d >= (1 << 22) else 4))
        X = rng.standard_normal((ns, d), dtype=np.float32)
        Xs = rng.random(ns, dtype=np.float32)
        dpad = orc.pad_pow2(d)
        diag = np.where(rng.random(dpad) < 0.5, -1.0, 1.0).astype(np.float32)
        if a.kind == "rotated":
            m = orc.m_for(R, dpad)
            def fn():
                qs = [orc.type_unbiased(orc.rht(X[c], diag), m, float(Xs[c]))["deq"] for c in range(ns)]
                return orc.irht(orc.mean_of(qs), diag)[:d]
        else:
            m = orc.m_for(R, d)
            fn = lambda: orc.mean_of([orc.type_biased(X[c], m)["deq"] for c in range(ns)])
        used, what = 1, "oracle/dme_oracle.c pieces driven from Python, 1 thread"
    times = []
    t_start = time.perf_counter()
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        fn()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        if budget_s and time.perf_counter() - t_start > budget_s and len(times) >= 1:
            break
    t = float(np.mean(times))
    return {"value": ns * d / t, "unit": UNIT, "cores": used, "kind": "port",
            "sample": f"{ns} clients x d={d} per step, {len(times)} timed steps, {what}",
            "ms_per_step": t * 1e3, "steps": len(times)}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, min(a.steps, 5)), max(1, min(a.warmup, 1))
    cb = cpu_arm(a, steps, warmup, budget_s=150.0)
    line = {"impl": "reference", "metric": metric_name(a), "value": cb["value"], "unit": UNIT, "n_gpus": a.gpus, "steps": cb["steps"],
            "warmup": warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32+f64acc", "data": "synthetic", "config": {"workload": workload_name(a), "sample": cb["sample"]},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ clocks sampler
class Clocks:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.rows.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------ our arm
def run_ours(a):
    import torch
    import torch.distributed as dist
    import dme_b200 as dme
    from dme_b200 import _cabi
    import ctypes as C

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback); use --impl reference for the CPU arm"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    n, d, R = a.n, a.d, rate_key(a.rate)
    n_total = n * world
    client0 = rank * n
    gen = torch.Generator(device=dev).manual_seed(42 + rank)
    copies = max(1, -(-L2_BYTES // (4 * n * d)))          # small inputs: rotate over enough copies to defeat L2
    Xs = []
    for k in range(copies):
        # client rows in a row-aligned buffer (row stride = d rounded up to 4 floats: 16-byte aligned rows, what the kernels read
        # in place; a dense (n, d) matrix with d % 4 != 0 would be staged into such a buffer on every call)
        X = torch.empty((n, (d + 3) // 4 * 4), dtype=torch.float32, device=dev)[:, :d]
        for c in range(n):                                   # row by row: keeps the generator's scratch small
            X[c].normal_(generator=gen)
        Xs.append(X)
    mean = torch.empty(d, dtype=torch.float32, device=dev)
    L = _cabi.lib()

    from dme_b200 import distributed as dmed

    def local_type(x, r, **kw):
        return dme.quantize_mean(x, r, check=False, **kw)

    # The plain type path replays one captured CUDA graph per input copy (dme_b200.MeanGraph: uniforms of the round, workspace
    # reset, l1, quantize, decode in one launch; the round seed lives on the device and advances with every replay); --eager and
    # the rotated workload call the eager API.  The all-reduce (N > 1) follows the replay.
    # N > 1: the partial means are summed by our own kernel over NVLink peer memory (distributed.PeerReduce: symmetric buffer,
    # every rank reduces and broadcasts its slice in rank order between two device-side barriers) -- the decode writes straight
    # into the symmetric buffer; --reduce nccl (or a box without peer access) uses ncclAllReduce.
    peer, reduce_name = None, "none (one rank)"
    if world > 1:
        reduce_name = "ncclAllReduce"
        if a.reduce == "peer" and a.kind != "rotated":
            try:
                # measured (profiles/r02f_peer_check_n2/n4.txt, r02g_peer_check_n8.txt; 64 MiB): peer loads win at N = 2 (0.126 ms vs
                # multimem 0.20 / NCCL 0.146), the in-switch reduction at N = 8 (0.212 vs 0.314 / 0.256); a tie at N = 4
                peer = dmed.PeerReduce.get(d, None, multicast=world >= 4)
                mean = peer.buffer()
                reduce_name = ("dme_peer_sum_slice over symmetric memory (" +
                               ("multimem.ld_reduce / multimem.st through the NVSwitch" if peer.mc else "peer loads in rank order") + ")")
            except Exception as ex:            # said in the JSON line, not silent
                reduce_name = f"ncclAllReduce (peer memory unavailable: {str(ex)[:80]})"

    def reduce_mean():
        if peer is not None:
            peer.sum_()
        elif world > 1:
            dist.all_reduce(mean, op=dist.ReduceOp.SUM)

    graphs = None
    if a.kind != "rotated" and not a.eager:
        graphs = [dme.MeanGraph(X, R, mode=a.mode, seed=1234 + 100000 * k, client0=client0, n_total=n_total, out=mean) for k, X in enumerate(Xs)]

    def step(i):
        if graphs is not None:
            graphs[i % copies]()
            reduce_mean()
            return
        step_eager(i)

    def step_eager(i):
        X = Xs[i % copies]
        if a.kind == "rotated":
            # every rank rotates its clients, averages them in the rotated domain with the global divisor, ONE all-reduce of the
            # rotated partial mean, then the single inverse rotation (linear) on every rank
            dmed.rotated_quantize_mean_sharded(X, R, n_total=n_total, client0=client0, seed=1234 + i, mode=a.mode, out=mean)
        else:
            # N > 1: every rank quantizes + decodes its own clients with the global divisor, then ONE all-reduce
            if peer is not None:
                local_type(X, R, n_total=n_total, client0=client0, seed=1234 + i, mode=a.mode, out=mean)
                peer.sum_()
            else:
                dmed.quantize_mean_sharded(X, R, n_total=n_total, client0=client0, seed=1234 + i, mode=a.mode, out=mean, local_fn=local_type)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(a.warmup):
        step(i)
    def status():
        if graphs is not None:
            for g in graphs:
                g.status()
        else:
            dme.Workspace.get(dev).status()

    status()                                             # a kernel-side error in warm-up fails loudly here
    sync()
    launches0 = L.dme_launch_count()
    clk = Clocks(local)
    if rank == 0:
        clk.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.steps):
        step(a.warmup + i)
    e1.record()
    sync()
    ms = e0.elapsed_time(e1)
    clocks = clk.stop() if rank == 0 else None
    launches = L.dme_launch_count() - launches0
    status()
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_step = ms / a.steps
    value = n_total * d / (ms_step * 1e-3)

    # ---- roofline leg: per-kernel CUDA-event times (events on the launching stream), outside the timed region
    peak, peak_src = peaks()
    B_alg = 4.0 * n * d + 4.0 * d
    runs = [dme.profile_kernels(lambda: step_eager(1000 + j), warm=0) for j in range(4)][1:]       # eager: one event per launch
    kern_ms = {}
    for run in runs:
        for name, t_ms in run:
            kern_ms[name] = kern_ms.get(name, 0.0) + t_ms / len(runs)
    dom = max(kern_ms, key=kern_ms.get) if kern_ms else "step"
    dom_ms = kern_ms.get(dom, ms_step)
    # algorithmic bytes of one step (every input coordinate read once, the mean written once) over the
    # time the dominant kernel takes per step; the whole step against the same bytes is step_achieved
    ach = B_alg / (dom_ms * 1e-3) / 1e9
    roof = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
            "peak_source": peak_src, "algorithmic_bytes_per_launch": B_alg,
            "kernel_ms": {nm: float(v) for nm, v in sorted(kern_ms.items(), key=lambda kv: -kv[1])},
            "step_achieved": B_alg / (ms_step * 1e-3) / 1e9, "step_frac": B_alg / (ms_step * 1e-3) / 1e9 / peak}
    tr = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr):
        try:
            roof["traffic"] = json.load(open(tr)).get(a.workload, {}).get(dom.split("<")[0])     # ncu dram read + write per launch, by workload
        except Exception:
            pass

    # ---- e2e: host buffers, H2D + D2H inside the timed region
    e2e = None
    if not a.no_e2e:
        try:
            Xh = torch.empty((n, d), dtype=torch.float32, pin_memory=True)
            Xh.copy_(Xs[0])
            outh = torch.empty(d, dtype=torch.float32, pin_memory=True)
            red = (lambda t: dist.all_reduce(t, op=dist.ReduceOp.SUM)) if world > 1 else None
            if a.kind == "rotated":
                stage = torch.empty((n, d), dtype=torch.float32, device=dev)
                def e2e_step(i):
                    # pinned host rows -> device, the rotated fused path, all-reduce (N > 1), D2H of the mean, stream synchronize
                    stage.copy_(Xh, non_blocking=True)
                    dmed.rotated_quantize_mean_sharded(stage, R, n_total=n_total, client0=client0, seed=77 + i, mode=a.mode, out=mean)
                    outh.copy_(mean, non_blocking=True)
                    torch.cuda.current_stream().synchronize()
                api_name = "dme_b200.distributed.rotated_quantize_mean_sharded on rows copied from pinned host memory, D2H of the mean"
            else:
                def e2e_step(i):
                    # the public host-buffer API: chunked H2D on a copy stream overlapped with quantize + decode of the previous
                    # chunk, the all-reduce of the partial mean (N > 1), D2H of the mean, stream synchronize
                    dme.quantize_mean_host(Xh, R, out_host=outh, mode=a.mode, seed=77 + i, client0=client0, n_total=n_total,
                                           check=False, reduce_fn=red)
                api_name = "dme_b200.quantize_mean_host on pinned host rows (chunked H2D overlapped with the fused path, D2H of the mean)"
            e2e_step(0)
            sync()
            t0 = time.perf_counter()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            for i in range(a.e2e_steps):
                e2e_step(1 + i)
            s1.record()
            sync()
            ems = max(s0.elapsed_time(s1), (time.perf_counter() - t0) * 1e3) / a.e2e_steps
            if world > 1:
                t = torch.tensor([ems], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ems = float(t.item())
            e2e = {"value": n_total * d / (ems * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(4 * n * d), "d2h_bytes_per_step": int(4 * d),
                   "ms_per_step": ems, "steps": a.e2e_steps, "api": api_name}
            del Xh
            if a.kind == "type":
                # the server-side deployment: clients upload packed codes; H2D of the codes, one decode-mean launch, D2H of the mean
                pch = dme.codes_to_host(dme.type_encode(Xs[0], R, mode=a.mode, seed=77, client0=client0))
                dme.decode_mean_host(pch, out_host=outh, n_total=n_total, reduce_fn=red)
                sync()
                t0 = time.perf_counter()
                for i in range(a.e2e_steps):
                    dme.decode_mean_host(pch, out_host=outh, n_total=n_total, reduce_fn=red)
                sync()
                sms = (time.perf_counter() - t0) * 1e3 / a.e2e_steps
                if world > 1:
                    t = torch.tensor([sms], device=dev, dtype=torch.float64)
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    sms = float(t.item())
                code_bytes = int(pch["codes"].numel() + 8 * pch["dir"].numel() + 4 * pch["l1"].numel())
                e2e["server_side"] = {"value": n_total * d / (sms * 1e-3), "unit": UNIT, "ms_per_step": sms, "h2d_bytes_per_step": code_bytes,
                                      "d2h_bytes_per_step": int(4 * d), "api": "dme_b200.decode_mean_host: packed codes in pinned host memory -> H2D -> decode-mean -> D2H of the mean"}
                del pch
        except Exception as ex:  # pinned allocation can fail on a small host
            e2e = {"value": None, "unit": UNIT, "error": str(ex)[:200]}

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu:
        cb = cpu_arm(a, steps=3, warmup=1, budget_s=a.cpu_seconds)
        cpu = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        hyg = (f"inputs ({4 * n * d / 2**30:.2f} GiB/GPU) larger than L2; no flush needed" if copies == 1 else
               f"inputs ({4 * n * d / 2**20:.0f} MiB/GPU) rotated over {copies} copies ({copies * 4 * n * d / 2**20:.0f} MiB > L2)")
        mm = dme.m_for_rate(R, (1 << (d - 1).bit_length()) if a.kind == "rotated" else d)
        line = {"metric": metric_name(a), "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (f64 accumulate)", "data": "synthetic",
                "config": {"workload": workload_name(a), "clients_total": n_total, "m": mm, "parallelism": f"clients sharded x{world}",
                           "l2_hygiene": hyg,
                           "reduce": reduce_name,
                           "launch": "one CUDA graph replay per step (dme_b200.MeanGraph)" if graphs is not None else "eager API calls"},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "e2e": e2e, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)

#!/usr/bin/env python
"""Benchmark of the cfm_b200 hot path on BASELINE.json's metric:
minibatch-OT couplings/sec (N=8192, d=784) [+ ODE samples/sec], 1..8 B200 of one node.

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's CPU path (oracle port) on host cores

One step = one full coupling `OTPlanSampler("sinkhorn", reg=0.05, normalize_cost=True).sample_plan`
(cost matrix + 100 log-domain Sinkhorn iterations, stopThr=0 + N pair draws + gather) on a fresh
pair of synthetic Gaussian batches; weak scaling: every rank couples its own N=8192 shard
(per-shard coupling, the only collective is the NCCL all-gather of the sampled index pairs).
Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N, D, REG, ITERS = 8192, 784, 0.05, 100
METRIC = "minibatch-OT couplings/sec (N=8192,d=784)"
WORKLOAD = ("C2: Sinkhorn minibatch coupling N=8192 d=784 fp32, 100 log-domain iterations "
            "(normalize_cost=True, reg=0.05, stopThr=0): cost matrix + solve + 8192 pair draws + gather")
ODE_B, ODE_DIM, ODE_W = 10000, 784, 256


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops", 1590.0)), "measured"
    return 6650.0, 1590.0, "fallback"


class ClockSampler:
    """nvidia-smi sampled DURING the timed region (B200_PROFILING.md's clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "50", "-i", str(index)], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(",") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, r[3:7]):
                if v.strip().lower().startswith("active"):
                    reasons.add(nm)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), samples=len(sm))
        out["reasons"] = sorted(reasons)
        return out


# ------------------------------------------------------------------------------ CPU reference arm
def cpu_reference_sample(n_iter_sample=3, seed=0):
    """The reference's CPU path for ONE C2 coupling, restated (oracle/coupling.py + oracle/ot):
    torch.cdist**2 -> POT's kernel-space sinkhorn_knopp (what OTPlanSampler('sinkhorn') calls;
    finite in this normalised regime) -> np.random.choice -> gather.  Bounded sample: the cost
    matrix, exp(-M/reg), plan formation and the pair draw are timed in full; the Sinkhorn loop is
    timed for `n_iter_sample` iterations and scaled to 100."""
    from oracle import coupling as oc
    from oracle import ot as oot  # noqa: F401
    g = torch.Generator().manual_seed(seed)
    x0, x1 = torch.randn(N, D, generator=g), torch.randn(N, D, generator=g)
    t = {}
    t0 = time.perf_counter()
    M = oc.cost_matrix(x0, x1, normalize_cost=True)
    Mn = M.detach().cpu().numpy()
    t["cost"] = time.perf_counter() - t0
    a = b = np.ones(N) / N
    t0 = time.perf_counter()
    K = np.exp(Mn / (-REG))
    Kp = (1 / a).reshape(-1, 1) * K
    u = np.ones(N, dtype=Mn.dtype) / N
    t["setup"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    for _ in range(n_iter_sample):
        v = b / np.dot(K.T, u)
        u = 1.0 / np.dot(Kp, v)
    t["iter"] = (time.perf_counter() - t0) / n_iter_sample
    t0 = time.perf_counter()
    P = u.reshape((-1, 1)) * K * v.reshape((1, -1))
    i, j = oc.draw_pairs(P, N)
    _ = x0[i], x1[j]
    t["draw"] = time.perf_counter() - t0
    total = t["cost"] + t["setup"] + ITERS * t["iter"] + t["draw"]
    return 1.0 / total, t


class _CpuState:
    """Inputs and the kernel matrix of the CPU reference sample, built once per process."""
    x0 = x1 = K = Kp = u = None
    fixed = None  # timings of the parts measured once: cost, setup, draw


def cpu_reference_iterations(n_iter):
    """Re-time only the Sinkhorn-Knopp loop of the reference CPU path (bounded step of --impl reference)."""
    S = _CpuState
    if S.K is None:
        from oracle import coupling as oc
        g = torch.Generator().manual_seed(0)
        S.x0, S.x1 = torch.randn(N, D, generator=g), torch.randn(N, D, generator=g)
        t0 = time.perf_counter()
        Mn = oc.cost_matrix(S.x0, S.x1, normalize_cost=True).detach().cpu().numpy()
        t_cost = time.perf_counter() - t0
        t0 = time.perf_counter()
        S.K = np.exp(Mn / (-REG))
        S.Kp = (N * 1.0) * S.K
        S.u = np.ones(N, dtype=Mn.dtype) / N
        t_setup = time.perf_counter() - t0
        b = np.ones(N) / N
        v = b / np.dot(S.K.T, S.u)
        u = 1.0 / np.dot(S.Kp, v)
        t0 = time.perf_counter()
        P = u.reshape((-1, 1)) * S.K * v.reshape((1, -1))
        i, j = oc.draw_pairs(P, N)
        _ = S.x0[i], S.x1[j]
        del P
        S.fixed = {"cost": t_cost, "setup": t_setup, "draw": time.perf_counter() - t0}
    b = np.ones(N) / N
    u = S.u
    t0 = time.perf_counter()
    for _ in range(n_iter):
        v = b / np.dot(S.K.T, u)
        u = 1.0 / np.dot(S.Kp, v)
    t_iter = (time.perf_counter() - t0) / n_iter
    parts = dict(S.fixed, iter=t_iter)
    return 1.0 / (parts["cost"] + parts["setup"] + ITERS * t_iter + parts["draw"]), parts


def run_reference_arm(args, rank):
    """`bench.py --impl reference`: the reference's own CPU implementation of the path (POT is not
    installable here, so its restatement in oracle/: kind "port") on all host cores.  The cost matrix,
    exp(-M/reg), plan formation, np.random.choice and the gather are timed in full ONCE per process;
    every step re-times 2 Sinkhorn-Knopp iterations (the part that is 100x per coupling) and the value
    of a step is 1 / (cost + setup + 100 * iter + draw)."""
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    for _ in range(max(1, args.warmup)):
        cpu_reference_iterations(1)
    vals, parts = [], None
    t0 = time.perf_counter()
    for _ in range(args.steps):
        v, parts = cpu_reference_iterations(2)
        vals.append(v)
    wall = time.perf_counter() - t0
    value = float(np.mean(vals))
    cores = os.cpu_count() or 1
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "couplings/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 / value, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "CPU path, one coupling stream regardless of --gpus; "
                   f"wall time of the {args.steps} bounded steps: {wall:.1f} s"},
        "cpu_baseline": {"value": value, "unit": "couplings/s", "cores": cores, "kind": "port",
                         "sample": "cdist**2, exp(-M/reg), plan, np.random.choice, gather timed in full once; per step "
                                   "2 Sinkhorn-Knopp iterations timed and scaled to 100; "
                                   f"parts(s)={ {k: round(v, 4) for k, v in parts.items()} }"},
        "e2e": {"value": value, "unit": "couplings/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cfm_b200", choices=["cfm_b200", "reference"])
    ap.add_argument("--no-ode", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "cfm_b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return run_reference_arm(args, rank)

    import torch.distributed as dist

    import cfm_b200
    from cfm_b200 import _ffi
    from cfm_b200 import dist as cdist

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    _ffi.require_device()  # raises if the library / an sm_100 device is missing: no fallback
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    L = _ffi.lib()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # synthetic inputs: generated on the CPU (identical bits for the CPU baseline), distinct per rank
    g = torch.Generator().manual_seed(rank)
    x0_h = torch.randn(N, D, generator=g).pin_memory()
    x1_h = torch.randn(N, D, generator=g).pin_memory()
    x0, x1 = x0_h.to(dev), x1_h.to(dev)
    sampler = cfm_b200.OTPlanSampler("sinkhorn", reg=REG, normalize_cost=True, num_iter_max=ITERS,
                                     stop_thr=0.0, warn=False)
    np.random.seed(1234 + rank)

    def step():
        if world > 1:
            return cdist.sharded_sample_plan(sampler, x0, x1)
        return sampler.sample_plan(x0, x1)

    # nvidia-smi is started BEFORE the warm-up (its start-up stalls the GPU for a few ms) and keeps
    # sampling through the timed region
    clocks = ClockSampler(local_rank) if rank == 0 else None
    if clocks is not None:
        time.sleep(0.7)  # let nvidia-smi finish initialising NVML before any timed work
    for _ in range(args.warmup):
        step()
    barrier()

    # ---- timed region: K steps, device events, barrier + synchronize both sides ----
    sampler.stage_events = []
    launches0 = L.cfm_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        out = step()
    ev1.record()
    barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    launches = L.cfm_launch_count() - launches0
    clk = clocks.stop() if clocks else None
    stage_ms = {}
    for name, a, b in sampler.stage_events:
        stage_ms.setdefault(name, []).append(a.elapsed_time(b))
    sampler.stage_events = None
    tmax = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    elapsed_ms = float(tmax.item())
    value = world * args.steps / (elapsed_ms * 1e-3)

    # ---- parity gate reported with the number: marginals of the implied plan (float64, device)
    cp = sampler._couple(x0, x1, dev)
    Mn = (cp.M[:, :N] / cp.cost_max).double()
    P = (-(Mn / REG) + cp.log_u[:, None] + cp.log_v[None, :]).exp()
    parity = {"row_marginal_max_rel_err": float((P.sum(1) * N - 1).abs().max().item()),
              "col_marginal_max_rel_err": float((P.sum(0) * N - 1).abs().max().item()),
              "iterations": int(cp.status[1].item())}
    del P, Mn, cp

    # ---- e2e: public API with HOST (pinned) inputs; H2D + D2H of every step inside the timed region ----
    # (1) blocking call: sampler.sample_plan(host, host) -> host, one batch at a time
    for _ in range(2):
        sampler.sample_plan(x0_h, x1_h)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        a_h, b_h = sampler.sample_plan(x0_h, x1_h)
    torch.cuda.synchronize(dev)
    single_s = time.perf_counter() - t0
    # (2) the streaming form of the same call (cfm_b200.CouplingStream): every step still uploads its
    # inputs and downloads its coupled batch, but batch k+1's upload and batch k-1's download overlap
    # batch k's solve on separate streams.  This is the headline e2e number.
    from cfm_b200 import CouplingStream
    pipe = CouplingStream(sampler, dev, depth=2)
    # one continuous stream of warm + K batches; the clock starts when result number `warm` is delivered, so
    # the timed window holds exactly K steady-state steps (each with its own upload and download) and not the
    # pipeline fill or the one-off growth of torch's pinned / device allocator pools
    warm = 6
    n_out = -warm
    t0 = None
    barrier()
    for a_h, b_h in pipe.map((x0_h, x1_h) for _ in range(warm + args.steps)):
        n_out += 1
        if n_out == 0:
            t0 = time.perf_counter()
    e2e_s = time.perf_counter() - t0
    torch.cuda.synchronize(dev)
    assert n_out == args.steps
    te = torch.tensor([e2e_s, single_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e = {"value": world * args.steps / float(te[0].item()), "unit": "couplings/s",
           "h2d_bytes_per_step": int(2 * N * D * 4 + N * 8), "d2h_bytes_per_step": int(2 * N * D * 4),
           "api": "CouplingStream(sampler).map(host batches), depth 2; K consecutive steady-state results timed "
                  "after 6 untimed ones of the same stream",
           "blocking_call_value": world * args.steps / float(te[1].item())}
    assert a_h.device.type == "cpu" and a_h.shape == (N, D)

    # ---- roofline of the dominant kernel (the persistent Sinkhorn sweep kernel) ----
    hbm_peak, tc_peak, which = peaks()
    solve_ms = float(np.mean(stage_ms["solve"]))
    alg_bytes = ITERS * 2 * N * N * 4  # two passes over M per iteration
    achieved = alg_bytes / (solve_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("sinkhorn_kernel_dram_bytes_per_launch")
    roofline = {"bound": "hbm", "kernel": "sinkhorn_v2_kernel<512,4,2> (100 fused sweeps, one cooperative launch)",
                "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": traffic, "peak_source": which,
                "note": "algorithmic bytes = iters*2*N^2*4 (two passes per iteration); the fused sweep "
                        "reads M from HBM once per iteration, so frac can exceed 1",
                "stage_ms": {k: float(np.mean(v)) for k, v in stage_ms.items()}}

    line = {
        "metric": METRIC, "value": value, "unit": "couplings/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "per_gpu_batch": N, "dim": D, "parallelism": f"shard{world}",
                   "l2": "inputs larger than L2: the cost matrix streamed every iteration is 268 MB"},
        "clocks": clk, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
        "parity": parity,
    }

    # ---- second half of the metric: ODE samples/sec (BASELINE config 3) ----
    if not args.no_ode:
        torch.manual_seed(0)
        mlp = cfm_b200.MLP(dim=ODE_DIM, w=ODE_W, time_varying=True).to(dev)
        node = cfm_b200.NeuralODE(cfm_b200.torch_wrapper(mlp), solver="dopri5", atol=1e-4, rtol=1e-4)
        xo = torch.randn(ODE_B, ODE_DIM, generator=torch.Generator().manual_seed(100 + rank)).to(dev)
        span = torch.linspace(0, 1, 2)
        for _ in range(3):
            node.trajectory(xo, span)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            node.trajectory(xo, span)
        e1.record()
        barrier()
        tm = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ode_s = float(tm.item()) * 1e-3 / reps
        nfe = node.stats["nfe"]
        flops = 2.0 * ODE_B * (785 * 256 + 256 * 256 + 256 * 256 + 256 * 784) * nfe
        line["ode"] = {"metric": "ODE samples/sec (MLP 785-256-256-256-784 SELU, dopri5 atol=rtol=1e-4, B=10000/GPU)",
                       "value": world * ODE_B / ode_s, "unit": "samples/s", "nfe": nfe,
                       "accepted": node.stats["accepted"], "rejected": node.stats["rejected"],
                       "ms_per_trajectory": ode_s * 1e3, "mlp_tflops_algorithmic": flops / ode_s / 1e12,
                       "mlp_rows_per_s_per_nfe": world * ODE_B * nfe / ode_s}
        if rank == 0:
            # comparator on the same GPU: stock PyTorch eager (cuBLAS sgemm, TF32 off) through the
            # oracle's torchdyn-style driver
            from oracle import vector_field as vf
            torch.backends.cuda.matmul.allow_tf32 = False
            ref_m = vf.make_mlp(ODE_DIM, w=ODE_W, time_varying=True).to(dev)
            ref_m.load_state_dict(mlp.state_dict())
            f = lambda t, z: vf.wrapped_forward(ref_m, t, z)  # noqa: E731
            for _ in range(2):
                vf.dopri5_trajectory(f, xo, span.to(dev))
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            for _ in range(3):
                vf.dopri5_trajectory(f, xo, span.to(dev))
            torch.cuda.synchronize(dev)
            line["ode"]["torch_eager_same_gpu_samples_per_s"] = ODE_B / ((time.perf_counter() - t0) / 3)

        # ---- BASELINE config 1 sampling: the 2-D tutorial model, launch-bound regime ----
        torch.manual_seed(0)
        small = cfm_b200.MLP(dim=2, w=64, time_varying=True).to(dev)
        nsm = cfm_b200.NeuralODE(cfm_b200.torch_wrapper(small), solver="dopri5", atol=1e-4, rtol=1e-4)
        xs = torch.randn(1024, 2, generator=torch.Generator().manual_seed(7)).to(dev)
        span100 = torch.linspace(0, 1, 100)
        for _ in range(3):
            nsm.trajectory(xs, span100)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for _ in range(20):
            nsm.trajectory(xs, span100)
        torch.cuda.synchronize(dev)
        c1_s = (time.perf_counter() - t0) / 20
        line["ode_c1"] = {"metric": "ODE samples/sec (MLP 3-64-64-64-2 SELU, dopri5 1e-4, B=1024, 100 t_span points)",
                          "value": world * 1024 / c1_s, "unit": "samples/s", "ms_per_trajectory": c1_s * 1e3,
                          "nfe": nsm.stats["nfe"], "launches_per_trajectory": 1 if nsm.stats.get("fused") else None}
        if rank == 0:
            ref_s = vf.make_mlp(2, w=64, time_varying=True).to(dev)
            ref_s.load_state_dict(small.state_dict())
            fs = lambda t, z: vf.wrapped_forward(ref_s, t, z)  # noqa: E731
            with torch.no_grad():
                vf.dopri5_trajectory(fs, xs, span100.to(dev))
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                vf.dopri5_trajectory(fs, xs, span100.to(dev))
                torch.cuda.synchronize(dev)
            line["ode_c1"]["torch_eager_same_gpu_samples_per_s"] = 1024 / (time.perf_counter() - t0)

    # ---- CPU baseline on this box's host cores (rank 0, N=1 only) ----
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count() or 1)
        cpu_reference_sample(1)
        v, parts = cpu_reference_sample(5)
        line["cpu_baseline"] = {
            "value": v, "unit": "couplings/s", "cores": os.cpu_count() or 1, "kind": "port",
            "sample": "one N=8192,d=784 coupling of the reference CPU path restated (torch.cdist**2, POT "
                      "sinkhorn_knopp in NumPy, np.random.choice, gather): everything timed in full except the "
                      "Sinkhorn loop = 5 iterations timed and scaled to 100; "
                      f"parts(s)={ {k: round(x, 4) for k, x in parts.items()} }"}
    elif rank == 0:
        line["cpu_baseline"] = None

    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

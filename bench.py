#!/usr/bin/env python
"""bench.py -- the headline benchmark of the ICP hot path on B200 (contract: see the task statement / DESIGN.md).

  python bench.py --gpus N --steps K --warmup W            # this framework (CUDA, sm_100a)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port) on the host cores

Metric (BASELINE.json): frame pairs/s of full registrations (buildRBC + 40 ICP iterations, power method,
weighted residuals) at |F|=|M|=16384, |R|=256, alpha=2e2, c=1e-6; plus, at N=1, the device-timed
us per ICP iteration of ONE pair (latency mode) -- both reported in the same JSON line.

One "step" = one batch of PAIRS_PER_GPU independent synthetic frame pairs per GPU registered end to end.
Multi-GPU: one process per GPU (torchrun), independent pairs per rank, no collective on the hot path
(only the final poses are gathered); value = all pairs of all ranks / max-over-ranks device time.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

M_POINTS, N_REPS, ALPHA, SCALE_C, ITERS = 16384, 256, 2e2, 1e-6, 40
PAIRS_PER_GPU = 256
FLOP_PER_EVAL = 25            # 8 sub, 8 mul, 6 add, 2 mul, 1 add (no FMA contraction)
BYTES_PER_ITER = 32 * M_POINTS + 32 * M_POINTS + 32 * N_REPS + 8 * N_REPS + 64   # 1 058 880 B


def workload_name():
    return (f"batched registration of independent synthetic frame pairs (known transform + noise), |F|=|M|={M_POINTS}, "
            f"|R|={N_REPS}, a={ALPHA:g}, c={SCALE_C:g}, {ITERS} fixed iterations, power method + weighted residuals")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        sm, mx, reasons, power = [], [], set(), []
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_sample(n_pairs, threads, seed0=9000):
    """Time the oracle port of the reference CPU path on `n_pairs` full registrations. Returns (pairs/s, seconds)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle as po
    from icp_b200 import synth
    po.set_threads(threads)
    pairs = [synth.batch_pair(seed0 + i) for i in range(n_pairs)]
    t0 = time.perf_counter()
    for F, Mv, _, _ in pairs:
        po.icp_register(F, Mv, 128, 128, N_REPS, a=ALPHA, c=SCALE_C, rot="power", weighted=True, fixed_iters=ITERS)
    dt = time.perf_counter() - t0
    return n_pairs / dt, dt


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of the path (oracle port; the reference itself
    cannot be built offline: no OpenCL / RBC / Eigen) with all host threads, bounded sample per step."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle as po
    threads = po.hw_threads()
    sample_pairs = 1
    for _ in range(max(1, min(args.warmup, 1))):
        cpu_reference_sample(sample_pairs, threads)
    times = []
    for s in range(args.steps):
        _, dt = cpu_reference_sample(sample_pairs, threads, seed0=9100 + s)
        times.append(dt)
    total = sum(times)
    value = sample_pairs * args.steps / total
    line = {"impl": "reference", "metric": "frame_pairs_per_s", "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(), "sample": f"{sample_pairs} frame pair per step (40 iterations)"},
            "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": threads, "kind": "port",
                             "sample": f"{sample_pairs * args.steps} full registrations, all host threads (std::thread over the NN searches; reductions serial)"},
            "us_per_icp_iteration": 1e6 * total / (args.steps * sample_pairs * ITERS),
            "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs", type=int, default=PAIRS_PER_GPU, help="frame pairs per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    dist = None
    torch = None
    if world > 1:
        import torch
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", rank=rank, world_size=world)

    from icp_b200 import algorithms as alg, capi, parallel, synth
    L = capi.lib()
    ctx = capi.Context(local_rank)           # raises without a GPU / without the extension: no fallback
    n_pairs = args.pairs

    # ---------------- data: synthetic pairs generated on the device, then staged in PINNED host memory for e2e
    base = ctx.upload(synth.base_landmarks())
    batch = alg.ICPBatch(ctx, n_pairs, M_POINTS, N_REPS, a=ALPHA, c=SCALE_C, rot=capi.ROT_POWER_METHOD, weighting=capi.W_WEIGHTED)
    batch.synthesize(base, 5000 + 100003 * rank)
    ctx.sync()
    pair_bytes = M_POINTS * 8 * 4
    hF = capi.PinnedArray((n_pairs, M_POINTS, 8), np.float32)
    hM = capi.PinnedArray((n_pairs, M_POINTS, 8), np.float32)
    capi.check(L.icp_memcpy_d2h(ctx.h, hF.ptr, L.icp_batch_F(batch.h), n_pairs * pair_bytes, 1))
    capi.check(L.icp_memcpy_d2h(ctx.h, hM.ptr, L.icp_batch_M(batch.h), n_pairs * pair_bytes, 1))

    def barrier():
        ctx.sync()
        if dist is not None:
            torch.cuda.synchronize()
            dist.barrier()

    def max_over_ranks(x):
        return parallel.max_over_ranks(x, dist=dist, device="cuda" if dist is not None else None)

    # ---------------- device-resident throughput (value)
    for _ in range(args.warmup):
        batch.register(ITERS)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ctx.timer_start()
    for _ in range(args.steps):
        batch.register(ITERS)              # working set (n_pairs x ~2.4 MB) >> 126 MB L2: no flush needed
    ms_total = ctx.timer_stop()
    barrier()
    clocks = sampler.stop()
    ms_total = max_over_ranks(ms_total)
    ms_per_step = ms_total / args.steps
    value = world * n_pairs * args.steps / (ms_total * 1e-3)
    poses = batch.read_poses()

    # ---------------- end to end through the public API with HOST buffers (pinned): h2d inputs + register + d2h poses
    def e2e_step():
        batch.upload_ptr(0, n_pairs, hF.ptr, hM.ptr, block=False)
        batch.register(ITERS)
        return batch.read_poses()          # blocking d2h of the step's result
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        poses_e2e = e2e_step()
    ctx.sync()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_value = world * n_pairs * args.steps / e2e_s
    assert np.array_equal(poses_e2e.view(np.uint32), poses.view(np.uint32)), "e2e poses differ from the device-resident run"

    # gather of the final poses (the only inter-GPU traffic; off the hot path)
    all_poses = parallel.gather_poses(poses, world * n_pairs, dist=dist, device="cuda" if dist is not None else None)
    if rank == 0:
        assert all_poses.shape == (world * n_pairs, 8) and np.isfinite(all_poses).all()

    line = None
    if rank == 0:
        cfgk = batch.config()
        # ---------------- roofline of the dominant kernel (k_assign: stage-1 nearest representative), timed live
        ms_A = batch.time_kernel(0, 20)
        ms_B = batch.time_kernel(1, 20)
        ms_C = batch.time_kernel(2, 20)
        ms_D = batch.time_kernel(3, 5)
        rates = (C.c_double * 4)()
        capi.check(L.icp_measure_fp32_rates(ctx.h, rates))
        fp32_peak = rates[0]                                        # measured non-fused mul/add issue rate (flop/s)
        flops_A = FLOP_PER_EVAL * M_POINTS * N_REPS * n_pairs       # algorithmic flop per launch of k_assign
        achieved = flops_A / (ms_A * 1e-3)
        traffic = None
        prof = os.path.join(ROOT, "profiles", "ncu_summary.json")
        if os.path.exists(prof):
            try:
                traffic = json.load(open(prof)).get("k_assign_batch", {}).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        step_kernel_ms = ITERS * (ms_A + ms_B + ms_C + ms_D)
        roofline = {"bound": "fp32", "kernel": f"k_assign<S={cfgk['S']},QPT=2> (RBC stage 1: transform + nearest representative)",
                    "achieved": achieved / 1e12, "peak": fp32_peak / 1e12, "unit": "TFLOP/s", "frac": achieved / fp32_peak,
                    "traffic": traffic,
                    "peak_source": "in-run micro-benchmark of the non-fused FMUL/FADD issue rate (the distance may not contract "
                                   "into FMA); tensor cores / HBM are not the bound of this path (SURVEY.md 8d)",
                    "kernel_ms_per_launch": {"A_assign": ms_A, "B_colscan": ms_B, "C_search": ms_C, "D_reduce_solve": ms_D},
                    "kernel_share_of_step": {"A_assign": ITERS * ms_A / ms_per_step, "C_search": ITERS * ms_C / ms_per_step,
                                             "D_reduce_solve": ITERS * ms_D / ms_per_step, "sum_of_kernels_ms": step_kernel_ms},
                    "hbm": {"bound": "hbm", "achieved": BYTES_PER_ITER * n_pairs * ITERS / (ms_per_step * 1e-3) / 1e9, "peak": hbm_peak,
                            "unit": "GB/s", "frac": BYTES_PER_ITER * n_pairs * ITERS / (ms_per_step * 1e-3) / 1e9 / hbm_peak,
                            "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6650 GB/s",
                            "note": "whole iteration, algorithmic bytes 1 058 880 B/iteration/pair; expected << 1: the path is FP32-issue bound"}}

        # ---------------- latency mode: ONE pair owns the GPU (us per ICP iteration, device timed)
        latency = {}
        F0, M0 = hF.array[0], hM.array[0]
        for name, rot in (("power_method", capi.ROT_POWER_METHOD), ("svd", capi.ROT_EIGEN)):
            s = alg.ICPStep(ctx, rot, capi.W_WEIGHTED)
            s.init(M_POINTS, N_REPS, ALPHA, SCALE_C)
            s.write(capi.MEM_D_IN_F, F0); s.write(capi.MEM_D_IN_M, M0)
            ts = []
            for rep in range(8):
                s.reset(); s.buildRBC(); ctx.sync()
                if rep % 2 == 0:
                    ctx.flush_l2(); ctx.sync()                  # cold L2 on the even repetitions
                ctx.timer_start(); s.run(ITERS); ts.append(ctx.timer_stop() * 1e3 / ITERS)
            latency[name] = {"us_per_iteration_warm_l2": min(ts[1::2]), "us_per_iteration_flushed_l2": min(ts[2::2])}
            s.close()
        a, b = C.c_float(), C.c_float()
        capi.check(L.icp_measure_launch_floor(ctx.h, C.byref(a), C.byref(b)))
        latency["launch_floor_us"] = {"stream_launch": a.value, "graph_node": b.value, "kernels_per_iteration": 4}
        latency["readme_r9_270x_us_per_iteration"] = 1100.0

        cpu_baseline = None
        if not args.no_cpu_baseline:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import pyoracle as po
            threads = po.hw_threads()
            v, dt = cpu_reference_sample(2, threads)
            v1, dt1 = cpu_reference_sample(1, 1)
            cpu_baseline = {"value": v, "unit": "pairs/s", "cores": threads, "kind": "port",
                            "sample": f"2 full registrations (40 iterations each) of the same workload, {threads} host threads, {dt:.2f} s",
                            "us_per_iteration": 1e6 * dt / (2 * ITERS),
                            "single_thread": {"value": v1, "us_per_iteration": 1e6 * dt1 / ITERS, "seconds": dt1}}

        line = {"metric": "frame_pairs_per_s", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload_name(), "pairs_per_gpu_per_step": n_pairs, "iterations": ITERS,
                           "parallelism": f"{world} x independent pair batches (no collective on the hot path)",
                           "l2": f"inputs larger than L2: {n_pairs} pairs x ~2.4 MB working set per GPU vs 126 MB L2 (no flush)",
                           "kernel_config": cfgk},
                "us_per_icp_iteration": latency["power_method"]["us_per_iteration_warm_l2"],
                "us_per_pair_iteration_batched": 1e3 * ms_per_step / (n_pairs * ITERS),
                "latency": latency,
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": 2 * n_pairs * pair_bytes,
                        "d2h_bytes_per_step": n_pairs * 8 * 4},
                "gpu_launches": args.steps * (1 + 4 + 4 * ITERS),
                "roofline": roofline,
                "cpu_baseline": cpu_baseline}
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    batch.close()
    ctx.close()


if __name__ == "__main__":
    main()

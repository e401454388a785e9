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


def bench_config(n_pairs, world):
    """`config` of the JSON line: IDENTICAL in both arms (the reference arm times a bounded sample of this workload and says
    so in cpu_baseline.sample), so that the driver's same_config comparison holds."""
    return {"workload": workload_name(), "pairs_per_gpu_per_step": n_pairs, "iterations": ITERS,
            "parallelism": f"{world} x independent pair batches (replicas; no collective on the hot path)",
            "l2": f"inputs larger than L2: {n_pairs} pairs x ~2.4 MB working set per GPU vs 126 MB L2 (no flush)"}


def workload_name():
    return (f"batched registration of independent synthetic frame pairs (known transform + noise), |F|=|M|={M_POINTS}, "
            f"|R|={N_REPS}, a={ALPHA:g}, c={SCALE_C:g}, {ITERS} fixed iterations, power method + weighted residuals")


class ClockSampler:
    """SM clocks / throttle reasons sampled DURING the timed region: NVML in a thread (5 ms period), nvidia-smi as fallback."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        try:
            ids = [int(x) for x in vis.split(",") if x.strip() != ""]
            self.gpu = ids[gpu_index] if ids else gpu_index
        except (ValueError, IndexError):
            self.gpu = gpu_index
        self.samples = []          # (sm_mhz, max_mhz, power_w, reason bits)
        self.proc = None
        self.stop_flag = False
        self.thread = None
        self.nvml = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
            self.nvml = pynvml
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        mx = n.nvmlDeviceGetMaxClockInfo(self.h, n.NVML_CLOCK_SM)
        getr = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(n, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self.stop_flag:
            try:
                self.samples.append((float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)), float(mx),
                                     n.nvmlDeviceGetPowerUsage(self.h) / 1e3, int(getr(self.h))))
            except Exception:
                pass
            time.sleep(0.005)

    def _pump(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.strip().split(",")]
            if len(f) < 9:
                continue
            try:
                bits = 0
                for b, v in zip((0x8, 0x40, 0x20, 0x4), f[5:9]):
                    if v.lower().startswith("active"):
                        bits |= b
                self.samples.append((float(f[1]), float(f[2]), float(f[3]), bits))
            except ValueError:
                continue

    def stop(self):
        self.stop_flag = True
        if self.nvml is None and self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no NVML / nvidia-smi"], "samples": 0}
        if self.proc is not None:
            time.sleep(0.05)
            self.proc.terminate()
        if self.thread is not None:
            self.thread.join(timeout=1.0)
        sm = [x[0] for x in self.samples]
        reasons = set()
        for x in self.samples:
            for b, name in self.BITS.items():
                if x[3] & b:
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(x[1] for x in self.samples) if sm else None,
                "power_w_max": max(x[2] for x in self.samples) if sm else None, "samples": len(sm), "reasons": sorted(reasons),
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def cpu_reference_sample(n_pairs, threads, seed0=9000):
    """Time the oracle port of the reference CPU path on `n_pairs` full registrations. Returns (pairs/s, seconds)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle as po
    from icp_b200 import synth
    po.set_threads(threads)
    distinct = [synth.batch_pair(seed0 + i) for i in range(min(n_pairs, 8))]      # generation (numpy) is not timed
    t0 = time.perf_counter()
    for i in range(n_pairs):
        F, Mv, _, _ = distinct[i % len(distinct)]
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
    sample_pairs = 16
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
            "config": bench_config(args.pairs, args.gpus),
            "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": threads, "kind": "port",
                             "sample": f"{sample_pairs} frame pairs per step x {args.steps} steps = {sample_pairs * args.steps} full registrations "
                                       f"(40 iterations each; 8 distinct pairs of the workload cycled), all {threads} host threads "
                                       "(std::thread over the NN searches; reductions serial), oracle built -O3 -march=native -ffp-contract=off",
                             "us_per_icp_iteration": 1e6 * total / (args.steps * sample_pairs * ITERS)},
            "roofline": {"us_per_icp_iteration": 1e6 * total / (args.steps * sample_pairs * ITERS),
                         "note": "CPU arm: no roofline; the per-iteration latency of the CPU port is reported for comparison"},
            "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def check_sampled_poses(hF, hM, poses, n_check, seed):
    """Registers `n_check` randomly drawn pairs of this rank with the oracle (checker only; outside every timed region) and
    compares the 8-float pose bit for bit with what the GPU produced.  Raises on any difference."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle as po
    po.set_threads(po.hw_threads())
    rng = np.random.default_rng(seed)
    picks = rng.choice(len(poses), size=min(n_check, len(poses)), replace=False)
    for p in picks:
        ref = po.icp_register(np.asarray(hF[p]), np.asarray(hM[p]), 128, 128, N_REPS, a=ALPHA, c=SCALE_C, rot="power", weighted=True, fixed_iters=ITERS)
        if not np.array_equal(ref["T"].view(np.uint32), np.asarray(poses[p]).view(np.uint32)):
            raise AssertionError(f"parity: pose of pair {p} differs from the oracle after {ITERS} iterations: {poses[p]} vs {ref['T']}")
    return len(picks)


def run_config5(args, ctx, alg, capi, parallel, synth, base, rank, world, dist, barrier, max_over_ranks):
    """BASELINE.json configs[4]: 4096 independent synthetic pairs partitioned over the ranks (parallel.pair_range), poses
    gathered at the end and a sample from every rank checked against the oracle.  Device-resident (the pairs are generated on
    the device from their seeds: 4 GiB would have to be staged from the host otherwise).  Strong scaling."""
    total = args.pairs_total
    if total <= 0:
        return None
    lo, hi = parallel.pair_range(total, world, rank)
    n_loc = hi - lo
    b5 = alg.ICPBatch(ctx, n_loc, M_POINTS, N_REPS, a=ALPHA, c=SCALE_C, rot=capi.ROT_POWER_METHOD, weighting=capi.W_WEIGHTED)
    b5.synthesize(base, 7000 + 100003 * rank)
    b5.register(ITERS)                                      # warm-up (graph capture)
    barrier()
    ctx.timer_start()
    for _ in range(2):
        b5.register(ITERS)
    ms = max_over_ranks(ctx.timer_stop()) / 2
    barrier()
    p5 = b5.read_poses()
    # parity: 2 pairs of this rank against the oracle (inputs read back from the device)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle as po
    po.set_threads(po.hw_threads())
    checked = 0
    for p in sorted({0, n_loc - 1}):
        F = b5.debug("F", np.float32, (M_POINTS, 8), pair=p)
        Mv = b5.debug("M", np.float32, (M_POINTS, 8), pair=p)
        ref = po.icp_register(F, Mv, 128, 128, N_REPS, a=ALPHA, c=SCALE_C, rot="power", weighted=True, fixed_iters=ITERS)
        assert np.array_equal(ref["T"].view(np.uint32), p5[p].view(np.uint32)), f"config 5 parity: pair {lo + p} differs from the oracle"
        checked += 1
    slices = b5.slices()
    b5.close()
    allp = parallel.gather_poses(p5, total, dist=dist, device="cuda" if dist is not None else None)
    if rank != 0:
        return None
    assert allp.shape == (total, 8) and np.isfinite(allp).all()
    return {"workload": f"batched registration of {total} independent synthetic frame pairs partitioned over {world} GPU(s) "
                        f"(|F|=|M|={M_POINTS}, |R|={N_REPS}, {ITERS} iterations), poses gathered on rank 0",
            "pairs_total": total, "pairs_per_gpu": n_loc, "value": total / (ms * 1e-3), "unit": "pairs/s", "ms_per_job": ms,
            "scaling": "strong", "data": "synthetic, generated on the device (device-resident; no h2d in this leg)",
            "slices_per_gpu": slices, "parity_checked_pairs": checked * world, "poses_gathered": int(allp.shape[0])}


def run_scaled_configs(ctx, alg, capi, synth, fp32_peak, hbm_peak):
    """BASELINE.json configs[3]: scaled landmark sets on one GPU, one registration at a time (latency engine), device-timed
    us per ICP iteration and fraction of the FP32 (non-fused mul/add) and HBM rooflines, executed and algorithmic."""
    out = []
    iters = 20
    for m, nr, lm in [(65536, 512, (256, 256)), (65536, 1024, (256, 256)), (307200, 512, (640, 480)), (307200, 1024, (640, 480))]:
        F = synth.grid_cloud(*lm)
        F2, M_, _, _ = synth.known_transform_pair(seed=77, deg=2.0, t=(10, -5, 8), F=F)
        s = alg.ICPStep(ctx, capi.ROT_POWER_METHOD, capi.W_WEIGHTED)
        s.init(m, nr, ALPHA, SCALE_C, lm[0], lm[1])
        s.write(capi.MEM_D_IN_F, F2); s.write(capi.MEM_D_IN_M, M_)
        ts = []
        for rep in range(4):
            s.reset(); s.buildRBC(); ctx.sync(); ctx.timer_start(); s.run(iters); ts.append(ctx.timer_stop() * 1e3 / iters)
        us = float(np.median(ts[1:]))
        s.set_count_evals(True)
        s.reset(); s.buildRBC(); s.run(iters); ctx.sync()
        e1, e2 = s.eval_counts()
        e1x, e2x = s.stage1_executed(), s.stage2_executed()
        s.close()
        flop_alg = FLOP_PER_EVAL * (e1 + e2) / iters + 75.0 * m
        flop_exec = FLOP_PER_EVAL * (e1x + e2x) / iters + 75.0 * m
        bytes_alg = 32.0 * m * 2 + 40.0 * nr + 64
        out.append({"m": m, "nr": nr, "us_per_icp_iteration": round(us, 2), "iterations_timed": iters,
                    "stage1_evals_per_iter": e1 // iters, "stage1_executed_per_iter": e1x // iters,
                    "stage2_evals_per_iter": e2 // iters, "stage2_executed_per_iter": e2x // iters,
                    "frac": round(flop_exec / (us * 1e-6) / fp32_peak, 4), "frac_algorithmic": round(flop_alg / (us * 1e-6) / fp32_peak, 4),
                    "bound": "fp32", "hbm_gbs_algorithmic": round(bytes_alg / (us * 1e-6) / 1e9, 1),
                    "hbm_frac": round(bytes_alg / (us * 1e-6) / 1e9 / hbm_peak, 4)})
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs", type=int, default=PAIRS_PER_GPU, help="frame pairs per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--slices", type=int, default=0, help="concurrent slices of the batch (0 = library default: pairs/64 in [1, 8])")
    ap.add_argument("--pairs-total", type=int, default=4096, help="BASELINE config 5: independent pairs partitioned over the GPUs (0 = skip that leg)")
    ap.add_argument("--no-scaled", action="store_true", help="skip the BASELINE config 4 legs (65536 / 307200 landmarks)")
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
    if args.slices > 0:
        batch.set_slices(args.slices)
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

    # ---------------- end to end through the public API with HOST buffers (pinned): h2d inputs + register + d2h poses.
    # A streaming caller alternates two batches through the asynchronous pair of the host-buffer entry (enqueue / collect):
    # the sliced h2d of step k+1 runs while step k registers; every step's inputs are uploaded and every step's poses are
    # read back inside the timed region.
    batch2 = alg.ICPBatch(ctx, n_pairs, M_POINTS, N_REPS, a=ALPHA, c=SCALE_C, rot=capi.ROT_POWER_METHOD, weighting=capi.W_WEIGHTED)
    if args.slices > 0:
        batch2.set_slices(args.slices)
    ring = [batch, batch2]

    def e2e_run(n_steps):
        out = None
        for s in range(n_steps):
            bcur = ring[s % 2]
            if s >= 2:
                out = bcur.collect()                           # poses of step s-2 (blocking d2h read)
            bcur.register_host_async(hF.ptr, hM.ptr, ITERS, 0)
        for s in range(max(0, n_steps - 2), n_steps):            # drain
            out = ring[s % 2].collect()
        return out
    e2e_run(2)
    barrier()
    t0 = time.perf_counter()
    poses_e2e = e2e_run(args.steps)
    ctx.sync()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_value = world * n_pairs * args.steps / e2e_s
    assert np.array_equal(poses_e2e.view(np.uint32), poses.view(np.uint32)), "e2e poses differ from the device-resident run"
    # the blocking single-call entry (one batch, no overlap across steps), for reference
    for _ in range(2):
        batch.register_host(hF.ptr, hM.ptr, ITERS, 0)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        poses_blk = batch.register_host(hF.ptr, hM.ptr, ITERS, 0)
    e2e_blocking_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    assert np.array_equal(poses_blk.view(np.uint32), poses.view(np.uint32)), "blocking e2e poses differ from the device-resident run"
    batch2.close()

    # gather of the final poses (the only inter-GPU traffic; off the hot path)
    all_poses = parallel.gather_poses(poses, world * n_pairs, dist=dist, device="cuda" if dist is not None else None)
    if rank == 0:
        assert all_poses.shape == (world * n_pairs, 8) and np.isfinite(all_poses).all()

    # ---------------- parity of what was just timed: sampled pairs of EVERY rank against the oracle (bit-exact pose after 40 iterations)
    parity_n = check_sampled_poses(hF.array, hM.array, poses, n_check=4, seed=1234 + rank)
    parity_total = parity_n * world            # every rank checks its own sample and raises on a mismatch (a failing rank fails the job)

    # ---------------- BASELINE config 5 as named: 4096 independent pairs PARTITIONED over the ranks (strong scaling), generated on the device
    config5 = run_config5(args, ctx, alg, capi, parallel, synth, base, rank, world, dist, barrier, max_over_ranks)

    line = None
    if rank == 0:
        cfgk = batch.config()
        # ---------------- per-kernel timing (CUDA events on the library's stream, batch of n_pairs per launch) + roofline
        # averages over the ITERS iterations of a fresh registration (the work per iteration shrinks as the pairs converge)
        ms = {"A_assign": batch.time_kernel(0, ITERS), "B_colscan": batch.time_kernel(1, ITERS),
              "C_search": batch.time_kernel(2, ITERS), "D_reduce_solve": batch.time_kernel(3, ITERS)}
        rates = (C.c_double * 4)()
        capi.check(L.icp_measure_fp32_rates(ctx.h, rates))
        fp32_peak = rates[0]                                        # measured non-fused mul/add issue rate (flop/s)
        # distance evaluations per pair-iteration, counted on the device over full registrations of the first pairs of the batch:
        # E1 = m*nr (what stage 1 is algorithmically), E1x = what the pruned kernel A executes, E2 = sum of searched list sizes
        # counted by the batch kernels themselves on the first pairs of the batch (a second small batch with the per-pair
        # counters attached; >= 10 pairs select the same batch-mode kernels)
        n_cnt = min(24, n_pairs)
        os.environ["ICP_B200_BATCH_EVALS"] = "1"
        cb = alg.ICPBatch(ctx, n_cnt, M_POINTS, N_REPS, a=ALPHA, c=SCALE_C, rot=capi.ROT_POWER_METHOD, weighting=capi.W_WEIGHTED)
        os.environ.pop("ICP_B200_BATCH_EVALS", None)
        cb.upload_ptr(0, n_cnt, hF.ptr, hM.ptr, block=True)
        cb.register(ITERS); ctx.sync()
        ev = np.stack([cb.debug("evals", np.uint64, 4, pair=p) for p in range(n_cnt)]).astype(np.float64).sum(0) / (n_cnt * ITERS)
        e1, e2, e1x, e2x = float(ev[0]), float(ev[1]), float(ev[2]), float(ev[3])
        same_cfg = cb.config() == cfgk and cb.cmode() == batch.cmode()
        cb.close()
        prof = {}
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", "ncu_summary.json")))
        except Exception:
            pass
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)

        def fp32_entry(name, kernel, evals_alg, evals_exec, t_ms, prof_key, note):
            alg_flops = FLOP_PER_EVAL * evals_alg * n_pairs
            ach = alg_flops / (t_ms * 1e-3)
            tr = prof.get(prof_key, {})
            traffic = tr.get("dram_bytes_per_launch")
            if traffic is not None and tr.get("pairs_per_launch"):
                traffic = traffic * n_pairs / tr["pairs_per_launch"]      # ncu capture was taken at another batch size
            ach_exec = FLOP_PER_EVAL * evals_exec * n_pairs / (t_ms * 1e-3)
            # frac = what the FP32 pipe actually does (executed evaluations); the algorithmic figure (evaluations the stage IS,
            # including the ones the exact pruning proves unnecessary) is kept beside it
            return {"bound": "fp32", "kernel": kernel, "achieved": ach_exec / 1e12, "peak": fp32_peak / 1e12, "unit": "TFLOP/s",
                    "frac": ach_exec / fp32_peak, "achieved_algorithmic": ach / 1e12, "frac_algorithmic": ach / fp32_peak,
                    "traffic": traffic, "ms_per_launch": t_ms,
                    "algorithmic_evals_per_pair": evals_alg, "executed_evals_per_pair": evals_exec,
                    "flop_per_eval": FLOP_PER_EVAL, "note": note}
        kernels_per_iter = 3 if batch.cmode() in (2, 3) and os.environ.get("ICP_B200_FUSED", "1") != "0" else 4      # D runs in the tail of C'
        c_kernel_name, c_prof_key = {
            3: ("k_search_span (RBC stage 2 over the sorted query records written by k_colscan_sort: the CTA's list span staged in shared "
                "memory by bulk-async copies, list scans + weights, coalesced sorted outputs)", "k_search_span_batch"),
            2: ("k_search_sorted (RBC stage 2 over the queries sorted by representative by k_colscan_sort: list scans + weights, "
                "coalesced sorted outputs)", "k_search_sorted_batch"),
            1: ("k_search_grouped (RBC stage 2: list scans + weights + scatter)", "k_search_grouped_batch"),
        }.get(batch.cmode(), ("k_search<L> (RBC stage 2, L lanes per query)", "k_search_batch"))
        kern = {
            "A_assign": fp32_entry("A", "k_assign_tri (RBC stage 1: transform + nearest representative, triangle-inequality pruning)",
                                   e1, e1x, ms["A_assign"], "k_assign_tri_batch",
                                   "frac counts the evaluations the exact triangle-inequality pruning executes (executed_evals_per_pair); "
                                   "frac_algorithmic counts the m*nr evaluations the stage is algorithmically (SURVEY 8d) and may exceed 1"),
            "C_search": fp32_entry("C", c_kernel_name, e2, e2x, ms["C_search"], c_prof_key,
                                   "25 flop per evaluation (19 issued: the two constant homogeneous lanes are skipped bit-exactly); frac counts "
                                   "the evaluations left after the exact temporal pruning (DESIGN 4.5), frac_algorithmic every evaluation stage 2 "
                                   "is algorithmically; counters of a 24-pair batch with the same kernel configuration: " + str(same_cfg)),
        }
        dominant = max(ms, key=ms.get)
        step_kernel_ms = ITERS * sum(ms.values())
        hbm_ach = BYTES_PER_ITER * n_pairs * ITERS / (ms_per_step * 1e-3) / 1e9
        roofline = dict(kern[dominant]) if dominant in kern else {"bound": "latency", "kernel": dominant}
        roofline.update({
            "dominant_kernel": dominant,
            "peak_source": "in-run micro-benchmark of the non-fused FMUL/FADD issue rate (the distance may not contract into FMA); "
                           "tensor cores / HBM are not the bound of this path (SURVEY.md 8d); traffic = dram bytes of the committed "
                           "ncu --set full capture (profiles/ncu_summary.json) scaled to this batch size",
            "kernel_ms_per_launch": ms,
            "kernel_share_of_step": {k: ITERS * v / ms_per_step for k, v in ms.items()},
            "sum_of_kernels_ms": step_kernel_ms,
            "kernels": kern,
            "hbm": {"bound": "hbm", "achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_ach / hbm_peak,
                    "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6650 GB/s",
                    "note": "whole iteration, algorithmic bytes 1 058 880 B/iteration/pair; expected << 1: the path is FP32-issue / latency bound"}})

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
        latency["launch_floor_us"] = {"stream_launch": a.value, "graph_node": b.value, "kernels_per_iteration": 4, "kernels_per_iteration_batched": kernels_per_iter}
        latency["readme_r9_270x_us_per_iteration"] = 1100.0

        cpu_baseline = None
        if not args.no_cpu_baseline:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import pyoracle as po
            threads = po.hw_threads()
            cpu_reference_sample(2, threads)                 # warm-up (thread pool, page faults)
            n_mt, n_st = 96, 6
            v, dt = cpu_reference_sample(n_mt, threads)
            v1, dt1 = cpu_reference_sample(n_st, 1)
            cpu_baseline = {"value": v, "unit": "pairs/s", "cores": threads, "kind": "port",
                            "sample": f"{n_mt} full registrations (40 iterations each; 8 distinct pairs of the same workload cycled), "
                                      f"{threads} host threads, {dt:.2f} s",
                            "us_per_iteration": 1e6 * dt / (n_mt * ITERS),
                            "single_thread": {"value": v1, "us_per_iteration": 1e6 * dt1 / (n_st * ITERS), "seconds": dt1,
                                              "sample": f"{n_st} full registrations"}}

        scaled = None
        if not args.no_scaled:
            scaled = run_scaled_configs(ctx, alg, capi, synth, fp32_peak, hbm_peak)
        us_iter = latency["power_method"]["us_per_iteration_warm_l2"]
        # BASELINE.json's first metric (us per ICP iteration, one pair, latency mode) and the launch/sync floor it is reported
        # against live INSIDE `roofline` (and `config`), which the driver's record keeps
        roofline.update({"us_per_icp_iteration": us_iter, "latency": latency, "scaled": scaled,
                         "us_per_pair_iteration_batched": 1e3 * ms_per_step / (n_pairs * ITERS)})
        cfg = bench_config(n_pairs, world)
        line = {"metric": "frame_pairs_per_s", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": cfg,
                "run": {"slices": f"{batch.slices()} concurrent slices of the batch on separate streams (one graph each; fork/join on the "
                                  "library stream): kernel D / B of one slice overlap A / C of the others; in the e2e entry slice i+1 "
                                  "uploads while slice i registers",
                        "kernel_config": cfgk, "us_per_icp_iteration": us_iter},
                "us_per_icp_iteration": us_iter,
                "us_per_pair_iteration_batched": 1e3 * ms_per_step / (n_pairs * ITERS),
                "parity_checked_pairs": parity_total,
                "config5": config5,
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "pairs/s", "api": "icp_batch_register_host_async + icp_batch_collect on two alternating batches",
                        "blocking_single_call_value": world * n_pairs * args.steps / e2e_blocking_s, "h2d_bytes_per_step": 2 * n_pairs * pair_bytes,
                        "d2h_bytes_per_step": n_pairs * 8 * 4, "parity_checked_pairs": parity_total},
                "gpu_launches": args.steps * batch.slices() * (1 + 5 + kernels_per_iter * ITERS),
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

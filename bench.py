#!/usr/bin/env python
"""bench.py — the hot path's headline benchmark (BASELINE.json: "grasp candidates/sec end-to-end (15ch)").

One "step" = one pass of the whole path (sample -> local frame -> hand search -> grasp image -> LeNet score)
over the batch of sample indices of BASELINE config 3: synthetic 300k-point cluttered cloud (seed 3),
num_samples = 100000 per GPU, 15-channel images, the reference's 15-channel LeNet weights.

  value : samples/s with inputs resident in HBM (gpdb_detect_resident), CUDA events on the launching stream
  e2e   : the same through gpdb_detect with HOST buffers (H2D of the sample indices, D2H of every result)
  N > 1 : one process per GPU (torchrun); the sample indices are sharded by contiguous slice over the same
          cloud (weak scaling: 100k samples per GPU), ONE NCCL all-gather of fixed-stride score slots.
  --impl reference : the CPU restatement of the reference path (oracle/, all host threads) on a bounded sample.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from gpd_b200 import abi, scenes  # noqa: E402

METRIC = "grasp candidates/sec end-to-end (15ch)"
UNIT = "samples/s (1 sample = 8 hand poses swept, ~1.5 classified)"
SAMPLES_PER_GPU = 100000
FLOPS_PER_IMAGE = {"conv1": 2 * 56 * 56 * 20 * 375, "conv2": 2 * 24 * 24 * 50 * 500, "ip1": 2 * 7200 * 500 + 2 * 500 * 2}


def load_weights(ch=15):
    z = np.load(os.path.join(ROOT, "gpd_b200", "weights", f"lenet_{ch}ch.npz"))
    names = ["conv1_weights", "conv1_biases", "conv2_weights", "conv2_biases", "ip1_weights", "ip1_biases",
             "ip2_weights", "ip2_biases"]
    return [z[n] for n in names]


def make_workload(n_gpus, samples_per_gpu):
    cloud = scenes.synthetic_table_scene(3)
    n_total = samples_per_gpu * n_gpus
    ncl = len(cloud["xyz"])
    if n_total <= ncl:
        sidx = np.random.default_rng(3).choice(ncl, n_total, replace=False).astype(np.int32)
    else:  # config 4 style: with replacement
        sidx = np.random.default_rng(4).integers(0, ncl, n_total).astype(np.int32)
    return cloud, sidx


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        self.cmd = ["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(index)]
        self.rows = []
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(self.cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        busy = sorted(sm)[len(sm) // 2:]
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline(cloud, sidx, weights, target_seconds=15.0, nthreads=0):
    """Times the CPU restatement of the reference path (oracle/) on a bounded sample of the same workload."""
    from oracle import oracle

    p = abi.default_params(15)
    oc = oracle.OracleCloud(cloud["xyz"], cloud["normals"], cloud["cam_source"], cloud["view_points"])
    wp = oracle.WeightPack(weights)
    nt = nthreads or oracle.num_threads()
    if not nthreads and nt > 32:
        # SMT siblings often hurt this memory-bound code: time a small sample with all logical CPUs and with half of
        # them and keep the faster setting (the CPU arm gets its best configuration)
        probe = sidx[: min(len(sidx), 1024)]
        rates = {}
        for cand in (nt, nt // 2):
            oc.detect(p, wp, probe[:256], nthreads=cand)
            t = time.perf_counter()
            oc.detect(p, wp, probe, nthreads=cand)
            rates[cand] = len(probe) / (time.perf_counter() - t)
        nt = max(rates, key=rates.get)
    n1 = min(len(sidx), 16 * nt)
    while True:  # grow the sample until it takes about target_seconds (bounded: at most 4 rounds)
        t = time.perf_counter()
        r = oc.detect(p, wp, sidx[:n1], nthreads=nt)
        dt = time.perf_counter() - t
        if dt >= 0.5 * target_seconds or n1 >= len(sidx):
            break
        n1 = int(min(len(sidx), max(2 * n1, n1 * target_seconds / max(dt, 1e-3))))
    return {"value": n1 / dt, "unit": UNIT, "cores": nt, "kind": "port",
            "sample": f"first {n1} of the step's sample indices, {r['n_candidates']} candidates classified, {dt:.1f} s; "
                      f"stage seconds candidates/images/classify = "
                      f"{r['stage_seconds'][0]:.2f}/{r['stage_seconds'][1]:.2f}/{r['stage_seconds'][2]:.2f}"}, r


def bench_preprocess(ctx, hbm_peak, with_cpu):
    """Secondary measurement: CandidatesGenerator::preprocessPointCloud on the device —
    NaN / workspace filter, voxelisation at 0.003, normal estimation r = 0.03 — for the RAW cloud of the same scene
    family (seed 3, ~0.9 M points -> ~0.5 M voxels), through gpdb_preprocess with HOST buffers."""
    from gpd_b200 import lib
    raw = scenes.synthetic_raw_scene(3)
    pp = lib.preprocess_params()
    n_out = 0
    for _ in range(2):
        n_out = ctx.preprocess(raw["xyz"], raw["cam_source"], raw["view_points"], pp, read_back=False)
    reps, wall, dev_ms = 3, 0.0, np.zeros(6)
    for _ in range(reps):
        t0 = time.perf_counter()
        ctx.preprocess(raw["xyz"], raw["cam_source"], raw["view_points"], pp, read_back=False)
        wall += time.perf_counter() - t0
        dev_ms += ctx.preprocess_timings()
    wall /= reps
    dev_ms /= reps
    m = len(raw["xyz"])
    out = {"metric": "raw points preprocessed / s (filter + voxelise 0.003 + normals r=0.03), host buffers in, processed cloud resident",
           "value": m / wall, "unit": "raw points/s", "raw_points": m, "processed_points": int(n_out),
           "ms_per_call_wall": round(wall * 1e3, 3),
           "device_ms": {k: round(float(v), 3) for k, v in zip(["upload", "filter", "voxelise", "grid", "normals", "total"], dev_ms)},
           "h2d_bytes_per_call": int(m * (12 + 1))}
    cloud = ctx.get_cloud()
    from oracle import oracle
    oc = oracle.OracleCloud(cloud["xyz"], cloud["normals"], cloud["cam_source"], raw["view_points"])
    probe = np.arange(0, n_out, max(1, n_out // 256))[:256]
    n_nb = float(np.mean([len(oc.radius_search(cloud["xyz"][i], pp.normals_radius)[0]) for i in probe]))
    alg = n_out * (n_nb * 16 + 24)  # float4 gather per neighbour + one float64 normal out
    ach = alg / (dev_ms[4] * 1e-3) / 1e9
    out["k_normals"] = {"ms": round(float(dev_ms[4]), 3), "mean_neighbours": n_nb, "algorithmic_bytes": alg, "GB/s": round(ach, 1),
                        "frac_hbm": round(ach / hbm_peak, 4)}
    if with_cpu:
        t0 = time.perf_counter()
        ro = oracle.preprocess(raw["xyz"], raw["cam_source"], raw["view_points"], pp)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": m / dt, "unit": "raw points/s", "cores": oracle.num_threads(), "kind": "port",
                               "sample": f"the same {m} raw points, once: {dt:.2f} s (voxelise {ro['seconds'][0]:.2f} s, normals {ro['seconds'][1]:.2f} s)"}
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cloud, sidx = make_workload(args.gpus, args.samples)
    weights = load_weights()
    times, n_used, cores = [], 0, 0
    for it in range(args.warmup + args.steps):
        cb, _ = cpu_baseline(cloud, sidx, weights, target_seconds=10.0)
        if it >= args.warmup:
            times.append(cb["value"])
        cores = cb["cores"]
        n_used = cb["sample"]
    v = float(np.mean(times))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * args.samples * args.gpus / v, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64 geometry / f32 LeNet", "data": "synthetic",
            "config": {"workload": "BASELINE config 3: synthetic 300k-pt cluttered cloud (seed 3), 15-channel, CPU "
                                   "restatement of the reference path (oracle/, OpenMP, all host threads); each step times a "
                                   "bounded sample and ms_per_step is extrapolated to the full step",
                       "num_samples": args.samples * args.gpus},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": n_used},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def _divert_stdout():
    """Send everything written to file descriptor 1 (Python AND native libraries: NCCL prints its version banner to
    stdout when NCCL_DEBUG >= VERSION) to stderr; the ONE JSON line is printed after _restore_stdout."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    return saved


def _restore_stdout(saved):
    sys.stdout.flush()
    try:
        ctypes.CDLL(None).fflush(None)  # flush C stdio buffers into the diverted descriptor first
    except Exception:
        pass
    os.dup2(saved, 1)
    os.close(saved)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--samples", type=int, default=SAMPLES_PER_GPU, help="samples per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--lenet-impl", type=int, default=0)
    ap.add_argument("--no-preprocess", action="store_true", help="skip the secondary gpdb_preprocess measurement")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    from gpd_b200 import lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    saved_stdout = _divert_stdout()
    final_line = None
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # NCCL's banner / warnings: not on stdout
        dist.init_process_group("nccl", device_id=dev)
    n_gpus = world
    cloud, sidx_all = make_workload(n_gpus, args.samples)
    weights = load_weights()
    if world > 1:  # the cloud is broadcast from rank 0 over NVLink; ranks then hold identical copies
        for key in ("xyz", "normals"):
            t = torch.from_numpy(cloud[key]).to(dev)
            dist.broadcast(t, 0)
            cloud[key] = t.cpu().numpy()
    per = args.samples
    sidx = np.ascontiguousarray(sidx_all[rank * per:(rank + 1) * per])
    n = len(sidx)

    params = lib.default_params(channels=15, device=local, lenet_impl=args.lenet_impl)
    P = params.num_hand_axes * params.num_orientations
    ctx = lib.Context(params)
    stream = torch.cuda.current_stream(dev)
    ctx.set_stream(stream.cuda_stream)
    ctx.set_weights(weights)
    ctx.set_cloud(cloud["xyz"], cloud["normals"], cloud["cam_source"], cloud["view_points"])

    d_sidx = torch.from_numpy(sidx).to(dev)
    d_flags = torch.zeros(n * P, dtype=torch.uint8, device=dev)
    d_scores = torch.zeros(n * P, dtype=torch.float32, device=dev)
    g_scores = torch.zeros(world * n * P, dtype=torch.float32, device=dev) if world > 1 else None
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    stats = abi.Result()

    def step_resident():
        nc = ctx.detect_resident(d_sidx.data_ptr(), n, d_flags.data_ptr(), d_scores.data_ptr(), stats)
        if world > 1:
            dist.all_gather_into_tensor(g_scores, d_scores)
        return nc

    for _ in range(args.warmup):
        step_resident()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    stage_ms = np.zeros(8)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches = 0
    ncand = 0
    torch.cuda.synchronize()
    for k in range(args.steps):
        flush.fill_(k)  # evict L2 between timed steps
        ev[k][0].record(stream)
        ncand = step_resident()
        ev[k][1].record(stream)
        stage_ms += ctx.last_timings()
        launches += int(stats.kernel_launches)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop()
    total_ms = sum(a.elapsed_time(b) for a, b in ev)
    if world > 1:
        t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
        cnt = torch.tensor([float(ncand)], device=dev, dtype=torch.float64)
        dist.all_reduce(cnt)
        ncand_all = int(cnt.item())
    else:
        ncand_all = ncand
    ms_per_step = total_ms / args.steps
    value = n * world / (ms_per_step * 1e-3)

    # ---- end to end through the public C-ABI call with HOST buffers
    h_sidx = torch.from_numpy(sidx).pin_memory()
    h_np = h_sidx.numpy()
    res = abi.Result()
    for _ in range(2):
        ctx.detect_raw(h_np, res)
        lib.free_result(res)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_e2e = 0.0
    d2h = 0
    for k in range(args.steps):
        flush.fill_(k)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        nc = ctx.detect_raw(h_np, res)
        if world > 1:
            sc = torch.from_numpy(np.ctypeslib.as_array(res.pose_scores, (n * P,))).to(dev)
            dist.all_gather_into_tensor(g_scores, sc)
            torch.cuda.synchronize()
        t_e2e += time.perf_counter() - t0
        d2h = n * 9 * 8 + n + n * P + n * P * 4 + nc * ctypes.sizeof(abi.Pose)
        lib.free_result(res)
    if world > 1:
        t = torch.tensor([t_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e = float(t.item())
    e2e_value = n * world / (t_e2e / args.steps)
    # the reference-facing call of GraspDetector::detectGrasps proper returns the num_selected best grasps only
    # (selectGrasps, cfg default 100): gpdb_detect_select picks them on the device (secondary number, N = 1)
    e2e_select = None
    if world == 1:
        for _ in range(2):
            ctx.detect_select_raw(h_np, 100, res)
            lib.free_result(res)
        t_sel = 0.0
        for k in range(args.steps):
            flush.fill_(k)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            nsel = ctx.detect_select_raw(h_np, 100, res)
            t_sel += time.perf_counter() - t0
            lib.free_result(res)
        e2e_select = {"value": n / (t_sel / args.steps), "unit": UNIT, "num_selected": 100, "h2d_bytes_per_step": int(n * 4),
                      "d2h_bytes_per_step": int(nsel * ctypes.sizeof(abi.Pose)),
                      "call": "gpdb_detect_select (detectGrasps + selectGrasps, top-100 picked on the device)"}

    if rank == 0:
        st = stage_ms / args.steps  # per step, this rank
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        tf_peak = peaks.get("bf16_tflops_sustained", 1400.0)
        peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
        # neighbourhood statistics for the algorithmic byte counts, from the oracle's grid
        from oracle import oracle
        oc = oracle.OracleCloud(cloud["xyz"], cloud["normals"], cloud["cam_source"], cloud["view_points"])
        probe = sidx[:: max(1, n // 256)][:256]
        n_hs = float(np.mean([len(oc.radius_search(cloud["xyz"][i], 0.11)[0]) for i in probe]))
        n_img = float(np.mean([len(oc.radius_search(cloud["xyz"][i], 0.10)[0]) for i in probe]))
        kernels = {
            "k_hands": {"ms": st[1], "bound": "hbm", "bytes": n * (n_hs * 24 + P * (ctypes.sizeof(abi.Pose) + 1))},
            "k_images": {"ms": st[2], "bound": "hbm", "bytes": ncand * (n_img * 24 + 60 * 60 * 15 + ctypes.sizeof(abi.Pose))},
            "lenet_conv1": {"ms": st[5], "bound": "tensor", "flops": ncand * FLOPS_PER_IMAGE["conv1"]},
            "lenet_conv2": {"ms": st[6], "bound": "tensor", "flops": ncand * FLOPS_PER_IMAGE["conv2"]},
            "lenet_ip": {"ms": st[7], "bound": "tensor", "flops": ncand * FLOPS_PER_IMAGE["ip1"]},
        }
        dom = max(kernels, key=lambda k: kernels[k]["ms"])
        kd = kernels[dom]
        # measured DRAM traffic (ncu --set full capture, profiles/traffic.json: bytes per image / per sample) scaled to
        # the units of one step, like `achieved` (which sums all launches of the kernel in the step)
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(dom)
            if tj:
                traffic = tj["bytes_per_unit"] * (n if tj["unit"] == "sample" else ncand)
        except Exception:
            pass
        if kd["bound"] == "hbm":
            ach = kd["bytes"] / (kd["ms"] * 1e-3) / 1e9
            roof = {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                    "frac": ach / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                    "per": "step: all launches of the kernel summed (algorithmic bytes, CUDA-event time, ncu DRAM bytes)",
                    "algorithmic_bytes": kd["bytes"]}
        else:
            ach = kd["flops"] / (kd["ms"] * 1e-3) / 1e12
            roof = {"kernel": dom, "bound": "tensor", "achieved": ach, "peak": tf_peak, "unit": "TFLOP/s",
                    "frac": ach / tf_peak, "traffic": traffic, "peak_source": peak_src,
                    "per": "step: all launches of the kernel summed (algorithmic flops, CUDA-event time, ncu DRAM bytes)",
                    "algorithmic_flops": kd["flops"]}
        per_kernel = {}
        for k, v in kernels.items():
            if v["bound"] == "hbm":
                a = v["bytes"] / max(v["ms"], 1e-9) / 1e6
                per_kernel[k] = {"ms_per_step": round(v["ms"], 3), "GB/s": round(a, 1), "frac_hbm": round(a / hbm_peak, 4)}
            else:
                a = v["flops"] / max(v["ms"], 1e-9) / 1e9
                per_kernel[k] = {"ms_per_step": round(v["ms"], 3), "TFLOP/s": round(a, 2), "frac_tensor": round(a / tf_peak, 4)}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64 geometry / f32 LeNet", "data": "synthetic",
            "config": {"workload": "BASELINE config 3 (configs[2], the one north_star's 200k/s target is quoted on): synthetic "
                                   "300k-pt cluttered cloud seed 3, num_samples=100000 per GPU, 15-channel images, reference "
                                   "15-ch LeNet weights", "num_samples": n * world, "poses_per_sample": P,
                       "parallelism": f"samples sharded over {world} GPU(s), one all-gather of score slots",
                       "l2": "512 MB flush write between timed steps; per-step CUDA events summed",
                       "lenet": lib.lib().gpdb_build_info().decode()},
            "rates": {"samples_per_s": value, "poses_evaluated_per_s": value * P,
                      "candidates_classified_per_s": ncand_all / (ms_per_step * 1e-3)},
            "stage_ms_per_step": {"frames": round(st[0], 3), "hand_search": round(st[1], 3), "images": round(st[2], 3),
                                  "lenet": round(st[3], 3), "call_total": round(st[4], 3)},
            "kernels": per_kernel,
            "roofline": roof,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(n * 4), "d2h_bytes_per_step": int(d2h),
                    "timing": "wall clock around gpdb_detect (host buffers, synchronous), max over ranks"},
            "e2e_select": e2e_select,
            "gpu_launches": launches,
            "neighbourhood": {"mean_r0.11": n_hs, "mean_r0.10": n_img, "candidates_per_step": ncand_all},
        }
        if not args.no_cpu_baseline:
            cb, _ = cpu_baseline(cloud, sidx, weights)
            line["cpu_baseline"] = cb
        if world == 1 and not args.no_preprocess:
            line["preprocess"] = bench_preprocess(ctx, hbm_peak, not args.no_cpu_baseline)
        final_line = json.dumps(line)
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    _restore_stdout(saved_stdout)
    if final_line is not None:
        print(final_line, flush=True)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- ORB frames/s on BASELINE.json's configs[1]: 640x480 mono, 1000 features, 8 levels,
scale 1.2, FAST 20/7, a batch of 4096 synthetic frames per step and per GPU.

    python bench.py --gpus 1 --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU extractor (oracle/_ref)
    torchrun --nproc-per-node N ... bench.py --gpus N ...    # one rank per GPU, frames sharded, no data collective

One "step" = one pass of the hot path (ORBextractor::operator(), reference ORBextractor.cc:1078-1162) over
one batch of frames.  `value` = whole-job frames/s with the frames resident in HBM (device in, device out,
CUDA-event time on the launching stream, max over ranks).  `e2e` = the same metric through the C-ABI with
pinned HOST buffers, H2D/D2H inside the timed region (host wall clock around the calls).  Inputs (1.26 GB
per step) are larger than L2 (126 MB), so no explicit L2 flush is needed between steps.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

W, H = 640, 480
NFEATURES, SCALE, NLEVELS, INI_TH, MIN_TH = 1000, 1.2, 8, 20, 7
WORKLOAD = "TUM RGB-D 640x480 mono, 1000 features, 8 levels, scale 1.2, FAST 20/7"
METRIC = "ORB frames/s @640x480 1000kp"
N_BASE = 64

# Level geometry of the workload: used for the algorithmic byte counts.
LEVEL_SIZES = [(640, 480), (533, 400), (444, 333), (370, 278), (309, 231), (257, 193), (214, 161), (179, 134)]
PIXELS_ALL_LEVELS = sum(w * h for w, h in LEVEL_SIZES)                       # 950 532
BORDERED_BYTES = sum((w + 38) * (h + 38) for w, h in LEVEL_SIZES)            # 1 158 012
PATH_BYTES_PER_FRAME = W * H + BORDERED_BYTES + NFEATURES * 60               # 1 525 212


def stage_algorithmic_bytes(mean_candidates):
    """Compulsory HBM bytes per frame of each stage when run as its own kernel."""
    return {
        "pyramid": W * H + BORDERED_BYTES,                 # read the frame once, write every bordered plane once
        "fast": PIXELS_ALL_LEVELS + 8 * mean_candidates,   # read every level pixel once, write 8-byte candidates
        "octree": 8 * mean_candidates + 24 * NFEATURES,    # read candidates, write kept-keypoint records
        "blur": 2 * PIXELS_ALL_LEVELS,                     # read level, write blurred level
        "describe": 24 * NFEATURES + NFEATURES * 60,       # read records, write cv::KeyPoint + descriptor
    }


def natural_base():
    """The reference's own fixture images (decoded copies in tests/golden/images.npz): two 640x480 robot frames
    and 640x480 centre crops of the 2x-upsampled 512x512 luna / TUM frames."""
    with np.load(os.path.join(ROOT, "tests", "golden", "images.npz")) as z:
        imgs = [z["robot866"], z["robot2196"]]
        for k in ("luna", "tum_room4"):
            up = np.kron(z[k], np.ones((2, 2), np.uint8))
            imgs.append(up[272:272 + H, 192:192 + W])
    return np.stack(imgs)


def make_frames(n_frames, seed, dataset="synthetic"):
    """Deterministic frame set: base frames (N_BASE corner-rich synthetic frames from tests/common.synth_frame,
    or the natural fixture images), each output frame a cyclic shift / flip / +-20 % gain of one of them.
    uint8 torch tensor on the CPU."""
    import torch
    from common import synth_frame
    if dataset == "natural":
        base = torch.from_numpy(natural_base())
    else:
        base = torch.from_numpy(np.stack([synth_frame(seed * 1000 + i, W, H) for i in range(N_BASE)]))
    nb = base.shape[0]
    g = torch.Generator().manual_seed(1234 + seed)
    dy = torch.randint(0, H, (n_frames,), generator=g).tolist()
    dx = torch.randint(0, W, (n_frames,), generator=g).tolist()
    flip = torch.randint(0, 2, (n_frames,), generator=g).tolist()
    gain = (0.8 + 0.4 * torch.rand(n_frames, generator=g)).tolist()
    out = torch.empty((n_frames, H, W), dtype=torch.uint8)
    for i in range(n_frames):
        f = torch.roll(base[i % nb], (dy[i], dx[i]), (0, 1))
        if flip[i]:
            f = torch.flip(f, (1,))
        out[i] = (f.float() * gain[i]).clamp_(0, 255).to(torch.uint8)
    return out


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs (one looping nvidia-smi
    process, 50 ms period; only samples taken between __enter__ and __exit__ are kept)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self._p = index, [], None

    def __enter__(self):
        try:
            self._p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                        "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self._p.stdout.readline()          # first sample = the loop is running before the timed region starts
        except Exception:
            self._p = None
        return self

    def __exit__(self, *a):
        if self._p is None:
            return
        time.sleep(0.06)
        self._p.terminate()
        try:
            out, _ = self._p.communicate(timeout=5)
        except Exception:
            self._p.kill()
            out = ""
        for line in out.splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) >= 7:
                self.rows.append(c)

    def summary(self):
        def num(x):
            try:
                return float(x)
            except ValueError:
                return None
        sm = [num(r[0]) for r in self.rows if num(r[0]) is not None]
        mx = [num(r[1]) for r in self.rows if num(r[1]) is not None]
        pw = [num(r[2]) for r in self.rows if num(r[2]) is not None]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": reasons, "samples": len(self.rows)}


def cpu_reference_run(frames_np, seconds, threads):
    """Time the reference's own CPU implementation (oracle/_ref/ref_extract: ORBextractor.cc compiled
    verbatim + shim primitives, normal allocator) or, if that binary is absent, the C port."""
    from oracle import refio
    if refio.have_ref(bump=False):
        r = refio.bench_reference(frames_np, threads, seconds, NFEATURES, SCALE, NLEVELS, INI_TH, MIN_TH, (0, 0))
        return {"kind": "reference", "fps": r["fps"], "frames": r["frames"], "wall_s": r["wall_s"], "cores": threads,
                "p50_ms": r["p50_ms"], "mean_keypoints": r["mean_keypoints"]}
    from oracle import pyoracle
    o = pyoracle.OracleExtractor(NFEATURES, SCALE, NLEVELS, INI_TH, MIN_TH)
    t0 = time.perf_counter()
    n = 0
    lat = []
    while time.perf_counter() - t0 < seconds or n == 0:
        a = time.perf_counter()
        o.extract(frames_np[n % len(frames_np)], (0, 0))
        lat.append((time.perf_counter() - a) * 1e3)
        n += 1
    wall = time.perf_counter() - t0
    return {"kind": "port", "fps": n / wall, "frames": n, "wall_s": wall, "cores": 1, "p50_ms": float(np.median(lat)),
            "mean_keypoints": None}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    sample = make_frames(64, seed=0).numpy()
    steps, warm = args.steps, args.warmup
    per_step = max(2.0, min(20.0, 150.0 / max(1, steps + warm)))
    for _ in range(warm):
        cpu_reference_run(sample, per_step, threads)
    tot_frames, tot_wall, last = 0, 0.0, None
    for _ in range(steps):
        last = cpu_reference_run(sample, per_step, threads)
        tot_frames += last["frames"]
        tot_wall += last["wall_s"]
    fps = tot_frames / tot_wall
    sample_desc = "64 synthetic 640x480 frames cycled for %.1f s per step, one extractor + one frame per thread" % per_step
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": 1e3 * tot_wall / max(1, steps), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step": tot_frames // max(1, steps), "where": "host CPU",
                   "implementation": "reference ORBextractor.cc compiled verbatim against oracle/shim (scalar C primitives)"
                   if last["kind"] == "reference" else "oracle C port"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": last["cores"], "kind": last["kind"], "sample": sample_desc,
                         "p50_ms_per_frame": last["p50_ms"]},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=4096, help="frames per step and per GPU")
    ap.add_argument("--group", type=int, default=256, help="frames per launch group (OrbxParams.max_batch)")
    ap.add_argument("--e2e-steps", type=int, default=4)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--latency-iters", type=int, default=300)
    ap.add_argument("--no-natural", action="store_true")
    ap.add_argument("--prof-steps", type=int, default=3)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.warmup < 3:
        args.warmup = 3   # timing rule: at least 3 warm-up steps

    import torch
    import torch.distributed as dist
    import extractorb_b200 as ex

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    F = args.frames
    host_frames = make_frames(F, seed=rank).pin_memory()          # rank-private frames: weak scaling, no exchange
    dev_frames = host_frames.cuda(non_blocking=False)
    ext = ex.ORBextractor(NFEATURES, SCALE, NLEVELS, INI_TH, MIN_TH, device=local, max_batch=args.group, profile=False)
    cap = ext.max_keypoints(W, H)
    d_kps = torch.empty((F, cap, 7), dtype=torch.float32, device="cuda")
    d_desc = torch.empty((F, cap, 32), dtype=torch.uint8, device="cuda")
    d_counts = torch.zeros((F, 2), dtype=torch.int32, device="cuda")
    stream = torch.cuda.Stream()          # the kernels are launched on this stream; events are recorded on it too

    def step_device():
        ext.extract_batch_raw(dev_frames.data_ptr(), ex.MEM_DEVICE, F, W, H, W, W * H, (0, 0), d_kps.data_ptr(), d_desc.data_ptr(),
                              cap, d_counts.data_ptr(), ex.MEM_DEVICE, stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    barrier()
    launches0 = ext.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        e0.record(stream)
        for _ in range(args.steps):
            step_device()
        e1.record(stream)
        barrier()
    ms = e0.elapsed_time(e1)
    launches = ext.launch_count() - launches0
    # Per-stage CUDA-event times (and the dominant kernel's launch duration for the roofline) come from a second,
    # profiled handle: profiling keeps every launch group on one stream so that stage times do not overlap, while
    # the timed region above overlaps consecutive groups on two streams.
    prof = ex.ORBextractor(NFEATURES, SCALE, NLEVELS, INI_TH, MIN_TH, device=local, max_batch=args.group, profile=True)
    def step_prof():
        prof.extract_batch_raw(dev_frames.data_ptr(), ex.MEM_DEVICE, F, W, H, W, W * H, (0, 0), d_kps.data_ptr(), d_desc.data_ptr(),
                               cap, d_counts.data_ptr(), ex.MEM_DEVICE, stream.cuda_stream)
    step_prof()
    prof.stage_times()
    for _ in range(args.prof_steps):
        step_prof()
    stage_ms, _ = prof.stage_times()
    prof.close()
    ncand = sum(len(ext.level_candidates(l, frame=0)[0]) for l in range(NLEVELS))   # of one frame of the workload
    counts = d_counts.cpu().numpy()
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    kp_sum = torch.tensor([float(counts[:, 0].sum())], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)     # timing: max over ranks
        dist.all_reduce(kp_sum, op=dist.ReduceOp.SUM)  # statistics only: frames need no data collective
    ms_max = float(t.item())
    value = F * world * args.steps / (ms_max * 1e-3)

    # ---- end to end through the C-ABI with pinned host buffers ----
    h_kps = torch.empty((F, cap, 7), dtype=torch.float32).pin_memory()
    h_desc = torch.empty((F, cap, 32), dtype=torch.uint8).pin_memory()
    h_counts = torch.zeros((F, 2), dtype=torch.int32).pin_memory()

    def step_host():
        ext.extract_batch_raw(host_frames.data_ptr(), ex.MEM_HOST, F, W, H, W, W * H, (0, 0), h_kps.data_ptr(), h_desc.data_ptr(),
                              cap, h_counts.data_ptr(), ex.MEM_HOST, None)

    step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        step_host()
    ext.synchronize()
    e2e_s = time.perf_counter() - t0
    # what bounds e2e: the plain pinned host->device copy rate of the same bytes on this box
    barrier()                                  # all ranks copy at the same time: the box's aggregate host->device rate is what counts
    tp0 = time.perf_counter()
    for _ in range(3):
        dev_frames.copy_(host_frames, non_blocking=True)
    torch.cuda.synchronize()
    h2d_gbs = 3 * F * W * H / (time.perf_counter() - tp0) / 1e9
    th = torch.tensor([h2d_gbs], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(th, op=dist.ReduceOp.SUM)
    h2d_gbs_all = float(th.item())
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = F * world * args.e2e_steps / float(te.item())
    assert np.array_equal(h_counts.numpy(), counts), "host and device paths disagree"

    # ---- the same device-resident measurement on the natural-image set (fewer candidates per frame) ----
    natural = None
    if world == 1 and not args.no_natural:
        nat = make_frames(F, seed=7, dataset="natural").cuda()
        def step_nat():
            ext.extract_batch_raw(nat.data_ptr(), ex.MEM_DEVICE, F, W, H, W, W * H, (0, 0), d_kps.data_ptr(), d_desc.data_ptr(),
                                  cap, d_counts.data_ptr(), ex.MEM_DEVICE, stream.cuda_stream)
        for _ in range(3):
            step_nat()
        torch.cuda.synchronize()
        n0, n1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0.record(stream)
        for _ in range(5):
            step_nat()
        n1e.record(stream)
        torch.cuda.synchronize()
        nat_ms = n0.elapsed_time(n1e)
        natural = {"value": F * 5 / (nat_ms * 1e-3), "unit": "frames/s", "mean_keypoints_per_frame": float(d_counts[:, 0].float().mean().item()),
                   "what": "4 natural fixture images (robot x2, luna, TUM room4) with the same shift/flip/gain augmentation, device-resident"}
        del nat

    # ---- single-frame latency (the reference's own calling pattern: one operator() per frame, host to host) ----
    lat = None
    if rank == 0:
        one = ex.ORBextractor(NFEATURES, SCALE, NLEVELS, INI_TH, MIN_TH, device=local, max_batch=1)
        f0 = host_frames[0].numpy()
        k1 = np.zeros(cap, ex.KP_DTYPE)
        d1 = np.zeros((cap, 32), np.uint8)
        import ctypes as C
        n1, m1 = C.c_int(0), C.c_int(0)
        ts = []
        for i in range(args.latency_iters + 50):
            t0 = time.perf_counter()
            one._check(one._L.orbx_extract(one._h, f0.ctypes.data, W, H, W, 0, 0, k1.ctypes.data, d1.ctypes.data, cap,
                                           C.byref(n1), C.byref(m1)))
            ts.append((time.perf_counter() - t0) * 1e6)
        ts = np.array(ts[50:])
        lat = {"p50_us": float(np.percentile(ts, 50)), "p99_us": float(np.percentile(ts, 99)), "iters": int(len(ts)),
               "what": "orbx_extract on one 640x480 host frame, host buffers in and out, wall clock"}
        one.close()

    if rank == 0:
        mean_kp = float(kp_sum.item()) / (F * world)
        alg = stage_algorithmic_bytes(ncand)
        groups_per_step = (F + args.group - 1) // args.group
        n_group_launches = groups_per_step * args.prof_steps
        dom = max(stage_ms, key=lambda k: stage_ms[k])
        kernel_launches = {"pyramid": NLEVELS, "fast": 1, "octree": 1, "blur": 1, "describe": 1}
        dom_ms_per_launch = stage_ms[dom] / (n_group_launches * kernel_launches[dom])
        frames_per_launch = min(args.group, F)
        bytes_per_launch = alg[dom] * frames_per_launch / kernel_launches[dom]
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        achieved = bytes_per_launch / (dom_ms_per_launch * 1e-3) / 1e9
        traffic, pipe = None, None
        try:   # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture (profiles/, 256-frame launch)
            cap_ = json.load(open(os.path.join(ROOT, "profiles", "ncu_capture.json")))[dom]
            traffic = cap_["dram_bytes_per_launch"] * frames_per_launch / cap_["frames_per_launch"]
            pipe = cap_.get("pipes")
        except Exception:
            pass
        per_rank_fps = value / world
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_step_per_gpu": F, "frames_per_launch_group": args.group,
                       "input": "device-resident uint8 frames (%.2f GB per GPU, larger than the 126 MB L2: no flush needed)" % (F * W * H / 1e9),
                       "sharding": "independent frames per rank, no data collective (NCCL: timing/statistics all-reduce only)",
                       "mean_keypoints_per_frame": mean_kp, "fast_candidates_frame0": ncand},
            "p50_us_per_frame_amortised": 1e6 / per_rank_fps,
            "single_frame_latency": lat,
            "natural_set": natural,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "pipes_from_ncu": pipe, "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback",
                         "algorithmic_bytes_per_frame": alg[dom], "ms_per_launch": dom_ms_per_launch,
                         "frames_per_launch": frames_per_launch,
                         "path": {"algorithmic_bytes_per_frame": PATH_BYTES_PER_FRAME, "achieved": PATH_BYTES_PER_FRAME * per_rank_fps / 1e9,
                                  "frac": PATH_BYTES_PER_FRAME * per_rank_fps / 1e9 / peak}},
            "stage_ms_per_step": {k: v / args.prof_steps for k, v in stage_ms.items()},
            "stage_timing": "CUDA events around each stage, %d extra steps on a profiled handle (single compute stream)" % args.prof_steps,
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": F * W * H, "d2h_bytes_per_step": F * (cap * 60 + 8),
                    "steps": args.e2e_steps, "timing": "host wall clock, pinned host buffers, max over ranks",
                    "h2d_copy_gbs_measured": h2d_gbs, "h2d_copy_gbs_all_ranks_concurrent": h2d_gbs_all,
                    "h2d_gbs_needed": e2e_value / world * W * H / 1e9,
                    "note": "H2D of group g+1, kernels of group g and D2H of group g-1 overlap; the plain pinned H2D copy rate of this box bounds e2e"},
            "gpu_launches": int(launches),
            "clocks": clocks.summary(),
        }
        if world == 1 and not args.no_cpu_baseline:
            sample = host_frames[:64].numpy()
            threads = os.cpu_count() or 1
            r = cpu_reference_run(sample, args.cpu_seconds, threads)
            line["cpu_baseline"] = {"value": r["fps"], "unit": "frames/s", "cores": r["cores"], "kind": r["kind"],
                                    "sample": "first 64 frames of the workload cycled for %.0f s, one extractor + one frame per thread" % args.cpu_seconds,
                                    "p50_ms_per_frame": r["p50_ms"]}
        print(json.dumps(line))
    ext.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

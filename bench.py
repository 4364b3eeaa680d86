#!/usr/bin/env python
"""bench.py -- ORB frames/s on BASELINE.json's configs[1]: 640x480 mono, 1000 features, 8 levels,
scale 1.2, FAST 20/7, a batch of 4096 synthetic frames per step and per GPU.

    python bench.py --gpus 1 --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU extractor (oracle/_ref)
    torchrun --nproc-per-node N ... bench.py --gpus N ...    # one rank per GPU, frames sharded, no data collective

One "step" = one pass of the hot path (ORBextractor::operator(), reference ORBextractor.cc:1078-1162) over
one batch of frames.
  value  whole-job frames/s with the frames resident in HBM (device in, device out, CUDA-event time on the
         launching stream, max over ranks).
  e2e    the same metric through the C-ABI with pinned HOST buffers, H2D of the frames and D2H of keypoints +
         descriptors inside the timed region (host wall clock, max over ranks, >= 20 steps).  e2e.bound_fps is
         what the same copies alone reach on this box (same call, same slots and streams, ORBX_FLAG_COPY_ONLY),
         all ranks copying at once; e2e.frac_of_bound = value / bound.  mvImagePyramid stays on the device in
         that number (e2e.pyramid_download = false); e2e_with_pyramid copies it out as the reference leaves it.
  e2e_single_process (N > 1)  rank 0 alone feeds all N GPUs through orbx_extract_batch_multi (one host thread +
         handle per device, launch groups pulled from a shared cursor) while the other ranks sleep on the store.
Inputs (1.26 GB per step) are larger than L2 (126 MB), so no explicit L2 flush is needed between steps.
"""
import argparse
import json
import os
import struct
import subprocess
import sys
import tempfile
import threading
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

W, H = 640, 480
NFEATURES, SCALE, NLEVELS, INI_TH, MIN_TH = 1000, 1.2, 8, 20, 7
WORKLOAD = "TUM RGB-D 640x480 mono, 1000 features, 8 levels, scale 1.2, FAST 20/7"
METRIC = "ORB frames/s @640x480 1000kp"
N_BASE = 64
FRAMES_PER_STEP = 4096

# Level geometry of the workload: used for the algorithmic byte counts.
LEVEL_SIZES = [(640, 480), (533, 400), (444, 333), (370, 278), (309, 231), (257, 193), (214, 161), (179, 134)]
PIXELS_ALL_LEVELS = sum(w * h for w, h in LEVEL_SIZES)                       # 950 532
BORDERED_BYTES = sum((w + 38) * (h + 38) for w, h in LEVEL_SIZES)            # 1 158 012
PATH_BYTES_PER_FRAME = W * H + BORDERED_BYTES + NFEATURES * 60               # 1 525 212


def common_config(frames_per_step):
    """The workload definition: identical in both arms (ours and --impl reference)."""
    return {"workload": WORKLOAD, "width": W, "height": H, "nfeatures": NFEATURES, "nlevels": NLEVELS, "scale_factor": SCALE,
            "ini_th_fast": INI_TH, "min_th_fast": MIN_TH, "frames_per_step_per_gpu": frames_per_step,
            "frames": "synthetic: 64 seeded corner-rich base frames, each output frame a cyclic shift / flip / +-20 % gain of one of them",
            "sharding": "independent frames per rank, no data collective (NCCL: timing/statistics all-reduce only)"}


def stage_algorithmic_bytes(mean_candidates):
    """Compulsory HBM bytes per frame of each stage when run as its own kernel."""
    return {
        "pyramid": W * H + sum(w * h for w, h in LEVEL_SIZES),  # read the frame, write every level (the border is written only on demand)
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


def make_frames(n_frames, seed, dataset="synthetic", w=W, h=H, n_base=N_BASE):
    """Deterministic frame set: base frames (corner-rich synthetic frames from tests/common.synth_frame, or the natural
    fixture images), each output frame a cyclic shift / flip / +-20 % gain of one of them.  uint8 torch tensor on the CPU."""
    import torch
    from common import synth_frame
    if dataset == "natural":
        base = torch.from_numpy(natural_base())
    else:
        base = torch.from_numpy(np.stack([synth_frame(seed * 1000 + i, w, h) for i in range(n_base)]))
    nb = base.shape[0]
    g = torch.Generator().manual_seed(1234 + seed)
    dy = torch.randint(0, h, (n_frames,), generator=g).tolist()
    dx = torch.randint(0, w, (n_frames,), generator=g).tolist()
    flip = torch.randint(0, 2, (n_frames,), generator=g).tolist()
    gain = (0.8 + 0.4 * torch.rand(n_frames, generator=g)).tolist()
    out = torch.empty((n_frames, h, w), dtype=torch.uint8)
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


def cv2_primitives_ms(frames_np, reps=3):
    """Lower bound for ANY CPU implementation of the path built on OpenCV: only the OpenCV primitives
    the reference calls -- 7 x resize, 8 x copyMakeBorder, cv::FAST over every level (whole level, iniThFAST, with NMS),
    8 x GaussianBlur 7x7 -- python-cv2 (SIMD build), one thread, per frame.  No cell loop, quadtree, orientation or descriptors."""
    try:
        import cv2
    except Exception:
        return None
    cv2.setNumThreads(1)
    fast = cv2.FastFeatureDetector_create(INI_TH, True)
    ts = []
    for _ in range(reps):
        for img in frames_np:
            t0 = time.perf_counter()
            lv = [img]
            for (w, h) in LEVEL_SIZES[1:]:
                lv.append(cv2.resize(lv[-1], (w, h), interpolation=cv2.INTER_LINEAR))
            for im in lv:
                cv2.copyMakeBorder(im, 19, 19, 19, 19, cv2.BORDER_REFLECT_101)
                fast.detect(im, None)
                cv2.GaussianBlur(im, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
            ts.append((time.perf_counter() - t0) * 1e3)
    return float(np.median(ts))


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
    sample_desc = ("64 frames of the workload cycled for %.1f s per step (%d frames per step on this box), one extractor + one frame per thread"
                   % (per_step, tot_frames // max(1, steps)))
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": 1e3 * tot_wall / max(1, steps), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": common_config(args.frames),
        "where": "host CPU",
        "implementation": ("reference ORBextractor.cc compiled verbatim against oracle/shim (scalar C restatements of the OpenCV primitives)"
                           if last["kind"] == "reference" else "oracle C port"),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": last["cores"], "kind": last["kind"], "sample": sample_desc,
                         "p50_ms_per_frame": last["p50_ms"]},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    ms = cv2_primitives_ms(sample[:16])
    if ms is not None:
        line["cpu_baseline"]["cv2_primitives_ms"] = ms
        line["cpu_baseline"]["cv2_primitives_fps_all_cores"] = threads * 1e3 / ms
    print(json.dumps(line))
    return 0


def write_orbf(path, frames_np):
    """Frame file of tests/cpp/dropin_main.cpp / oracle/ref_main.cpp: int32 'ORBF', n, w, h; n*h*w bytes."""
    n, h, w = frames_np.shape
    with open(path, "wb") as f:
        f.write(struct.pack("<4i", 0x4642524f, n, w, h))
        f.write(np.ascontiguousarray(frames_np).tobytes())


def cpp_operator_latency(frames_np, iters):
    """Latency of the C++ drop-in operator() itself (tests/cpp/dropin_main latency): 5- / 6-argument, with / without the
    download of mvImagePyramid."""
    exe = os.path.join(ROOT, "tests", "cpp", "dropin_main")
    if not os.access(exe, os.X_OK):
        return None
    with tempfile.TemporaryDirectory() as td:
        fin = os.path.join(td, "lat.orbf")
        write_orbf(fin, frames_np)
        r = subprocess.run([exe, "latency", fin, str(iters), str(NFEATURES), repr(SCALE), str(NLEVELS), str(INI_TH), str(MIN_TH)],
                           capture_output=True, text=True, timeout=120)
    if r.returncode != 0:
        return {"error": r.stderr.strip()[-200:]}
    return json.loads(r.stdout)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=FRAMES_PER_STEP, help="frames per step and per GPU")
    ap.add_argument("--group", type=int, default=256, help="frames per launch group (OrbxParams.max_batch)")
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--latency-iters", type=int, default=300)
    ap.add_argument("--no-natural", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the other BASELINE configs, the pyramid / single-process e2e variants")
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

    def mk(**kw):
        return ex.ORBextractor(NFEATURES, SCALE, NLEVELS, INI_TH, MIN_TH, device=local, max_batch=args.group, **kw)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(v):
        t = torch.tensor([float(v)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(v):
        t = torch.tensor([float(v)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    F = args.frames
    host_frames = make_frames(F, seed=rank).pin_memory()          # rank-private frames: weak scaling, no exchange
    dev_frames = host_frames.cuda(non_blocking=False)
    ext = mk()
    cap = ext.max_keypoints(W, H)
    d_kps = torch.empty((F, cap, 7), dtype=torch.float32, device="cuda")
    d_desc = torch.empty((F, cap, 32), dtype=torch.uint8, device="cuda")
    d_counts = torch.zeros((F, 2), dtype=torch.int32, device="cuda")
    stream = torch.cuda.Stream()          # the kernels are launched on this stream; events are recorded on it too

    def step_device(e=ext, frames=dev_frames, n=F, w=W, h=H, c=cap, k=d_kps, d=d_desc, cn=d_counts):
        e.extract_batch_raw(frames.data_ptr(), ex.MEM_DEVICE, n, w, h, w, w * h, (0, 0), k.data_ptr(), d.data_ptr(), c, cn.data_ptr(),
                            ex.MEM_DEVICE, stream.cuda_stream)

    # ---- value: device-resident frames, CUDA events on the launching stream ----
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
    counts = d_counts.cpu().numpy()
    # whole-job statistics: SUM of frames / keypoints, MAX of the elapsed time over ranks (extractorb_b200/sharding.py; NCCL here, gloo in
    # the CPU tests) -- the only collectives of the job, frames need no data exchange
    from extractorb_b200 import sharding
    stats = sharding.reduce_stats(frames=F * args.steps, keypoints=int(counts[:, 0].sum()), elapsed_ms=ms, device="cuda")
    ms_max = stats["elapsed_ms_max"]
    value = stats["frames"] / (ms_max * 1e-3)
    kp_sum = float(stats["keypoints"])
    ncand = sum(len(ext.level_candidates(l, frame=0)[0]) for l in range(NLEVELS))   # of one frame of the workload

    # Per-stage CUDA-event times (and the dominant kernel's launch duration for the roofline) come from a second,
    # profiled handle: profiling keeps every launch group on one stream so that stage times do not overlap, while
    # the timed region above overlaps consecutive groups on two streams.
    prof = mk(profile=True)
    step_device(prof)
    prof.stage_times()
    for _ in range(args.prof_steps):
        step_device(prof)
    stage_ms, _ = prof.stage_times()
    prof.close()

    # ---- e2e: the same call with pinned HOST buffers, copies inside the timed region ----
    h_kps = torch.empty((F, cap, 7), dtype=torch.float32).pin_memory()
    h_desc = torch.empty((F, cap, 32), dtype=torch.uint8).pin_memory()
    h_counts = torch.zeros((F, 2), dtype=torch.int32).pin_memory()

    def step_host(e):
        e.extract_batch_raw(host_frames.data_ptr(), ex.MEM_HOST, F, W, H, W, W * H, (0, 0), h_kps.data_ptr(), h_desc.data_ptr(),
                            cap, h_counts.data_ptr(), ex.MEM_HOST, None)

    def timed_host_steps(e, steps):
        step_host(e)
        step_host(e)
        barrier()
        per = []
        t0 = time.perf_counter()
        for _ in range(steps):
            a = time.perf_counter()
            step_host(e)
            per.append(time.perf_counter() - a)
        total = time.perf_counter() - t0
        return total, per

    e2e_total, e2e_per = timed_host_steps(ext, args.e2e_steps)
    assert np.array_equal(h_counts.numpy(), counts), "host and device paths disagree"
    e2e_total_max = allmax(e2e_total)
    e2e_value = F * world * args.e2e_steps / e2e_total_max
    e2e_p50_max = allmax(float(np.median(e2e_per)))
    # what bounds it: the same copies alone (same call, same staging slots / streams / events, no kernel), all ranks at once
    probe = mk(flags=ex.FLAG_COPY_ONLY)
    probe_steps = max(5, args.e2e_steps // 2)
    pr_total, pr_per = timed_host_steps(probe, probe_steps)
    probe.close()
    bound_fps = F * world * probe_steps / allmax(pr_total)
    # plain one-direction copy rates, for the record (all ranks at once)
    barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        dev_frames.copy_(host_frames, non_blocking=True)
    torch.cuda.synchronize()
    h2d_gbs = 3 * F * W * H / (time.perf_counter() - t0) / 1e9
    barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        h_desc.copy_(d_desc, non_blocking=True)
    torch.cuda.synchronize()
    d2h_gbs = 3 * d_desc.numel() / (time.perf_counter() - t0) / 1e9
    h2d_gbs_all, d2h_gbs_all = allsum(h2d_gbs), allsum(d2h_gbs)
    e2e = {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": F * W * H, "d2h_bytes_per_step": F * (cap * 60 + 8),
           "steps": args.e2e_steps, "timing": "host wall clock over all steps, pinned host buffers, max over ranks",
           "p50_ms_per_step": 1e3 * e2e_p50_max, "p50_fps": F * world / e2e_p50_max,
           "pyramid_download": False,
           "bound_fps": bound_fps, "frac_of_bound": e2e_value / bound_fps,
           "bound": "the same orbx_extract_batch call with ORBX_FLAG_COPY_ONLY (identical H2D / D2H copies through the same staging "
                    "slots, streams and events, no kernel), all ranks at once",
           "h2d_copy_gbs_per_rank": h2d_gbs, "h2d_copy_gbs_all_ranks": h2d_gbs_all, "d2h_copy_gbs_per_rank": d2h_gbs,
           "d2h_copy_gbs_all_ranks": d2h_gbs_all, "h2d_gbs_used": e2e_value / world * W * H / 1e9}

    extras = {}
    if not args.no_extras:
        # ---- e2e with mvImagePyramid copied out as well (the reference leaves it on the host, ORBextractor.cc:1173-1177) ----
        try:
            Fp = min(F, 1024)
            fb = ext.pyramid_layout(W, H)[0]
            h_pyr = torch.empty((Fp, fb), dtype=torch.uint8).pin_memory()
            ext.set_pyramid_output(h_pyr.data_ptr(), fb)

            def step_pyr():
                ext.extract_batch_raw(host_frames.data_ptr(), ex.MEM_HOST, Fp, W, H, W, W * H, (0, 0), h_kps.data_ptr(), h_desc.data_ptr(),
                                      cap, h_counts.data_ptr(), ex.MEM_HOST, None)
            step_pyr()
            barrier()
            t0 = time.perf_counter()
            for _ in range(4):
                step_pyr()
            tp_ = allmax(time.perf_counter() - t0)
            ext.set_pyramid_output(None, 0)
            extras["e2e_with_pyramid"] = {"value": Fp * world * 4 / tp_, "unit": "frames/s", "frames_per_step_per_gpu": Fp, "steps": 4,
                                          "d2h_bytes_per_frame": fb + cap * 60 + 8, "pyramid_download": True,
                                          "what": "as e2e, plus every frame's bordered pyramid block (border written on the device) copied to pinned host memory"}
            del h_pyr
        except Exception as exc:   # never lose the headline line to an optional measurement
            extras["e2e_with_pyramid"] = {"error": str(exc)[:200]}
            ext.set_pyramid_output(None, 0)

        # ---- cross-rank identity: every rank extracts one common probe set; the result CRCs must agree ----
        try:
            Pn = 64
            pf = make_frames(Pn, seed=4242).cuda()
            step_device(frames=pf, n=Pn)
            torch.cuda.synchronize()
            cn = d_counts[:Pn].cpu().numpy()
            kk, dd = d_kps[:Pn].cpu().numpy(), d_desc[:Pn].cpu().numpy()
            crc = 0
            for f in range(Pn):
                n = int(cn[f, 0])
                crc = zlib.crc32(cn[f].tobytes() + kk[f, :n].tobytes() + dd[f, :n].tobytes(), crc)
            t = torch.tensor([crc], dtype=torch.int64, device="cuda")
            allc = [torch.zeros_like(t) for _ in range(world)]
            if world > 1:
                dist.all_gather(allc, t)
            else:
                allc = [t]
            crcs = [int(x.item()) for x in allc]
            extras["cross_rank_identity"] = {"frames": Pn, "crc32": crc, "identical_on_all_ranks": len(set(crcs)) == 1, "ranks": world}
            assert len(set(crcs)) == 1, "ranks disagree on the common probe set: %r" % crcs
            del pf
        except AssertionError:
            raise
        except Exception as exc:
            extras["cross_rank_identity"] = {"error": str(exc)[:200]}

        # ---- the other BASELINE configs, per rank, device-resident (frames/s summed over ranks) ----
        try:
            others = {}
            for name, (w_, h_, nf_, nl_, fr_, grp_, nb_) in {"C4_kitti_1241x376_2000kp": (1241, 376, 2000, 8, 1024, 256, 16),
                                                             "C5_4k_3840x2160_8000kp_12lv_batch64": (3840, 2160, 8000, 12, 64, 32, 4)}.items():
                fr = make_frames(fr_, seed=100 + rank, w=w_, h=h_, n_base=nb_).cuda()
                e_ = ex.ORBextractor(nf_, SCALE, nl_, INI_TH, MIN_TH, device=local, max_batch=grp_)
                c_ = e_.max_keypoints(w_, h_)
                k_ = torch.empty((fr_, c_, 7), dtype=torch.float32, device="cuda")
                d_ = torch.empty((fr_, c_, 32), dtype=torch.uint8, device="cuda")
                n_ = torch.zeros((fr_, 2), dtype=torch.int32, device="cuda")
                for _ in range(2):
                    step_device(e_, fr, fr_, w_, h_, c_, k_, d_, n_)
                barrier()
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record(stream)
                reps = 3
                for _ in range(reps):
                    step_device(e_, fr, fr_, w_, h_, c_, k_, d_, n_)
                a1.record(stream)
                barrier()
                t_ = allmax(a0.elapsed_time(a1))
                others[name] = {"value": fr_ * world * reps / (t_ * 1e-3), "unit": "frames/s", "frames_per_step_per_gpu": fr_,
                                "mean_keypoints_per_frame": float(n_[:, 0].float().mean().item())}
                e_.close()
                del fr, k_, d_, n_
            extras["other_configs"] = others
        except Exception as exc:
            extras["other_configs"] = {"error": str(exc)[:200]}

        # ---- C3: EuRoC 752x480 stereo, 1200 features per image: left / right extractors on two host threads (the reference's own
        # pattern, src/Frame.cc:109-112), then Frame::ComputeStereoMatches on the device-resident results (orbx_stereo_match_batch) ----
        try:
            from common import make_stereo_pair, STEREO_MB, STEREO_MBF
            ws_, hs_, nfs_, Ps, G = 752, 480, 1200, 1024, 256
            lbase = make_frames(16, seed=300 + rank, w=ws_, h=hs_, n_base=16).numpy()
            rbase = np.stack([make_stereo_pair(l, 1 + i) for i, l in enumerate(lbase)])
            left = torch.from_numpy(lbase).repeat(Ps // 16, 1, 1).cuda()
            right = torch.from_numpy(rbase).repeat(Ps // 16, 1, 1).cuda()
            eL = ex.ORBextractor(nfs_, SCALE, NLEVELS, INI_TH, MIN_TH, device=local, max_batch=G)
            eR = ex.ORBextractor(nfs_, SCALE, NLEVELS, INI_TH, MIN_TH, device=local, max_batch=G)
            cs = eL.max_keypoints(ws_, hs_)
            bufs = {s: (torch.empty((G, cs, 7), dtype=torch.float32, device="cuda"), torch.empty((G, cs, 32), dtype=torch.uint8, device="cuda"),
                        torch.zeros((G, 2), dtype=torch.int32, device="cuda")) for s in "lr"}
            u_ = torch.empty((G, cs), dtype=torch.float32, device="cuda")
            z_ = torch.empty((G, cs), dtype=torch.float32, device="cuda")
            nm_ = torch.zeros(G, dtype=torch.int32, device="cuda")

            def one_side(e, imgs, g0, b):
                e.extract_batch_raw(imgs[g0:g0 + G].data_ptr(), ex.MEM_DEVICE, G, ws_, hs_, ws_, ws_ * hs_, (0, 0), b[0].data_ptr(), b[1].data_ptr(),
                                    cs, b[2].data_ptr(), ex.MEM_DEVICE, None)

            def stereo_pass():
                matched = 0
                for g0 in range(0, Ps, G):
                    tl = threading.Thread(target=one_side, args=(eL, left, g0, bufs["l"]))
                    tr = threading.Thread(target=one_side, args=(eR, right, g0, bufs["r"]))
                    tl.start(); tr.start(); tl.join(); tr.join()
                    ex.stereo_match_batch_raw(eL, eR, G, bufs["l"][0].data_ptr(), bufs["l"][1].data_ptr(), bufs["l"][2].data_ptr(),
                                              bufs["r"][0].data_ptr(), bufs["r"][1].data_ptr(), bufs["r"][2].data_ptr(), cs, STEREO_MB, STEREO_MBF,
                                              u_.data_ptr(), z_.data_ptr(), nm_.data_ptr())
                    eL.synchronize()
                    matched = int(nm_.sum().item())
                return matched
            stereo_pass()
            barrier()
            t0 = time.perf_counter()
            reps = 3
            for _ in range(reps):
                matched = stereo_pass()
            ts_ = allmax(time.perf_counter() - t0)
            extras["C3_euroc_752x480_stereo_1200kp"] = {
                "value": Ps * world * reps / ts_, "unit": "stereo pairs/s", "pairs_per_step_per_gpu": Ps, "launch_group": G,
                "mean_matches_per_pair": matched / G,
                "what": "two extractors on two host threads per launch group of 256 pairs (device-resident frames), then orbx_stereo_match_batch "
                        "on the device-resident keypoints / descriptors / pyramids; host wall clock, max over ranks"}
            eL.close(); eR.close()
            del left, right, bufs, u_, z_, nm_
        except Exception as exc:
            extras["C3_euroc_752x480_stereo_1200kp"] = {"error": str(exc)[:200]}

    # ---- the same device-resident measurement on the natural-image set (fewer candidates per frame) ----
    natural = None
    if world == 1 and not args.no_natural:
        nat = make_frames(F, seed=7, dataset="natural").cuda()
        for _ in range(3):
            step_device(frames=nat)
        torch.cuda.synchronize()
        n0, n1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0.record(stream)
        for _ in range(5):
            step_device(frames=nat)
        n1e.record(stream)
        torch.cuda.synchronize()
        nat_ms = n0.elapsed_time(n1e)
        natural = {"value": F * 5 / (nat_ms * 1e-3), "unit": "frames/s", "mean_keypoints_per_frame": float(d_counts[:, 0].float().mean().item()),
                   "what": "4 natural fixture images (robot x2, luna, TUM room4) with the same shift/flip/gain augmentation, device-resident"}
        del nat

    # ---- single-frame latency (the reference's own calling pattern: one operator() per frame, host to host) ----
    lat, cpp_lat = None, None
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
        if world == 1 and not args.no_extras:
            try:
                cpp_lat = cpp_operator_latency(host_frames[:8].numpy(), args.latency_iters)
            except Exception as exc:
                cpp_lat = {"error": str(exc)[:200]}

    # ---- N > 1: the same host-buffer job from ONE process (rank 0 feeds every GPU through orbx_extract_batch_multi) ----
    single_process = None
    if world > 1 and not args.no_extras:
        barrier()
        store = dist.distributed_c10d._get_default_store()
        if rank == 0:
            try:
                exts = [ext] + [ex.ORBextractor(NFEATURES, SCALE, NLEVELS, INI_TH, MIN_TH, device=d, max_batch=args.group) for d in range(1, world)]
                Fm = F * world
                m_frames = host_frames.repeat(world, 1, 1).pin_memory()
                m_kps = torch.empty((Fm, cap, 7), dtype=torch.float32).pin_memory()
                m_desc = torch.empty((Fm, cap, 32), dtype=torch.uint8).pin_memory()
                m_counts = torch.zeros((Fm, 2), dtype=torch.int32).pin_memory()

                def step_multi():
                    return ex.extract_batch_multi_raw(exts, m_frames.data_ptr(), Fm, W, H, W, W * H, (0, 0), m_kps.data_ptr(), m_desc.data_ptr(),
                                                      cap, m_counts.data_ptr())
                step_multi()
                step_multi()
                per = []
                t0 = time.perf_counter()
                msteps = max(5, args.e2e_steps // 2)
                for _ in range(msteps):
                    a = time.perf_counter()
                    shares = step_multi()
                    per.append(time.perf_counter() - a)
                tm = time.perf_counter() - t0
                ok = bool(np.array_equal(m_counts.numpy()[:F], counts))
                single_process = {"value": Fm * msteps / tm, "unit": "frames/s", "steps": msteps, "p50_fps": Fm / float(np.median(per)),
                                  "frames_per_device_last_step": shares, "results_equal_per_rank_run": ok,
                                  "what": "rank 0 alone: orbx_extract_batch_multi, one host thread + handle per GPU, launch groups pulled from a "
                                          "shared cursor, the same pinned host buffers and copies as e2e; the other ranks sleep on the store"}
                for e_ in exts[1:]:
                    e_.close()
                del m_frames, m_kps, m_desc, m_counts
            except Exception as exc:
                single_process = {"error": str(exc)[:200]}
            store.set("orbx_multi_done", "1")
        else:
            store.wait(["orbx_multi_done"])
        barrier()

    if rank == 0:
        mean_kp = kp_sum / (F * world)
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
        dom_achieved = bytes_per_launch / (dom_ms_per_launch * 1e-3) / 1e9
        traffic, pipe, path_traffic = None, None, None
        try:   # dram__bytes_read.sum + dram__bytes_write.sum of one ncu capture per kernel (profiles/, 256-frame launch group)
            cap_all = json.load(open(os.path.join(ROOT, "profiles", "ncu_capture.json")))
            cap_ = cap_all[dom]      # entries are per STAGE: all kernels of the stage over one launch group
            traffic = cap_["dram_bytes_per_launch"] / kernel_launches[dom] * frames_per_launch / cap_["frames_per_launch"]
            pipe = cap_.get("pipes")
            path_traffic = sum(v["dram_bytes_per_launch"] for k, v in cap_all.items() if k in kernel_launches) \
                * frames_per_launch / cap_["frames_per_launch"]
        except Exception:
            pass
        per_rank_fps = value / world
        path_achieved = PATH_BYTES_PER_FRAME * per_rank_fps / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": common_config(F),
            "frames_per_launch_group": args.group,
            "input": "device-resident uint8 frames (%.2f GB per GPU, larger than the 126 MB L2: no flush needed)" % (F * W * H / 1e9),
            "mean_keypoints_per_frame": mean_kp, "fast_candidates_frame0": ncand,
            "p50_us_per_frame_amortised": 1e6 / per_rank_fps,
            "single_frame_latency": lat,
            "cpp_operator_latency": cpp_lat,
            "natural_set": natural,
            # the whole path's algorithmic bytes per frame x frames/s against the measured HBM peak; the dominant
            # kernel's own figure (its stage bytes / its launch time) is the sub-object
            "roofline": {"bound": "hbm", "achieved": path_achieved, "peak": peak, "unit": "GB/s", "frac": path_achieved / peak,
                         "traffic": path_traffic, "algorithmic_bytes_per_frame": PATH_BYTES_PER_FRAME, "frames_per_launch": frames_per_launch,
                         "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback",
                         "what": "whole path: (W*H + sum (w_l+38)(h_l+38) + 60*nfeatures) bytes/frame x frames/s; traffic = ncu DRAM bytes of all "
                                 "kernels of one launch group",
                         "dominant_kernel": {"kernel": dom, "achieved": dom_achieved, "frac": dom_achieved / peak, "traffic": traffic,
                                             "algorithmic_bytes_per_frame": alg[dom], "ms_per_launch": dom_ms_per_launch, "pipes_from_ncu": pipe}},
            "stage_ms_per_step": {k: v / args.prof_steps for k, v in stage_ms.items()},
            "stage_timing": "CUDA events around each stage, %d extra steps on a profiled handle (single compute stream)" % args.prof_steps,
            "e2e": e2e,
            "gpu_launches": int(launches),
            "clocks": clocks.summary(),
        }
        line.update(extras)
        if single_process is not None:
            line["e2e_single_process"] = single_process
        if world == 1 and not args.no_cpu_baseline:
            sample = host_frames[:64].numpy()
            threads = os.cpu_count() or 1
            r = cpu_reference_run(sample, args.cpu_seconds, threads)
            line["cpu_baseline"] = {"value": r["fps"], "unit": "frames/s", "cores": r["cores"], "kind": r["kind"],
                                    "sample": "first 64 frames of the workload cycled for %.0f s, one extractor + one frame per thread" % args.cpu_seconds,
                                    "p50_ms_per_frame": r["p50_ms"],
                                    "note": "the reference's own source; its OpenCV primitives are this repo's scalar C restatements (no OpenCV C++ in the image)"}
            msp = cv2_primitives_ms(sample[:16])
            if msp is not None:
                line["cpu_baseline"]["cv2_primitives_ms"] = msp
                line["cpu_baseline"]["cv2_primitives_fps_all_cores"] = threads * 1e3 / msp
                line["cpu_baseline"]["cv2_primitives_what"] = ("python-cv2 (SIMD) resize x7 + copyMakeBorder x8 + FAST(20, nms) per level + GaussianBlur x8, one "
                                                               "thread: a lower bound for any OpenCV-based CPU implementation, scaled to all cores")
        print(json.dumps(line))
    ext.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

#!/usr/bin/env python
"""tools/e2e_knobs.py -- the host-buffer pipeline of orbx_extract_batch under its knobs, one process per setting (the knobs are read
at orbx_create): staging slots (ORBX_SLOTS), group-size ramp (ORBX_RAMP), launch-group size, write-combined input memory
(orbx_host_alloc), CPU affinity.  4096 frames per step, 10 steps each; prints frames/s and the copy-only bound of the same setting.

    python tools/e2e_knobs.py            # runs every setting in a child process
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def child(group, wc, pin_cpu):
    import ctypes as C
    import numpy as np
    import torch
    import bench
    import extractorb_b200 as ex
    if pin_cpu >= 0:
        os.sched_setaffinity(0, set(range(pin_cpu, pin_cpu + 4)))   # before the pinned allocations: first touch on these cores' node
    F, W, H = 4096, bench.W, bench.H
    frames = bench.make_frames(F, seed=0)
    L = ex.load_library()
    nbytes = F * W * H
    if wc:
        ptr = L.orbx_host_alloc(nbytes, 1)
        assert ptr
        C.memmove(ptr, frames.data_ptr(), nbytes)
        in_ptr = ptr
    else:
        host = frames.pin_memory()
        in_ptr = host.data_ptr()
    res = {}
    for name, flags in (("e2e", 0), ("copy_only", ex.FLAG_COPY_ONLY)):
        ext = ex.ORBextractor(1000, 1.2, 8, 20, 7, max_batch=group, flags=flags)
        cap = ext.max_keypoints(W, H)
        hk = torch.empty((F, cap, 7), dtype=torch.float32).pin_memory()
        hd = torch.empty((F, cap, 32), dtype=torch.uint8).pin_memory()
        hc = torch.zeros((F, 2), dtype=torch.int32).pin_memory()

        def step():
            ext.extract_batch_raw(in_ptr, ex.MEM_HOST, F, W, H, W, W * H, (0, 0), hk.data_ptr(), hd.data_ptr(), cap, hc.data_ptr(), ex.MEM_HOST, None)
        step(); step()
        t0 = time.perf_counter()
        for _ in range(10):
            step()
        res[name] = F * 10 / (time.perf_counter() - t0)
        ext.close()
    print(json.dumps(res))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]))
        sys.exit(0)
    settings = [("default (4 slots, ramp, group 256)", {}, 256, 0, -1), ("3 slots", {"ORBX_SLOTS": "3"}, 256, 0, -1), ("2 slots", {"ORBX_SLOTS": "2"}, 256, 0, -1),
                ("no ramp", {"ORBX_RAMP": "0"}, 256, 0, -1), ("group 128", {}, 128, 0, -1), ("group 512", {}, 512, 0, -1),
                ("write-combined input", {}, 256, 1, -1), ("pinned to cores 4-7", {}, 256, 0, 4)]
    for name, env, group, wc, cpu in settings:
        e = dict(os.environ); e.update(env)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "child", str(group), str(wc), str(cpu)], env=e, capture_output=True, text=True,
                               timeout=150)
        except subprocess.TimeoutExpired:
            print("%-36s TIMED OUT" % name)
            continue
        line = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if line:
            d = json.loads(line[-1])
            print("%-36s e2e %8.0f frames/s   copies only %8.0f   ratio %.3f" % (name, d["e2e"], d["copy_only"], d["e2e"] / d["copy_only"]))
        else:
            print("%-36s FAILED: %s" % (name, r.stderr.strip()[-300:]))

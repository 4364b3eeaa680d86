import sys, time, ctypes as C, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import extractorb_b200 as ex
from common import synth_frame
import bench
f0 = bench.make_frames(4, seed=0).numpy()[0]
one = ex.ORBextractor(1000, 1.2, 8, 20, 7, max_batch=1)
cap = one.max_keypoints(640, 480)
k1 = np.zeros(cap, ex.KP_DTYPE); d1 = np.zeros((cap, 32), np.uint8); n1, m1 = C.c_int(0), C.c_int(0)
ts = []
for i in range(700):
    t0 = time.perf_counter()
    one._check(one._L.orbx_extract(one._h, f0.ctypes.data, 640, 480, 640, 0, 0, k1.ctypes.data, d1.ctypes.data, cap, C.byref(n1), C.byref(m1)))
    ts.append((time.perf_counter() - t0) * 1e6)
ts = np.array(ts[100:])
print("latency p50 %.1f us p99 %.1f us  n=%d tma=%s" % (np.percentile(ts, 50), np.percentile(ts, 99), n1.value, one.uses_tma()))

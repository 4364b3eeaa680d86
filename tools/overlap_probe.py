import sys, time, threading
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import torch, bench, extractorb_b200 as ex
F = 4096
frames = bench.make_frames(F, seed=0).cuda()
def mk():
    e = ex.ORBextractor(1000, 1.2, 8, 20, 7, max_batch=256)
    cap = e.max_keypoints(640, 480)
    return e, cap, torch.empty((F, cap, 7), dtype=torch.float32, device='cuda'), torch.empty((F, cap, 32), dtype=torch.uint8, device='cuda'), torch.zeros((F, 2), dtype=torch.int32, device='cuda')
def run(h, lo, n, reps):
    e, cap, k, d, c = h
    for _ in range(reps):
        e.extract_batch_raw(frames[lo:lo + n].data_ptr(), ex.MEM_DEVICE, n, 640, 480, 640, 640 * 480, (0, 0), k.data_ptr(), d.data_ptr(), cap, c.data_ptr(), ex.MEM_DEVICE, None)
a, b = mk(), mk()
run(a, 0, F, 2); run(b, 0, F, 2)
torch.cuda.synchronize(); t = time.perf_counter(); run(a, 0, F, 6); torch.cuda.synchronize()
print("one handle (2 streams): %.0f frames/s" % (F * 6 / (time.perf_counter() - t)))
torch.cuda.synchronize(); t = time.perf_counter()
ta = threading.Thread(target=run, args=(a, 0, F // 2, 6)); tb = threading.Thread(target=run, args=(b, F // 2, F // 2, 6))
ta.start(); tb.start(); ta.join(); tb.join(); torch.cuda.synchronize()
print("two handles concurrently (4 streams): %.0f frames/s" % (F * 6 / (time.perf_counter() - t)))

"""tools/e2e_probe.py -- which direction bounds the host-buffer pipeline: the same batch through orbx_extract_batch with device or
pinned host memory on either side.  1024 frames (frame generation is CPU work: keep it short when several copies run at once)."""
import sys, time, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
import torch, numpy as np
import bench, extractorb_b200 as ex
F=int(os.environ.get('PROBE_FRAMES', '1024')); W,H=bench.W,bench.H
host=bench.make_frames(F,0).pin_memory(); dev=host.cuda()
ext=ex.ORBextractor(1000,1.2,8,20,7,max_batch=256)
cap=ext.max_keypoints(W,H)
dk=torch.empty((F,cap,7),dtype=torch.float32,device='cuda'); dd=torch.empty((F,cap,32),dtype=torch.uint8,device='cuda'); dc=torch.zeros((F,2),dtype=torch.int32,device='cuda')
hk=torch.empty((F,cap,7),dtype=torch.float32).pin_memory(); hd=torch.empty((F,cap,32),dtype=torch.uint8).pin_memory(); hc=torch.zeros((F,2),dtype=torch.int32).pin_memory()
def run(inp,imem,k,d,c,omem,n=4):
    ext.extract_batch_raw(inp.data_ptr(),imem,F,W,H,W,W*H,(0,0),k.data_ptr(),d.data_ptr(),cap,c.data_ptr(),omem,None)
    torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(n): ext.extract_batch_raw(inp.data_ptr(),imem,F,W,H,W,W*H,(0,0),k.data_ptr(),d.data_ptr(),cap,c.data_ptr(),omem,None)
    torch.cuda.synchronize(); return F*n/(time.perf_counter()-t)
print('dev->dev  %.0f'%run(dev,1,dk,dd,dc,1))
print('host->dev %.0f'%run(host,0,dk,dd,dc,1))
print('dev->host %.0f'%run(dev,1,hk,hd,hc,0))
print('host->host %.0f'%run(host,0,hk,hd,hc,0))

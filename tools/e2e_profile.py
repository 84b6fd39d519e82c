import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, synth
import jxl_rs_b200 as j
from jxl_rs_b200 import abi
from concurrent.futures import ThreadPoolExecutor
n=64
with ThreadPoolExecutor(64) as ex:
    files=list(ex.map(lambda s: synth.encode_synthetic(3840,2160,2000+s,0.5,2,1,1), range(n)))
pool=ThreadPoolExecutor(64)
ctx=j.JxgContext(0)
outs=[torch.empty((2160,3840,3),dtype=torch.uint8).pin_memory() for _ in range(n)]
for it in range(4):
    t0=time.perf_counter()
    parsed=list(pool.map(j.ParsedFrame, files)); t1=time.perf_counter()
    b=j.Batch(ctx,n)
    for fr,o in zip(parsed,outs): b.add(fr,o.data_ptr(),3840*3,abi.FORMAT_RGB_U8,False)
    t2=time.perf_counter()
    b.run(); t3=time.perf_counter()
    b.wait(); t4=time.perf_counter()
    st=b.stats(); b.close(); t5=time.perf_counter()
    print('parse %.1f add %.1f run %.1f wait %.1f close %.1f | device_ms %.1f'%((t1-t0)*1e3,(t2-t1)*1e3,(t3-t2)*1e3,(t4-t3)*1e3,(t5-t4)*1e3, st['device_ms']))
# raw D2H bandwidth
d=torch.empty(1592524800,dtype=torch.uint8,device='cuda'); h=torch.empty(1592524800,dtype=torch.uint8).pin_memory()
torch.cuda.synchronize(); t=time.perf_counter(); h.copy_(d,non_blocking=True); torch.cuda.synchronize(); print('D2H GB/s', 1.5925/(time.perf_counter()-t))
t=time.perf_counter(); d.copy_(h,non_blocking=True); torch.cuda.synchronize(); print('H2D GB/s', 1.5925/(time.perf_counter()-t))

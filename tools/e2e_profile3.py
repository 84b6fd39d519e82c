import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, synth
import jxl_rs_b200 as j
from jxl_rs_b200 import abi
from concurrent.futures import ThreadPoolExecutor
from collections import deque
n=64
with ThreadPoolExecutor(64) as ex:
    files=list(ex.map(lambda s: synth.encode_synthetic(3840,2160,2000+s,0.5,2,1,1), range(n)))
outs=[[torch.empty((2160,3840,3),dtype=torch.uint8).pin_memory() for _ in range(n)] for _ in range(2)]
ptrs=[[(o.data_ptr(),3840*3) for o in oo] for oo in outs]
pool=ThreadPoolExecutor(64)
ctxs=[j.JxgContext(0) for _ in range(2)]
inflight=deque()
T0=time.perf_counter()
def now(): return (time.perf_counter()-T0)*1e3
for k in range(7):
    t=[now()]
    futs=[pool.submit(j.ParsedFrame,f) for f in files]
    if len(inflight)==2:
        b=inflight.popleft(); b.wait(); b.close()
    t.append(now())
    parsed=[f.result() for f in futs]; t.append(now())
    b=j.Batch(ctxs[k%2],n)
    for fr,(p,s) in zip(parsed,ptrs[k%2]): b.add(fr,p,s,abi.FORMAT_RGB_U8,False)
    t.append(now()); b.run(); t.append(now())
    inflight.append(b)
    print('k',k,'start %.0f retire %.0f parsewait %.0f add %.0f run %.0f'%(t[0],t[1]-t[0],t[2]-t[1],t[3]-t[2],t[4]-t[3]))
while inflight:
    b=inflight.popleft(); b.wait(); print('final wait done at %.0f dev_ms %.0f'%(now(), b.stats()['device_ms'])); b.close()

import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, synth
import jxl_rs_b200 as j
from jxl_rs_b200 import abi
from concurrent.futures import ThreadPoolExecutor
n=64
with ThreadPoolExecutor(64) as ex:
    files=list(ex.map(lambda s: synth.encode_synthetic(3840,2160,2000+s,0.5,2,1,1), range(n)))
outs=[[torch.empty((2160,3840,3),dtype=torch.uint8).pin_memory() for _ in range(n)] for _ in range(2)]
ptrs=[[(o.data_ptr(),3840*3) for o in oo] for oo in outs]
for depth in (2,3):
    dec=j.PipelinedDecoder(0,depth=depth)
    orig_retire=dec._retire
    log=[]
    def retire():
        t=time.perf_counter(); b=dec.inflight[0]; orig_retire(); log.append(('retire',(time.perf_counter()-t)*1e3, dec.last_stats['device_ms']))
    dec._retire=retire
    for i in range(3): dec.submit(files, ptrs[i%2])
    dec.drain(); log.clear()
    t0=time.perf_counter()
    for i in range(6):
        t=time.perf_counter(); dec.submit(files, ptrs[i%2]); log.append(('submit',(time.perf_counter()-t)*1e3,0))
    dec.drain(); torch.cuda.synchronize()
    tot=(time.perf_counter()-t0)/6*1e3
    print('depth',depth,'ms/step %.1f'%tot, ' '.join('%s:%.0f/%.0f'%l for l in log))
    dec.close()

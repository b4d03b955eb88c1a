import ctypes, sys, os, numpy as np, torch
sys.path.insert(0, '/root/repo')
import zkmember_b200 as zkm
from zkmember_b200 import _lib
from oracle import capi
zkm.init(0); L=_lib.lib()
st=torch.cuda.Stream(); torch.cuda.set_stream(st); sp=ctypes.c_void_p(st.cuda_stream)
for lg in (16,20,24):
    n=1<<lg
    x=torch.from_numpy(capi.random_field_elements(0,n,seed=lg).view(np.int64)).cuda(); y=torch.empty_like(x)
    for radix in (8,9,10,11,12):
        zkm.set_option("ntt_max_radix_log", radix)
        f=lambda: _lib.check(L.zkm_ntt_device(0, ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(y.data_ptr()), lg,0,0,sp))
        for _ in range(3): f()
        torch.cuda.synchronize()
        ts=[]
        for _ in range(7):
            e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
            e0.record(); f(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        print(lg, radix, "ms=%.4f"%sorted(ts)[3], flush=True)

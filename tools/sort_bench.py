import sys, torch, numpy as np
sys.path.insert(0,'.')
from ocrfdet_b200 import _lib
L=_lib.lib()
n=3_847_121; end_bit=45
rng=np.random.default_rng(0)
tile=rng.integers(0,4224,size=n,dtype=np.uint64)
depth=rng.uniform(0.2,60,size=n).astype(np.float32).view(np.uint32).astype(np.uint64)
keys=(tile<<np.uint64(32))|depth
k=torch.from_numpy(keys.view(np.int64)).cuda(); v=torch.arange(n,dtype=torch.int32,device='cuda')
ko,vo,kt,vt=torch.empty_like(k),torch.empty_like(v),torch.empty_like(k),torch.empty_like(v)
ws=torch.empty(int(L.ocrf_sort_workspace_bytes(n))+256,dtype=torch.uint8,device='cuda')
def run():
    _lib.check(L.ocrf_sort_pairs(_lib.current_stream(),n,end_bit,_lib.ptr(k),_lib.ptr(v),_lib.ptr(ko),_lib.ptr(vo),_lib.ptr(kt),_lib.ptr(vt),_lib.ptr(ws)),"sort")
for _ in range(3): run()
torch.cuda.synchronize()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): run()
e1.record(); torch.cuda.synchronize()
print("sort_pairs n=%d end_bit=%d: %.1f us per call (incl. 2 staging copies, histogram, %d passes)"%(n,end_bit,e0.elapsed_time(e1)/20*1e3,(end_bit+7)//8))
order=np.argsort(keys,kind='stable')
assert np.array_equal(vo.cpu().numpy().view(np.uint32),order.astype(np.uint32)); print("correct")

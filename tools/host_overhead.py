import sys, time, torch, numpy as np
sys.path.insert(0,'.')
from ocrfdet_b200 import rasterizer as R
from ocrfdet_b200.scenes import ring_scene
W,H,P,V=704,256,100000,6
g,cams=ring_scene(P=P,seed=1234,width=W,height=H,n_views=V)
names=("means3D","scales","rotations","opacities","colors")
dev={k:torch.from_numpy(g[k]).unsqueeze(0).cuda().requires_grad_(True) for k in names}
cam_t=R.pack_camera_dicts(cams,"cuda"); bg=torch.zeros(3,device="cuda")
gcol=torch.randn(V,3,H,W,device="cuda"); gop=torch.randn(V,1,H,W,device="cuda")
def step(cap=None):
    for k in names: dev[k].grad=None
    c,r,d,o=R.render_batch(dev["means3D"],dev["opacities"],cam_t,H,W,bg,colors_precomp=dev["colors"],scales=dev["scales"],rotations=dev["rotations"],pair_capacity=cap)
    torch.autograd.backward([c,o],[gcol,gop])
R.KEEP_STATE=True
step(); torch.cuda.synchronize()
N=R.last_state()["num_pairs"]; R.KEEP_STATE=False; R._LAST_STATE=None
print("N",N)
for cap in (None, int(N*1.3)):
    for _ in range(5): step(cap)
    torch.cuda.synchronize()
    t0=time.perf_counter()
    for _ in range(50): step(cap)
    t1=time.perf_counter(); torch.cuda.synchronize(); t2=time.perf_counter()
    print("cap",cap,"cpu issue ms/step %.3f  total ms/step %.3f"%((t1-t0)/50*1e3,(t2-t0)/50*1e3))
R.check_overflow()
import cProfile,pstats
cap=int(N*1.3)
pr=cProfile.Profile(); pr.enable()
for _ in range(50): step(cap)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(25)

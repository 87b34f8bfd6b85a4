import sys, time, torch, numpy as np
sys.path.insert(0,'.')
from ocrfdet_b200 import rasterizer as R
from ocrfdet_b200.scenes import ring_scene
def run(name,S,P,V,W,H,C,bev=128,reps=5):
    gs=[ring_scene(P=P,seed=1234+s,width=W,height=H,channels=C,n_views=V,bev=bev) for s in range(S)]
    cams=sum([g[1] for g in gs],[])
    names=("means3D","scales","rotations","opacities","colors")
    dev={k:torch.from_numpy(np.stack([g[0][k] for g in gs])).cuda().requires_grad_(True) for k in names}
    cam_t=R.pack_camera_dicts(cams,"cuda"); bg=torch.zeros(C,device="cuda")
    gcol=torch.randn(S*V,C,H,W,device="cuda"); gop=torch.randn(S*V,1,H,W,device="cuda")
    def step():
        for k in names: dev[k].grad=None
        c,r,d,o=R.render_batch(dev["means3D"],dev["opacities"],cam_t,H,W,bg,colors_precomp=dev["colors"],scales=dev["scales"],rotations=dev["rotations"])
        torch.autograd.backward([c,o],[gcol,gop]); return c
    R.KEEP_STATE=True; c=step(); torch.cuda.synchronize(); st=R.last_state(); N=st["num_pairs"]; R.KEEP_STATE=False; R._LAST_STATE=None
    marks=[]
    def rec(n):
        e=torch.cuda.Event(enable_timing=True); e.record(); marks.append((n,e))
    R.STAGE_HOOK=rec
    for _ in range(reps):
        rec("begin"); step()
    torch.cuda.synchronize(); R.STAGE_HOOK=None
    agg={}
    for (n0,e0),(n1,e1) in zip(marks[:-1],marks[1:]):
        if n1 in("begin","backward_begin"): continue
        agg[n1]=agg.get(n1,0)+e0.elapsed_time(e1)/reps
    tot=sum(agg.values())
    print(name,"views",S*V,"N_dup %.2fM"%(N/1e6),"finite",bool(torch.isfinite(c).all()),"ms/step %.3f -> %.0f views/s"%(tot,S*V/tot*1e3),{k:round(v,3) for k,v in agg.items()}, "mem GB %.1f"%(torch.cuda.max_memory_allocated()/1e9))
run("config2",1,100000,6,704,256,3)
run("config4 (C=80, 2 frames)",2,100000,6,704,256,80)
run("config5 (1 of 8 samples: 6 views 512x1408, 1M)",1,1000000,6,1408,512,3,bev=277)
run("config3-like (8 samples x 6 views in one call)",8,100000,6,704,256,3)

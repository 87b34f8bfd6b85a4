import numpy as np, torch, sys
sys.path.insert(0,'.')
from ocrfdet_b200 import rasterizer as R
from oracle import oracle, ref
from tests import util
for kind,kw in [("frustum", dict(P=500, seed=2, W=50, H=37)),("ring", dict(P=20000, seed=3, W=352, H=128)),("ring", dict(P=30000, seed=22, W=704, H=256))]:
    g,cams=util.small_scene(kind,**kw); cam=cams[0]; W,H=kw['W'],kw['H']
    bg=[0.3,0.1,0.6]
    rng=np.random.default_rng(7)
    gcol=rng.normal(size=(3,H,W)).astype(np.float32)
    gc=util.to_cuda(g)
    for k in gc: gc[k].requires_grad_(True)
    st=util.settings_for(cam,bg)
    m2=torch.zeros_like(gc['means3D'],requires_grad=True)
    R.KEEP_STATE=True
    color,radii,depth,opac=R.GaussianRasterizer(st,return_opacity=True)(means3D=gc['means3D'],means2D=m2,opacities=gc['opacities'],colors_precomp=gc['colors'],scales=gc['scales'],rotations=gc['rotations'])
    (color*torch.from_numpy(gcol).cuda()).sum().backward()
    want,wst=util.oracle_forward(g,cam,W,H,bg)
    ms=R.last_state()
    nc=ms['n_contrib'][0].cpu().numpy(); fT=ms['final_T'][0].cpu().numpy()
    print(kind,kw,'n_contrib mismatches',(nc!=want['n_contrib']).sum(),'amb',want['ambiguous'].sum(), 'finalT maxdiff',np.abs(fT-want['final_T']).max())
    gw=util.oracle_backward(g,cam,W,H,bg,want,wst,gcol)
    # oracle backward using GPU forward state
    want2=dict(want); want2['n_contrib']=nc.astype(np.uint32); want2['final_T']=fT
    gw2=util.oracle_backward(g,cam,W,H,bg,want2,wst,gcol)
    rr=ref.RefRasterizer()
    d=lambda t:t.detach()
    rcol,rrad,rN=rr.forward(d(gc['means3D']),d(gc['opacities']),d(gc['colors']),st.viewmatrix,st.projmatrix,st.campos,W,H,st.tanfovx,st.tanfovy,st.bg,scales=d(gc['scales']),rotations=d(gc['rotations']))
    gr=rr.backward(d(gc['means3D']),d(gc['colors']),st.viewmatrix,st.projmatrix,st.campos,st.tanfovx,st.tanfovy,st.bg,rrad,torch.from_numpy(gcol).cuda(),scales=d(gc['scales']),rotations=d(gc['rotations']))
    torch.cuda.synchronize()
    mine=dict(means3D=gc['means3D'].grad,scales=gc['scales'].grad,rotations=gc['rotations'].grad,opacities=gc['opacities'].grad.reshape(-1),colors=gc['colors'].grad,means2D=m2.grad[:,:2])
    refg=dict(means3D=gr['means3D'],scales=gr['scales'],rotations=gr['rotations'],opacities=gr['opacities'].reshape(-1),colors=gr['colors'],means2D=gr['means2D'][:,:2])
    for k in mine:
        a=mine[k].cpu().numpy(); b=refg[k].cpu().numpy(); o=np.asarray(gw[k],np.float64); o2=np.asarray(gw2[k],np.float64)
        print('  %-10s mine-vs-oracle %.2e  mine-vs-oracle(gpu fwd state) %.2e  ref-vs-oracle %.2e  mine-vs-ref %.2e   max|g| %.3g'%(k,util.rel_err(a,o),util.rel_err(a,o2),util.rel_err(b,o),util.rel_err(a,b),np.abs(o).max()))

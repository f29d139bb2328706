import sys, time, torch, numpy as np
sys.path.insert(0,'.')
import dhfk
from dhfk import synthetic, tables, _cabi
dev=torch.device('cuda:0')
blk=tables.camera_block('S1',0)
for n in (1024, 4608):
    d={k:torch.tensor(v,device=dev) for k,v in synthetic.gan_like(n,seed=1).items()}
    gw=torch.randn(n,16,3,device=dev); gu=torch.randn(n,16,2,device=dev)
    a,g,r=d['ang'].requires_grad_(True),d['grot'].requires_grad_(True),d['root'].requires_grad_(True)
    def step():
        w,_,uv=dhfk.fk_project(a,g,d['bone'],r,blk,return_cam=False)
        torch.autograd.backward((w,uv),(gw,gu))
    for _ in range(50): step()
    torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(500): step()
    torch.cuda.synchronize(); dt=(time.perf_counter()-t)/500
    # raw C ABI
    lib=_cabi.load(); st=torch.cuda.current_stream().cuda_stream
    w=torch.empty(n,16,3,device=dev); uv=torch.empty(n,16,2,device=dev); ga=torch.empty(n,33,device=dev); gg=torch.empty(n,3,device=dev); gr=torch.empty(n,3,device=dev)
    P=lambda t:t.data_ptr()
    def raw():
        lib.dhfk_forward(P(d['ang']),33,P(d['grot']),3,P(d['bone']),15,P(d['root']),3,blk.ctypes.data,P(w),None,P(uv),n,0,st)
        lib.dhfk_backward(P(d['ang']),33,P(d['grot']),3,P(d['bone']),15,P(d['root']),3,blk.ctypes.data,P(gw),None,P(gu),P(ga),33,P(gg),3,P(gr),3,None,15,n,0,st)
    for _ in range(50): raw()
    torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(2000): raw()
    torch.cuda.synchronize(); dr=(time.perf_counter()-t)/2000
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200): raw()
    e1.record(); torch.cuda.synchronize()
    print('n=%d  autograd fwd+bwd %.1f us/step   raw C-ABI fwd+bwd %.1f us/step (host-bound)   GPU time %.1f us/step'%(n,dt*1e6,dr*1e6,e0.elapsed_time(e1)/200*1e3))

import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch
from oracle import np_ref, torch_ref
from action_conditioned_gans_b200 import engine as E
cuda=torch.device('cuda:0')
def _params(spec, seed):
    rng = np.random.RandomState(seed)
    p = np_ref.init_params(spec, rng)
    for k in p:
        if not k.endswith("weights"):
            p[k] = (rng.randn(*p[k].shape) * 0.1).astype(np.float32)
    return p
B=3
p=_params(np_ref.g_direct_spec(),4)
rng=np.random.RandomState(1)
img = rng.uniform(-1, 1, (B, 64, 64, 3)).astype(np.float32)
act = rng.randn(B, 10).astype(np.float32)
r = rng.randn(B, 64, 64, 3).astype(np.float32)
dt=torch.float64
pt = {k: torch.tensor(v, dtype=dt, requires_grad=True) for k, v in p.items()}
keep={}
def layer(name,x,kind,act=torch.relu):
    w=pt[name+'/weights']
    z = torch_ref.conv2d(x,w,2) if kind=='conv' else torch_ref.conv2d_transpose(x,w,2)
    if name+'/biases' in pt: z=z+pt[name+'/biases']
    z.retain_grad(); keep[name+'.z']=z
    y=z
    if name+'/BatchNorm/beta' in pt: y=torch_ref.batch_norm(z,pt[name+'/BatchNorm/beta'])
    a=act(y) if act is not None else y
    a.retain_grad(); keep[name+'.a']=a
    return a
x=torch.tensor(img,dtype=dt)
out=layer('g/conv1',x,'conv'); out=layer('g/conv2',out,'conv'); out=layer('g/conv3',out,'conv'); out=layer('g/conv4',out,'conv')
out=torch.cat([out, torch_ref.tile_actions(torch.tensor(act,dtype=dt),4)],3)
out=layer('g/tconv1',out,'deconv'); out=layer('g/tconv2',out,'deconv'); out=layer('g/tconv3',out,'deconv')
frame=layer('g/tconv4',out,'deconv',act=torch.tanh)
(frame*torch.tensor(r,dtype=dt)).sum().backward()
store = E.ParamStore(E.g_direct_spec(), cuda, p)
run = E.GeneratorRun(store, B, cuda, False, 5)
t = lambda a: torch.from_numpy(a).to(cuda)
g_out,_=run.forward(t(img),t(act))
run.dg_out.copy_(t(r)); store.grad.zero_()
d=run.layer_bwd("g/tconv4", run.dg_out, 3)
d3=run.layer_bwd("g/tconv3", d, 64)
torch.cuda.synchronize()
snap=run.layers['g/tconv3'].dz.clone()
ref=keep['g/tconv3.z'].grad.numpy()
dd=np.abs(snap.double().cpu().numpy()-ref)
flat=np.argwhere(dd.reshape(-1)>3e-4).reshape(-1)
print('right after tconv3 bwd: nbad',len(flat), 'flat range', flat.min() if len(flat) else None, flat.max() if len(flat) else None)
print('got',snap.reshape(-1)[:8].cpu().numpy(),'ref',ref.reshape(-1)[:8])
print('ptrs z %x dz %x dx %x  t4.dx %x red %x'%(run.layers['g/tconv3'].z.data_ptr(),run.layers['g/tconv3'].dz.data_ptr(),run.layers['g/tconv3'].dx.data_ptr(),run.layers['g/tconv4'].dx.data_ptr(), run.layers['g/tconv3'].red.data_ptr()))
d=run.layer_bwd("g/tconv2", d3, 128)
torch.cuda.synchronize()
dd=np.abs(run.layers['g/tconv3'].dz.double().cpu().numpy()-ref)
print('after tconv2 bwd: nbad',(dd>3e-4).sum())
sys.exit(0)
torch.cuda.synchronize()
def cmp(name,got,ref):
    got=got.double().cpu().numpy(); ref=ref.numpy()
    d=np.abs(got-ref); print('%-22s maxdiff %.3g scale %.3g  nbad %d/%d'%(name,d.max(),np.abs(ref).max(),(d>1e-4*np.abs(ref).max()).sum(),d.size))
    return d
for n in ['g/tconv4','g/tconv3','g/tconv2']:
    st=run.layers[n]
    cmp(n+'.z', st.z, keep[n+'.z'].detach())
    cmp(n+'.dz', st.dz, keep[n+'.z'].grad)
    if n!='g/tconv4': cmp(n+'.a', st.a, keep[n+'.a'].detach())
    d=cmp(n+'.dx(=dA prev)', st.dx, keep[{'g/tconv4':'g/tconv3','g/tconv3':'g/tconv2','g/tconv2':'g/tconv1'}[n]+'.a'].grad)
    bad=np.argwhere(d>1e-4*d.max()) if d.max()>0 else []
    if len(bad): print('  bad idx sample', bad[:5], 'unique b',np.unique(bad[:,0]), 'h', np.unique(bad[:,1])[:10], 'w',np.unique(bad[:,2])[:10], 'c', np.unique(bad[:,3])[:10])

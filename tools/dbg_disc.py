import sys, copy, torch, torch.nn as nn, torch.nn.functional as F
sys.path.insert(0,'climate-super-resolution_b200'); sys.path.insert(0,'.')
from climsr_b200.models import discriminator as D
from oracle import synth
def rel(a,b): return float((a-b).norm()/(b.norm()+1e-12))
d=D.Discriminator(); d.load_state_dict(synth.make_discriminator_state_dict(seed=3), strict=True)
ref=copy.deepcopy(d).train(); d=d.cuda().train()
g=torch.Generator().manual_seed(7)
x=torch.rand((4,1,128,128),generator=g)*2-1; w=torch.randn((4,1),generator=g)
outs={}
def hook(name):
    def f(m,i,o):
        o.retain_grad(); outs[name]=o
    return f
for i,m in enumerate(ref.feature_extraction):
    if isinstance(m,(nn.Conv2d,nn.BatchNorm2d,nn.LeakyReLU)): m.register_forward_hook(hook(i))
xr=x.clone().requires_grad_(True)
f=ref.feature_extraction(xr); outr=ref.classification(f.view(4,-1)); (outr*w).sum().backward()
# ours: monkeypatch to capture g buffers
cap={}
orig_collect=D._collect
def my_collect(dpad, view, n, pad, act, gate_neg):
    gbuf=orig_collect(dpad, view, n, pad, act, gate_neg)
    cap.setdefault('collect',[]).append((gbuf, view, dpad))
    return gbuf
D._collect=my_collect
convs=[]
orig_conv=D._conv
def my_conv(p,w,b,slope):
    o=orig_conv(p,w,b,slope); convs.append((p,o,slope)); return o
D._conv=my_conv
xg=x.cuda().requires_grad_(True); out=d(xg); (out*w.cuda()).sum().backward()
def logical(buf, v):
    return buf[:, v.off:v.off+v.step*v.hl:v.step, v.off:v.off+v.step*v.wl:v.step, :].float().permute(0,3,1,2).cpu()
# collect calls order: g4 (conv idx 28 output grad), then per stage reversed: gb (conv_b output grad)
names=[28, 26, 19, 12, 5]
for (gbuf, v, dpad), idx in zip(cap['collect'], names):
    refg = outs[idx].grad   # grad wrt conv output (pre-activation)
    print('g of conv', idx, 'rel', rel(logical(gbuf, v), refg), 'shape', tuple(refg.shape))
# dp5: gradient wrt input of conv 30 = output of LeakyReLU idx 29
gb, v, dpad = cap['collect'][0]
print('dp5 vs grad of lrelu29 out', rel(dpad.float().permute(0,3,1,2).cpu(), outs[29].grad))

# forward S4 vs ref post-activation
p4,s4,sl = convs[8]
v=cap['collect'][0][1]
print('slope', sl, 'S4 vs ref lrelu out', rel(logical(s4, v), outs[29].detach()))
sgn_ours=(logical(s4,v)>0); sgn_ref=(outs[28].detach()>0)
print('sign mismatch frac', float((sgn_ours!=sgn_ref).float().mean()), 'neg frac', float((~sgn_ref).float().mean()))
gb,_,dpad=cap['collect'][0]
man = dpad.float().permute(0,3,1,2).cpu() * torch.where(sgn_ref, torch.tensor(1.0), torch.tensor(0.2))
print('manual gate of our dp5 vs ref', rel(man, outs[28].grad))
print('ours vs manual', rel(logical(gb,v), man))
lg=logical(gb,v); r=outs[28].grad
ratio=(lg/ (r+1e-30))
print('ratio where ref neg', float(ratio[~sgn_ref].median()), 'pos', float(ratio[sgn_ref].median()))

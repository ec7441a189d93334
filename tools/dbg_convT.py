import sys, torch, torch.nn.functional as F
sys.path.insert(0,'climate-super-resolution_b200'); sys.path.insert(0,'.')
from climsr_b200 import ops
def bf(t): return t.to(torch.bfloat16).float()
torch.manual_seed(0)
for (n,h,w,cf_out,cf_in) in ((2,6,6,512,512),(2,6,6,128,64),(2,10,10,256,128),(2,6,6,64,64),(1,8,8,512,256)):
    g = torch.randn(n,cf_out,h,w)*0.1
    wt = torch.randn(cf_out,cf_in,3,3)*0.05
    ref = F.conv_transpose2d(bf(g).double(), bf(wt).double(), padding=1).float()
    gn = bf(g).permute(0,2,3,1).contiguous().to(torch.bfloat16).cuda()
    out = ops.conv2d_nhwc(gn, wt.cuda(), None, transposed=True)
    got = out[..., :cf_in].float().permute(0,3,1,2).cpu()
    print('convT', (n,h,w,cf_out,cf_in), 'rel', float((got-ref).norm()/ref.norm()), 'max', float((got-ref).abs().max()), float(ref.abs().max()))
    # forward conv too
    x = torch.randn(n,cf_in,h,w)*0.5
    reff = F.conv2d(bf(x).double(), bf(wt).double(), None, padding=1).float()
    xn = bf(x).permute(0,2,3,1).contiguous().to(torch.bfloat16).cuda()
    o2 = ops.conv2d_nhwc(xn, wt.cuda(), None)
    got2 = o2[..., :cf_out].float().permute(0,3,1,2).cpu()
    print('conv ', 'rel', float((got2-reff).norm()/reff.norm()))
    # wgrad
    dw, db = ops.conv2d_wgrad(xn, gn[..., :128].contiguous() if cf_out>128 else gn, (min(cf_out,128), cf_in, 3, 3))
    xr = bf(x).double().requires_grad_(True); wr = bf(wt).double().requires_grad_(True)
    y = F.conv2d(xr, wr, None, padding=1); (y*bf(g).double()).sum().backward()
    print('wgrad', 'rel', float((dw.cpu()-wr.grad[:128].float()).norm()/wr.grad[:128].float().norm()))

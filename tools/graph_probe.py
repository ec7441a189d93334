import sys, torch, time
sys.path.insert(0, "/root/repo/climate-super-resolution_b200")
from climsr_b200._lib import lib
from climsr_b200.models import ESRGANGenerator
torch.manual_seed(0)
net = ESRGANGenerator(4, 1, 64, 11, 16).cuda().eval()
n=8
x = torch.rand(n, 4, 32, 32, device="cuda") * 2 - 1
e = torch.rand(n, 1, 128, 128, device="cuda")
m = (torch.rand(n, 1, 128, 128, device="cuda") > 0.3).float()
for opt in (1, 0, 1):
    lib.csr_set_option(11, opt)
    net._plans.clear()
    with torch.no_grad():
        for _ in range(4): o = net(x, e, m)
        torch.cuda.synchronize()
        plan = list(net._plans.values())[0][0]
        st = lib.csr_plan_graph_status(plan, 0)
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50): o = net(x, e, m)
        e1.record(); torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / 50 * 1e3
    print(lib.csr_last_error())
    print(f"graphs={opt} status={st} gpu {e0.elapsed_time(e1)/50:.3f} ms wall {wall:.3f} ms sum {float(o.sum()):.4f}")

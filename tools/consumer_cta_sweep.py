import sys, torch
sys.path.insert(0, "/root/repo")
import mdqe_cvpr2023_b200 as pkg
from mdqe_cvpr2023_b200 import _lib
flush = torch.ones(160 * 1024 * 1024, device="cuda")
def timed(fn, iters=10):
    fn(); fn(); torch.cuda.synchronize(); ts=[]
    for _ in range(iters):
        flush.sum()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b)*1e3)
    ts.sort(); return ts[len(ts)//2]
g = torch.Generator().manual_seed(0)
Q,K,G,T,H,W = 196,32,10,4,96,160
coeff = torch.tanh(torch.randn(Q,K,generator=g)).cuda(); proto = torch.randn(K,T,H,W,generator=g).cuda(); tgt=(torch.rand(G,T,H,W,generator=g)>0.8).float().cuda()
mask_pred = (torch.randn(50,T,H,W,generator=g)*2-0.5).cuda()
sv, iv = torch.rand(20,T,H,W,generator=g).cuda(), torch.rand(12,T,H,W,generator=g).cuda()
for n in (0, 37, 74, 148, 296, 444, 592):
    _lib.set_option("consumer_ctas", n)
    print(n, "match_cost %.1f  nms_siou %.1f  track_siou %.1f" % (timed(lambda: pkg.mask_match_cost(coeff, proto, tgt)), timed(lambda: pkg.mask_nms_siou(mask_pred)), timed(lambda: pkg.mask_track_siou(sv, iv))))

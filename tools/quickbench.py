import torch, time, sys, json
sys.path.insert(0, '.')
from skeleton_action_recognition_b200 import VirtualRadar, _cabi
layer = VirtualRadar(wavelength=5e-4, device='cuda:0').to('cuda:0')
res = {}
for N in (256, 1024, 4096, 16384):
    nb = max(2, int(300e6 // (N*180000)) + 1)
    xs = [torch.randn(N,3,300,25,2, device='cuda')*0.3 for _ in range(nb)]
    for i in range(3): layer(xs[i % nb])
    torch.cuda.synchronize()
    K = 20
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(K+1)]
    ev[0].record()
    for i in range(K):
        layer(xs[i % nb]); ev[i+1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i+1]) for i in range(K))
    med = ts[K//2]
    res[N] = dict(ms=med, sps=N/med*1e3, frac=N/med*1e3*199456/6550.1e9)
    print(N, res[N], flush=True)
    del xs
print(json.dumps(res))

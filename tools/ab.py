"""A/B timing aid (not a test): python tools/ab.py lib1.so lib2.so ...  -- times vr_forward_f32 of each library
(plain ctypes, only the symbols every version has) on the same box, interleaved, N=256 and N=16384."""
import ctypes, sys
import torch
E_SRC = [0, 1, 20, 2, 20, 4, 5, 6, 7, 7, 20, 8, 9, 10, 11, 11, 0, 0, 12, 13, 14, 16, 17, 18]
E_DST = [1, 20, 2, 3, 4, 5, 6, 7, 21, 22, 8, 9, 10, 11, 23, 24, 16, 12, 13, 14, 15, 17, 18, 19]
libs = []
for path in sys.argv[1:]:
    L = ctypes.CDLL(path)
    vp, i64, i32, u32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_uint32
    L.vr_forward_f32.argtypes = [vp, i64, i64, i32, i32, ctypes.POINTER(i32), ctypes.POINTER(i32), i32, vp, vp, i32, i32, u32, vp, vp]
    libs.append((path.split('/')[-1], L))
src = (ctypes.c_int32 * 24)(*E_SRC); dst = (ctypes.c_int32 * 24)(*E_DST)
lam = torch.tensor(5e-4, device='cuda'); loc = torch.zeros(3, device='cuda')
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for N, K in ((256, 1000), (16384, 30)):
    nb = max(2, int(320e6 // (N * 199456)) + 1)
    xs = [torch.randn(N, 3, 300, 25, 2, device='cuda') * 0.3 for _ in range(nb)]
    outs = [torch.empty(N, 256, 19, device='cuda') for _ in range(nb)]
    for rep in range(3):
        for name, L in libs:
            def step(i):
                rc = L.vr_forward_f32(xs[i % nb].data_ptr(), N, 300, 25, 2, src, dst, 24, lam.data_ptr(), loc.data_ptr(), 256, 16, 0, outs[i % nb].data_ptr(), st)
                assert rc == 0
            for i in range(5): step(i)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(K): step(i)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / K
            print("N=%d rep%d %-28s %.2f us  %.2f M/s" % (N, rep, name, ms * 1e3, N / ms / 1e3), flush=True)
    del xs, outs

"""A/B timing aid for the tcgen05 STFT (not a test, not shipped): every tools/_ab/lib_*.so (built by tools/ab.py build)
and the product library, interleaved on one box: forward (inference), forward (saving Re/Im) and backward at
N = 4096 and N = 256 sequences of 300 samples, n_fft 256, hop 16.   python tools/ab_stft.py"""
import ctypes, glob, os, sys
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
paths = sorted(glob.glob(os.path.join(ROOT, "tools", "_ab", "lib_*.so"))) + [os.path.join(ROOT, "skeleton_action_recognition_b200", "lib", "libvirtual_radar_b200.so")]
vp, i64, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32
libs = []
for path in paths:
    L = ctypes.CDLL(path)
    L.vr_stft_general_workspace_floats.restype = i64
    L.vr_stft_general_workspace_floats.argtypes = [i64, i64, i32, i32, ctypes.POINTER(i64)]
    L.vr_stft_general_f32.argtypes = [vp, i64, i64, i32, i32, vp, vp, vp, vp, vp, vp, vp]
    L.vr_stft_general_backward_f32.argtypes = [vp, vp, vp, vp, i64, i64, i32, i32, vp, vp, vp, vp, vp, vp, vp]
    libs.append((os.path.basename(path)[4:-3] if "_ab" in path else "product", L))
n_fft, hop, T = 256, 16, 300
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
g = torch.Generator(device="cuda").manual_seed(0)
wsin = torch.randn(n_fft, 1, n_fft, device="cuda", generator=g); wcos = torch.randn(n_fft, 1, n_fft, device="cuda", generator=g)
for N, reps in ((4096, 20), (256, 100)):
    F = T // hop + 1
    iq = torch.randn(N, T, 2, device="cuda", generator=g)
    gout = torch.randn(N, n_fft, F, device="cuda", generator=g)
    best = {}
    for rep in range(3):
        for name, L in libs:
            parts = (i64 * 3)()
            L.vr_stft_general_workspace_floats(N, T, n_fft, hop, parts)
            fr = torch.empty(int(parts[0]), device="cuda"); bt = torch.empty(int(parts[1]), device="cuda"); cs = torch.empty(int(parts[2]), device="cuda")
            dc = torch.empty_like(cs); da = torch.empty(N * F * 2 * n_fft, device="cuda"); dbt = torch.empty_like(bt)
            out = torch.empty(N, n_fft, F, device="cuda"); giq = torch.empty(N, T, 2, device="cuda"); gs = torch.empty_like(wsin); gc = torch.empty_like(wcos)
            def fwd(save):
                rc = L.vr_stft_general_f32(iq.data_ptr(), N, T, n_fft, hop, wsin.data_ptr(), wcos.data_ptr(), fr.data_ptr(), bt.data_ptr(), cs.data_ptr() if save else None, out.data_ptr(), st)
                assert rc == 0, rc
            def bwd():
                rc = L.vr_stft_general_backward_f32(gout.data_ptr(), fr.data_ptr(), bt.data_ptr(), cs.data_ptr(), N, T, n_fft, hop, dc.data_ptr(), da.data_ptr(), dbt.data_ptr(), giq.data_ptr(), gs.data_ptr(), gc.data_ptr(), st)
                assert rc == 0, rc
            fwd(True)
            for key, fn in (("fwd", lambda: fwd(False)), ("fwd+save", lambda: fwd(True)), ("bwd", bwd)):
                for _ in range(3): fn()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps): fn()
                e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / reps
                best[(name, key)] = min(best.get((name, key), 1e9), ms)
            if rep == 0: print("   check %-14s out %.6e giq %.6e gsin %.6e" % (name, float(out.double().sum()), float(giq.double().sum()), float(gs.double().sum())), flush=True)
    for name, _ in libs:
        print("N=%5d %-14s fwd %.3f ms | fwd+save %.3f ms | bwd %.3f ms" % (N, name, best[(name, "fwd")], best[(name, "fwd+save")], best[(name, "bwd")]), flush=True)

"""A/B timing aid (not a test, not shipped): builds variants of the library with compile-time macros and times them on
the same box, interleaved.

    python tools/ab.py build  name1:-DFOO=1 name2:-DBAR ...     (here, no GPU: nvcc -> tools/_ab/lib_<name>.so)
    python tools/ab.py run [--flag]                             (on the GPU box: every tools/_ab/*.so + the product library;
                                                                 libraries with vr_set_schedule are timed under both schedules)

`run` times vr_forward_f32 (plain ctypes, only symbols every variant has) at N = 256 / 1024 / 4096 / 16384, stream
ordered (flags 0) or as independent batches (--flag: VR_FLAG_INPUTS_READY), best of 3 interleaved repetitions."""
import ctypes, glob, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
AB = os.path.join(ROOT, "tools", "_ab")
E_SRC = [0, 1, 20, 2, 20, 4, 5, 6, 7, 7, 20, 8, 9, 10, 11, 11, 0, 0, 12, 13, 14, 16, 17, 18]
E_DST = [1, 20, 2, 3, 4, 5, 6, 7, 21, 22, 8, 9, 10, 11, 23, 24, 16, 12, 13, 14, 15, 17, 18, 19]


def build(specs):
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    os.makedirs(AB, exist_ok=True)
    procs = []
    for spec in specs:
        name, _, flags = spec.partition(":")
        out = os.path.join(AB, "lib_%s.so" % name)
        cmd = [ge.NVCC] + ge.NVCC_FLAGS + [f for f in flags.split(",") if f] + ["-o", out, ge.SRC]
        procs.append((name, subprocess.Popen(cmd)))
    for name, pr in procs:
        assert pr.wait() == 0, name


def run(argv):
    import torch
    flag = 2 if "--flag" in argv else 0
    paths = sorted(glob.glob(os.path.join(AB, "lib_*.so"))) + [os.path.join(ROOT, "skeleton_action_recognition_b200", "lib", "libvirtual_radar_b200.so")]
    libs = []
    vp, i64, i32, u32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_uint32
    for path in paths:
        L = ctypes.CDLL(path)
        L.vr_forward_f32.argtypes = [vp, i64, i64, i32, i32, ctypes.POINTER(i32), ctypes.POINTER(i32), i32, vp, vp, i32, i32, u32, vp, vp]
        name = os.path.basename(path)[4:-3] if "_ab" in path else "product"
        if hasattr(L, "vr_set_schedule"):
            libs += [(name + "/coop", L, 0), (name + "/team", L, 1)]
        else:
            libs.append((name, L, None))
    src = (ctypes.c_int32 * 24)(*E_SRC); dst = (ctypes.c_int32 * 24)(*E_DST)
    lam = torch.tensor(5e-4, device="cuda"); loc = torch.zeros(3, device="cuda")
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    base = (torch.randn(256, 3, 300, 25, 2) * 0.3).cuda()
    ref = {}
    for N, K in ((256, 1500), (1024, 400), (4096, 100), (16384, 30)):
        nb = max(2, int(320e6 // (N * 199456)) + 1)
        xs = [base.repeat(N // 256, 1, 1, 1, 1).roll(i, 0).contiguous() for i in range(nb)]
        outs = [torch.empty(N, 256, 19, device="cuda") for _ in range(nb)]
        best = {}
        for rep in range(3):
            for name, L, sched in libs:
                if sched is not None: L.vr_set_schedule(sched)
                def step(i):
                    rc = L.vr_forward_f32(xs[i % nb].data_ptr(), N, 300, 25, 2, src, dst, 24, lam.data_ptr(), loc.data_ptr(), 256, 16, flag, outs[i % nb].data_ptr(), st)
                    assert rc == 0, rc
                for i in range(5): step(i)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(K): step(i)
                e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / K
                best[name] = min(best.get(name, 1e9), ms)
                if rep == 0:     # all variants must produce the same bits
                    got = outs[(K - 1) % nb].clone()
                    key = (N, (K - 1) % nb)
                    if key not in ref: ref[key] = got
                    elif not torch.equal(ref[key], got): print("!! %s differs from %s at N=%d" % (name, libs[0][0], N), flush=True)
        for name, _, _ in libs:
            ms = best[name]
            print("N=%6d %-22s %9.2f us  %6.2f M/s  %.3f of 6541 GB/s" % (N, name, ms * 1e3, N / ms / 1e3, N / ms * 1e3 * 199456 / 6541.1e9), flush=True)
        del xs, outs


if __name__ == "__main__":
    if sys.argv[1] == "build": build(sys.argv[2:])
    else: run(sys.argv[2:])

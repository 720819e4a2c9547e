"""Run on the GPU box (not a test): the reference's algorithm (oracle port, stock PyTorch ops) on CUDA against the same
port on the CPU, and our kernels against both -- SURVEY 7 asked for this number: it bounds what any tolerance against
"the reference" can mean, because stock PyTorch on the GPU rounds the radar range differently from stock PyTorch on the
CPU.  python tools/ref_cuda_spread.py > gpurun_out/ref_cuda_spread.md"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import virtual_radar_oracle as vro
from tests import fixtures as fx
from skeleton_action_recognition_b200 import VirtualRadar

def on_cuda(x, kw):
    o = vro.OracleVirtualRadar(**kw)
    o.stft = o.stft.to("cuda")
    o.wavelength, o.radar_location = o.wavelength.cuda(), o.radar_location.cuda()
    return o(x.cuda(), "aten").cpu().numpy()

def row(name, what, rep):
    t1, t2 = rep["t1"], rep["t2"]
    print("| %s | %s | %.1e | %.1e | %.1e | %.4f | %.2e | %.4f | %.1e |" % (name, what, t1["rel_median"], t1["rel_p99"], t1["rel_max"],
          t1["frac_rel_1e-4"], t1["db_max"], t2["frac_db_0.01"], rep["global_abs_over_peak"]), flush=True)

print("# Stock-PyTorch reference on CUDA vs on CPU, and the B200 kernels vs both (%s, torch %s)\n" % (torch.cuda.get_device_name(0), torch.__version__))
print("| input | comparison | t1 rel median | t1 rel p99 | t1 rel max | t1 frac <= 1e-4 | t1 max dB | t2 frac <= 0.01 dB | max abs / peak |")
print("|---|---|---|---|---|---|---|---|---|")
cases = [(n,) + fx.full_case(n)[:2] for n in sorted(fx.FULL)]
cases.append(("randn N=8", fx.s1_iid(8), dict(wavelength=5e-4)))
cases.append(("ntu raw", torch.from_numpy(fx.load("ntu_raw.npz")["x"]), dict(wavelength=5e-4)))
for name, x, kw in cases:
    cpu = vro.forward(x, **kw).numpy()
    gpu = on_cuda(x, kw)
    ours = VirtualRadar(device="cuda:0", **kw).to("cuda:0")(x.cuda()).cpu().numpy()
    row(name, "reference on CUDA vs reference on CPU", vro.parity_report(gpu, cpu))
    row(name, "B200 kernels vs reference on CPU", vro.parity_report(ours, cpu))
    row(name, "B200 kernels vs reference on CUDA", vro.parity_report(ours, gpu))

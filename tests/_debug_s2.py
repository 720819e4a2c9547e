import sys; sys.path.insert(0,'.')
import numpy as np, torch
from tests import fixtures as fx
from oracle import virtual_radar_oracle as vro
from skeleton_action_recognition_b200 import VirtualRadar
x = fx.s2_ntu_like(64)
layer = VirtualRadar(wavelength=5e-4, device='cuda:0').to('cuda:0')
out, iq = layer.forward_debug(x.cuda())
iq = iq.cpu().numpy()
o = vro.OracleVirtualRadar(wavelength=5e-4)
iq_ref = o.iq(x, 'seq').numpy()
err = np.abs(iq - iq_ref).max(-1)
bad = np.argwhere(err > 1e-3)
print('n bad', len(bad), 'of', err.size)
nz = (x.abs().sum(dim=(1,3)) > 0)   # (N,T,M)
for n, t in bad[:40]:
    print(n, t, 'err %.3g' % err[n, t], 'gpu', iq[n, t], 'ref', iq_ref[n, t], 'body present', nz[n, t].tolist(), 'len', int(nz[n,:,0].sum()))
print('bad per sample', np.bincount(bad[:,0], minlength=64) if len(bad) else None)

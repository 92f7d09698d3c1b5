import sys, numpy as np, torch
sys.path.insert(0, '.')
from oracle import ddsp_oracle as O
from ddsp_b200 import core
from ddsp_b200.synthetic import make_inputs
F = 25840
d = make_inputs(2, F, 1539, seed=78, zero_f0_fraction=0.02)
b = 1
f0f = d['f0_frames'][b:b+1]
ctrl = torch.from_numpy(d['ctrl'][b:b+1]).cuda()
hm, hp, nm = torch.split(ctrl, [513]*3, -1)
f0 = torch.from_numpy(f0f).cuda()[..., None]
pf, prefix, full = core.phase_stage(f0, 512, 44100, full_rate=True)
sig = core.combsubfast_stage(hm, hp, nm, f0, prefix, 512, 44100, noise_u=torch.from_numpy(d['U'][b:b+1]).cuda()).cpu().numpy()
ref, pf_ref = O.combsubfast_forward(d['ctrl'][b:b+1, :, :513], d['ctrl'][b:b+1, :, 513:1026], d['ctrl'][b:b+1, :, 1026:], f0f, d['U'][b:b+1])
err = np.abs(sig - ref)[0]
idx = np.argsort(err)[-10:]
print('worst samples', idx, err[idx], 'frames', idx // 512)
f0u, rot, _ = O.stage_a(f0f, 44100, 512)
ph = (np.float32(2*np.pi) * rot).astype(np.float32)
dphi = np.abs(full.cpu().numpy().astype(np.float64) - ph)
dphi = np.minimum(dphi, np.abs(dphi - 2*np.pi))
print('phase err max', dphi.max(), 'at', dphi.argmax(), 'frame', dphi.argmax() // 512)
w = idx[-1] // 512
print('f0 frames around worst', f0f[0, w-3:w+4])
print('err per frame (top)', np.sort(err.reshape(F, 512).max(1))[-10:], np.argsort(err.reshape(F, 512).max(1))[-10:])
# comb excitation comparison at worst frame
comb_ref = O.combtooth(f0u, rot, 44100, True)
print('x range at worst frame', (44100 * rot[0, w*512:(w+1)*512] / (f0u[0, w*512:(w+1)*512] + 1e-3)).min(), (44100 * rot[0, w*512:(w+1)*512] / (f0u[0, w*512:(w+1)*512] + 1e-3)).max())

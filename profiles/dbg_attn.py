import sys, os
sys.path.insert(0, os.getcwd())
import torch
from ddsp_b200 import core, _cabi
from ddsp_b200.control import _orthogonal_gaussian_features
torch.manual_seed(0)
B, F, H = 3, 301, 8
x = torch.randn(B, F, 256, device='cuda')
w = torch.randn(1536, 256, device='cuda') / 16
bias = torch.randn(1536, device='cuda') * 0.1
proj = _orthogonal_gaussian_features(266, 64).cuda()
ps = (64 ** -0.25 * proj).contiguous()
ws = core.favor_workspace(B, H, F, x.device)
Fp, Z = ws['Fp'], B * H
L = _cabi.lib()
st = torch.cuda.current_stream().cuda_stream
x2 = x.reshape(B * F, 256)
def run(name, fn):
    rc = fn()
    try:
        torch.cuda.synchronize()
        print(name, 'rc', rc, 'ok', flush=True)
    except Exception as e:
        print(name, 'rc', rc, 'FAILED', str(e)[:100], flush=True)
        sys.exit(1)
run('qkv', lambda: L.ddsp_b200_qkv_heads(x2.data_ptr(), 256, w.data_ptr(), 0, 256, bias.data_ptr(), ws['q'].data_ptr(), ws['k'].data_ptr(), ws['vt'].data_ptr(), 0, B, F, Fp, H, 256, st))
qkv = torch.nn.functional.linear(x.double(), w.double(), bias.double()).view(B, F, 3, H, 64)
qr = qkv[:, :, 0].permute(0, 2, 1, 3); kr = qkv[:, :, 1].permute(0, 2, 1, 3); vr = qkv[:, :, 2].permute(0, 2, 3, 1)
print('q err', (ws['q'].double() - qr).abs().max().item(), 'k err', (ws['k'].double() - kr).abs().max().item(),
      'vt err', (ws['vt'][:, :, :64, :F].double() - vr).abs().max().item(), 'ones', ws['vt'][:, :, 64, :F].min().item(), ws['vt'][:, :, 65:].abs().max().item())
run('featq', lambda: L.ddsp_b200_favor_features(ws['q'].data_ptr(), ps.data_ptr(), 266, 1, 1e-4, ws['qf'].data_ptr(), Z, F, Fp, st))
def feat(t, is_q):
    P = proj.double(); scale = 64 ** -0.25
    dash = torch.einsum('bhnd,jd->bhnj', scale * t, P)
    diag = (t * t).sum(-1, keepdim=True) * (0.5 * scale * scale)
    if is_q:
        return 266 ** -0.5 * (torch.exp(dash - diag - dash.amax(dim=-1, keepdim=True)) + 1e-4)
    return 266 ** -0.5 * torch.exp(dash - diag + 1e-4)
qf = feat(qr, True)
got = ws['qf'].view(B, H, F, 272)
print('qf err', (got[..., :266].double() - qf).abs().max().item(), 'scale', qf.abs().max().item(), 'pad', got[..., 266:].abs().max().item())
run('featk', lambda: L.ddsp_b200_favor_features(ws['k'].data_ptr(), ps.data_ptr(), 266, 0, 1e-4, ws['kt'].data_ptr(), Z, F, Fp, st))
kf = feat(kr, False)
gk = ws['kt'].view(B, H, 272, Fp)
print('kt rel err', ((gk[:, :, :266, :F].double() - kf.transpose(2, 3)).abs().max() / kf.abs().max()).item(), 'pad', gk[:, :, 266:].abs().max().item(), gk[..., F:].abs().max().item() if Fp > F else 0)
run('ctx', lambda: L.ddsp_b200_favor_context(ws['vt'].data_ptr(), 0, ws['kt'].data_ptr(), ws['ctx'].data_ptr(), ws['ctx_lo'].data_ptr(), Z, Fp, st))
ctx_ref = torch.einsum('bhen,bhjn->bhej', ws['vt'].double(), gk.double())
gc = ws['ctx'].view(B, H, 80, 272)
print('ctx rel err', ((gc.double() - ctx_ref).abs().max() / ctx_ref.abs().max()).item())
out = torch.empty(B, F, H * 64, device='cuda')
run('out', lambda: L.ddsp_b200_favor_output(ws['qf'].data_ptr(), ws['ctx'].data_ptr(), ws['ctx_lo'].data_ptr(), out.data_ptr(), B, H, F, st))
num = torch.einsum('bhnj,bhej->bhne', got.double(), gc.double())
ref = (num[..., :64] / (num[..., 64:65] + 1e-8)).transpose(1, 2).reshape(B, F, H * 64)
print('out err', (out.double() - ref).abs().max().item(), 'scale', ref.abs().max().item())

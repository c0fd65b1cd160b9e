"""GPU calibration of the tridiagonalisation cost model: times the stages of kfacb_eigh_batched for one matrix of
size n on a forced team size (KFACB_TRD_TEAM is read once per process, so each point is its own process).
    python tools/trd_calib.py n team [count]      -> one line: n team count trd_ms dc_ms bt_ms"""
import os
import sys

n, team = int(sys.argv[1]), int(sys.argv[2])
count = int(sys.argv[3]) if len(sys.argv) > 3 else 1
if team > 0:
    os.environ['KFACB_TRD_TEAM'] = str(team)
import torch  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kfac_pytorch_b200.native import NativeEngine  # noqa: E402

dev = torch.device('cuda:0')
eng = NativeEngine(dev)
gen = torch.Generator(device=dev).manual_seed(0)
mats = []
for _ in range(count):
    x = torch.randn(max(1568, n), n, device=dev, generator=gen)
    mats.append((0.05 * torch.eye(n, device=dev) + 0.95 * (x.t() @ x) / x.shape[0]).contiguous())
Qs = [torch.empty_like(m) for m in mats]
ds = [torch.empty(n, device=dev) for _ in mats]
best = None
for it in range(3):
    eng.eigh(mats, Qs, ds, 1e-10, None)
    torch.cuda.synchronize()
    p = eng.eigh_last_profile()
    if best is None or p['ms_trd'] < best['ms_trd']:
        best = p
print('calib n=%d team=%d count=%d trd_ms=%.3f dc_ms=%.3f bt_ms=%.3f grid=%d' % (n, team, count, best['ms_trd'], best['ms_dc'], best['ms_bt'], int(best['grid'])))

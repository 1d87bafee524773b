import sys, os
sys.path.insert(0, "/root/repo")
from simplestereo_b200 import _cabi
if len(sys.argv) > 1 and sys.argv[1] != "default":
    _cabi.LIB_PATH = sys.argv[1]
import torch, numpy as np
from simplestereo_b200.synth import synth_pair
L = _cabi.lib(); _cabi.check(L.ss_init(0))
W, H = 1242, 375
l, r, _ = synth_pair(W, H, 127, 0)
dl, dr = torch.from_numpy(l).cuda(), torch.from_numpy(r).cuda()
out = torch.empty((H, W), dtype=torch.int16, device="cuda")
st = torch.cuda.current_stream().cuda_stream
def step():
    _cabi.check(L.ss_asw_compute_device(dl.data_ptr(), dr.data_ptr(), W, H, 35, 127, 0, 5.0, 17.5, 0, 0, H, out.data_ptr(), st))
for _ in range(3): step()
torch.cuda.synchronize()
L.ss_profile_reset(); L.ss_profile_enable(1)
for _ in range(10): step()
torch.cuda.synchronize()
ms, n, tot = _cabi.profile_read()
print(f"{sys.argv[1] if len(sys.argv)>1 else 'default'} FREERUN={os.environ.get('SS_FREERUN','0')}: aggregate kernel {ms/n:.3f} ms")

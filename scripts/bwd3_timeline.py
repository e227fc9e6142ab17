"""NOTE (round 2): the clock64 timeline is compiled out of release builds; rebuild the library with -DAGB_TIMELINE
(add it to FLAGS in attention-gan_b200/agb_native/build_native.py) before running this script."""
"""Timeline of CTA 0 of damsm_bwd3_kernel (AGB_DAMSM_DEBUG=16): when the MMA issuer, the TMA producer and three
epilogue warps reach each region tile.  Times in microseconds from the first stamp (SM clock at 1.965 GHz)."""
import ctypes, os, sys
os.environ["AGB_DAMSM_DEBUG"] = "16"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import attention_gan_b200 as pkg
from attention_gan_b200.agb_native import ops
lib = pkg.native.lib()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
g = torch.Generator().manual_seed(0)
img = torch.randn(B, 256, 289, generator=g).cuda()
wrd = torch.randn(B, 18, 256, generator=g).cuda().transpose(1, 2)
lens = torch.randint(2, 19, (B,), generator=g).to(torch.int32).cuda()
dm = torch.randn(B, B, generator=g).cuda() * 1e-3
for _ in range(2):
    m = ops.damsm_fwd(img, wrd, lens, 4.0, 5.0, 1e-8, 0, False, 1)[0]
    ops.damsm_bwd(img, wrd, lens, 4.0, 5.0, 1e-8, dm, None, True, 1, m)
torch.cuda.synchronize()
K = 96
buf = np.zeros((6, K, 2), np.int64)
fn = lib.agb_damsm_debug_timeline
fn.restype = ctypes.c_int
fn.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert fn(buf.ctypes.data, buf.size) == 0
t0 = buf[0, 0, 0]
us = lambda x: (x - t0) / 1965.0
print("tile  mma_issue   mma_done_issue | w0 wait  w0 start  w0 end | w3 wait   w3 end | w8 wait   w8 end")
for k in range(3, 48):
    print(f"{k:4d}  {us(buf[0,k,0]):9.2f}  {us(buf[0,k,1]):9.2f}      | {us(buf[1,k,0]):8.2f} {us(buf[5,k,1]):8.2f} {us(buf[1,k,1]):8.2f} |"
          f" {us(buf[2,k,0]):8.2f} {us(buf[2,k,1]):8.2f} | {us(buf[3,k,0]):8.2f} {us(buf[3,k,1]):8.2f}")
print("item  producer: at item start, after mma_done wait | MMA: v_full passed")
for i in range(1, 16):
    print(f"{i:4d}  {us(buf[4,i,0]):9.2f} {us(buf[4,i,1]):9.2f} | {us(buf[5,i,0]):9.2f}")

cta = np.zeros((160, 4), np.int64)
fn2 = lib.agb_damsm_debug_cta_times
fn2.restype = ctypes.c_int
fn2.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert fn2(cta.ctypes.data, cta.size) == 0
cta = cta[:148]
g0 = cta[:, 2].min()
dur_gt = (cta[:, 3] - cta[:, 2]) / 1e3
dur_ck = (cta[:, 1] - cta[:, 0])
print("per-CTA duration (globaltimer, us): min %.1f median %.1f max %.1f; kernel span %.1f us" %
      (dur_gt.min(), np.median(dur_gt), dur_gt.max(), (cta[:, 3].max() - g0) / 1e3))
print("SM clock during the kernel (clock64 / globaltimer): median %.0f MHz" % np.median(dur_ck / dur_gt))
print("start skew (us): max %.1f" % ((cta[:, 2].max() - g0) / 1e3))
order = np.argsort(dur_gt)
print("slowest CTAs:", [(int(i), round(float(dur_gt[i]), 1)) for i in order[-6:]], "fastest:", [(int(i), round(float(dur_gt[i]), 1)) for i in order[:4]])
np.save(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "bwd3_cta_us.npy"), dur_gt)
np.save(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "bwd3_lens.npy"), lens.cpu().numpy())

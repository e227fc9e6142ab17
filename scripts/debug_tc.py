import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import attention_gan_b200 as pkg
from attention_gan_b200.agb_native import ops
from oracle import closed_form as cf, ref_port as rp
np.set_printoptions(precision=4, linewidth=200, suppress=True)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 7
full = (sys.argv[2] == "full") if len(sys.argv) > 2 else True
img, wrd, _, _, _, lens, _ = rp.synth_damsm(B, seed=200 + B, full_len=full)
img3 = img.cuda().reshape(B, 256, -1).contiguous()
m, _, _ = ops.damsm_fwd(img3, wrd.cuda(), lens.cuda().to(torch.int32), 4.0, 5.0, 1e-8, 0, False, 1)
m32, _, _ = ops.damsm_fwd(img3, wrd.cuda(), lens.cuda().to(torch.int32), 4.0, 5.0, 1e-8, 0, False, 0)
print("lens", lens.tolist())
print("tc\n", m.cpu().numpy()[:8, :8])
print("fp32\n", m32.cpu().numpy()[:8, :8])
print("diff\n", (m - m32).cpu().numpy()[:8, :8])

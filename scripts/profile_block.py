"""One tensor-core DAMSM forward + backward of a Bi x Bc row block through the ops layer (for ncu / timing).
    python scripts/profile_block.py [Bi] [Bc] [math] [reps]
Prints per-kernel times from the library's CUDA-event hooks."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import attention_gan_b200 as pkg  # noqa: E402
from attention_gan_b200.agb_native import native, ops  # noqa: E402

Bi = int(sys.argv[1]) if len(sys.argv) > 1 else 512
Bc = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
math = native.MATH_NAMES[sys.argv[3]] if len(sys.argv) > 3 else native.AGB_MATH_TC_F16
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
g = torch.Generator().manual_seed(0)
img = torch.randn(Bi, 256, 289, generator=g).cuda()
wrd = torch.randn(Bc, 18, 256, generator=g).cuda().transpose(1, 2)
lens = torch.randint(2, 19, (Bc,), generator=g)
lens[0] = 18
l32 = lens.cuda().to(torch.int32)
dm = (torch.randn(Bi, Bc, generator=g) * (1.0 / Bc)).cuda()
lib = native.lib()


def once():
    m, _, _, ws = ops.damsm_fwd(img, wrd, l32, 4.0, 5.0, 1e-8, 0, False, math, keep_ws=True, save=True)
    return ops.damsm_bwd(img, wrd, l32, 4.0, 5.0, 1e-8, dm, None, True, math, m, ws, True)


once()
torch.cuda.synchronize()
lib.agb_prof_enable(1)
for _ in range(reps):
    once()
torch.cuda.synchronize()
names = {2: "fwd2", 3: "bwd3", 6: "dimg", 7: "dwords"}
tot = 0.0
for tag, nm in names.items():
    ms, n = ctypes.c_double(0), ctypes.c_longlong(0)
    lib.agb_prof_read(tag, ctypes.byref(ms), ctypes.byref(n))
    print(f"{nm:7s} {ms.value / reps:8.3f} ms/step  ({n.value // reps} launches)")
    tot += ms.value / reps
print(f"sum     {tot:8.3f} ms   pairs {Bi * Bc}  mean len {float(lens.float().mean()):.2f}")
lib.agb_prof_enable(0)

"""Times the DAMSM forward/backward native calls alone (CUDA events inside the library)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import attention_gan_b200 as pkg
from attention_gan_b200.agb_native import ops
from oracle import ref_port as rp
lib = pkg.native.lib()
def read(tag):
    ms = ctypes.c_double(); n = ctypes.c_longlong(); lib.agb_prof_read(tag, ctypes.byref(ms), ctypes.byref(n)); return ms.value, n.value
sizes = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [48, 256, 1024]
math = int(sys.argv[2]) if len(sys.argv) > 2 else 1
bwd = len(sys.argv) > 3 and sys.argv[3] == "bwd"
for B in sizes:
    g = torch.Generator().manual_seed(0)
    img = torch.randn(B, 256, 289, generator=g).cuda()
    wrd = torch.randn(B, 18, 256, generator=g).cuda().transpose(1, 2)
    lens = torch.randint(2, 19, (B,), generator=g).to(torch.int32).cuda()
    dm = torch.randn(B, B, generator=g).cuda() * 1e-3
    Lbar = lens.float().mean().item()
    for it in range(3):
        mfw = ops.damsm_fwd(img, wrd, lens, 4.0, 5.0, 1e-8, 0, False, math)[0]
        if bwd: ops.damsm_bwd(img, wrd, lens, 4.0, 5.0, 1e-8, dm, None, True, math, mfw)
    torch.cuda.synchronize()
    lib.agb_prof_enable(1)
    reps = 5
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    for it in range(reps):
        ops.damsm_fwd(img, wrd, lens, 4.0, 5.0, 1e-8, 0, False, math)
    e1.record()
    if bwd:
        for it in range(reps):
            ops.damsm_bwd(img, wrd, lens, 4.0, 5.0, 1e-8, dm, None, True, math, mfw)
    e2.record()
    torch.cuda.synchronize()
    out = {t: read(t) for t in (1, 2, 3)}
    lib.agb_prof_enable(0)
    fwd_ms = e0.elapsed_time(e1) / reps
    kms = out[2][0] / max(out[2][1], 1)
    flop = 4 * 289 * Lbar * 256 * B * B
    print(f"B={B} Lbar={Lbar:.1f} fwd call {fwd_ms:.3f} ms, fused fwd kernel {kms:.3f} ms -> {flop/kms/1e9:.1f} TFLOP/s algorithmic "
          f"({B*B/kms/1e3:.2f} Mpairs/s)" + (f"; bwd call {e1.elapsed_time(e2)/reps:.3f} ms (kernel tag3 {out[3][0]/max(out[3][1],1):.3f} ms, sgemm {out[1][0]/reps:.3f} ms/call)" if bwd else ""))

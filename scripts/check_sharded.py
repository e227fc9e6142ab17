"""torchrun script: the NCCL-sharded DAMSM losses equal the single-process losses on the concatenated
batch (value and every gradient).  Usage:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      scripts/check_sharded.py [global_batch] [math] [ragged]
With `ragged`, every rank but 0 hands over captions padded to T - 1 only (the reference's RNN pads to the local batch's
longest caption, rnn_encoder.py:89-92).  Rank 0 also checks the sharded loss and gradients against the fp64 oracle."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import attention_gan_b200 as pkg
from oracle import ref_port as rp      # seeded inputs only


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    math = sys.argv[2] if len(sys.argv) > 2 else "fp32"
    ragged = len(sys.argv) > 3 and sys.argv[3] == "ragged"
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    img, wrd, cnn, rnn, labels, lens, cls = rp.synth_damsm(B, seed=42, n_classes=max(2, B // 3))
    n = B // world
    sl = slice(rank * n, rank * n + n)
    T = wrd.shape[2]
    if ragged:
        lens[n:] = torch.clamp(lens[n:], max=T - 1)

    def leaf(t):
        return t.to(dev).requires_grad_(True)

    # sharded
    T_loc = T - 1 if (ragged and rank > 0) else T
    im, wd, cn, rn = leaf(img[sl]), leaf(wrd[sl][:, :, :T_loc].contiguous()), leaf(cnn[sl]), leaf(rnn[sl])
    L = pkg.DAMSMLoss(dev, math=math, process_group=dist.group.WORLD, att_maps="packed")
    wl, sls, att = L.get_losses(im, cn, wd, rn, torch.arange(n, device=dev), lens[sl].to(dev), cls[sl])
    (wl + sls).backward()
    # single process on the concatenated batch (every rank does it redundantly)
    im1, wd1, cn1, rn1 = leaf(img), leaf(wrd), leaf(cnn), leaf(rnn)
    L1 = pkg.DAMSMLoss(dev, math=math, att_maps="packed")
    wl1, sl1, att1 = L1.get_losses(im1, cn1, wd1, rn1, labels.to(dev), lens.to(dev), cls)
    (wl1 + sl1).backward()

    def rel(a, b):
        return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()

    errs = dict(wloss=abs(wl.item() - wl1.item()) / abs(wl1.item()), sloss=abs(sls.item() - sl1.item()) / abs(sl1.item()),
                dimg=rel(im.grad, im1.grad[sl]), dwords=rel(wd.grad, wd1.grad[sl][:, :, :T_loc]),
                dcnn=rel(cn.grad, cn1.grad[sl]), drnn=rel(rn.grad, rn1.grad[sl]), att=rel(att, att1[sl][:, :T_loc]))
    tol = 1e-5 if math == "fp32" else 2e-3
    ok = all(v <= tol for v in errs.values()) and wd.grad.shape[2] == T_loc
    if rank == 0 and B <= 64:
        # oracle parity of the SHARDED result (fp64 closed form on the concatenated batch)
        from oracle import closed_form as cf
        wl0, _, dc0, dw0 = cf.words_loss_fwd_bwd(img.numpy().reshape(B, img.shape[1], -1), wrd.numpy(), labels.numpy(),
                                                 lens.numpy(), cls)
        sl0, _, dcnn0, drnn0 = cf.sentence_loss_fwd_bwd(cnn.numpy(), rnn.numpy(), labels.numpy(), cls)
        gt = 1e-4 if math == "fp32" else 5e-3
        o = dict(wloss=abs(wl.item() - wl0) / abs(wl0), sloss=abs(sls.item() - sl0) / abs(sl0),
                 dimg=rel(im.grad.double().cpu(), torch.from_numpy(dc0.reshape(img.shape))[sl]),
                 dwords=rel(wd.grad.double().cpu(), torch.from_numpy(dw0)[sl][:, :, :T_loc]),
                 dcnn=rel(cn.grad.double().cpu(), torch.from_numpy(dcnn0)[sl]))
        ok = ok and o["wloss"] <= (1e-5 if math in ("fp32", "f16x2") else 1e-4) and o["sloss"] <= 1e-5 and \
            o["dimg"] <= gt and o["dwords"] <= gt and o["dcnn"] <= 1e-4
        print("sharded-vs-oracle: " + " ".join(f"{k}={v:.2e}" for k, v in o.items()), flush=True)
    t = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"sharded-vs-single world={world} B={B} math={math}: " + " ".join(f"{k}={v:.2e}" for k, v in errs.items()),
              "OK" if t.item() == 1.0 else "MISMATCH", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if t.item() == 1.0 else 1)


if __name__ == "__main__":
    main()

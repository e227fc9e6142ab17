"""GPU test of the product's host -> device prefetcher (agb_native/pipeline.py): double-buffered copies on a side
stream hand the right data to the compute stream, buffers are only overwritten after their consumer has finished,
and a DAMSM step fed through it gives the numbers of a step fed directly."""
import pytest
import torch

from oracle import ref_port as rp

pytestmark = pytest.mark.gpu


def test_prefetcher_delivers_batches_in_order_and_protects_buffers():
    import attention_gan_b200 as agb
    pf = agb.HostPrefetcher("cuda")
    n = 1 << 22
    hosts = [(torch.full((n,), float(i)).pin_memory(), torch.arange(8, dtype=torch.int32).add_(i).pin_memory()) for i in range(6)]
    sums = []
    pf.submit(hosts[0])
    for i in range(6):
        a, b = pf.acquire()
        if i + 1 < 6:
            pf.submit(hosts[i + 1])                         # overlaps the (slow) consumer below
        x = a
        for _ in range(20):                                 # keep the compute stream busy on THIS buffer
            x = x * 1.0 + 0.0
        sums.append((x.sum() / n, b.clone()))
        pf.release()
    torch.cuda.synchronize()
    for i, (s, b) in enumerate(sums):
        assert abs(s.item() - i) < 1e-6
        assert torch.equal(b.cpu(), torch.arange(8, dtype=torch.int32) + i)
    with pytest.raises(RuntimeError):
        for _ in range(3):
            pf.submit(hosts[0])                             # more sets in flight than buffers


def test_damsm_step_through_the_prefetcher_equals_direct_step():
    import attention_gan_b200 as agb
    B = 32
    img, wrd, cnn, rnn, labels, lens, cls = rp.synth_damsm(B, seed=77, n_classes=8)
    loss = agb.DAMSMLoss("cuda", math="f16", att_maps=None)
    host = [img.contiguous().pin_memory(), wrd.transpose(1, 2).contiguous().pin_memory(), cnn.pin_memory(),
            rnn.pin_memory(), lens.to(torch.int32).pin_memory(), torch.from_numpy(cls).to(torch.int32).pin_memory()]

    def step(ts):
        im, wd, cn, rn, ln, cl = ts
        for t in (im, wd, cn, rn):
            t.grad = None
            t.requires_grad_(True)
        wl, sl, _ = loss.get_losses(im, cn, wd.transpose(1, 2), rn, labels.cuda(), ln, cl)
        (wl + sl).backward()
        return wl.item(), sl.item(), im.grad.clone(), wd.grad.clone()

    direct = step([t.cuda() for t in host])
    pf = agb.HostPrefetcher("cuda")
    pf.submit(host)
    for i in range(3):
        ts = pf.acquire()
        if i < 2:
            pf.submit(host)
        out = step(ts)
        pf.release()
        assert out[0] == direct[0] and out[1] == direct[1]
        assert torch.equal(out[2], direct[2]) and torch.equal(out[3], direct[3])

"""Emulates the 16-bit operand roundings of the tensor-core DAMSM forward in numpy (fp64 arithmetic) and reports
the relative loss error per golden fixture for each subset of roundings.  Used to pick the operand splits."""
import itertools, math, sys
import numpy as np
sys.path.insert(0, ".")
from oracle import closed_form as cf

def r16(x, mode):
    if mode == "f16":
        return x.astype(np.float16).astype(np.float64)
    # hi+lo
    hi = x.astype(np.float16).astype(np.float64)
    lo = (x - hi).astype(np.float16).astype(np.float64)
    return hi + lo

def fwd(c, words, lens, rc1, rw1, re, rc2, rwn, g1=4.0, g2=5.0):
    """r*: None (exact), 'f16', 'split' for: C in GEMM1, W in GEMM1, e in GEMM2, C in GEMM2, w in cosine numerator"""
    R_ = lambda x, m: x if m is None else r16(x, m)
    Bi, D, R = c.shape
    m = np.zeros((Bi, len(lens)))
    c1, c2 = R_(c, rc1), R_(c, rc2)
    for i, L in enumerate(int(x) for x in lens):
        w = words[i, :, :L]
        s = np.einsum("bdr,dt->brt", c1, R_(w, rw1)) / math.sqrt(D)
        a = cf._softmax(s, 2)
        e = R_(np.exp(g1 * a), re)
        V = np.einsum("bdr,brt->bdt", c2, e)
        n = np.einsum("dt,bdt->bt", R_(w, rwn), V)
        den = np.linalg.norm(w, axis=0)[None] * np.linalg.norm(V, axis=1)
        m[:, i] = cf._lse(g2 * n / den, 1)
    return m

def loss_of(m, g, g3, lam):
    sim = g3 * m
    cm = cf.class_mask(g["class_ids"]) if bool(g["has_class_ids"]) else None
    if cm is not None: sim = np.where(cm, -np.inf, sim)
    return cf.two_way_ce_fwd_bwd(sim, g["labels"], lam=lam)[0]

names = ["damsm_small", "damsm_small_cls", "damsm_gammas", "damsm_real_cls", "damsm_real_trained"]
variants = {
  "all f16": ("f16",)*5,
  "C1 split": ("split","f16","f16","f16","f16"),
  "W split (g1+num)": ("f16","split","f16","f16","split"),
  "g1 both split": ("split","split","f16","f16","f16"),
  "g1 both split + wn exact": ("split","split","f16","f16",None),
  "only g1 rounding": ("f16","f16",None,None,None),
  "only e": (None,None,"f16",None,None),
  "only C2": (None,None,None,"f16",None),
  "only wn": (None,None,None,None,"f16"),
  "e+C2": (None,None,"f16","f16",None),
  "g2 both split": ("f16","f16","split","split","f16"),
  "all split": ("split",)*5,
  "g1 split, e split": ("split","split","split","f16",None),
  "g1 split, C2 split": ("split","split","f16","split",None),
}
for nm in names:
    g = np.load(f"tests/golden/{nm}.npz")
    print(nm, {k: g[k].shape for k in ("img","words")}, [k for k in g.files if "gamma" in k or "lam" in k])
    c = g["img"].astype(np.float64).reshape(g["img"].shape[0], g["img"].shape[1], -1)
    words = g["words"].astype(np.float64)
    if words.shape[1] != c.shape[1]: words = words.transpose(0, 2, 1)
    gam = [float(g[k]) if k in g.files else d for k, d in (("gamma1",4.0),("gamma2",5.0),("gamma3",10.0),("wlambda",5.0))]
    ref = float(g["wloss_f64"])
    for vn, v in variants.items():
        m = fwd(c, words, g["cap_lens"], *v, g1=gam[0], g2=gam[1])
        l = loss_of(m, g, gam[2], gam[3])
        print(f"   {vn:28s} rel err {abs(l-ref)/abs(ref):.2e}")

"""numpy prototype of the FWHT QSPA check node used by the CUDA kernel, checked against the golden frames
(development aid only)."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from oracle import pyoracle as po
from tools.proto_tems import load

def wht(a):
    a = a.copy(); n = a.shape[-1]; h = 1
    while h < n:
        a = a.reshape(a.shape[:-1] + (n // (2 * h), 2, h))
        a = np.stack([a[..., 0, :] + a[..., 1, :], a[..., 0, :] - a[..., 1, :]], axis=-2).reshape(a.shape[:-3] + (n,))
        h *= 2
    return a

def decode(code, mul, inv, L_ch, max_iter, relfloor):
    N, M, q = code["N"], code["M"], code["q"]
    Lch = np.zeros((N, q)); Lch[:, 1:] = L_ch
    edges = [(r, d, v, g) for r, links in enumerate(code["chk"]) for d, (v, g) in enumerate(links)]
    eid = {(r, v): i for i, (r, d, v, g) in enumerate(edges)}
    var_edges = [[eid[(r, v)] for (r, g) in code["var"][v]] for v in range(N)]
    E = len(edges); ar = np.arange(q)
    v2c = np.array([Lch[v] for (r, d, v, g) in edges]); c2v = np.zeros((E, q))
    def decide(x):
        k = int(np.argmax(x)); return k if x[k] > 0 else 0
    for it in range(1, max_iter + 1):
        post = Lch.copy()
        for v in range(N):
            for e in var_edges[v]: post[v] = post[v] + c2v[e]
        dec = np.array([decide(post[v]) for v in range(N)])
        ok = all(np.bitwise_xor.reduce([mul[g, dec[v]] for (v, g) in links]) == 0 for links in code["chk"])
        if ok: return 1, it, dec, post[:, 1:]
        if it == max_iter: break
        for i, (r, d, v, g) in enumerate(edges):
            before = decide(v2c[i]); new = post[v] - c2v[i]
            if decide(new) != before: new = 0.5 * v2c[i] + 0.5 * new
            v2c[i] = new
        for r, links in enumerate(code["chk"]):
            es = [eid[(r, v)] for (v, g) in links]; hs = [g for (v, g) in links]; dc = len(es)
            F = np.zeros((dc, q))
            for d in range(dc):
                V = v2c[es[d]]; p = np.exp(V - V.max())
                F[d] = p[mul[inv[hs[d]], ar]]
            Fh = wht(F)
            for d in range(dc):
                R = np.prod(np.delete(Fh, d, axis=0), axis=0)
                rr = wht(R[None])[0]
                fl = abs(rr[0]) * relfloor + 1e-300
                a = np.maximum(rr[mul[hs[d], ar]], fl)
                out = np.log(a) - np.log(max(rr[0], fl)); out[0] = 0
                c2v[es[d]] = out
    return 0, max_iter, dec, post[:, 1:]

if __name__ == "__main__":
    name = sys.argv[1]; relfloor = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-300
    z, meta = load(name)
    b = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "codes.json")))
    c = b["codes"][meta["code"]]
    code = dict(N=c["N"], M=c["M"], q=c["q"], var=[[tuple(e) for e in l] for l in c["var"]], chk=[[tuple(e) for e in l] for l in c["chk"]])
    q = meta["q"]; mul, inv = po.gf_tables(q)
    cons = np.array(b["constellations"][meta["constellation"][:-4]])
    bad = 0; worst = 0
    for f in range(len(z["ret"])):
        L = po.demap_bpsk(q, meta["N"], z["rx"][f, :, 0], meta["sigma"]) if meta["nqam"] == 2 else po.demap_qam(q, meta["N"], z["rx"][f], cons, meta["sigma"])
        r, it, dec, post = decode(code, mul, inv, L, meta["max_iter"], relfloor)
        same = (r, it) == (int(z["ret"][f]), int(z["iters"][f])) and np.array_equal(dec, z["dec_sym"][f])
        d = np.abs(post[:, z["lpost_idx"]] - z["lpost_sample"][f]).max()
        worst = max(worst, d); bad += not same
        if not same: print("frame", f, "MISMATCH", (r, it), (int(z["ret"][f]), int(z["iters"][f])))
    print(name, "frames", len(z["ret"]), "mismatches", bad, "worst |dLpost|", worst, "max|Lch|", np.abs(L).max())

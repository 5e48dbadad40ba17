"""Statistics of the T-EMS pair search on real frames (development aid): how many (a,b) leaves can matter."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from oracle import pyoracle as po
from tools.proto_tems import load

def run(name, nfr):
    z, meta = load(name)
    b = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "codes.json")))
    c = b["codes"][meta["code"]]
    code = dict(N=c["N"], M=c["M"], q=c["q"], var=[[tuple(e) for e in l] for l in c["var"]], chk=[[tuple(e) for e in l] for l in c["chk"]])
    q = meta["q"]; mul, inv = po.gf_tables(q); ar = np.arange(q)
    N, M = code["N"], code["M"]
    stats = []
    for f in range(nfr):
        L = po.demap_bpsk(q, N, z["rx"][f, :, 0], meta["sigma"])
        Lch = np.zeros((N, q)); Lch[:, 1:] = L
        edges = [(r, d, v, g) for r, links in enumerate(code["chk"]) for d, (v, g) in enumerate(links)]
        eid = {(r, v): i for i, (r, d, v, g) in enumerate(edges)}
        var_edges = [[eid[(r, v)] for (r, g) in code["var"][v]] for v in range(N)]
        E = len(edges)
        v2c = np.array([Lch[v] for (r, d, v, g) in edges]); c2v = np.zeros((E, q))
        def decide(x):
            k = int(np.argmax(x)); return k if x[k] > 0 else 0
        for it in range(1, 10):
            post = Lch.copy()
            for v in range(N):
                for e in var_edges[v]: post[v] = post[v] + c2v[e]
            dec = np.array([decide(post[v]) for v in range(N)])
            if all(np.bitwise_xor.reduce([mul[g, dec[v]] for (v, g) in links]) == 0 for links in code["chk"]): break
            for i, (r, d, v, g) in enumerate(edges):
                before = decide(v2c[i]); new = post[v] - c2v[i]
                if decide(new) != before: new = 0.25 * v2c[i] + 0.75 * new
                v2c[i] = new
            new_c2v = np.zeros_like(c2v)
            for r, links in enumerate(code["chk"]):
                dc = len(links); es = [eid[(r, v)] for (v, g) in links]; hs = [g for (v, g) in links]
                V = v2c[es]
                beta = [mul[hs[d], decide(V[d])] for d in range(dc)]
                syn = np.bitwise_xor.reduce(beta)
                dU = np.zeros((dc, q))
                for d in range(dc):
                    hi = inv[hs[d]]
                    dU[d] = V[d][mul[hi, beta[d]]] - V[d][mul[hi, ar ^ beta[d]]]
                order = np.argsort(dU, axis=0, kind="stable")
                c0, c1 = order[0], order[1]
                m0 = dU[c0, ar]; m1 = dU[c1, ar]
                # all allowed entries (eta>=1): (eta, col, cost)
                ent = [(e, c0[e], m0[e]) for e in range(1, q)] + [(e, c1[e], m1[e]) for e in range(1, q)]
                cost = np.array([x[2] for x in ent]); col = np.array([x[1] for x in ent]); eta = np.array([x[0] for x in ent])
                # brute force exact pair search
                S = cost[:, None] + cost[None, :]
                valid = (col[:, None] < col[None, :])
                tgt = eta[:, None] ^ eta[None, :]
                dW = m0.copy(); dW[0] = 0
                useful = valid & (S < m0[tgt]) & (tgt != 0)
                nvalid = valid.sum()
                # global-bound pruning: entry useful only if cost + min_other < Wmax
                Wmax = m0[1:].max()
                cmin = cost.min()
                keep = cost + cmin < Wmax
                n_keep = keep.sum()
                pairs_keep = (valid & keep[:, None] & keep[None, :] & (S < Wmax)).sum()
                # improvement count: how many s get a double path
                best2 = np.full(q, np.inf)
                np.minimum.at(best2, tgt[valid], S[valid])
                nimp = (best2[1:] < m0[1:]).sum()
                stats.append((it, nvalid, useful.sum(), n_keep, pairs_keep, nimp, Wmax, np.median(m0[1:]), np.median(m1[1:])))
                # produce c2v via the same formulas (reuse prototype pieces minimal): need full CN; call proto for correctness is heavy -> replicate quickly
                ndev = np.where(best2 < m0, 2, 1); ndev[0] = 0
                dWf = np.where(best2 < m0, best2, m0); dWf[0] = 0
                T = np.full((dc, q), np.inf)
                # recover paths
                flat = np.where(valid, S, np.inf)
                for s in range(1, q):
                    if ndev[s] == 2:
                        cand = np.where((tgt == s) & valid, S, np.inf)
                        i, j = np.unravel_index(np.argmin(cand), cand.shape)
                        a, ca, bb, cb = eta[i], col[i], eta[j], col[j]
                        for d in range(dc):
                            if d != ca and d != cb: T[d, s] = min(T[d, s], dWf[s])
                        T[ca, bb] = min(T[ca, bb], dWf[s] - dU[ca, a]); T[cb, a] = min(T[cb, a], dWf[s] - dU[cb, bb])
                    else:
                        for d in range(dc):
                            if d != c0[s]: T[d, s] = min(T[d, s], dWf[s])
                T[:, 0] = 0
                for d in range(dc):
                    fill = np.where(c0 == d, m1, m0)
                    Td = np.where(np.isinf(T[d]), fill, T[d])
                    bs = syn ^ beta[d]; t = mul[hs[d], ar] ^ bs
                    x = Td[bs] - Td[t]
                    out = np.where(x < -0.5, x + 0.5, np.where(x > 0.5, x - 0.5, 0)); out[0] = 0
                    new_c2v[es[d]] = out
            c2v = new_c2v
    st = np.array(stats)
    print(name, "CN updates", len(st))
    for it in sorted(set(st[:, 0])):
        s = st[st[:, 0] == it]
        print("iter %d: n=%4d valid pairs %.0f useful %.0f (%.1f%%) | entries kept %.0f/510 pairs kept %.0f (%.1f%%) | s improved %.0f | Wmax %.1f med m0 %.1f med m1 %.1f" %
              (it, len(s), s[:, 1].mean(), s[:, 2].mean(), 100 * s[:, 2].mean() / s[:, 1].mean(), s[:, 3].mean(), s[:, 4].mean(),
               100 * s[:, 4].mean() / s[:, 1].mean(), s[:, 5].mean(), s[:, 6].mean(), s[:, 7].mean(), s[:, 8].mean()))

if __name__ == "__main__":
    run(sys.argv[1] if len(sys.argv) > 1 else "cfg3_tems_gf256", int(sys.argv[2]) if len(sys.argv) > 2 else 3)

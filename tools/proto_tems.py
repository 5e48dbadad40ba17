"""numpy prototype of the PARALLEL T-EMS formulation the CUDA kernel uses (prefix-min column DP for n_c<=2,
scatter-min output stage, on-the-fly fill-in), checked against the C oracle on the golden frames.
Development aid only -- not product code, not part of the tests."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from oracle import pyoracle as po

def load(name):
    z = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", name + ".npz"))
    meta = json.loads(str(z["meta"]))
    return z, meta

def decode_proto(code, mul, inv, L_ch, max_iter, nr, nc, factor, offset, dtype=np.float64):
    N, M, q = code["N"], code["M"], code["q"]
    INF = np.inf
    Lch = np.zeros((N, q), dtype); Lch[:, 1:] = L_ch.astype(dtype)
    edges = [(r, d, v, g) for r, links in enumerate(code["chk"]) for d, (v, g) in enumerate(links)]
    eid = {(r, v): i for i, (r, d, v, g) in enumerate(edges)}
    var_edges = [[eid[(r, v)] for (r, g) in code["var"][v]] for v in range(N)]
    E = len(edges)
    v2c = np.zeros((E, q), dtype); c2v = np.zeros((E, q), dtype)
    for i, (r, d, v, g) in enumerate(edges): v2c[i] = Lch[v]
    def decide(x):
        k = int(np.argmax(x)); return k if x[k] > 0 else 0
    ar = np.arange(q)
    for it in range(1, max_iter + 1):
        post = Lch.copy()
        for v in range(N):
            for e in var_edges[v]: post[v] = post[v] + c2v[e]
        dec = np.array([decide(post[v]) for v in range(N)])
        ok = True
        for r, links in enumerate(code["chk"]):
            s = 0
            for (v, g) in links: s ^= mul[g, dec[v]]
            if s: ok = False; break
        if ok: return 1, it, dec, post[:, 1:]
        if it == max_iter: break
        for i, (r, d, v, g) in enumerate(edges):
            before = decide(v2c[i]); new = post[v] - c2v[i]
            if decide(new) != before: new = dtype(0.25) * v2c[i] + dtype(0.75) * new
            v2c[i] = new
        new_c2v = np.zeros_like(c2v)
        for r, links in enumerate(code["chk"]):
            dc = len(links); es = [eid[(r, v)] for (v, g) in links]; hs = [g for (v, g) in links]
            V = v2c[es]
            beta = [mul[hs[d], decide(V[d])] for d in range(dc)]
            syn = 0
            for b in beta: syn ^= b
            dU = np.zeros((dc, q), dtype)
            for d in range(dc):
                hi = inv[hs[d]]
                vmax = V[d][mul[hi, beta[d]]]
                dU[d] = vmax - V[d][mul[hi, ar ^ beta[d]]]
            # ranks (stable)
            rank = np.zeros((dc, q), int)
            for d in range(dc):
                for c in range(dc):
                    rank[d] += (dU[c] < dU[d]) | ((dU[c] == dU[d]) & (c < d))
            allowed = rank < nr
            c0 = np.argmax(rank == 0, axis=0); c1 = np.argmax(rank == 1, axis=0)
            m0 = dU[c0, ar]
            best = np.full(q, INF, dtype); bb = np.zeros(q, int); bd = np.zeros(q, int)
            P = np.full(q, INF, dtype); Pc = np.zeros(q, int)
            if nc >= 2:
                for d in range(dc):
                    if d >= 1:
                        for b in range(1, q):
                            if allowed[d, b]:
                                cand = P[ar ^ b] + dU[d, b]
                                imp = cand < best
                                best[imp] = cand[imp]; bb[imp] = b; bd[imp] = d
                    col = np.where(allowed[d], dU[d], INF); col[0] = INF
                    imp = col < P
                    P[imp] = col[imp]; Pc[imp] = d
            dW = np.zeros(q, dtype); ndev = np.zeros(q, int); pa = np.zeros(q, int); pc1 = np.zeros(q, int)
            for s in range(1, q):
                if best[s] < m0[s]:
                    a = s ^ bb[s]; m = INF; cc = 0
                    for c in range(bd[s]):
                        if allowed[c, a] and dU[c, a] < m: m = dU[c, a]; cc = c
                    dW[s] = best[s]; ndev[s] = 2; pa[s] = a; pc1[s] = cc
                else:
                    dW[s] = m0[s]; ndev[s] = 1
            T = np.full((dc, q), INF, dtype)
            for s in range(q):
                for d in range(dc):
                    uses = (ndev[s] == 1 and c0[s] == d) or (ndev[s] == 2 and (pc1[s] == d or bd[s] == d))
                    if not uses: T[d, s] = min(T[d, s], dW[s])
            T[:, 0] = 0
            for s in range(1, q):
                if ndev[s] == 2:
                    a, b, ca, cb = pa[s], bb[s], pc1[s], bd[s]
                    T[ca, b] = min(T[ca, b], dW[s] - dU[ca, a])
                    T[cb, a] = min(T[cb, a], dW[s] - dU[cb, b])
            for d in range(dc):
                fill = np.where(c0 == d, dU[c1, ar], dU[c0, ar])
                Td = np.where(np.isinf(T[d]), fill, T[d])
                bs = syn ^ beta[d]
                t = mul[hs[d], ar] ^ bs
                x = (Td[bs] - Td[t]) / dtype(factor)
                out = np.where(x < -offset, x + dtype(offset), np.where(x > offset, x - dtype(offset), 0)).astype(dtype)
                out[0] = 0
                new_c2v[es[d]] = out
        c2v = new_c2v
    return 0, max_iter, dec, post[:, 1:]

if __name__ == "__main__":
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg5_tems_hbit256"
    nfr = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    z, meta = load(name)
    b = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "codes.json")))
    if meta["code"] in b["codes"]:
        c = b["codes"][meta["code"]]
        code = dict(N=c["N"], M=c["M"], q=c["q"], vdeg=[len(l) for l in c["var"]], cdeg=[len(l) for l in c["chk"]],
                    var=[[tuple(e) for e in l] for l in c["var"]], chk=[[tuple(e) for e in l] for l in c["chk"]])
    else:
        code = po.hbit_to_code(b["h_bit"][meta["code"]], 256)
    q = meta["q"]; mul, inv = po.gf_tables(q)
    oc = po.OracleCode(code)
    for f in range(nfr):
        L = po.demap_bpsk(q, meta["N"], z["rx"][f, :, 0], meta["sigma"])
        r, it, dec, post = decode_proto(code, mul, inv, L, meta["max_iter"], meta["p0"], meta["p1"], meta["factor"], meta["offset"])
        ro, ito, deco, posto = oc.decode(4, meta["max_iter"], meta["p0"], meta["p1"], meta["factor"], meta["offset"], L)
        print(f, (r, it), (ro, ito), "dec equal", np.array_equal(dec, deco), "max |dLpost|", np.abs(post - posto).max(),
              "golden", int(z["ret"][f]), int(z["iters"][f]))

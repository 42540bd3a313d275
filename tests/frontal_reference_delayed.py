"""NumPy executor of a front plan WITH DELAYED PIVOTS in static shapes — test infrastructure and the executable
specification of the next factorisation kernels.

Same plan, same block algebra as ``frontal_reference.py`` (explicit pivot-block inverse, W = inverse x F12, Schur
complement, sweeps as matrix-vector products), plus:

* every front reserves ``DC`` INCOMING slots per child at the head of its pivot block and ``DC`` OUTGOING slots at the
  head of its update block, whatever happens numerically: all shapes are known when the plan is built;
* after the pivot block is inverted and W formed, the unknown with the largest row of W is handed to the parent
  (rank-1 downdate of the inverse and of W), at most DC times, while max|W| > tau; the handed-up unknown occupies an
  outgoing slot, its Schur-complement row is  F[k, rest] - F[k, pivot block] W  like every other row of the update block;
* an outgoing slot that stays unused carries a unit diagonal entry and nothing else, so the parent's incoming slot
  is a decoupled unknown with right-hand side 0: no flags on the factorisation path;
* the pivot-block inverse is SYMMETRISED (average with its transpose).  The product forms S = F22 - F12^T (F11^-1 F12) from
  the upper block alone, so the antisymmetric part of the computed inverse (cond x eps) lands in S, then in the parent's
  pivot block, and is amplified by |W|^2 at every level: measured, this — not the restricted pivoting — is what costs the
  product most of its digits (config 1 raw solve 7.9e-7 -> 6.4e-11, 250 x 250 structured cells 4.5e+2 -> 2.0e-9 with the
  symmetrisation alone; ``symmetrise=False`` reproduces the loss);
* sweeps: the forward sweep sends the assembled right-hand side of a handed-up unknown to the parent in its slot of the
  update vector; the backward sweep reads its solution from the parent's incoming slot (``x_in``).

Layout of front f (unknowns):  [ incoming slots: DC per child | own 2s | outgoing slots: DC | update 2u ].
"""
import numpy as np

DC = 4


def _unk(nodes):
    u = np.empty(2 * len(nodes), dtype=np.int64)
    u[0::2] = 2 * nodes
    u[1::2] = 2 * nodes + 1
    return u


def factor(Kp, plan, tau=30.0, symmetrise=True):
    first, s, sptr, strct = plan["first"], plan["s"], plan["sptr"], plan["strct"]
    parent, cmap_ptr, cmap = plan["parent"], plan["cmap_ptr"], plan["cmap"]
    nf = plan["nfronts"]
    children = [[] for _ in range(nf)]
    for f in range(nf):
        if parent[f] >= 0:
            children[parent[f]].append(f)
    Kc = Kp.tocsr()
    fronts = [None] * nf
    for f in range(nf):                                     # post-order: children first
        own = _unk(np.arange(first[f], first[f] + s[f]))
        upd = _unk(strct[sptr[f]:sptr[f + 1]].astype(np.int64))
        s2, u2, DS = len(own), len(upd), DC * len(children[f])
        p2, r2 = DS + s2, DC + u2
        m = p2 + r2
        F = np.zeros((m, m))
        # original entries of the own rows (and, symmetric, columns); the outgoing slots start empty
        blk_o = Kc[own, :][:, own].toarray()
        blk_u = Kc[own, :][:, upd].toarray()
        F[DS:p2, DS:p2] = blk_o
        F[DS:p2, p2 + DC:] = blk_u
        F[p2 + DC:, DS:p2] = blk_u.T
        for q, c in enumerate(children[f]):
            fc = fronts[c]
            cm = cmap[cmap_ptr[c]:cmap_ptr[c + 1]].astype(np.int64)           # child's update NODES -> [own | update] of f
            pos_upd = np.empty(2 * len(cm), dtype=np.int64)
            pos_upd[0::2] = 2 * cm
            pos_upd[1::2] = 2 * cm + 1
            pos_upd = np.where(pos_upd < s2, DS + pos_upd, DS + DC + pos_upd)  # outgoing slots sit between own and update
            pos = np.concatenate([q * DC + np.arange(DC), pos_upd])           # child's slot k -> incoming slot q*DC + k
            F[np.ix_(pos, pos)] += fc["S"]
            fc["S"] = None
        P = np.arange(p2)
        U = np.arange(p2 + DC, m)
        F11 = F[np.ix_(P, P)]
        Einv = np.linalg.inv(F11)
        if symmetrise:
            Einv = 0.5 * (Einv + Einv.T)
        W = Einv @ F[np.ix_(P, U)]
        dl = []                                             # pivot-block indices handed to the parent, in slot order
        while len(dl) < DC and u2 > 0:
            g = np.abs(W).max(axis=1)
            k = int(np.argmax(g))
            if g[k] <= tau:
                break
            piv = Einv[k, k]
            W = W - np.outer(Einv[:, k], W[k, :]) / piv
            Einv = Einv - np.outer(Einv[:, k], Einv[k, :]) / piv
            W[k, :] = 0.0
            Einv[k, :] = 0.0
            Einv[:, k] = 0.0
            dl.append(k)
        # final quantities from the ORIGINAL front rows (the downdates above only served the selection)
        E = np.array([i for i in range(p2) if i not in dl], dtype=np.int64)
        Ebar = np.zeros((p2, p2))
        if len(E):
            Ebar[np.ix_(E, E)] = np.linalg.inv(F11[np.ix_(E, E)])
            if symmetrise:
                Ebar = 0.5 * (Ebar + Ebar.T)
        # rows of the update block: outgoing slots are aliases of the handed-up pivot rows
        Frest_P = np.zeros((r2, p2))                        # F[rest, pivot block]
        Frest_rest = np.zeros((r2, r2))
        for a, k in enumerate(dl):
            Frest_P[a, :] = F[k, :p2]
            Frest_rest[a, DC:] = F[k, p2 + DC:]
            for b, k2 in enumerate(dl):
                Frest_rest[a, b] = F[k, k2]
        Frest_P[DC:, :] = F[p2 + DC:, :p2]
        Frest_rest[DC:, DC:] = F[p2 + DC:, p2 + DC:]
        for a, k in enumerate(dl):
            Frest_rest[DC:, a] = F[p2 + DC:, k]
        Wext = Ebar @ Frest_P.T                             # p2 x r2, zero rows at the handed-up unknowns
        S = Frest_rest - Frest_P @ Wext
        for a in range(len(dl), DC):                        # unused outgoing slots: decoupled unit entries
            S[a, :] = 0.0
            S[:, a] = 0.0
            S[a, a] = 1.0
        if parent[f] < 0:
            assert u2 == 0 and not dl
        fronts[f] = dict(Ebar=Ebar, Wext=Wext, S=S, dl=dl, DS=DS, s2=s2, u2=u2, own=own, upd=upd)
    return fronts, children


def solve(fronts, children, plan, b):
    nf = plan["nfronts"]
    parent, cmap_ptr, cmap = plan["parent"], plan["cmap_ptr"], plan["cmap"]
    slot_of = {}
    for f in range(nf):
        for q, c in enumerate(children[f]):
            slot_of[c] = q
    t_out = [None] * nf                                     # forward: update vectors [outgoing slots | update unknowns]
    zP = [None] * nf
    for f in range(nf):
        fr = fronts[f]
        DS, s2, u2 = fr["DS"], fr["s2"], fr["u2"]
        p2, r2 = DS + s2, DC + u2
        y = np.zeros(p2 + r2)
        y[DS:p2] = b[fr["own"]]
        for q, c in enumerate(children[f]):
            cm = cmap[cmap_ptr[c]:cmap_ptr[c + 1]].astype(np.int64)
            pos_upd = np.empty(2 * len(cm), dtype=np.int64)
            pos_upd[0::2] = 2 * cm
            pos_upd[1::2] = 2 * cm + 1
            pos_upd = np.where(pos_upd < s2, DS + pos_upd, DS + DC + pos_upd)
            pos = np.concatenate([q * DC + np.arange(DC), pos_upd])
            y[pos] += t_out[c]
            t_out[c] = None
        yP = y[:p2]
        yR = y[p2:].copy()
        for a, k in enumerate(fr["dl"]):
            yR[a] = yP[k]                                   # the handed-up unknown takes its assembled right-hand side along
        zP[f] = fr["Ebar"] @ yP
        t = yR - fr["Wext"].T @ yP
        t[len(fr["dl"]):DC] = 0.0
        t_out[f] = t
    x = np.zeros_like(b)
    x_in = [None] * nf                                      # backward: solution at the incoming slots of every front
    for f in range(nf - 1, -1, -1):
        fr = fronts[f]
        DS, s2, u2 = fr["DS"], fr["s2"], fr["u2"]
        p2 = DS + s2
        xR = np.zeros(DC + u2)
        xR[DC:] = x[fr["upd"]]
        if parent[f] >= 0:
            q = slot_of[f]
            xR[:DC] = x_in[parent[f]][q * DC:(q + 1) * DC]
        xP = zP[f] - fr["Wext"] @ xR
        for a, k in enumerate(fr["dl"]):
            xP[k] = xR[a]                                   # solved at the parent
        x[fr["own"]] = xP[DS:p2]
        x_in[f] = xP[:DS]
    return x

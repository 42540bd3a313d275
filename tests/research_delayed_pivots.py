"""Research prototype (test infrastructure, CPU only): multifrontal block elimination with DELAYED PIVOTS.

The product's factorisation inverts every pivot block F11 whole (pivoting stays inside the block).  On an
indefinite operator that loses accuracy level by level (|W| = |F11^-1 F12| of 1e2..1e4 squares
into the Schur complement), 8e-7 raw error on config 1 and divergence on coarse structured meshes.  This
script runs the same front plan in NumPy with the classical cure: an unknown of the pivot block is eliminated
at a front only if a threshold test against its WHOLE front column passes; the others are handed to the
parent front (delayed), where more of their couplings are fully summed.

    python tests/research_delayed_pivots.py cfg1|<nx> [u]

prints, per threshold, the raw solve error against SuperLU, the refinement contraction, the number of delayed
unknowns (total, per level, largest front growth): u < 1 = classical threshold test (1x1 / 2x2 pivots), u >= 1 = tau of the
growth-based selection (largest row of W).  Results are summarised in DESIGN.md 4.4a.  NOTE: this prototype forms the Schur
complement from BOTH off-diagonal blocks (F_RE E^-1 F_ER), which is why it does not show the asymmetry amplification of the
product's one-block form — see ``frontal_reference_delayed.py`` for that finding and for the static-shape version.
"""
import sys
import time

import numpy as np

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.abspath(__file__)))


def choose_pivots(F, nfs, u):
    """Symmetric threshold pivoting on the leading nfs (fully summed) unknowns of the symmetric front F.
    Greedy 1x1 / 2x2 pivots (Bunch-Kaufman tests against the whole active column, threshold u).  Works on a
    copy; returns (eliminated local indices in order, delayed local indices)."""
    A = F.copy()
    m = A.shape[0]
    active = np.ones(m, bool)
    cand = list(range(nfs))
    elim = []
    progress = True
    while progress and cand:
        progress = False
        for i in list(cand):
            if i not in cand:
                continue
            rows = np.nonzero(active)[0]
            col = np.abs(A[rows, i])
            col_off = col.copy()
            col_off[rows == i] = 0.0
            cmax = col_off.max() if len(col_off) else 0.0
            aii = abs(A[i, i])
            if aii > 0 and aii >= u * cmax:
                # 1x1 pivot
                r = rows[rows != i]
                w = A[r, i] / A[i, i]
                A[np.ix_(r, r)] -= np.outer(w, A[i, r])
                active[i] = False
                cand.remove(i)
                elim.append(i)
                progress = True
                continue
            # 2x2 pivot with the fully summed unknown j most strongly coupled to i
            others = [j for j in cand if j != i]
            if not others:
                continue
            j = others[int(np.argmax(np.abs(A[others, i])))]
            B = A[np.ix_([i, j], [i, j])]
            det = B[0, 0] * B[1, 1] - B[0, 1] * B[1, 0]
            if det == 0:
                continue
            Binv = np.array([[B[1, 1], -B[0, 1]], [-B[1, 0], B[0, 0]]]) / det
            r = rows[(rows != i) & (rows != j)]
            if len(r):
                c2 = np.abs(A[np.ix_(r, [i, j])])
                # multipliers of the 2x2 pivot must stay below 1/u
                mult = np.abs(A[np.ix_(r, [i, j])] @ Binv)
                if mult.max() > 1.0 / u:
                    continue
            if len(r):
                Wm = A[np.ix_(r, [i, j])] @ Binv
                A[np.ix_(r, r)] -= Wm @ A[np.ix_([i, j], r)]
            active[i] = active[j] = False
            cand.remove(i); cand.remove(j)
            elim.extend([i, j])
            progress = True
    return elim, cand


def choose_by_growth(F, nfs, tau, dmax=8):
    """GPU-friendly variant: invert the whole pivot block, look at W = E^-1 F(E, rest); while some row of W exceeds tau,
    hand the unknown of the largest row to the parent (E^-1 is downdated by a rank-1 correction) — at most dmax times."""
    m = F.shape[0]
    E = list(range(nfs))
    dly = []
    Einv = np.linalg.inv(F[np.ix_(E, E)])
    while len(dly) < dmax and len(E) > 1:
        R = dly + list(range(nfs, m))
        if not R:
            break
        W = Einv @ F[np.ix_(E, R)]
        g = np.abs(W).max(axis=1)
        if g.max() <= tau:
            break
        k = int(np.argmax(g))
        # remove unknown E[k] from the eliminated set: inverse of the principal submatrix by a rank-1 downdate
        Einv = Einv - np.outer(Einv[:, k], Einv[k, :]) / Einv[k, k]
        keep = [j for j in range(len(E)) if j != k]
        Einv = Einv[np.ix_(keep, keep)]
        dly.append(E[k])
        E = [E[j] for j in keep]
    return E, dly


def factor(Kp, plan, u):
    first, s, sptr, strct = plan["first"], plan["s"], plan["sptr"], plan["strct"]
    parent = plan["parent"]
    nf = plan["nfronts"]
    children = [[] for _ in range(nf)]
    for f in range(nf):
        if parent[f] >= 0:
            children[parent[f]].append(f)
    fronts = [None] * nf
    Kc = Kp.tocsr()
    n_delayed = np.zeros(nf, int)
    for f in range(nf):
        own_nodes = np.arange(first[f], first[f] + s[f])
        own = np.empty(2 * len(own_nodes), np.int64); own[0::2] = 2 * own_nodes; own[1::2] = 2 * own_nodes + 1
        st_nodes = strct[sptr[f]:sptr[f + 1]].astype(np.int64)
        upd = np.empty(2 * len(st_nodes), np.int64); upd[0::2] = 2 * st_nodes; upd[1::2] = 2 * st_nodes + 1
        delayed_in = np.concatenate([fronts[c]["delayed"] for c in children[f]]) if children[f] else np.zeros(0, np.int64)
        fs = np.concatenate([delayed_in, own])           # delayed unknowns first: they have waited longest
        if parent[f] < 0:
            assert len(upd) == 0
        idx = np.concatenate([fs, upd])
        pos = {int(g): k for k, g in enumerate(idx)}
        m = len(idx)
        F = np.zeros((m, m))
        # original entries: rows of the front's own unknowns (and, symmetric, their columns)
        o0 = len(delayed_in)
        # (entries coupling them to a delayed unknown were assembled where that unknown was first fully summed)
        blk = Kc[own, :][:, idx[o0:]].toarray()
        F[o0:o0 + len(own), o0:] = blk
        F[o0:, o0:o0 + len(own)] = blk.T
        for c in children[f]:
            fc = fronts[c]
            p = np.array([pos[int(g)] for g in fc["rest"]], np.int64)
            F[np.ix_(p, p)] += fc["S"]
            fc["S"] = None
        nfs = len(fs)
        if parent[f] < 0:
            elim, dly = list(range(nfs)), []             # the root eliminates everything (plain inverse)
        else:
            elim, dly = choose_pivots(F, nfs, u) if u < 1 else choose_by_growth(F, nfs, u)
        E = np.array(elim, np.int64)
        R = np.array(dly + list(range(nfs, m)), np.int64)
        FEE = F[np.ix_(E, E)]
        Einv = np.linalg.inv(FEE) if len(E) else np.zeros((0, 0))
        W = Einv @ F[np.ix_(E, R)]
        S = F[np.ix_(R, R)] - F[np.ix_(R, E)] @ W
        fronts[f] = dict(E=idx[E], rest=idx[R], delayed=idx[np.array(dly, np.int64)] if dly else np.zeros(0, np.int64),
                         Einv=Einv, W=W, S=S, size=m, nfs=nfs)
        n_delayed[f] = len(dly)
    return fronts, children, n_delayed


def solve(fronts, children, plan, b):
    nf = plan["nfronts"]
    z = np.zeros_like(b)
    upd = [None] * nf
    for f in range(nf):
        fr = fronts[f]
        idx = np.concatenate([fr["E"], fr["rest"]])
        pos = {int(g): k for k, g in enumerate(idx)}
        y = np.zeros(len(idx))
        ne = len(fr["E"])
        # right-hand side entries enter where the unknown is ELIMINATED (delayed ones travel in the update vector)
        y[:ne] = b[fr["E"]]
        for c in children[f]:
            fc = fronts[c]
            p = np.array([pos[int(g)] for g in fc["rest"]], np.int64)
            y[p] += upd[c]
            upd[c] = None
        z[fr["E"]] = fr["Einv"] @ y[:ne]
        upd[f] = y[ne:] - fr["W"].T @ y[:ne]
    # b of a delayed unknown must be added exactly once, at the front that eliminates it: done above via b[fr["E"]]
    x = z.copy()
    for f in range(nf - 1, -1, -1):
        fr = fronts[f]
        if len(fr["E"]):
            x[fr["E"]] -= fr["W"] @ x[fr["rest"]]
    return x


def main():
    import plfem_b200 as P
    from plfem_b200 import _cabi
    import oracle.fem_oracle as O
    import frontal_reference as FR
    from scipy.sparse.linalg import splu
    which = sys.argv[1] if len(sys.argv) > 1 else "cfg1"
    us = [float(a) for a in sys.argv[2:]] or [0.0, 0.01, 0.1]
    if which == "cfg1":
        g = P.PhotonicLanternGeometry(arrangement="hexagonal_1plus6_7", core_radius_um=1.5, pitch_um=8.0, n_core=1.535,
                                      n_clad=1.0, wavelength_nm=1550)
        mesh, _ = P.MeshGenerator.generate(g)
    elif which == "small":
        g = P.MCFGeometry(3, 6.0, 1.2, 1.53, 1.0, 1.55)
        mesh, _ = P.MeshGenerator.generate(g, refinement=0.4)
    else:
        nx = int(which)
        g = P.MCFGeometry(7, 8.0, 1.5, 1.535, 1.0, 1.55)
        L = 64.0
        xs = np.linspace(-L / 2, L / 2, nx + 1)
        X, Y = np.meshgrid(xs, xs, indexing="xy")
        p = np.vstack([X.ravel(), Y.ravel()])
        idx = np.arange((nx + 1) * (nx + 1)).reshape(nx + 1, nx + 1)
        a, b, c, d = idx[:-1, :-1].ravel(), idx[:-1, 1:].ravel(), idx[1:, :-1].ravel(), idx[1:, 1:].ravel()
        mesh = P.MeshTri(p, np.hstack([np.vstack([a, b, d]), np.vstack([a, d, c])]))
    s = O.interior_system(g, mesh)
    K = (s["A_int"] - O.sigma_estimate(g) * s["B_int"]).tocsr()
    pl = _cabi.Problem(mesh, host_only=True).plan()
    Kp, _ = FR.permuted_operator(K, pl)
    n2 = Kp.shape[0]
    bb = np.random.default_rng(1).standard_normal(n2)
    xr = splu(Kp.tocsc()).solve(bb)
    print(f"{which}: {n2} unknowns, {pl['nfronts']} fronts, {pl['nlevels']} levels", flush=True)
    for u in us:
        t0 = time.time()
        fronts, ch, nd = factor(Kp, pl, u) if u > 0 else (None, None, None)
        if u == 0:
            fr0, ch0 = FR.factor(Kp, pl)
            x = FR.solve(fr0, ch0, pl, bb)
            dx = FR.solve(fr0, ch0, pl, bb - Kp @ x)
            wmax = max((np.abs(f["W"]).max() if f["W"].size else 0) for f in fr0)
            extra = ""
        else:
            x = solve(fronts, ch, pl, bb)
            dx = solve(fronts, ch, pl, bb - Kp @ x)
            wmax = max((np.abs(f["W"]).max() if f["W"].size else 0) for f in fronts)
            lev = pl["level"]
            per_level = [int(nd[lev == l].sum()) for l in range(pl["nlevels"])]
            grow = max(f["nfs"] for f in fronts)
            ent = sum(f["Einv"].size + f["W"].size for f in fronts)
            extra = f" delayed {int(nd.sum())} (fronts with delays {int((nd > 0).sum())}, max per front {int(nd.max())}), largest pivot block {grow}, entries {ent / 1e6:.2f}M, per level {per_level}"
        e0 = np.linalg.norm(x - xr) / np.linalg.norm(xr)
        rho = np.linalg.norm(dx) / np.linalg.norm(x)
        e1 = np.linalg.norm(x + dx - xr) / np.linalg.norm(xr)
        print(f"u={u}: raw err {e0:.2e}, rho {rho:.2e}, refined err {e1:.2e}, max|W| {wmax:.1e}{extra}  ({time.time() - t0:.0f} s)", flush=True)


if __name__ == "__main__":
    main()

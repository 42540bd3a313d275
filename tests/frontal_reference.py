"""NumPy executor of a front plan — test infrastructure.

Runs the same block-LDL^T multifrontal factorisation and sweeps the CUDA kernels in
``csrc/factor.cu`` run, front by front with dense NumPy, from the plan exported by
``plfem_plan_export``.  Used to check the host-built plan (ordering, update sets,
child->parent maps) without a GPU: if the plan is right, solving with it reproduces
``scipy.sparse.linalg.spsolve``.
"""
import numpy as np


def permuted_operator(K_int, plan):
    """K_int is 2n x 2n in reference ordering [x-block, y-block]; returns it in the plan's
    node-interleaved nested-dissection ordering as CSR."""
    n = plan["n"]
    perm = plan["perm"].astype(np.int64)
    idx = np.empty(2 * n, dtype=np.int64)
    idx[0::2] = perm
    idx[1::2] = perm + n
    return K_int[idx, :][:, idx].tocsr(), idx


def gauss_jordan_inverse(F, symmetrise=False):
    """The algorithm of `invert_kernel` (factor.cu) in NumPy: in-place Gauss-Jordan with partial pivoting, the row swaps
    undone at write-back through the column permutation `src` (entry (i, j) of the inverse = a[i, src[j]]).  With
    ``symmetrise`` the write-back averages entry (i, j) with entry (j, i) = a[j, src[i]] — the one-line change of
    DESIGN.md 4.4a."""
    m = F.shape[0]
    a = np.array(F, dtype=float)
    piv = np.zeros(m, dtype=np.int64)
    for k in range(m):
        p = k + int(np.argmax(np.abs(a[k:, k])))
        piv[k] = p
        inv = 1.0 / a[p, k]
        if p != k:
            a[[k, p], :] = a[[p, k], :]
        rowk = a[k, :] * inv
        colk = a[:, k].copy()
        colk[k] = 0.0
        a -= np.outer(colk, rowk)
        a[:, k] = -colk * inv
        a[k, :] = rowk
        a[k, k] = inv
    src = np.arange(m)
    for k in range(m - 1, -1, -1):
        src[k], src[piv[k]] = src[piv[k]], src[k]
    out = a[:, src]
    return 0.5 * (out + out.T) if symmetrise else out


def factor(Kp, plan, inverse=np.linalg.inv):
    first, s, sptr, strct = plan["first"], plan["s"], plan["sptr"], plan["strct"]
    parent, cmap_ptr, cmap = plan["parent"], plan["cmap_ptr"], plan["cmap"]
    nf = plan["nfronts"]
    fronts = [None] * nf
    children = [[] for _ in range(nf)]
    for f in range(nf):
        if parent[f] >= 0:
            children[parent[f]].append(f)
    for f in range(nf):                     # post-order: children first
        own = np.arange(first[f], first[f] + s[f])
        st = strct[sptr[f]:sptr[f + 1]].astype(np.int64)
        nodes = np.concatenate([own, st])
        unk = np.empty(2 * len(nodes), dtype=np.int64)
        unk[0::2] = 2 * nodes
        unk[1::2] = 2 * nodes + 1
        s2 = 2 * s[f]
        F = np.zeros((len(unk), len(unk)))
        F[:s2, :] = Kp[unk[:s2], :][:, unk].toarray()
        for c in children[f]:
            cm = cmap[cmap_ptr[c]:cmap_ptr[c + 1]].astype(np.int64)
            pos = np.empty(2 * len(cm), dtype=np.int64)
            pos[0::2] = 2 * cm
            pos[1::2] = 2 * cm + 1
            F[np.ix_(pos, pos)] += fronts[c]["S"]
        F11inv = inverse(F[:s2, :s2])
        W = F11inv @ F[:s2, s2:]
        S = F[s2:, s2:] - F[:s2, s2:].T @ W
        fronts[f] = dict(F11inv=F11inv, W=W, S=S, unk=unk, s2=s2)
    return fronts, children


def solve(fronts, children, plan, b):
    nf = plan["nfronts"]
    cmap_ptr, cmap = plan["cmap_ptr"], plan["cmap"]
    z = np.zeros_like(b)
    upd = [None] * nf
    for f in range(nf):
        fr = fronts[f]
        s2 = fr["s2"]
        y = np.zeros(len(fr["unk"]))
        y[:s2] = b[fr["unk"][:s2]]
        for c in children[f]:
            cm = cmap[cmap_ptr[c]:cmap_ptr[c + 1]].astype(np.int64)
            pos = np.empty(2 * len(cm), dtype=np.int64)
            pos[0::2] = 2 * cm
            pos[1::2] = 2 * cm + 1
            y[pos] += upd[c]
        z[fr["unk"][:s2]] = fr["F11inv"] @ y[:s2]
        upd[f] = y[s2:] - fr["W"].T @ y[:s2]
    x = z.copy()
    for f in range(nf - 1, -1, -1):
        fr = fronts[f]
        s2 = fr["s2"]
        x[fr["unk"][:s2]] -= fr["W"] @ x[fr["unk"][s2:]]
    return x

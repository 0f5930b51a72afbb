"""CPU experiment (numpy / scipy, no GPU): how many Jacobi-PCG iterations a block-constant deflation space saves on
the Kirchhoff problem of one square mixed site/bond realization near the threshold.  Deflated PCG after
Saad, Yeung, Erhel, Guyomarc'h (SIAM J. Sci. Comput. 21, 2000): x0 = Z E^-1 Z^T b, p = z - Z E^-1 (A Z)^T z + beta p with
E = Z^T A Z, Z = indicator vectors of bs x bs blocks of unknowns.  usage: deflation_experiment.py L pb tol bs1,bs2,...
Results of round 1 (ps 0.80, pb 0.70, tol 1e-10): L = 256: x1.9 / x2.6 / x4.1 fewer iterations with 32 / 16 / 8-site blocks;
L = 1024: x3.3 / x4.3 / x5.7 with 64 / 32 / 16-site blocks (coarse dimension 256 / 1024 / 4096); L = 2048: 7662 / 5792
iterations with 128 / 64-site blocks (plain: about 35 000); G unchanged to ~1e-9."""
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla, scipy.sparse.csgraph as csg, sys, time
def build(L, ps, pb, seed):
    rng = np.random.default_rng(seed)
    m = n = L; t = m*n
    site = rng.random(t) < ps
    idx = np.arange(t).reshape(n, m)
    e_a = idx[:, :-1].ravel(); e_b = idx[:, 1:].ravel()
    n_a = idx[:-1, :].ravel(); n_b = idx[1:, :].ravel()
    a = np.concatenate([e_a, n_a]); b = np.concatenate([e_b, n_b])
    occ = (rng.random(a.size) < pb) & site[a] & site[b]
    g = sp.coo_matrix((np.ones(occ.sum()), (a[occ], b[occ])), shape=(t, t))
    nc, lab = csg.connected_components(g, directed=False)
    bot = set(lab[idx[0][site[idx[0]]]]); top = set(lab[idx[-1][site[idx[-1]]]])
    span = sorted(bot & top)
    if not span: return None
    cid = min(span, key=lambda c: np.flatnonzero(lab == c)[0])
    incl = (lab == cid)
    cond = occ & incl[a]
    w = np.where(cond, 1.0, 1e-12)
    W = sp.coo_matrix((np.concatenate([w, w]), (np.concatenate([a, b]), np.concatenate([b, a]))), shape=(t, t)).tocsr()
    d = np.asarray(W.sum(axis=1)).ravel()
    A = (sp.diags(d) - W).tocsr()
    inter = np.arange(m, t - m)
    Aii = A[inter][:, inter].tocsr()
    V = np.zeros(t); V[t-m:] = 1.0
    rhs = -(A[inter] @ V)
    return Aii, rhs, d[inter], A, m, n, incl[inter]
def readout(A, x, m, n):
    t = m*n; V = np.zeros(t); V[m:t-m] = x; V[t-m:] = 1.0
    I = A @ V
    return I[t-m:].sum()
def pcg(A, b, d, tol, itmax):
    x = np.zeros_like(b); r = b.copy(); bnrm = np.linalg.norm(b/d)
    z = r/d; p = np.zeros_like(b); bknum = r@z; bk = 0.0
    for it in range(1, itmax+1):
        p = z + bk*p; q = A@p; ak = bknum/(p@q); x += ak*p; r -= ak*q; z = r/d
        new = r@z; bk = new/bknum; bknum = new
        if np.linalg.norm(r)/bnrm <= tol: break
    return x, it
def dpcg(A, b, d, Z, tol, itmax):
    AZ = (A @ Z).tocsc(); E = (Z.T @ AZ).tocsc(); lu = spla.splu(E)
    bnrm = np.linalg.norm(b/d)
    x = Z @ lu.solve(Z.T @ b); r = b - A @ x
    z = r/d; p = z - Z @ lu.solve(AZ.T @ z); rz = r@z
    for it in range(1, itmax+1):
        q = A@p; ak = rz/(p@q); x += ak*p; r -= ak*q; z = r/d
        new = r@z; bk = new/rz; rz = new
        if np.linalg.norm(r)/bnrm <= tol: break
        p = z - Z @ lu.solve(AZ.T @ z) + bk*p
    return x, it
L = int(sys.argv[1]); pb = float(sys.argv[2]); tol = float(sys.argv[3]); blocks = [int(v) for v in sys.argv[4].split(',')]
B = build(L, 0.8, pb, 3)
Aii, rhs, d, A, m, n, incl = B
t0=time.time(); x0, it0 = pcg(Aii, rhs, d, tol, 2000000); g0 = readout(A, x0, m, n)
print(f"L={L} plain Jacobi-PCG: it={it0} G={g0:.12e} ({time.time()-t0:.0f}s)", flush=True)
N = Aii.shape[0]
yy, xx = np.divmod(np.arange(m, m*n - m), m)
for bs in blocks:
    for split in (0, 1):
        blk = (yy // bs) * ((m + bs - 1)//bs) + (xx // bs)
        col = blk * 2 + (incl.astype(int) if split else 0)
        u, col = np.unique(col, return_inverse=True)
        Z = sp.csc_matrix((np.ones(N), (np.arange(N), col)), shape=(N, len(u)))
        t0=time.time(); x1, it1 = dpcg(Aii, rhs, d, Z, tol, 2000000); g1 = readout(A, x1, m, n)
        print(f"  deflated, {bs}x{bs} blocks{' split cluster/rest' if split else ''}: coarse dim {Z.shape[1]}  it={it1}  (x{it0/it1:.2f} fewer)  relG={abs(g1-g0)/g0:.1e} ({time.time()-t0:.0f}s)", flush=True)

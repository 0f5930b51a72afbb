"""CPU check (numpy / scipy) of the DEFLATED one-pass recurrences proposed in DESIGN.md section 7 (4): state u = D^-1 r and
s = A p, per iteration ONE sweep (s = A u - (A Z) mu + beta s, u -= alpha D^-1 s, then w' = A u', sums r'.u', r'.r', u'.w'
and the block sums Z^T w') and one global stage (mu' = E^-1 Z^T w', delta~ = delta - mu'.Z^T w', beta' = gamma'/gamma,
alpha' = gamma' / (delta~ - beta' gamma'/alpha)).  Compared with plain Jacobi-PCG and with deflated PCG in its textbook form.
usage: deflation_onepass_experiment.py L pb tol blocksize.   Round 1: L = 256, 32-site blocks: 2581 iterations for both
deflated forms (plain 4922), G equal to 5e-11 relative."""
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla, sys
import os
exec(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'deflation_experiment.py')).read().split("L = int(sys.argv[1])")[0])
def dcgcg(A, b, d, Z, tol, itmax):
    """deflated PCG in the one-pass arrangement: state u = D^-1 r and s = A p; one global stage per iteration"""
    inv = 1.0/d
    AZ = (A @ Z).tocsc(); E = (Z.T @ AZ).tocsc(); lu = spla.splu(E)
    bnrm = np.linalg.norm(b*inv)
    nu = lu.solve(Z.T @ b); x = Z @ nu
    r = b - AZ @ nu
    u = r*inv
    w = A @ u; c = Z.T @ w                      # "prime" sweep
    gam = (d*u) @ u; delta = u @ w
    mu = lu.solve(c); dt = delta - mu @ c
    alpha = gam/dt; beta = 0.0
    s = np.zeros_like(b); p = np.zeros_like(b)
    for it in range(1, itmax+1):
        # ---- one sweep
        w = A @ u
        s = w - AZ @ mu + beta*s
        p = u - Z @ mu + beta*p                 # (only the read-out rows on the GPU)
        x = x + alpha*p
        u = u - alpha*(s*inv)
        r = d*u
        w2 = A @ u
        gam_new = r @ u; rr = r @ r; delta = u @ w2; c = Z.T @ w2
        # ---- global stage
        err = np.sqrt(rr)/bnrm
        if err <= tol: break
        mu = lu.solve(c)
        dt = delta - mu @ c
        beta = gam_new/gam
        alpha = gam_new/(dt - beta*gam_new/alpha)
        gam = gam_new
    return x, it
L = int(sys.argv[1]); pb = float(sys.argv[2]); tol = float(sys.argv[3]); bs = int(sys.argv[4])
Aii, rhs, d, A, m, n, incl = build(L, 0.8, pb, 3)
N = Aii.shape[0]
yy, xx = np.divmod(np.arange(m, m*n - m), m)
blk = (yy // bs) * ((m + bs - 1)//bs) + (xx // bs)
u_, col = np.unique(blk, return_inverse=True)
Z = sp.csc_matrix((np.ones(N), (np.arange(N), col)), shape=(N, len(u_)))
x0, it0 = pcg(Aii, rhs, d, tol, 10**6); g0 = readout(A, x0, m, n)
x1, it1 = dpcg(Aii, rhs, d, Z, tol, 10**6); g1 = readout(A, x1, m, n)
x2, it2 = dcgcg(Aii, rhs, d, Z, tol, 10**6); g2 = readout(A, x2, m, n)
print(f"L={L} bs={bs} coarse={Z.shape[1]}: plain it={it0} G={g0:.12e} | deflated PCG it={it1} G={g1:.12e} | deflated one-pass it={it2} G={g2:.12e} relG(one-pass vs plain)={abs(g2-g0)/g0:.1e}")

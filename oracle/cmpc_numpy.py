"""ORACLE — TEST INFRASTRUCTURE ONLY.

Independent numpy restatements used to cross-check (a) the C++ oracle's
condensation (cmpc_oracle.cpp, which has no reference-run pin) and (b) the
algorithm the CUDA kernels implement, at sizes numpy finishes in seconds.

  condense_dense()   SolverMPC.cpp:566-816 with scipy's expm on the 31x31
                     augmented matrix and dense B_qp^T S B_qp products (the
                     reference's formulation, fp64).
  condense_closed()  the closed form the CUDA path evaluates: the continuous
                     matrix is nilpotent (A^3 = 0), so exp() is a cubic
                     polynomial and H, g reduce to 3x3 foot-pair blocks times
                     scalar sums over the horizon.
  gi_solve()         Goldfarb-Idnani dual active set in range-space form with
                     an explicit H^-1 and an explicit inverse of the active
                     Schur complement -- the per-instance algorithm of the
                     CUDA QP kernel, statement for statement.
"""
import numpy as np


def quat_to_rot(q):
    w, x, y, z = [float(v) for v in q]
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def quat_to_rpy_ref(q):
    """SolverMPC.cpp:352-361; returns (rpy0, rpy1, rpy2) = (yaw, pitch, roll)."""
    w, x, y, z = [float(v) for v in q]
    as_ = min(-2.0 * (x * z - w * y), 0.99999)
    return np.array([np.arctan2(2 * (x * y + w * z), w * w + x * x - y * y - z * z), np.arcsin(as_),
                     np.arctan2(2 * (y * z + w * x), w * w - x * x - y * y + z * z)])


def skew(r):
    return np.array([[0, -r[2], r[1]], [r[2], 0, -r[0]], [-r[1], r[0], 0]], dtype=np.float64)


def _unpack(inst, i):
    f = lambda k: np.asarray(inst[k][i], dtype=np.float64)
    return f("p"), f("v"), f("q"), f("w"), f("r").reshape(3, 4), f("weights"), f("traj"), \
        float(inst["alpha"][i]), np.asarray(inst["gait"][i]), float(inst["x_drag"][i])


def condense_dense(inst, i, f_dist=None, mass=12.0, inertia=(0.07, 0.26, 0.242)):
    from scipy.linalg import expm
    h, dt = inst["horizon"], float(np.float32(inst["dt"]))
    p, v, q, w, rf, wt, traj, alpha, gait, xd = _unpack(inst, i)
    R = quat_to_rot(q)
    rpy = quat_to_rpy_ref(q)
    x0 = np.concatenate([[rpy[2], rpy[1], rpy[0]], p, w, v, [float(np.float32(-9.8))]])
    Iw = R @ np.diag(np.asarray(inertia, dtype=np.float32).astype(np.float64)) @ R.T
    Iinv = np.linalg.inv(Iw)
    M = np.zeros((31, 31))
    M[3, 9] = M[4, 10] = M[5, 11] = 1
    M[11, 9] = xd
    M[11, 12] = 1
    M[0:3, 6:9] = R.T
    for b in range(4):
        M[6:9, 13 + 3 * b:16 + 3 * b] = Iinv @ skew(rf[:, b])
        M[9:12, 13 + 3 * b:16 + 3 * b] = np.eye(3) / float(np.float32(mass))
    M[6:12, 25:31] = np.eye(6)
    E = expm(M * dt)
    Adt, Bdt, Qdt = E[:13, :13], E[:13, 13:25], E[:13, 25:31]
    pw = [np.eye(13)]
    for _ in range(h):
        pw.append(Adt @ pw[-1])
    Aqp = np.vstack([pw[r + 1] for r in range(h)])
    Bqp = np.zeros((13 * h, 12 * h))
    Qqp = np.zeros((13 * h, 6))
    for r in range(h):
        for c in range(r + 1):
            Bqp[13 * r:13 * r + 13, 12 * c:12 * c + 12] = pw[r - c] @ Bdt
            Qqp[13 * r:13 * r + 13] += pw[r - c] @ Qdt
    S = np.tile(np.concatenate([wt, [0.0]]), h)
    Xd = np.zeros(13 * h)
    for k in range(h):
        Xd[13 * k:13 * k + 12] = traj[12 * k:12 * k + 12]
    fd = np.zeros(6) if f_dist is None else np.asarray(f_dist, dtype=np.float64)
    H = 2 * (Bqp.T @ (S[:, None] * Bqp) + alpha * np.eye(12 * h))
    g = 2 * Bqp.T @ (S * (Aqp @ x0 + Qqp @ fd - Xd))
    return H, g, dict(Adt=Adt, Bdt=Bdt, Qdt=Qdt, x0=x0)


def condense_closed(inst, i, f_dist=None, mass=12.0, inertia=(0.07, 0.26, 0.242)):
    """Closed-form H (full 12h x 12h) and g; mirrors csrc/cmpc_kernels.cu."""
    h, dt = inst["horizon"], float(np.float32(inst["dt"]))
    p, v, q, w, rf, wt, traj, alpha, gait, xd = _unpack(inst, i)
    m = float(np.float32(mass))
    R = quat_to_rot(q)
    rpy = quat_to_rpy_ref(q)
    Iw = R @ np.diag(np.asarray(inertia, dtype=np.float32).astype(np.float64)) @ R.T
    Iinv = np.linalg.inv(Iw)
    W = [Iinv @ skew(rf[:, b]) for b in range(4)]
    RW = [R.T @ Wb for Wb in W]
    St, Sp, So, Sv = wt[0:3], wt[3:6], wt[6:9], wt[9:12]
    k = np.arange(h, dtype=np.float64)
    tau = k * dt
    c1 = np.full(h, dt)
    c2 = tau * dt + dt * dt / 2
    c3 = tau * tau * dt / 2 + tau * dt * dt / 2 + dt ** 3 / 6

    def sig(ca, cb):
        s = np.zeros((h, h))
        for a in range(h):
            for b in range(h):
                for r in range(max(a, b), h):
                    s[a, b] += ca[r - a] * cb[r - b]
        return s
    s11, s22, s23, s32, s33, s12, s21 = sig(c1, c1), sig(c2, c2), sig(c2, c3), sig(c3, c2), sig(c3, c3), \
        sig(c1, c2), sig(c2, c1)
    ex, ez = np.array([1.0, 0, 0]), np.array([0, 0, 1.0])
    H = np.zeros((12 * h, 12 * h))
    for a in range(h):
        for b in range(h):
            pv = (s22[a, b] * np.diag(Sp) + xd * Sp[2] * (s23[a, b] * np.outer(ez, ex) + s32[a, b] * np.outer(ex, ez))
                  + xd * xd * Sp[2] * s33[a, b] * np.outer(ex, ex)
                  + s11[a, b] * np.diag(Sv) + xd * Sv[2] * (s12[a, b] * np.outer(ez, ex) + s21[a, b] * np.outer(ex, ez))
                  + xd * xd * Sv[2] * s22[a, b] * np.outer(ex, ex)) / (m * m)
            for fi in range(4):
                for fj in range(4):
                    blk = s22[a, b] * (RW[fi].T @ (St[:, None] * RW[fj])) + s11[a, b] * (W[fi].T @ (So[:, None] * W[fj])) + pv
                    H[12 * a + 3 * fi:12 * a + 3 * fi + 3, 12 * b + 3 * fj:12 * b + 3 * fj + 3] = 2 * blk
    H += 2 * alpha * np.eye(12 * h)
    # free response + disturbance, then weighted error per step
    g0 = float(np.float32(-9.8))
    x0 = np.concatenate([[rpy[2], rpy[1], rpy[0]], p, w, v, [g0]])
    fd = np.zeros(6) if f_dist is None else np.asarray(f_dist, dtype=np.float64)
    e = np.zeros((h, 12))
    for r in range(h):
        T = (r + 1) * dt
        th = x0[0:3] + T * (R.T @ x0[6:9]) + (T * T / 2) * (R.T @ fd[0:3])
        om = x0[6:9] + T * fd[0:3]
        vv = x0[9:12].copy()
        pp = x0[3:6] + T * x0[9:12] + (T * T / 2) * fd[3:6]
        vv = vv + T * fd[3:6]
        # gravity and x_drag couple into z only (A[11,9]=x_drag, A[11,12]=1)
        vv[2] += T * (xd * x0[9] + g0) + (T * T / 2) * xd * fd[3]
        pp[2] += (T * T / 2) * (xd * x0[9] + g0) + (T ** 3 / 6) * xd * fd[3]
        xr = np.concatenate([th, pp, om, vv])
        e[r] = wt * (xr - traj[12 * r:12 * r + 12])
    g = np.zeros(12 * h)
    for c in range(h):
        aT = np.zeros(3); aO = np.zeros(3); aP = np.zeros(3); aV = np.zeros(3); azx = 0.0
        for r in range(c, h):
            kk = r - c
            aT += c2[kk] * e[r, 0:3]
            aO += c1[kk] * e[r, 6:9]
            aP += c2[kk] * e[r, 3:6]
            aV += c1[kk] * e[r, 9:12]
            azx += c3[kk] * e[r, 5] + c2[kk] * e[r, 11]
        for fi in range(4):
            gi = RW[fi].T @ aT + W[fi].T @ aO + (aP + aV) / m + (xd / m) * azx * ex
            g[12 * c + 3 * fi:12 * c + 3 * fi + 3] = 2 * gi
    return H, g


def contact_vars(gait, h):
    """Indices (into 12h) of the variables the reference keeps, SolverMPC.cpp:859-894."""
    keep = []
    for k in range(4 * h):
        if gait[k]:
            keep += [3 * k, 3 * k + 1, 3 * k + 2]
    return np.array(keep, dtype=np.int64)


# Constraint numbering of the reduced problem, per contact foot-step j (x = fx,fy,fz):
#   5j+0:  fx/mu + fz >= 0     5j+1: -fx/mu + fz >= 0
#   5j+2:  fy/mu + fz >= 0     5j+3: -fy/mu + fz >= 0
#   5j+4:  fz <= f_max  (its lower side fz >= 0 is implied by rows 0 and 1)
def gi_solve(H, g, mu, f_max, max_iter=400, tol=1e-9):
    n = len(g)
    nc = n // 3
    mu_inv = float(np.float32(1.0) / np.float32(mu))
    K = np.linalg.inv(H)
    K = 0.5 * (K + K.T)
    x = -K @ g
    m = 5 * nc
    # sparse normals: s_c(x) = n_c . x - b_c >= 0
    nrm = np.zeros((m, n))
    b = np.zeros(m)
    for j in range(nc):
        nrm[5 * j + 0, 3 * j + 0] = mu_inv;  nrm[5 * j + 0, 3 * j + 2] = 1
        nrm[5 * j + 1, 3 * j + 0] = -mu_inv; nrm[5 * j + 1, 3 * j + 2] = 1
        nrm[5 * j + 2, 3 * j + 1] = mu_inv;  nrm[5 * j + 2, 3 * j + 2] = 1
        nrm[5 * j + 3, 3 * j + 1] = -mu_inv; nrm[5 * j + 3, 3 * j + 2] = 1
        nrm[5 * j + 4, 3 * j + 2] = -1;      b[5 * j + 4] = -f_max
    act = []            # active constraint ids
    u = np.zeros(0)     # their multipliers
    P = np.zeros((0, 0))  # inverse of N^T K N
    it = 0
    status = 0
    while True:
        s = nrm @ x - b
        s_chk = s.copy()
        s_chk[act] = 0.0
        p = int(np.argmin(s_chk))
        if s_chk[p] >= -tol:
            break
        up = 0.0
        while True:
            it += 1
            if it > max_iter:
                status = 2
                break
            np_ = nrm[p]
            kn = K @ np_
            if act:
                N = nrm[act].T                      # n x q
                d = N.T @ kn                        # q
                r = P @ d
                z = kn - K @ (N @ r)
                rho2 = float(np_ @ kn - d @ r)
            else:
                d = np.zeros(0); r = np.zeros(0); z = kn; rho2 = float(np_ @ kn)
            scale = float(np_ @ kn)
            dependent = rho2 <= 1e-12 * scale
            # dual step length: largest t keeping all multipliers >= 0
            t1, kdrop = np.inf, -1
            for k in range(len(act)):
                if r[k] > 0 and u[k] / r[k] < t1:
                    t1, kdrop = u[k] / r[k], k
            t2 = np.inf if dependent else -(np_ @ x - b[p]) / rho2
            t = min(t1, t2)
            if not np.isfinite(t):
                status = 3  # infeasible
                break
            if not dependent:
                x = x + t * z
            u = u - t * r
            up += t
            if t == t2:
                # full step: constraint p becomes active; border the inverse Schur complement
                q = len(act)
                Pn = np.zeros((q + 1, q + 1))
                if q:
                    Pn[:q, :q] = P + np.outer(r, r) / rho2
                    Pn[:q, q] = -r / rho2
                    Pn[q, :q] = -r / rho2
                Pn[q, q] = 1.0 / rho2
                P = Pn
                act.append(p)
                u = np.append(u, up)
                break
            # partial step: drop constraint kdrop, keep working on p
            q = len(act)
            keep = [k for k in range(q) if k != kdrop]
            pk = P[keep, kdrop]
            P = P[np.ix_(keep, keep)] - np.outer(pk, pk) / P[kdrop, kdrop]
            act.pop(kdrop)
            u = np.delete(u, kdrop)
        if status:
            break
    lam = np.zeros(m)
    lam[act] = u
    obj = 0.5 * x @ H @ x + g @ x
    return x, dict(iters=it, active=sorted(act), lam=lam, status=status, objective=obj, n_active=len(act))


def active_mask(x, mu, f_max, tol=1e-6):
    """Primal activity of the reference's 5 rows per contact foot-step
    (fmat rows, SolverMPC.cpp:660): -1 at the lower bound, +1 at the upper, 0 inactive."""
    mu_inv = float(np.float32(1.0) / np.float32(mu))
    xs = np.asarray(x).reshape(-1, 3)
    rows = np.stack([xs[:, 0] * mu_inv + xs[:, 2], -xs[:, 0] * mu_inv + xs[:, 2],
                     xs[:, 1] * mu_inv + xs[:, 2], -xs[:, 1] * mu_inv + xs[:, 2], xs[:, 2]], -1)
    mask = np.zeros(rows.shape, dtype=np.int8)
    mask[rows <= tol] = -1
    mask[:, 4][rows[:, 4] >= f_max - tol] = 1
    return mask

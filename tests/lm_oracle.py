"""TEST INFRASTRUCTURE — independent dense-numpy restatement of the Ceres trust-region LM loop that
EventCalibSpline::optimize runs (src/EventCalibSpline.cpp:197-247), driven by the oracle's normal equations.
[external: Ceres 1.x TrustRegionMinimizer / LevenbergMarquardtStrategy, SURVEY.md Appendix C — parity unpinned]
Parameter order here is intrinsics-first and the solve is dense LAPACK, i.e. deliberately NOT the product's
banded-arrowhead code path."""
import numpy as np

import oracle


def _index_map(n_cp_list):
    maps, co = [], 0
    for n in n_cp_list:
        for k in range(n - 3):
            cp0 = co + k
            m = list(range(9))
            for j in range(4):
                m += [9 + 6 * (cp0 + j) + a for a in range(3)]
            for j in range(4):
                m += [9 + 6 * (cp0 + j) + 3 + a for a in range(3)]
            # local order: 9 intr | 4x3 rot | 4x3 trans
            loc = list(range(9)) + [9 + 6 * (cp0 + j) + a for j in range(4) for a in range(3)] + \
                  [9 + 6 * (cp0 + j) + 3 + a for j in range(4) for a in range(3)]
            maps.append(np.array(loc))
        co += n
    return maps, co


def assemble(H, g, maps, D):
    A = np.zeros((D, D))
    b = np.zeros(D)
    for s, m in enumerate(maps):
        A[np.ix_(m, m)] += H[s]
        b[m] += g[s]
    return A, b


SO3 = [False]   # set by solve(so3=...): LocalParameterizationSO3::Plus (T * exp(delta)) instead of the quaternion Plus


def plus(intr, rot, trans, d):
    C = len(rot)
    ni = intr + d[:9]
    nr = np.zeros_like(rot)
    nt = np.zeros_like(trans)
    for c in range(C):
        nr[c] = (oracle.so3_plus if SO3[0] else oracle.quat_plus)(rot[c], d[9 + 6 * c: 9 + 6 * c + 3])
        nt[c] = trans[c] + d[9 + 6 * c + 3: 9 + 6 * c + 6]
    return ni, nr, nt


def solve(P, n_cp_list, intr, rot, trans, max_iterations=50, ftol=1e-10, gtol=1e-10, ptol=1e-8, fixed=False, so3=False):
    SO3[0] = bool(so3)
    maps, C = _index_map(n_cp_list)
    D = 9 + 6 * C
    intr, rot, trans = np.array(intr, float), np.array(rot, float).reshape(C, 4), np.array(trans, float).reshape(C, 3)
    cost, H, g = P.normal_eq(intr, rot, trans)
    A, b = assemble(H, g, maps, D)
    scale = 1.0 / (1.0 + np.sqrt(np.diag(A)))
    radius, dec = 1e4, 2.0
    trace = []

    def gmax():
        pi, pr, pt = plus(intr, rot, trans, -b)
        return max(np.abs(intr - pi).max(), np.abs(rot - pr).max(), np.abs(trans - pt).max())

    trace.append((cost, gmax(), radius, 1.0))
    term = "running"
    if not fixed and trace[-1][1] <= gtol:
        return intr, rot, trans, trace, "gradient"
    it, reuse, diag = 0, False, None
    while True:
        if it >= max_iterations:
            term = "no_convergence"
            break
        it += 1
        As = A * scale[:, None] * scale[None, :]
        gs = b * scale
        if not reuse:
            diag = np.clip(np.diag(As), 1e-6, 1e32)
        M = As + np.diag(diag / radius)
        try:
            L = np.linalg.cholesky(M)
            y = np.linalg.solve(L.T, np.linalg.solve(L, gs))
            step = -y
            mcc = -(step @ gs + 0.5 * step @ (As @ step))
            ok = mcc > 0 and np.isfinite(mcc)
        except np.linalg.LinAlgError:
            ok = False
        if not ok:
            radius /= dec
            dec *= 2
            reuse = True
            trace.append((cost, trace[-1][1], radius, -1.0))
            continue
        delta = step * scale
        ci, cr, ct = plus(intr, rot, trans, delta)
        xn = np.sqrt((intr ** 2).sum() + (rot ** 2).sum() + (trans ** 2).sum())
        sn = np.sqrt(((intr - ci) ** 2).sum() + ((rot - cr) ** 2).sum() + ((trans - ct) ** 2).sum())
        cc = P.cost(ci, cr, ct)
        if not fixed:
            if sn <= ptol * (xn + ptol):
                term = "parameter"
                break
            if abs(cost - cc) <= ftol * cost:
                term = "function"
                break
        rho = (cost - cc) / mcc
        if rho > 1e-3:
            intr, rot, trans = ci, cr, ct
            radius = min(1e16, radius / max(1.0 / 3.0, 1.0 - (2 * rho - 1) ** 3))
            dec, reuse = 2.0, False
            cost, H, g = P.normal_eq(intr, rot, trans)
            A, b = assemble(H, g, maps, D)
            trace.append((cost, gmax(), radius, 1.0))
            if not fixed and trace[-1][1] <= gtol:
                term = "gradient"
                break
        else:
            radius /= dec
            dec *= 2
            reuse = True
            trace.append((cost, trace[-1][1], radius, 0.0))
            if radius < 1e-32:
                term = "min_radius"
                break
    return intr, rot, trans, trace, term

"""Host-side B-spline helpers mirroring BsplineReal (core/spline/include/opengv2/spline/BsplineReal.hpp).

Only what the cost-evaluation path needs on the host: knot placement (approximation(), :87-100), span search
(:208-231), the 4 cubic basis values (:107-145) and a plain least-squares control-point fit used to set up
synthetic problems (the reference's own fit, optimization() :329-449, is out of scope — SURVEY.md §2 row 5).
"""
import numpy as np

DEGREE = 3


def knot_vector(us, n_cp):
    us = np.asarray(us, np.float64)
    n = len(us)
    kn = np.empty(n_cp + DEGREE + 1)
    kn[:DEGREE + 1] = us[0]
    kn[-(DEGREE + 1):] = us[-1]
    d = n / float(n_cp - DEGREE)
    for j in range(1, n_cp - DEGREE):
        i = int(np.floor(j * d))
        alpha = j * d - i
        kn[DEGREE + j] = (1 - alpha) * us[i - 1] + alpha * us[i]
    return kn


def find_span(kn, u):
    n = len(kn) - 2 - DEGREE
    if u == kn[n + 1]:
        return n
    low, high = DEGREE, n + 1
    mid = (low + high) // 2
    while u < kn[mid] or u >= kn[mid + 1]:
        if u < kn[mid]:
            high = mid
        else:
            low = mid
        mid = (low + high) // 2
    return mid


def basis(kn, span, u):
    ndu = np.zeros((4, 4))
    left = np.zeros(4)
    right = np.zeros(4)
    ndu[0, 0] = 1
    for j in range(1, DEGREE + 1):
        left[j] = u - kn[span + 1 - j]
        right[j] = kn[span + j] - u
        saved = 0.0
        for r in range(j):
            ndu[j, r] = right[r + 1] + left[j - r]
            temp = ndu[r, j - 1] / ndu[j, r]
            ndu[r, j] = saved + right[r + 1] * temp
            saved = left[j - r] * temp
        ndu[j, j] = saved
    return ndu[:, DEGREE].copy()


def fit_control_points(kn, us, data, n_cp):
    """Least-squares control points (n_cp x dim) for samples data(us) with free end points — a builder of synthetic problems
    for tests / bench.  The reference's own fit (BsplineReal: first / last control point = first / last sample) is
    EventCalibSpline::fitSpline in include/ecb/event_calib.hpp."""
    A = np.zeros((len(us), n_cp))
    for r, u in enumerate(us):
        s = find_span(kn, u)
        A[r, s - 3:s + 1] = basis(kn, s, u)
    cp, *_ = np.linalg.lstsq(A, np.asarray(data, np.float64), rcond=None)
    return cp


def evaluate(kn, cp, u):
    s = find_span(kn, u)
    return basis(kn, s, u) @ cp[s - 3:s + 1]

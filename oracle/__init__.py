"""TEST INFRASTRUCTURE — ctypes access to the CPU oracle.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package.  The product (``eventcalib_b200``) never does.

Libraries (built by ``oracle/Makefile``; see the headers of the .cpp files for what each restates):
  * ``_build/libecb_oracle.so``  — the CPU restatement (always buildable)
  * ``_ref/libref_dbscan.so``    — the UNMODIFIED reference ``dbscan.h`` + ``kdtree.cpp`` compiled in place
  * ``_ref/libref_frontend.so``  — restated glue + verbatim reference DBSCAN (the "reference" CPU baseline)
  * ``_ref/libref_functor.so``   — the UNMODIFIED reference residual functor (EventCalibSpline.hpp), B-spline
                                   (BsplineReal.hpp), event window (EventFrame.cpp + utility.hpp hash), record reader
                                   (Event.hpp), CirclesEventFrame.cpp (extractFeatures / fitCircle / rectifyFeatures /
                                   findCenter, with dbscan.h + kdtree.cpp), EventCalibSpline.cpp (constructor: set-up,
                                   association, Ceres problem assembly into a recording stand-in) and EventCalibIni.cpp
                                   (tracking gate, checkPose, cvCalibration flow) compiled in place
                                   against the stand-in headers of ``shim_functor/``
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PORT = os.path.join(_HERE, "_build", "libecb_oracle.so")
_REF_DB = os.path.join(_HERE, "_ref", "libref_dbscan.so")
_REF_FE = os.path.join(_HERE, "_ref", "libref_frontend.so")
_REF_FN = os.path.join(_HERE, "_ref", "libref_functor.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_up = C.POINTER(C.c_uint)
_bp = C.POINTER(C.c_uint8)
_lp = C.POINTER(C.c_int64)


def build(force=False):
    """Compile the restatement and, when /root/reference is present, oracle/_ref."""
    if force or not os.path.exists(_PORT) or (os.path.isdir("/root/reference") and not (os.path.exists(_REF_DB) and os.path.exists(_REF_FN))):
        subprocess.check_call(["make", "-C", _HERE, "all"], stdout=subprocess.DEVNULL)
    return _PORT


def _p(a, t):
    return a.ctypes.data_as(t)


_libs = {}


def _load(path):
    if path not in _libs:
        if not os.path.exists(path):
            build()
        _libs[path] = C.CDLL(path)
    return _libs[path]


def port():
    lib = _load(_PORT)
    lib.orc_hash_double.restype = C.c_uint64
    lib.orc_hash_double.argtypes = [C.c_double]
    lib.orc_hash_p2.restype = C.c_uint64
    lib.orc_hash_p2.argtypes = [C.c_double, C.c_double]
    lib.orc_radius_threshold.restype = C.c_double
    lib.orc_radius_threshold.argtypes = [C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double]
    lib.orc_frontend_windows.restype = C.c_int64
    return lib


def have_ref():
    return os.path.exists(_REF_DB)


def have_ref_functor():
    return os.path.exists(_REF_FN)


_REF_DROPIN = os.path.join(_HERE, "_ref", "libref_dropin.so")


def have_ref_dropin():
    return os.path.exists(_REF_DROPIN)


def ref_dropin_lib():
    """The reference's sources compiled where they lie with `#include <dbscan.h>` resolved to the PRODUCT's include/ecb/dbscan.h:
    its DBSCAN::Run calls run on the GPU through libecb.so (oracle/Makefile, target _ref/libref_dropin.so)."""
    return C.CDLL(_REF_DROPIN)


def ref_functor_lib():
    lib = _load(_REF_FN)
    lib.ref_residual.restype = C.c_double
    lib.ref_residual_jac.restype = C.c_double
    return lib


def ref_residual_jac(intr, rcp, tcp, obs, lm, radius, b):
    """The reference's own CalibReprojectionError::operator() (EventCalibSpline.hpp:168-229) on Jet<37> and on double:
    returns (value on Jet, 1x37 ambient Jacobian, value on double)."""
    a = [np.ascontiguousarray(v, np.float64) for v in (intr, rcp, tcp, obs, lm, b)]
    jac = np.zeros(37)
    lib = ref_functor_lib()
    r = lib.ref_residual_jac(_p(a[0], _dp), _p(a[1], _dp), _p(a[2], _dp), _p(a[3], _dp), _p(a[4], _dp), C.c_double(radius),
                             _p(a[5], _dp), _p(jac, _dp))
    rd = lib.ref_residual(_p(a[0], _dp), _p(a[1], _dp), _p(a[2], _dp), _p(a[3], _dp), _p(a[4], _dp), C.c_double(radius), _p(a[5], _dp))
    return r, jac, rd


def ref_residual_jac_so3(intr, rcp, tcp, obs, lm, radius, beta3, basis4):
    """The reference's own CalibReprojectionError_SO3::operator() (EventCalibSpline.hpp:65-156) on Jet<37> and on double,
    compiled where it lies against the stand-in Sophus (oracle/shim_functor/sophus/so3.hpp): (value on Jet, 1x37 ambient
    Jacobian, value on double).  beta3: the cumulative rotation basis, basis4: the translation basis."""
    a = [np.ascontiguousarray(v, np.float64) for v in (intr, rcp, tcp, obs, lm, beta3, basis4)]
    jac = np.zeros(37)
    lib = ref_functor_lib()
    lib.ref_residual_jac_so3.restype = C.c_double
    lib.ref_residual_so3.restype = C.c_double
    r = lib.ref_residual_jac_so3(_p(a[0], _dp), _p(a[1], _dp), _p(a[2], _dp), _p(a[3], _dp), _p(a[4], _dp), C.c_double(radius),
                                 _p(a[5], _dp), _p(a[6], _dp), _p(jac, _dp))
    rd = lib.ref_residual_so3(_p(a[0], _dp), _p(a[1], _dp), _p(a[2], _dp), _p(a[3], _dp), _p(a[4], _dp), C.c_double(radius),
                              _p(a[5], _dp), _p(a[6], _dp))
    return r, jac, rd


def ref_so3_basis(knots, u):
    """The reference's BsplineSO3::findSpan + derBasisFuns(u, span, 0) (core/spline/src/BsplineSO3.cpp:73-109, compiled where
    it lies): (span, the 3 cumulative basis values)."""
    kn = np.ascontiguousarray(knots, np.float64)
    span = C.c_int()
    b = np.zeros(3)
    ref_functor_lib().ref_so3_basis(_p(kn, _dp), C.c_int(len(kn)), C.c_double(u), C.byref(span), _p(b, _dp))
    return span.value, b


def ref_so3_plus(x, d):
    """The reference's LocalParameterizationSO3::Plus (BsplineSO3.hpp:196-204): T * exp(delta)."""
    x = np.ascontiguousarray(x, np.float64)
    d = np.ascontiguousarray(d, np.float64)
    out = np.zeros(4)
    ref_functor_lib().ref_so3_plus(_p(x, _dp), _p(d, _dp), _p(out, _dp))
    return out


def ref_so3_plus_jacobian(x):
    """The reference's LocalParameterizationSO3::ComputeJacobian (BsplineSO3.hpp:209-217), 4 x 3."""
    x = np.ascontiguousarray(x, np.float64)
    J = np.zeros(12)
    ref_functor_lib().ref_so3_plus_jacobian(_p(x, _dp), _p(J, _dp))
    return J.reshape(4, 3)


def ref_undistort(intr, obs):
    a = [np.ascontiguousarray(v, np.float64) for v in (intr, obs)]
    out = np.zeros(3)
    ref_functor_lib().ref_undistort(_p(a[0], _dp), _p(a[1], _dp), _p(out, _dp))
    return out


def ref_spline_fit(us, data, n_cp):
    """The reference's BsplineReal<dim>(3, samples, cpNum, timestamps): (knot vector, control points)."""
    us = np.ascontiguousarray(us, np.float64)
    data = np.ascontiguousarray(data, np.float64)
    dim = data.shape[1]
    kn, cp = np.zeros(n_cp + 4), np.zeros((n_cp, dim))
    n = ref_functor_lib().ref_spline_fit(C.c_int(dim), _p(us, _dp), _p(data, _dp), C.c_int(len(us)), C.c_int(n_cp), _p(kn, _dp), _p(cp, _dp))
    return kn, cp, n


def ref_basis(knots, u):
    """The reference's BsplineReal::findSpan + dersBasisFuns(u, span, 0): (span, 4 basis values)."""
    kn = np.ascontiguousarray(knots, np.float64)
    span = C.c_int()
    N = np.zeros(4)
    ref_functor_lib().ref_basis(_p(kn, _dp), C.c_int(len(kn)), C.c_double(u), C.byref(span), _p(N, _dp))
    return span.value, N


def ref_event_frame(t, x, y, pol, t0, t1):
    """The reference's own EventFrame constructor (event/src/EventFrame.cpp:10-36): (positive xy, negative xy) in the
    iteration order of its hash sets."""
    t, x, y = (np.ascontiguousarray(v, np.float64) for v in (t, x, y))
    pol = np.ascontiguousarray(pol, np.uint8)
    n = len(t)
    pos, neg = np.zeros((max(n, 1), 2)), np.zeros((max(n, 1), 2))
    n_pos, n_neg = C.c_longlong(), C.c_longlong()
    ref_functor_lib().ref_event_frame(_p(t, _dp), _p(x, _dp), _p(y, _dp), _p(pol, _bp), C.c_longlong(n), C.c_double(t0),
                                      C.c_double(t1), _p(pos, _dp), _p(neg, _dp), C.byref(n_pos), C.byref(n_neg))
    return pos[:n_pos.value].copy(), neg[:n_neg.value].copy()


def ref_read_bin(path, cap):
    """The reference's own record reader (Event.hpp:41-47 operator>>) over a .bin file."""
    t, x, y, pol = np.zeros(cap), np.zeros(cap), np.zeros(cap), np.zeros(cap, np.uint8)
    lib = ref_functor_lib()
    lib.ref_read_bin.restype = C.c_longlong
    n = lib.ref_read_bin(path.encode(), C.c_longlong(cap), _p(t, _dp), _p(x, _dp), _p(y, _dp), _p(pol, _bp))
    return t[:n], x[:n], y[:n], pol[:n]


def ref_extract(t, x, y, pol, t0, t1, W, H, fitCircle, eps=4.0, minS=2, clusterMin=5, knn_num=3, rows=9, cols=4, asym=True,
                square=5.5, radius=1.75, lib=None):
    """The reference's own CirclesEventFrame constructor + extractFeatures() (event_camera_calib/src/CirclesEventFrame.cpp:16-359,
    with its DBSCAN) on raw events.  Returns dict(found, cand_f32 = the candidate centres handed to findCirclesGrid as
    cv::Point2f (None when :127-129 returned before), features = rows*cols x (cx, cy, r) in board order when found, rthr)."""
    t, x, y = (np.ascontiguousarray(v, np.float64) for v in (t, x, y))
    pol = np.ascontiguousarray(pol, np.uint8)
    prm = np.array([cols, rows, square, 1.0 if asym else 0.0, radius, eps, minS, clusterMin, knn_num, fitCircle], np.float64)
    cap = 1024
    cand = np.zeros((cap, 2), np.float32)
    n_cand, rthr = C.c_int(), C.c_double()
    feats = np.zeros((rows * cols, 3))
    ok = (lib or ref_functor_lib()).ref_extract(_p(t, _dp), _p(x, _dp), _p(y, _dp), _p(pol, _bp), C.c_longlong(len(t)), C.c_double(t0),
                                       C.c_double(t1), C.c_int(W), C.c_int(H), _p(prm, _dp), cand.ctypes.data_as(C.c_void_p),
                                       C.c_int(cap), C.byref(n_cand), _p(feats, _dp), C.byref(rthr))
    return dict(found=bool(ok), cand_f32=cand[:n_cand.value].copy() if n_cand.value >= 0 else None, features=feats, rthr=rthr.value)


def ref_fit_circle(pxy, nxy):
    """The reference's own CirclesEventFrame::fitCircle (CirclesEventFrame.cpp:361-415)."""
    pxy = np.ascontiguousarray(pxy, np.float64).reshape(-1, 2)
    nxy = np.ascontiguousarray(nxy, np.float64).reshape(-1, 2)
    out = np.zeros(3)
    ref_functor_lib().ref_fit_circle(_p(pxy, _dp), C.c_int(len(pxy)), _p(nxy, _dp), C.c_int(len(nxy)), _p(out, _dp))
    return out


def ref_rectify(t, x, y, pol, t0, t1, W, H, fitCircle, image_points, find_xy=None, rows=9, cols=4, lib=None):
    """The reference's own extractFeatures() + rectifyFeatures() (CirclesEventFrame.cpp:417-638) with the caller's projections
    image_points[rows*cols][5][2]; then findCenter() (CirclesEventFrame.hpp:50-65) for the pixels find_xy.
    Returns (verdict: -1 extractFeatures failed / 0 / 1, out[rows*cols][3] with r = -1 for deleted features, landmark ids)."""
    t, x, y = (np.ascontiguousarray(v, np.float64) for v in (t, x, y))
    pol = np.ascontiguousarray(pol, np.uint8)
    img = np.ascontiguousarray(image_points, np.float64)
    prm = np.array([cols, rows, 5.5, 1.0, 1.75, 4.0, 2, 5, 3, fitCircle], np.float64)
    out = np.zeros((rows * cols, 3))
    fxy = np.ascontiguousarray(find_xy if find_xy is not None else np.zeros((0, 2)), np.float64)
    fid = np.full(max(len(fxy), 1), -2, np.int32)
    rc = (lib or ref_functor_lib()).ref_rectify(_p(t, _dp), _p(x, _dp), _p(y, _dp), _p(pol, _bp), C.c_longlong(len(t)), C.c_double(t0),
                                       C.c_double(t1), C.c_int(W), C.c_int(H), _p(prm, _dp), _p(img, _dp), _p(out, _dp), _p(fxy, _dp),
                                       C.c_int(len(fxy)), _p(fid, _ip))
    return rc, out, fid[:len(fxy)]


def ref_calib_spline(t, x, y, pol, kf_t, kf_q, kf_twb, circles, board, cam9, W, H, step, radius, res_cap=None):
    """The reference's own EventCalibSpline constructor (event_camera_calib/src/EventCalibSpline.cpp: reduceMap segmentation,
    spline set-up, intrinsics, association loop, Ceres problem assembly with a recording ceres::Problem and a no-op Solve,
    updateMap) compiled in place.  Returns a dict, or raises RuntimeError with the std::logic_error text."""
    t, x, y = (np.ascontiguousarray(v, np.float64) for v in (t, x, y))
    pol = np.ascontiguousarray(pol, np.uint8)
    kf_t, kf_q, kf_twb = (np.ascontiguousarray(v, np.float64) for v in (kf_t, kf_q, kf_twb))
    circles = np.ascontiguousarray(circles, np.float64)
    board = np.ascontiguousarray(board, np.float64)
    cam9 = np.ascontiguousarray(cam9, np.float64)
    K, n_circ = circles.shape[0], circles.shape[1]
    cap = int(res_cap or len(t))
    info = np.zeros(8, np.int32)
    ncp = np.zeros(K + 1, np.int32)
    knots, rot, trans = np.zeros(K + 64 * 8), np.zeros(4 * (K + 64)), np.zeros(3 * (K + 64))
    ranges, intr, ht = np.zeros((K + 1, 2)), np.zeros(9), np.zeros(3)
    obs, lm, basis = np.zeros((cap, 2)), np.zeros((cap, 3)), np.zeros((cap, 4))
    span, spl, rcp = np.zeros(cap, np.int32), np.zeros(cap, np.int32), np.zeros((cap, 2), np.int32)
    pose = np.zeros((K, 8))
    err = C.create_string_buffer(256)
    rc = ref_functor_lib().ref_calib_spline(
        _p(t, _dp), _p(x, _dp), _p(y, _dp), _p(pol, _bp), C.c_longlong(len(t)), _p(kf_t, _dp), _p(kf_q, _dp), _p(kf_twb, _dp),
        _p(circles, _dp), C.c_int(K), C.c_int(n_circ), _p(board, _dp), _p(cam9, _dp), C.c_int(W), C.c_int(H), C.c_double(step),
        C.c_double(radius), _p(info, _ip), _p(ncp, _ip), _p(knots, _dp), _p(rot, _dp), _p(trans, _dp), _p(ranges, _dp), _p(intr, _dp),
        _p(ht, _dp), C.c_longlong(cap), _p(obs, _dp), _p(lm, _dp), _p(basis, _dp), _p(span, _ip), _p(spl, _ip), _p(rcp, _ip),
        _p(pose, _dp), err, C.c_int(256))
    if rc != 0:
        raise RuntimeError(err.value.decode())
    S, n = int(info[0]), int(info[1])
    ncp = ncp[:S].copy()
    ko = np.concatenate([[0], np.cumsum(ncp + 4)])
    co = np.concatenate([[0], np.cumsum(ncp)])
    return dict(n_splines=S, n_residuals=n, solve_calls=int(info[2]), param_blocks=int(info[3]), quaternion_blocks=int(info[4]),
                linear_solver=int(info[5]), frames_left=int(info[6]), n_cp=ncp,
                knots=[knots[ko[s]:ko[s + 1]].copy() for s in range(S)],
                rot_cp=[rot[4 * co[s]:4 * co[s + 1]].reshape(-1, 4).copy() for s in range(S)],
                trans_cp=[trans[3 * co[s]:3 * co[s + 1]].reshape(-1, 3).copy() for s in range(S)],
                ranges=ranges[:S].copy(), intrinsics=intr, huber=ht[0], gradient_tolerance=ht[1], function_tolerance=ht[2],
                obs=obs[:n], lm=lm[:n], basis=basis[:n], span=span[:n], spline=spl[:n], first_cp=rcp[:n], kf_pose=pose)


class RefIni:
    """The reference's own EventCalibIni (event_camera_calib/src/EventCalibIni.cpp) compiled in place: tracking gate
    (TrackingBase::process -> track), checkPose, and the front-to-back flow per window (CirclesEventFrame + extractFeatures,
    gate, cvCalibration) with the product's host header as the OpenCV hooks."""

    def __init__(self, W, H, step, n_use=200, fitCircle=0, rows=9, cols=4):
        self.lib = ref_functor_lib()
        self.lib.ref_ini_new.restype = C.c_void_p
        self.prm = np.array([cols, rows, 5.5, 1.0, 1.75, 4.0, 2, 5, 3, fitCircle], np.float64)
        self.n_feat = rows * cols
        self.fit = int(fitCircle)
        self.h = C.c_void_p(self.lib.ref_ini_new(C.c_int(W), C.c_int(H), C.c_double(step), _p(self.prm, _dp), C.c_int(n_use)))

    def __del__(self):
        try:
            self.lib.ref_ini_free(self.h)
        except Exception:
            pass

    def add_events(self, t, x, y, pol):
        t, x, y = (np.ascontiguousarray(v, np.float64) for v in (t, x, y))
        pol = np.ascontiguousarray(pol, np.uint8)
        self.lib.ref_ini_add_events(self.h, _p(t, _dp), _p(x, _dp), _p(y, _dp), _p(pol, _bp), C.c_longlong(len(t)))

    def gate(self, stamp, xy):
        xy = np.ascontiguousarray(xy, np.float64)
        return int(self.lib.ref_ini_gate(self.h, C.c_double(stamp), _p(xy, _dp), C.c_int(len(xy))))

    def run(self, windows):
        """-> dict(ok, cam9, status[w] (0 no features / 1 gate rejected / 2 dropped by cvCalibration / 3 kept), pose[w][7] =
        twb + Qwb (x y z w), feat[w][n_feat][3], frames_before, frames_after, calibrate_views, calibrate_flags)"""
        win = np.ascontiguousarray(windows, np.float64)
        nw = len(win)
        cam9, status = np.zeros(9), np.zeros(nw, np.int32)
        pose, feat, counts = np.zeros((nw, 7)), np.zeros((nw, self.n_feat, 3)), np.zeros(4, np.int32)
        ok = self.lib.ref_ini_run(self.h, _p(win, _dp), C.c_int(nw), C.c_int(self.fit), _p(cam9, _dp), _p(status, _ip), _p(pose, _dp),
                                  _p(feat, _dp), _p(counts, _ip))
        return dict(ok=bool(ok), cam9=cam9, status=status, pose=pose, feat=feat, frames_before=int(counts[0]),
                    frames_after=int(counts[1]), calibrate_views=int(counts[2]), calibrate_flags=int(counts[3]))


def ref_check_pose(ref_stamp, ref_q, ref_t, cur_stamp, cur_q, cur_t, step):
    """The reference's own EventCalibIni::checkPose (EventCalibIni.cpp:328-346) against a map whose last key frame is ref."""
    a = [np.ascontiguousarray(v, np.float64) for v in (ref_q, ref_t, cur_q, cur_t)]
    f = ref_functor_lib().ref_check_pose
    f.argtypes = [C.c_double, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_double]
    return int(f(ref_stamp, a[0].ctypes.data, a[1].ctypes.data, cur_stamp, a[2].ctypes.data, a[3].ctypes.data, step))


def ref_dbscan_lib():
    return _load(_REF_DB)


def ref_frontend_lib():
    lib = _load(_REF_FE)
    lib.orc_frontend_windows.restype = C.c_int64
    return lib


def _dbscan(fn, xy, eps, minpts):
    xy = np.ascontiguousarray(xy, dtype=np.float64).reshape(-1, 2)
    n = xy.shape[0]
    labels = np.full(max(n, 1), -1, np.int32)
    members = np.zeros(max(n, 1), np.uint32)
    off = np.zeros(n + 2, np.uint32)
    noise = np.zeros(max(n, 1), np.uint32)
    nc = C.c_int(0)
    nn = C.c_int(0)
    rc = fn(_p(xy, _dp), C.c_int(n), C.c_double(eps), C.c_uint(minpts), _p(labels, _ip), C.byref(nc),
            _p(members, _up), _p(off, _up), _p(noise, _up), C.byref(nn))
    clusters = [members[off[c]:off[c + 1]].copy() for c in range(nc.value)]
    return dict(rc=rc, labels=labels[:n].copy(), clusters=clusters, noise=noise[:nn.value].copy())


def dbscan(xy, eps, minpts):
    """Restated ordered DBSCAN (oracle/ecb_oracle_frontend.cpp)."""
    return _dbscan(port().orc_dbscan_run, xy, eps, minpts)


def ref_dbscan(xy, eps, minpts):
    """The unmodified reference DBSCAN<Vector2d,double>::Run (oracle/_ref)."""
    return _dbscan(ref_dbscan_lib().ref_dbscan_run, xy, eps, minpts)


def ref_dbscan_nd(pts, eps, minpts):
    """The unmodified reference DBSCAN<T,double>::Run(V, dim, ...) on n x dim points, dim = 1..4 (oracle/_ref)."""
    pts = np.ascontiguousarray(pts, dtype=np.float64)
    n, dim = pts.shape
    labels = np.full(max(n, 1), -1, np.int32)
    members = np.zeros(max(n, 1), np.uint32)
    off = np.zeros(n + 2, np.uint32)
    noise = np.zeros(max(n, 1), np.uint32)
    nc = C.c_int(0)
    nn = C.c_int(0)
    rc = ref_dbscan_lib().ref_dbscan_run_nd(_p(pts, _dp), C.c_int(n), C.c_int(dim), C.c_double(eps), C.c_uint(minpts),
                                            _p(labels, _ip), C.byref(nc), _p(members, _up), _p(off, _up), _p(noise, _up),
                                            C.byref(nn))
    clusters = [members[off[c]:off[c + 1]].copy() for c in range(nc.value)]
    return dict(rc=rc, labels=labels[:n].copy(), clusters=clusters, noise=noise[:nn.value].copy())


def kd_range(xy, q, eps, ref=False):
    xy = np.ascontiguousarray(xy, dtype=np.float64).reshape(-1, 2)
    out = np.zeros(xy.shape[0], np.uint32)
    fn = ref_dbscan_lib().ref_kd_range if ref else port().orc_kd_range
    k = fn(_p(xy, _dp), C.c_int(xy.shape[0]), C.c_int(q), C.c_double(eps), _p(out, _up))
    return out[:k].copy()


def kd_flags(xy):
    xy = np.ascontiguousarray(xy, dtype=np.float64).reshape(-1, 2)
    fl = np.zeros(max(xy.shape[0], 1), np.uint8)
    port().orc_kd_flags(_p(xy, _dp), C.c_int(xy.shape[0]), _p(fl, _bp))
    return fl[:xy.shape[0]]


def event_frame(t, x, y, pol, t0, t1):
    """EventFrame ctor: window [t0,t1] closed, per-pixel dedupe, +/- cancel, libstdc++ hash-set order."""
    t = np.ascontiguousarray(t, np.float64)
    x = np.ascontiguousarray(x, np.float64)
    y = np.ascontiguousarray(y, np.float64)
    pol = np.ascontiguousarray(pol, np.uint8)
    n = t.shape[0]
    lo = C.c_int64(0)
    hi = C.c_int64(0)
    lib = port()
    npos = C.c_int(0)
    nneg = C.c_int(0)
    # first call sizes
    cap = n
    pos = np.zeros((max(cap, 1), 2))
    neg = np.zeros((max(cap, 1), 2))
    lib.orc_event_frame(_p(t, _dp), _p(x, _dp), _p(y, _dp), _p(pol, _bp), C.c_int64(n), C.c_double(t0), C.c_double(t1),
                        _p(pos, _dp), C.byref(npos), _p(neg, _dp), C.byref(nneg), C.byref(lo), C.byref(hi))
    return pos[:npos.value].copy(), neg[:nneg.value].copy(), lo.value, hi.value


def uset_order(xy):
    xy = np.ascontiguousarray(xy, dtype=np.float64).reshape(-1, 2)
    out = np.zeros_like(xy)
    k = port().orc_uset_order(_p(xy, _dp), C.c_int(xy.shape[0]), _p(out, _dp))
    return out[:k].copy()


def radius_threshold(W, H, rows, cols, asym, square, radius):
    return port().orc_radius_threshold(W, H, rows, cols, asym, square, radius)


def fit_circle(pxy, nxy):
    pxy = np.ascontiguousarray(pxy, np.float64).reshape(-1, 2)
    nxy = np.ascontiguousarray(nxy, np.float64).reshape(-1, 2)
    out = np.zeros(3)
    port().orc_fit_circle(_p(pxy, _dp), C.c_int(len(pxy)), _p(nxy, _dp), C.c_int(len(nxy)), _p(out, _dp))
    return out


def extract(pxy, nxy, eps=4.0, minS=2, clusterMin=5, knn_num=3, fitCircle=0, Rthr=15.51, rows_cols=36,
            canonical_median=False, lib=None):
    """extractFeatures up to findCirclesGrid on explicit point sets (V order given)."""
    pxy = np.ascontiguousarray(pxy, np.float64).reshape(-1, 2)
    nxy = np.ascontiguousarray(nxy, np.float64).reshape(-1, 2)
    np_, nn = len(pxy), len(nxy)
    pl = np.full(max(np_, 1), -1, np.int32)
    nl = np.full(max(nn, 1), -1, np.int32)
    pm = np.zeros(max(np_, 1), np.uint32)
    nm = np.zeros(max(nn, 1), np.uint32)
    po = np.zeros(np_ + 2, np.uint32)
    no = np.zeros(nn + 2, np.uint32)
    kp = np.zeros(max(np_, 1), np.int32)
    kn = np.zeros(max(nn, 1), np.int32)
    mp = np.zeros(max(np_, 1), np.int32)
    mn = np.zeros(max(nn, 1), np.int32)
    cap = max(np_, 1)
    cand = np.zeros((cap, 5))
    info = np.zeros(8, np.int32)
    lib = lib or port()
    lib.orc_extract(_p(pxy, _dp), C.c_int(np_), _p(nxy, _dp), C.c_int(nn), C.c_double(eps), C.c_uint(minS),
                    C.c_uint(clusterMin), C.c_int(knn_num), C.c_int(fitCircle), C.c_double(Rthr), C.c_uint(rows_cols),
                    C.c_int(int(canonical_median)), _p(pl, _ip), _p(nl, _ip), _p(pm, _up), _p(po, _up), _p(nm, _up),
                    _p(no, _up), _p(kp, _ip), _p(kn, _ip), _p(mp, _ip), _p(mn, _ip), _p(cand, _dp), C.c_int(cap),
                    _p(info, _ip))
    return dict(p_labels=pl[:np_].copy(), n_labels=nl[:nn].copy(),
                p_clusters=[pm[po[c]:po[c + 1]].copy() for c in range(info[0])],
                n_clusters=[nm[no[c]:no[c + 1]].copy() for c in range(info[1])],
                kept_p=kp[:info[2]].copy(), kept_n=kn[:info[3]].copy(), enough=int(info[4]),
                med_p=mp[:info[2]].copy() if info[4] else np.zeros(0, np.int32),
                med_n=mn[:info[3]].copy() if info[4] else np.zeros(0, np.int32),
                cand=cand[:info[5]].copy())


def rectify(pxy, nxy, img, W, H, eps=4.0, minS=2, clusterMin=5, rows=9, cols=4, asym=True, fitCircle=0):
    """rectifyFeatures (CirclesEventFrame.cpp:417-609) on explicit point sets; img [n_feat][5][2] projected points.
    Returns (out [n_feat][3] = cx, cy, r with r < 0 for deleted features, frame verdict)."""
    pxy = np.ascontiguousarray(pxy, np.float64).reshape(-1, 2)
    nxy = np.ascontiguousarray(nxy, np.float64).reshape(-1, 2)
    img = np.ascontiguousarray(img, np.float64)
    nf = img.shape[0]
    out = np.zeros((nf, 3))
    ok = port().orc_rectify(_p(pxy, _dp), C.c_int(len(pxy)), _p(nxy, _dp), C.c_int(len(nxy)), C.c_double(eps),
                            C.c_uint(minS), C.c_uint(clusterMin), _p(img, _dp), C.c_int(nf), C.c_double(W), C.c_double(H),
                            C.c_int(rows), C.c_int(cols), C.c_int(int(asym)), C.c_int(fitCircle), _p(out, _dp))
    return out, int(ok)


def frontend_windows(t, x, y, pol, windows, eps=4.0, minS=2, clusterMin=5, knn_num=3, fitCircle=0, Rthr=15.51,
                     rows_cols=36, threads=1, ref=True):
    """CPU baseline: reference-shaped front end over a list of windows with `threads` std::threads.
    ref=True uses oracle/_ref (verbatim reference DBSCAN) when available. Returns (candidates, events, per-window)."""
    lib = ref_frontend_lib() if (ref and os.path.exists(_REF_FE)) else port()
    t = np.ascontiguousarray(t, np.float64)
    x = np.ascontiguousarray(x, np.float64)
    y = np.ascontiguousarray(y, np.float64)
    pol = np.ascontiguousarray(pol, np.uint8)
    win = np.ascontiguousarray(windows, np.float64).reshape(-1, 2)
    nev = C.c_int64(0)
    per = np.zeros(max(len(win), 1), np.int32)
    tot = lib.orc_frontend_windows(_p(t, _dp), _p(x, _dp), _p(y, _dp), _p(pol, _bp), C.c_int64(len(t)), _p(win, _dp),
                                   C.c_int(len(win)), C.c_double(eps), C.c_uint(minS), C.c_uint(clusterMin),
                                   C.c_int(knn_num), C.c_int(fitCircle), C.c_double(Rthr), C.c_uint(rows_cols),
                                   C.c_int(threads), C.byref(nev), _p(per, _ip))
    return int(tot), int(nev.value), per[:len(win)].copy()


def frontend_windows_detail(t, x, y, pol, windows, eps=4.0, minS=2, clusterMin=5, knn_num=3, fitCircle=0, Rthr=15.51,
                            rows_cols=36, threads=1, ref=True, cand_cap=64):
    """frontend_windows with the per-window results kept: returns (events, counts [n_win][6] = points -, +, raw clusters
    -, +, kept clusters -, +; candidates per window; cand [n_win][cand_cap][5] = pi, ni, cx, cy, r)."""
    lib = ref_frontend_lib() if (ref and os.path.exists(_REF_FE)) else port()
    lib.orc_frontend_windows_detail.restype = C.c_int64
    t = np.ascontiguousarray(t, np.float64)
    x = np.ascontiguousarray(x, np.float64)
    y = np.ascontiguousarray(y, np.float64)
    pol = np.ascontiguousarray(pol, np.uint8)
    win = np.ascontiguousarray(windows, np.float64).reshape(-1, 2)
    nev = C.c_int64(0)
    per = np.zeros(max(len(win), 1), np.int32)
    counts = np.zeros((max(len(win), 1), 6), np.int32)
    cand = np.zeros((max(len(win), 1), cand_cap, 5))
    lib.orc_frontend_windows_detail(_p(t, _dp), _p(x, _dp), _p(y, _dp), _p(pol, _bp), C.c_int64(len(t)), _p(win, _dp),
                                    C.c_int(len(win)), C.c_double(eps), C.c_uint(minS), C.c_uint(clusterMin),
                                    C.c_int(knn_num), C.c_int(fitCircle), C.c_double(Rthr), C.c_uint(rows_cols),
                                    C.c_int(threads), C.byref(nev), _p(per, _ip), _p(counts, _ip), _p(cand, _dp),
                                    C.c_int(cand_cap))
    return int(nev.value), counts[:len(win)].copy(), per[:len(win)].copy(), cand[:len(win)]


# ------------------------------------------------------------------------------- cost evaluation ----
def inverse_radial(k4):
    k = np.ascontiguousarray(k4, np.float64)
    b = np.zeros(5)
    port().orc_inverse_radial(_p(k, _dp), _p(b, _dp))
    return b


def inverse_distortion_roundtrip():
    f = port().orc_inverse_distortion_roundtrip
    f.restype = C.c_double
    return f()


def basis(knots, u):
    kn = np.ascontiguousarray(knots, np.float64)
    N = np.zeros(4)
    sp = C.c_int(0)
    port().orc_basis(_p(kn, _dp), C.c_int(len(kn)), C.c_double(u), C.byref(sp), _p(N, _dp))
    return sp.value, N


def knots(us, n_cp):
    us = np.ascontiguousarray(us, np.float64)
    kn = np.zeros(n_cp + 4)
    port().orc_knots(_p(us, _dp), C.c_int(len(us)), C.c_int(n_cp), _p(kn, _dp))
    return kn


def residual_jac(intr, rcp, tcp, obs, lm, radius, b):
    a = [np.ascontiguousarray(v, np.float64) for v in (intr, rcp, tcp, obs, lm, b)]
    jac = np.zeros(37)
    f = port().orc_residual_jac
    f.restype = C.c_double
    r = f(_p(a[0], _dp), _p(a[1], _dp), _p(a[2], _dp), _p(a[3], _dp), _p(a[4], _dp), C.c_double(radius), _p(a[5], _dp),
          _p(jac, _dp))
    return r, jac


def residual_jac_so3(intr, Q, T, obs, lm, radius, basis):
    """CalibReprojectionError_SO3 on Jet<37>: value and 1x37 ambient Jacobian (intrinsics | 4x4 SO3 coeffs | 4x3 translations)"""
    a = [np.ascontiguousarray(v, np.float64) for v in (intr, Q, T, obs, lm, basis)]
    jac = np.zeros(37)
    lib = port()
    lib.orc_residual_jac_so3.restype = C.c_double
    r = lib.orc_residual_jac_so3(_p(a[0], _dp), _p(a[1], _dp), _p(a[2], _dp), _p(a[3], _dp), _p(a[4], _dp), C.c_double(radius),
                                 _p(a[5], _dp), _p(jac, _dp))
    return r, jac


def so3_plus(x, d):
    x = np.ascontiguousarray(x, np.float64)
    d = np.ascontiguousarray(d, np.float64)
    out = np.zeros(4)
    port().orc_so3_plus(_p(x, _dp), _p(d, _dp), _p(out, _dp))
    return out


def so3_plus_jacobian(x):
    x = np.ascontiguousarray(x, np.float64)
    J = np.zeros(12)
    port().orc_so3_plus_jacobian(_p(x, _dp), _p(J, _dp))
    return J.reshape(4, 3)


def quat_plus(x, d):
    x = np.ascontiguousarray(x, np.float64)
    d = np.ascontiguousarray(d, np.float64)
    out = np.zeros(4)
    port().orc_quat_plus(_p(x, _dp), _p(d, _dp), _p(out, _dp))
    return out


class CostProblem:
    """Oracle of the cost path: association, cost, per-span normal equations (oracle/ecb_oracle_cost.cpp)."""

    def __init__(self, n_cp, knots_list, radius=1.75, huber=0.35, so3=False):
        self.lib = port()
        self.lib.orc_problem_create.restype = C.c_void_p
        self.lib.orc_cost.restype = C.c_double
        self.lib.orc_normal_eq.restype = C.c_double
        self.lib.orc_associate.restype = C.c_int64
        self.lib.orc_problem_num_residuals.restype = C.c_int64
        ncp = np.ascontiguousarray(np.atleast_1d(n_cp), np.int32)
        kn = np.ascontiguousarray(np.concatenate([np.asarray(k, np.float64).ravel() for k in knots_list]), np.float64)
        self.h = C.c_void_p(self.lib.orc_problem_create(C.c_int(len(ncp)), _p(ncp, _ip), _p(kn, _dp), C.c_double(radius),
                                                        C.c_double(huber)))
        self.n_spans = int(self.lib.orc_problem_num_spans(self.h))
        self.lib.orc_problem_set_so3(self.h, C.c_int(int(so3)))   # useSO3: CalibReprojectionError_SO3

    def __del__(self):
        try:
            self.lib.orc_problem_free(self.h)
        except Exception:
            pass

    def set_residuals(self, obs, lm, t, spline):
        obs = np.ascontiguousarray(obs, np.float64)
        lm = np.ascontiguousarray(lm, np.float64)
        t = np.ascontiguousarray(t, np.float64)
        sp = np.ascontiguousarray(spline, np.int32)
        self.lib.orc_problem_set_residuals(self.h, _p(obs, _dp), _p(lm, _dp), _p(t, _dp), _p(sp, _ip), C.c_int64(len(t)))

    def associate(self, ev_t, ev_x, ev_y, kf_t, circles, landmarks, step):
        ev_t = np.ascontiguousarray(ev_t, np.float64)
        ev_x = np.ascontiguousarray(ev_x, np.float64)
        ev_y = np.ascontiguousarray(ev_y, np.float64)
        kf_t = np.ascontiguousarray(kf_t, np.float64)
        circles = np.ascontiguousarray(circles, np.float64)
        lm = np.ascontiguousarray(landmarks, np.float64)
        oe = np.zeros(max(len(ev_t), 1), np.int64)
        oc = np.zeros(max(len(ev_t), 1), np.int32)
        n = self.lib.orc_associate(self.h, _p(ev_t, _dp), _p(ev_x, _dp), _p(ev_y, _dp), C.c_int64(len(ev_t)), _p(kf_t, _dp),
                                   _p(circles, _dp), C.c_int(len(kf_t)), C.c_int(circles.shape[1]), _p(lm, _dp),
                                   C.c_double(step), _p(oe, _lp), _p(oc, _ip))
        return oe[:n].copy(), oc[:n].copy()

    @property
    def n_residuals(self):
        return int(self.lib.orc_problem_num_residuals(self.h))

    def cost(self, intr, rot, trans):
        a = [np.ascontiguousarray(v, np.float64) for v in (intr, rot, trans)]
        return self.lib.orc_cost(self.h, _p(a[0], _dp), _p(a[1], _dp), _p(a[2], _dp))

    def eval_mt(self, intr, rot, trans, threads):
        """CPU baseline: one Jacobian evaluation + one cost-only evaluation on `threads` std::threads."""
        a = [np.ascontiguousarray(v, np.float64) for v in (intr, rot, trans)]
        self.lib.orc_eval_mt.restype = C.c_double
        c2 = C.c_double(0)
        c = self.lib.orc_eval_mt(self.h, _p(a[0], _dp), _p(a[1], _dp), _p(a[2], _dp), C.c_int(threads), C.byref(c2))
        return c, c2.value

    def normal_eq(self, intr, rot, trans, want_rows=False):
        a = [np.ascontiguousarray(v, np.float64) for v in (intr, rot, trans)]
        H = np.zeros((self.n_spans, 33, 33))
        g = np.zeros((self.n_spans, 33))
        n = self.n_residuals
        r = np.zeros(max(n, 1)) if want_rows else None
        J = np.zeros((max(n, 1), 33)) if want_rows else None
        c = self.lib.orc_normal_eq(self.h, _p(a[0], _dp), _p(a[1], _dp), _p(a[2], _dp), _p(H, _dp), _p(g, _dp),
                                   _p(r, _dp) if want_rows else None, _p(J, _dp) if want_rows else None)
        if want_rows:
            return c, H, g, r[:n], J[:n]
        return c, H, g

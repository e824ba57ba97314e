"""eventcalib_b200 — Python mirror of the C ABI in ``include/eventcalib_b200.h``.

The product is the CUDA library ``libecb.so`` (built in-tree by ``eventcalib_b200.build``); this module only
binds it with ctypes for the tests, ``bench.py`` and ``__graft_entry__``.  There is no CPU fallback: if the
library is missing or no CUDA device is present, every compute call raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libecb.so")

OK, FAILED, ERR_CUDA, ERR_ARG, ERR_UNSUPPORTED, ERR_STATE = 0, 1, -1, -2, -3, -4
PB_DUPLICATE, PB_CLUSTER_CAP, PB_RANGE = 2, 4, 8


class EcbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("ecb error %d: %s" % (code, msg))
        self.code = code


class FrontendParams(C.Structure):
    _fields_ = [("dbscan_eps", C.c_double), ("dbscan_min_pts", C.c_uint32), ("cluster_min", C.c_uint32),
                ("knn_num", C.c_int32), ("fit_circle", C.c_int32), ("radius_threshold", C.c_double),
                ("rows_cols", C.c_uint32), ("order_mode", C.c_int32), ("max_clusters", C.c_uint32),
                ("median_mode", C.c_uint32)]


class LmOptions(C.Structure):
    _fields_ = [("max_iterations", C.c_int32), ("jacobi_scaling", C.c_int32), ("fixed_iterations", C.c_int32),
                ("rotation_model", C.c_int32), ("function_tolerance", C.c_double), ("gradient_tolerance", C.c_double),
                ("parameter_tolerance", C.c_double), ("initial_radius", C.c_double), ("max_radius", C.c_double),
                ("min_radius", C.c_double), ("min_relative_decrease", C.c_double), ("min_lm_diagonal", C.c_double),
                ("max_lm_diagonal", C.c_double)]


class LmSummary(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("successful_steps", C.c_int32), ("termination", C.c_int32),
                ("reserved", C.c_int32), ("initial_cost", C.c_double), ("final_cost", C.c_double),
                ("gradient_max_norm", C.c_double), ("radius", C.c_double)]


class WindowSummary(C.Structure):
    _fields_ = [("ev_lo", C.c_int64), ("ev_hi", C.c_int64), ("n_points", C.c_int32 * 2), ("n_clusters", C.c_int32 * 2),
                ("n_kept", C.c_int32 * 2), ("n_candidates", C.c_int32), ("status", C.c_uint32),
                ("point_offset", C.c_int64 * 2)]


SUMMARY_DTYPE = np.dtype([("ev_lo", "<i8"), ("ev_hi", "<i8"), ("n_points", "<i4", 2), ("n_clusters", "<i4", 2),
                          ("n_kept", "<i4", 2), ("n_candidates", "<i4"), ("status", "<u4"), ("point_offset", "<i8", 2)])
assert SUMMARY_DTYPE.itemsize == C.sizeof(WindowSummary)

# every symbol include/eventcalib_b200.h declares (checked by tests/test_abi.py)
SYMBOLS = ["ecb_ctx_create", "ecb_ctx_destroy", "ecb_last_error", "ecb_launch_count", "ecb_synchronize", "ecb_version", "ecb_device_count",
           "ecb_set_sensor", "ecb_load_events_host", "ecb_load_events_device", "ecb_num_events", "ecb_frontend_run",
           "ecb_frontend_summary", "ecb_frontend_total_points", "ecb_frontend_points", "ecb_frontend_candidates",
           "ecb_frontend_clusters", "ecb_frontend_rectify", "ecb_frontend_device_ptrs", "ecb_dbscan_run", "ecb_dbscan_run_batch",
           "ecb_dbscan_run_ordered", "ecb_dbscan_run_batch_ordered", "ecb_dbscan_run_nd",
           "ecb_fit_circles", "ecb_set_profiling", "ecb_stage_ms", "ecb_cost_setup", "ecb_cost_set_rotation_model", "ecb_cost_layout",
           "ecb_cost_associate", "ecb_cost_associate_device", "ecb_cost_get_association", "ecb_cost_set_residuals", "ecb_cost_eval", "ecb_cost_normal_eq",
           "ecb_exchange_buffer_bytes", "ecb_cost_normal_eq_exchange", "ecb_device_alloc", "ecb_device_free", "ecb_ipc_export",
           "ecb_ipc_open", "ecb_ipc_close", "ecb_enable_peer_access", "ecb_lm_default_options", "ecb_lm_create",
           "ecb_lm_destroy", "ecb_lm_dimension", "ecb_lm_begin", "ecb_lm_propose", "ecb_lm_feedback", "ecb_lm_update",
           "ecb_lm_state", "ecb_lm_trace", "ecb_calibrate", "ecb_lm_device_create", "ecb_lm_device_destroy",
           "ecb_lm_device_set_exchange", "ecb_lm_device_begin", "ecb_lm_device_iterate", "ecb_lm_device_running",
           "ecb_lm_device_result", "ecb_calibrate_device"]

_lib = None


def load_library():
    """Loads libecb.so; raises if the CUDA extension was not built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("ECB_LIBRARY", LIB_PATH)  # A/B builds of the same library (profiles/tools/ab_build.py)
    if not os.path.exists(path):
        raise EcbError(ERR_CUDA, "libecb.so not built (run `python -m eventcalib_b200.build`); there is no CPU fallback")
    lib = C.CDLL(path)
    vp, i32, i64, u32, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_uint32, C.c_double
    lib.ecb_ctx_create.argtypes = [i32, vp, C.POINTER(vp)]
    lib.ecb_ctx_destroy.argtypes = [vp]
    lib.ecb_ctx_destroy.restype = None
    lib.ecb_last_error.argtypes = [vp]
    lib.ecb_last_error.restype = C.c_char_p
    lib.ecb_launch_count.argtypes = [vp]
    lib.ecb_launch_count.restype = C.c_uint64
    lib.ecb_synchronize.argtypes = [vp]
    lib.ecb_version.restype = C.c_char_p
    lib.ecb_set_sensor.argtypes = [vp, i32, i32]
    lib.ecb_load_events_host.argtypes = [vp, vp, i64]
    lib.ecb_load_events_device.argtypes = [vp, vp, i64]
    lib.ecb_num_events.argtypes = [vp]
    lib.ecb_num_events.restype = i64
    lib.ecb_frontend_run.argtypes = [vp, vp, i32, C.POINTER(FrontendParams)]
    lib.ecb_frontend_summary.argtypes = [vp, vp, i32]
    lib.ecb_frontend_total_points.argtypes = [vp, i32]
    lib.ecb_frontend_total_points.restype = i64
    lib.ecb_frontend_points.argtypes = [vp, i32, vp, vp]
    lib.ecb_frontend_candidates.argtypes = [vp, vp, i32]
    lib.ecb_frontend_clusters.argtypes = [vp, i32, i32, vp, vp, vp, i32]
    lib.ecb_frontend_device_ptrs.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(i32)]
    lib.ecb_frontend_rectify.argtypes = [vp, vp, i32, i32, vp, dbl, i32, i32, i32, vp, vp]
    lib.ecb_cost_set_rotation_model.argtypes = [vp, i32]
    lib.ecb_exchange_buffer_bytes.argtypes = [vp, i32]
    lib.ecb_exchange_buffer_bytes.restype = C.c_size_t
    lib.ecb_cost_normal_eq_exchange.argtypes = [vp, vp, vp, vp, i32, i32, vp, C.c_uint64, i32, vp, vp]
    lib.ecb_device_alloc.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
    lib.ecb_device_free.argtypes = [vp, vp]
    lib.ecb_ipc_export.argtypes = [vp, vp, vp]
    lib.ecb_ipc_open.argtypes = [vp, vp, C.POINTER(vp)]
    lib.ecb_ipc_close.argtypes = [vp, vp]
    lib.ecb_enable_peer_access.argtypes = [vp, i32]
    lib.ecb_dbscan_run.argtypes = [vp, vp, i32, dbl, u32, vp, C.POINTER(C.c_int32)]
    lib.ecb_dbscan_run_batch.argtypes = [vp, vp, vp, i32, dbl, u32, vp, vp, vp]
    lib.ecb_dbscan_run_ordered.argtypes = [vp, vp, i32, dbl, u32, vp, C.POINTER(C.c_int32), vp, vp]
    lib.ecb_dbscan_run_batch_ordered.argtypes = [vp, vp, vp, i32, dbl, u32, vp, vp, vp, vp, vp]
    lib.ecb_dbscan_run_nd.argtypes = [vp, vp, i32, i32, dbl, u32, vp, C.POINTER(C.c_int32), vp, vp]
    lib.ecb_fit_circles.argtypes = [vp, vp, vp, i32, vp]
    lib.ecb_set_profiling.argtypes = [vp, i32]
    lib.ecb_stage_ms.argtypes = [vp, vp]
    lib.ecb_cost_setup.argtypes = [vp, i32, vp, vp, dbl, dbl]
    lib.ecb_cost_layout.argtypes = [vp, vp, vp, vp, vp]
    lib.ecb_cost_associate.argtypes = [vp, vp, vp, i32, i32, vp, dbl, vp]
    lib.ecb_cost_associate_device.argtypes = [vp, vp, vp, i32, i32, vp, dbl, vp]
    lib.ecb_cost_get_association.argtypes = [vp, vp, vp, i64]
    lib.ecb_cost_set_residuals.argtypes = [vp, vp, vp, vp, vp, i64]
    lib.ecb_cost_eval.argtypes = [vp, vp, vp, vp, vp]
    lib.ecb_cost_normal_eq.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    lib.ecb_lm_default_options.argtypes = [C.POINTER(LmOptions)]
    lib.ecb_lm_default_options.restype = None
    lib.ecb_lm_create.argtypes = [i32, vp, C.POINTER(LmOptions)]
    lib.ecb_lm_create.restype = vp
    lib.ecb_lm_destroy.argtypes = [vp]
    lib.ecb_lm_destroy.restype = None
    lib.ecb_lm_dimension.argtypes = [vp]
    lib.ecb_lm_begin.argtypes = [vp, vp, vp, vp, vp]
    lib.ecb_lm_propose.argtypes = [vp, vp, vp, vp]
    lib.ecb_lm_feedback.argtypes = [vp, dbl]
    lib.ecb_lm_update.argtypes = [vp, vp]
    lib.ecb_lm_state.argtypes = [vp, vp, vp, vp, C.POINTER(LmSummary)]
    lib.ecb_lm_trace.argtypes = [vp, vp, i32]
    lib.ecb_calibrate.argtypes = [vp, i32, vp, vp, vp, vp, C.POINTER(LmOptions), C.POINTER(LmSummary), vp, i32]
    lib.ecb_lm_device_create.argtypes = [vp, i32, vp, C.POINTER(LmOptions), C.POINTER(vp)]
    lib.ecb_lm_device_destroy.argtypes = [vp]
    lib.ecb_lm_device_destroy.restype = None
    lib.ecb_lm_device_set_exchange.argtypes = [vp, i32, i32, vp]
    lib.ecb_lm_device_begin.argtypes = [vp, vp, vp, vp]
    lib.ecb_lm_device_iterate.argtypes = [vp, i32]
    lib.ecb_lm_device_running.argtypes = [vp]
    lib.ecb_lm_device_result.argtypes = [vp, vp, vp, vp, C.POINTER(LmSummary), vp, i32]
    lib.ecb_calibrate_device.argtypes = [vp, vp, vp, vp, C.POINTER(LmSummary), vp, i32]
    _lib = lib
    return lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def default_params(eps=4.0, min_pts=2, cluster_min=5, knn_num=3, fit_circle=0, radius_threshold=15.511363636363637,
                   rows_cols=36, order_mode=0, max_clusters=0, median_mode=0):
    """CirclesEventFrame::Params defaults + parameter/event_calibration/example.yaml.
    order_mode=1, median_mode=1 reproduce the reference's pid order and std::nth_element medians."""
    return FrontendParams(eps, min_pts, cluster_min, knn_num, fit_circle, radius_threshold, rows_cols, order_mode,
                          max_clusters, median_mode)


def radius_threshold(width, height, rows, cols, asymmetric, square, radius):
    """circleRadiusThreshold_ of the CirclesEventFrame ctor (CirclesEventFrame.cpp:16-33)."""
    c2 = 2.0 * cols if asymmetric else float(cols)
    W, H = float(width), float(height)
    return min(max(W, H) / max(float(rows), c2), min(W, H) / min(float(rows), c2)) / square * radius * 1.5


class Context:
    """One CUDA device + stream + buffers (``ecb_ctx``)."""

    def __init__(self, device=0, stream=None):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.ecb_ctx_create(device, C.c_void_p(stream) if stream else None, C.byref(h))
        if rc != OK:
            raise EcbError(rc, "ecb_ctx_create failed: no usable CUDA device %d (there is no CPU fallback)" % device)
        self.h = h
        self.n_win = 0

    def close(self):
        if getattr(self, "h", None):
            self.lib.ecb_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc < 0 or rc == FAILED:
            raise EcbError(rc, self.lib.ecb_last_error(self.h).decode())
        return rc

    @property
    def launches(self):
        return int(self.lib.ecb_launch_count(self.h))

    def synchronize(self):
        self._chk(self.lib.ecb_synchronize(self.h))

    STAGES = ["ingest", "bounds", "window", "cluster", "pair", "assoc", "normal_eq", "cost", "order", "bfs"]

    def set_profiling(self, on=True):
        self._chk(self.lib.ecb_set_profiling(self.h, int(on)))

    def stage_ms(self):
        out = np.zeros(len(self.STAGES), np.float32)
        self._chk(self.lib.ecb_stage_ms(self.h, _ptr(out)))
        return dict(zip(self.STAGES, out.tolist()))

    # ---- ingest ----
    def set_sensor(self, width, height):
        self._chk(self.lib.ecb_set_sensor(self.h, width, height))

    def load_events(self, records):
        """records: numpy structured array / bytes of packed 25-byte reference records (host)."""
        buf = np.ascontiguousarray(records)
        n = buf.nbytes // 25
        self._keep = buf
        self._chk(self.lib.ecb_load_events_host(self.h, _ptr(buf), n))
        return n

    def load_events_ptr(self, host_ptr, n):
        """host_ptr: address of n packed records in (ideally pinned) host memory."""
        self._chk(self.lib.ecb_load_events_host(self.h, C.c_void_p(host_ptr), n))
        return n

    def load_events_device(self, dptr, n):
        self._chk(self.lib.ecb_load_events_device(self.h, C.c_void_p(dptr), n))

    # ---- front end ----
    def frontend_run(self, windows, params=None):
        win = np.ascontiguousarray(windows, np.float64).reshape(-1, 2)
        params = params or default_params()
        self._chk(self.lib.ecb_frontend_run(self.h, _ptr(win), len(win), C.byref(params)))
        self.n_win = len(win)

    def summary(self, out=None):
        """per-window summaries; `out` (optional): a caller-owned SUMMARY_DTYPE array of n_win entries to fill"""
        if out is None:
            out = np.zeros(self.n_win, SUMMARY_DTYPE)
        assert out.dtype == SUMMARY_DTYPE and out.flags.c_contiguous and len(out) == self.n_win
        if self.n_win:
            self._chk(self.lib.ecb_frontend_summary(self.h, _ptr(out), self.n_win))
        return out

    def points(self, polarity):
        n = int(self.lib.ecb_frontend_total_points(self.h, polarity))
        xy = np.zeros((max(n, 1), 2))
        lab = np.zeros(max(n, 1), np.int32)
        self._chk(self.lib.ecb_frontend_points(self.h, polarity, _ptr(xy), _ptr(lab)))
        return xy[:n], lab[:n]

    def candidates(self, max_cand=64, out=None):
        if out is None:
            out = np.zeros((self.n_win, max_cand, 5))
        assert out.dtype == np.float64 and out.flags.c_contiguous and out.shape == (self.n_win, max_cand, 5)
        if self.n_win:
            self._chk(self.lib.ecb_frontend_candidates(self.h, _ptr(out), max_cand))
        return out

    def rectify(self, window_index, image_points, rows=9, cols=4, asymmetric=True, inlier_threshold=3.0):
        """rectifyFeatures for frames of the last front-end run; image_points [n_frames][n_circles][5][2]"""
        wi = np.ascontiguousarray(window_index, np.int32)
        img = np.ascontiguousarray(image_points, np.float64)
        nf, nc = img.shape[0], img.shape[1]
        assert img.shape == (nf, nc, 5, 2) and len(wi) == nf
        out = np.zeros((max(nf, 1), nc, 3))
        ok = np.zeros(max(nf, 1), np.int32)
        self._chk(self.lib.ecb_frontend_rectify(self.h, _ptr(wi), nf, nc, _ptr(img), float(inlier_threshold), rows, cols,
                                                int(asymmetric), _ptr(out), _ptr(ok)))
        return out[:nf], ok[:nf]

    def clusters(self, window, polarity, cap=512):
        raw = np.zeros(cap, np.int32)
        size = np.zeros(cap, np.int32)
        med = np.zeros(cap, np.int32)
        n = self._chk(self.lib.ecb_frontend_clusters(self.h, window, polarity, _ptr(raw), _ptr(size), _ptr(med), cap))
        return raw[:n], size[:n], med[:n]

    # ---- DBSCAN::Run boundary ----
    def dbscan(self, xy, eps, min_pts):
        """Returns (status, labels, n_clusters) like DBSCAN::Run: status 0 SUCCESS / 1 FAILED."""
        xy = np.ascontiguousarray(xy, np.float64).reshape(-1, 2)
        n = len(xy)
        labels = np.full(max(n, 1), -1, np.int32)
        nc = C.c_int32(0)
        rc = self.lib.ecb_dbscan_run(self.h, _ptr(xy), n, float(eps), int(min_pts), _ptr(labels), C.byref(nc))
        if rc == FAILED:
            return FAILED, labels[:0], 0
        self._chk(rc)
        return OK, labels[:n], nc.value

    def dbscan_ordered(self, xy, eps, min_pts):
        """DBSCAN::Run with `Clusters` as ordered lists (the reference's member order) and `Noise`."""
        xy = np.ascontiguousarray(xy, np.float64).reshape(-1, 2)
        n = len(xy)
        labels = np.full(max(n, 1), -1, np.int32)
        sizes = np.zeros(max(n, 1), np.int32)
        members = np.zeros(max(n, 1), np.uint32)
        nc = C.c_int32(0)
        rc = self.lib.ecb_dbscan_run_ordered(self.h, _ptr(xy), n, float(eps), int(min_pts), _ptr(labels), C.byref(nc),
                                             _ptr(sizes), _ptr(members))
        if rc == FAILED:
            return FAILED, labels[:0], [], labels[:0]
        self._chk(rc)
        off = np.concatenate([[0], np.cumsum(sizes[:nc.value])])
        clusters = [members[off[c]:off[c + 1]].copy() for c in range(nc.value)]
        return OK, labels[:n], clusters, np.nonzero(labels[:n] < 0)[0].astype(np.uint32)

    def dbscan_nd(self, pts, eps, min_pts, ordered=True):
        """DBSCAN<T,Float>::Run(V, dim, eps, min) for dim = pts.shape[1] (1..4): (status, labels, clusters, noise)."""
        pts = np.ascontiguousarray(pts, np.float64)
        n, dim = pts.shape
        labels = np.full(max(n, 1), -1, np.int32)
        sizes = np.zeros(max(n, 1), np.int32)
        members = np.zeros(max(n, 1), np.uint32)
        nc = C.c_int32(0)
        rc = self.lib.ecb_dbscan_run_nd(self.h, _ptr(pts), n, dim, float(eps), int(min_pts), _ptr(labels), C.byref(nc),
                                        _ptr(sizes) if ordered else None, _ptr(members) if ordered else None)
        if rc == FAILED:
            return FAILED, labels[:0], [], labels[:0]
        self._chk(rc)
        off = np.concatenate([[0], np.cumsum(sizes[:nc.value])])
        clusters = [members[off[c]:off[c + 1]].copy() for c in range(nc.value)] if ordered else nc.value
        return OK, labels[:n], clusters, np.nonzero(labels[:n] < 0)[0].astype(np.uint32)

    def dbscan_batch_ordered(self, xy, offsets, eps, min_pts):
        xy = np.ascontiguousarray(xy, np.float64).reshape(-1, 2)
        off = np.ascontiguousarray(offsets, np.int64)
        k = len(off) - 1
        labels = np.full(max(len(xy), 1), -1, np.int32)
        sizes = np.zeros(max(len(xy), 1), np.int32)
        members = np.zeros(max(len(xy), 1), np.uint32)
        nc = np.zeros(max(k, 1), np.int32)
        st = np.zeros(max(k, 1), np.uint32)
        self._chk(self.lib.ecb_dbscan_run_batch_ordered(self.h, _ptr(xy), _ptr(off), k, float(eps), int(min_pts),
                                                        _ptr(labels), _ptr(nc), _ptr(st), _ptr(sizes), _ptr(members)))
        out = []
        for j in range(k):
            o = np.concatenate([[0], np.cumsum(sizes[off[j]:off[j] + nc[j]])]) + off[j]
            out.append([members[o[c]:o[c + 1]].copy() for c in range(nc[j])])
        return labels[:len(xy)], nc[:k], st[:k], out

    def dbscan_batch(self, xy, offsets, eps, min_pts):
        xy = np.ascontiguousarray(xy, np.float64).reshape(-1, 2)
        off = np.ascontiguousarray(offsets, np.int64)
        k = len(off) - 1
        labels = np.full(max(len(xy), 1), -1, np.int32)
        nc = np.zeros(max(k, 1), np.int32)
        st = np.zeros(max(k, 1), np.uint32)
        self._chk(self.lib.ecb_dbscan_run_batch(self.h, _ptr(xy), _ptr(off), k, float(eps), int(min_pts), _ptr(labels),
                                                _ptr(nc), _ptr(st)))
        return labels[:len(xy)], nc[:k], st[:k]

    def fit_circles(self, xy, offsets):
        xy = np.ascontiguousarray(xy, np.float64).reshape(-1, 2)
        off = np.ascontiguousarray(offsets, np.int64)
        k = len(off) - 1
        out = np.zeros((max(k, 1), 3))
        self._chk(self.lib.ecb_fit_circles(self.h, _ptr(xy), _ptr(off), k, _ptr(out)))
        return out[:k]

    # ---- cost evaluation ----
    def cost_setup(self, n_cp, knots, radius=1.75, huber=0.35):
        n_cp = np.ascontiguousarray(np.atleast_1d(n_cp), np.int32)
        kn = np.ascontiguousarray(np.concatenate([np.asarray(k, np.float64).ravel() for k in knots])
                                  if isinstance(knots, (list, tuple)) else knots, np.float64)
        assert len(kn) == int(n_cp.sum()) + 4 * len(n_cp)
        self._chk(self.lib.ecb_cost_setup(self.h, len(n_cp), _ptr(n_cp), _ptr(kn), float(radius), float(huber)))

    def cost_set_rotation_model(self, use_so3):
        """0: normalised quaternion spline (useSO3: 0); 1: cumulative SO(3) spline + LocalParameterizationSO3 (useSO3: 1)"""
        self._chk(self.lib.ecb_cost_set_rotation_model(self.h, int(use_so3)))

    # ---- fused reduce + inter-GPU exchange of the normal equations ----
    def exchange_buffer_bytes(self, n_ranks):
        return int(self.lib.ecb_exchange_buffer_bytes(self.h, n_ranks))

    def device_alloc(self, nbytes):
        p = C.c_void_p()
        self._chk(self.lib.ecb_device_alloc(self.h, nbytes, C.byref(p)))
        return p.value

    def device_free(self, ptr):
        self._chk(self.lib.ecb_device_free(self.h, C.c_void_p(ptr)))

    def ipc_export(self, ptr):
        h = np.zeros(64, np.uint8)
        self._chk(self.lib.ecb_ipc_export(self.h, C.c_void_p(ptr), _ptr(h)))
        return h

    def ipc_open(self, handle):
        h = np.ascontiguousarray(handle, np.uint8)
        p = C.c_void_p()
        self._chk(self.lib.ecb_ipc_open(self.h, _ptr(h), C.byref(p)))
        return p.value

    def enable_peer_access(self, peer_device):
        """several GPUs driven from ONE process: lets this context's kernels store into buffers of `peer_device`"""
        self._chk(self.lib.ecb_enable_peer_access(self.h, int(peer_device)))

    def ipc_close(self, ptr):
        self._chk(self.lib.ecb_ipc_close(self.h, C.c_void_p(ptr)))

    def cost_normal_eq_exchange(self, intr, rot, trans, rank, recv_ptrs, epoch, d_out, want_cost=True, phases=3):
        """normal equations with the inter-GPU sum fused in; recv_ptrs: device pointers of every rank's receive buffer"""
        a = [np.ascontiguousarray(v, np.float64) for v in (intr, rot, trans)]
        ptrs = (C.c_void_p * len(recv_ptrs))(*[C.c_void_p(p) for p in recv_ptrs])
        cost = C.c_double(0)
        self._chk(self.lib.ecb_cost_normal_eq_exchange(self.h, _ptr(a[0]), _ptr(a[1]), _ptr(a[2]), rank, len(recv_ptrs), ptrs,
                                                       C.c_uint64(epoch), int(phases), C.c_void_p(d_out),
                                                       C.cast(C.byref(cost), C.c_void_p) if want_cost else None))
        return cost.value if want_cost else None

    def cost_layout(self):
        cp, sp = C.c_int32(), C.c_int32()
        nr, nd = C.c_int64(), C.c_int64()
        self._chk(self.lib.ecb_cost_layout(self.h, C.byref(cp), C.byref(sp), C.byref(nr), C.byref(nd)))
        return dict(total_cp=cp.value, total_spans=sp.value, n_residuals=nr.value, out_doubles=nd.value)

    def cost_associate(self, kf_t, circles, landmarks, step):
        kf_t = np.ascontiguousarray(kf_t, np.float64)
        circles = np.ascontiguousarray(circles, np.float64)
        lm = np.ascontiguousarray(landmarks, np.float64)
        n = C.c_int64(0)
        self._chk(self.lib.ecb_cost_associate(self.h, _ptr(kf_t), _ptr(circles), len(kf_t), circles.shape[1], _ptr(lm),
                                              float(step), C.byref(n)))
        return n.value

    def cost_associate_device(self, d_kf_t, d_circles, n_keyframes, n_circles, d_landmarks, step):
        """the same with device pointers (integers) of the three tables"""
        n = C.c_int64(0)
        self._chk(self.lib.ecb_cost_associate_device(self.h, C.c_void_p(d_kf_t), C.c_void_p(d_circles), int(n_keyframes),
                                                     int(n_circles), C.c_void_p(d_landmarks), float(step), C.byref(n)))
        return n.value

    def cost_association(self):
        n = self.cost_layout()["n_residuals"]
        ev = np.zeros(max(n, 1), np.int64)
        ci = np.zeros(max(n, 1), np.int32)
        self._chk(self.lib.ecb_cost_get_association(self.h, _ptr(ev), _ptr(ci), n))
        return ev[:n], ci[:n]

    def cost_set_residuals(self, obs, lm, t, spline):
        obs = np.ascontiguousarray(obs, np.float64)
        lm = np.ascontiguousarray(lm, np.float64)
        t = np.ascontiguousarray(t, np.float64)
        sp = np.ascontiguousarray(spline, np.int32)
        self._chk(self.lib.ecb_cost_set_residuals(self.h, _ptr(obs), _ptr(lm), _ptr(t), _ptr(sp), len(t)))

    def cost_eval(self, intr, rot, trans):
        intr = np.ascontiguousarray(intr, np.float64)
        rot = np.ascontiguousarray(rot, np.float64)
        trans = np.ascontiguousarray(trans, np.float64)
        c = C.c_double(0)
        self._chk(self.lib.ecb_cost_eval(self.h, _ptr(intr), _ptr(rot), _ptr(trans), C.byref(c)))
        return c.value

    def cost_normal_eq(self, intr, rot, trans, d_out=None, host=True):
        """Returns (cost, H[n_spans,33,33], g[n_spans,33]) when host=True, else leaves the packed result in d_out."""
        intr = np.ascontiguousarray(intr, np.float64)
        rot = np.ascontiguousarray(rot, np.float64)
        trans = np.ascontiguousarray(trans, np.float64)
        lay = self.cost_layout()
        c = C.c_double(0)
        if not host:
            self._chk(self.lib.ecb_cost_normal_eq(self.h, _ptr(intr), _ptr(rot), _ptr(trans), C.c_void_p(d_out), None, None))
            return None
        out = np.zeros(lay["out_doubles"])
        self._chk(self.lib.ecb_cost_normal_eq(self.h, _ptr(intr), _ptr(rot), _ptr(trans),
                                              C.c_void_p(d_out) if d_out else None, _ptr(out), C.byref(c)))
        ns = lay["total_spans"]
        blk = out[:ns * 1122].reshape(ns, 1122)
        return c.value, blk[:, :1089].reshape(ns, 33, 33).copy(), blk[:, 1089:].copy()

    def calibrate(self, n_cp, intr, rot, trans, options=None, trace_rows=256):
        """EventCalibSpline::optimize on one GPU: returns (intr, rot, trans, summary dict, trace[rows,4])."""
        n_cp = np.ascontiguousarray(np.atleast_1d(n_cp), np.int32)
        intr = np.array(intr, np.float64, copy=True)
        rot = np.array(rot, np.float64, copy=True)
        trans = np.array(trans, np.float64, copy=True)
        opt = options or lm_options()
        summ = LmSummary()
        tr = np.zeros((trace_rows, 4))
        self._chk(self.lib.ecb_calibrate(self.h, len(n_cp), _ptr(n_cp), _ptr(intr), _ptr(rot), _ptr(trans), C.byref(opt),
                                         C.byref(summ), _ptr(tr), trace_rows))
        rows = int(np.count_nonzero(tr[:, 2]))
        return intr, rot, trans, {f[0]: getattr(summ, f[0]) for f in LmSummary._fields_}, tr[:rows]


def lm_options(**kw):
    o = LmOptions()
    load_library().ecb_lm_default_options(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


class DeviceLm:
    """EventCalibSpline::optimize with the LM state machine and the band-arrow Cholesky solve on the device (ecb_lm_device_*):
    the host enqueues the iterations and reads the result once.  ctx must hold the residual set (cost_setup + association)."""

    def __init__(self, ctx, n_cp, options=None):
        self.ctx = ctx
        self.lib = ctx.lib
        self.n_cp = np.ascontiguousarray(np.atleast_1d(n_cp), np.int32)
        self.C = int(self.n_cp.sum())
        self.opt = options or lm_options()
        h = C.c_void_p()
        ctx._chk(self.lib.ecb_lm_device_create(ctx.h, len(self.n_cp), _ptr(self.n_cp), C.byref(self.opt), C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            if getattr(self.ctx, "h", None):   # the context owns the device: destroy before it (a closed context took it along)
                self.lib.ecb_lm_device_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_exchange(self, rank, recv_ptrs):
        """several GPUs: device pointers of every rank's receive buffer (own buffer at [rank]), see Context.exchange_buffer_bytes"""
        ptrs = (C.c_void_p * len(recv_ptrs))(*[C.c_void_p(p) for p in recv_ptrs])
        self._keep = ptrs
        self.ctx._chk(self.lib.ecb_lm_device_set_exchange(self.h, rank, len(recv_ptrs), ptrs))

    def begin(self, intr, rot, trans):
        a = [np.ascontiguousarray(v, np.float64) for v in (intr, rot, trans)]
        self.ctx._chk(self.lib.ecb_lm_device_begin(self.h, _ptr(a[0]), _ptr(a[1]), _ptr(a[2])))

    def iterate(self, n):
        self.ctx._chk(self.lib.ecb_lm_device_iterate(self.h, int(n)))

    def running(self):
        rc = self.lib.ecb_lm_device_running(self.h)
        if rc < 0:
            self.ctx._chk(rc)
        return rc == 1

    def result(self, trace_rows=256):
        i, r, t = np.zeros(9), np.zeros((self.C, 4)), np.zeros((self.C, 3))
        s = LmSummary()
        tr = np.zeros((trace_rows, 4))
        self.ctx._chk(self.lib.ecb_lm_device_result(self.h, _ptr(i), _ptr(r), _ptr(t), C.byref(s), _ptr(tr), trace_rows))
        out = {f[0]: getattr(s, f[0]) for f in LmSummary._fields_}
        out.update(intrinsics=i, rot_cp=r, trans_cp=t, trace=tr[:int(np.count_nonzero(tr[:, 2]))])
        return out

    def run(self, intr, rot, trans, trace_rows=256):
        """begin + all iterations + result; with fixed_iterations the whole loop is enqueued without a host round trip"""
        self.begin(intr, rot, trans)
        total = self.opt.max_iterations + 1
        if self.opt.fixed_iterations:
            self.iterate(total)
        else:
            done = 0
            while done < total:
                k = min(8, total - done)
                self.iterate(k)
                done += k
                if done < total and not self.running():
                    break
        return self.result(trace_rows)


class LmState:
    """Host-only LM state machine (ecb_lm_*), for callers that all-reduce the normal equations between GPUs."""

    def __init__(self, n_cp, options=None):
        self.lib = load_library()
        self.n_cp = np.ascontiguousarray(np.atleast_1d(n_cp), np.int32)
        self.C = int(self.n_cp.sum())
        self.h = C.c_void_p(self.lib.ecb_lm_create(len(self.n_cp), _ptr(self.n_cp), C.byref(options or lm_options())))

    def __del__(self):
        try:
            self.lib.ecb_lm_destroy(self.h)
        except Exception:
            pass

    def begin(self, intr, rot, trans, packed):
        a = [np.ascontiguousarray(v, np.float64) for v in (intr, rot, trans, packed)]
        return self.lib.ecb_lm_begin(self.h, *[_ptr(v) for v in a])

    def propose(self):
        ci, cr, ct = np.zeros(9), np.zeros((self.C, 4)), np.zeros((self.C, 3))
        st = self.lib.ecb_lm_propose(self.h, _ptr(ci), _ptr(cr), _ptr(ct))
        return st, ci, cr, ct

    def feedback(self, cost):
        return self.lib.ecb_lm_feedback(self.h, float(cost))

    def update(self, packed):
        p = np.ascontiguousarray(packed, np.float64)
        return self.lib.ecb_lm_update(self.h, _ptr(p))

    def state(self):
        i, r, t = np.zeros(9), np.zeros((self.C, 4)), np.zeros((self.C, 3))
        s = LmSummary()
        self.lib.ecb_lm_state(self.h, _ptr(i), _ptr(r), _ptr(t), C.byref(s))
        return i, r, t, {f[0]: getattr(s, f[0]) for f in LmSummary._fields_}

    def trace(self, rows=512):
        tr = np.zeros((rows, 4))
        n = self.lib.ecb_lm_trace(self.h, _ptr(tr), rows)
        return tr[:n]

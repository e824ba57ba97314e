"""TEST INFRASTRUCTURE — numpy restatement of the reference's tracking gate: TrackingBase::process
(core/tracking/src/TrackingBase.cpp:16-46) + EventCalibIni::track (event_camera_calib/src/EventCalibIni.cpp:23-97), with the
line fit done like the reference (last right singular vector of [x y 1], Eigen JacobiSVD -> numpy SVD)."""
import bisect

import numpy as np


class Gate:
    def __init__(self, rows, cols, step):
        self.rows, self.cols, self.step = rows, cols, step
        self.ts, self.frames = [], []

    def _dirs(self, f):
        out = []
        for i in range(self.rows):
            A = np.ones((self.cols, 3))
            A[:, :2] = f[i * self.cols:(i + 1) * self.cols]
            v = np.linalg.svd(A)[2][-1]
            d = np.array([v[1], -v[0]])
            if d @ (A[-1, :2] - A[0, :2]) < 0:
                d = -d
            out.append(d)
        return out

    def process(self, ts, f):
        f = np.asarray(f, float)
        if not self.ts:
            self.ts.append(ts)
            self.frames.append(f)
            return True
        k = bisect.bisect_left(self.ts, ts)           # map::lower_bound
        ref_t, ref = (self.ts[k], self.frames[k]) if k < len(self.ts) else (self.ts[-1], self.frames[-1])
        duration = abs(ts - ref_t)
        th = sorted(np.arccos(np.clip(a @ b / (np.linalg.norm(a) * np.linalg.norm(b)), -1, 1))
                    for a, b in zip(self._dirs(ref), self._dirs(f)))
        if th[len(th) // 2] / duration < (5e-4 * np.pi) / self.step:
            if ts not in self.ts:
                k = bisect.bisect_left(self.ts, ts)
                self.ts.insert(k, ts)
                self.frames.insert(k, f)
            return True
        return False

// Host-only pieces of the C++ façade (include/ecb/event_calib.hpp) for the CPU test suite: txt2bin, spline evaluation,
// TUM trajectory writer.  Links libecb.so (loads without a GPU) but makes no compute call.
#include "../../include/ecb/event_calib.hpp"
using namespace opengv2;
extern "C" {
long long fh_txt2bin(const char *path, double mag) { return EventStream::txt2bin(path, mag); }
int fh_eval(const double *knots, int n_cp, const double *rot, const double *trans, int so3, double t, double *q4, double *t3) {
    std::vector<EventCalibSpline::Segment> seg(1);
    seg[0].knots.assign(knots, knots + n_cp + 4);
    seg[0].rot_cp.assign(rot, rot + 4 * n_cp);
    seg[0].trans_cp.assign(trans, trans + 3 * n_cp);
    return EventCalibSpline::evaluate(seg, so3 != 0, t, q4, t3) ? 1 : 0;
}
int fh_tum(const double *knots, int n_cp, const double *rot, const double *trans, int so3, const char *file, const double *ts, int n) {
    std::vector<EventCalibSpline::Segment> seg(1);
    seg[0].knots.assign(knots, knots + n_cp + 4);
    seg[0].rot_cp.assign(rot, rot + 4 * n_cp);
    seg[0].trans_cp.assign(trans, trans + 3 * n_cp);
    return EventCalibSpline::saveKeyFrameTrajectoryTUM(seg, so3 != 0, file, std::vector<double>(ts, ts + n));
}
}

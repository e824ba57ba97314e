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
// tracking gate (TrackingGate): feed frames one by one; features = n x 2 centres in board order
extern "C" {
void *fh_gate_new(int rows, int cols, double step) { return new TrackingGate(rows, cols, step); }
void fh_gate_free(void *g) { delete (TrackingGate *) g; }
int fh_gate_process(void *g, double ts, const double *xy, int n) {
    std::vector<CalibCircleLite> f((size_t) n);
    for (int i = 0; i < n; ++i) f[(size_t) i] = CalibCircleLite{{{xy[2 * i], xy[2 * i + 1]}}, 1.0, -1, -1};
    return ((TrackingGate *) g)->process(ts, f) ? 1 : 0;
}
}
extern "C" void fh_inverse_radial(const double *k4, double *b5) {
    const auto b = inverseRadialDistortion({k4[0], k4[1], k4[2], k4[3]});
    for (int i = 0; i < 5; ++i) b5[i] = b[(size_t) i];
}
// spline set-up from key frames (EventCalibSpline::segmentsFromKeyframes): returns the number of segments; segment `which`
// is copied out (knots n_cp+4, rot 4 n_cp, trans 3 n_cp); n_cp_out = its control point count
extern "C" int fh_segments(const double *ts, const double *q, const double *t, int n, double step, int which, int *n_cp_out,
                           double *knots, double *rot, double *trans) {
    std::vector<EventCalibSpline::KeyPose> kf((size_t) n);
    for (int i = 0; i < n; ++i) {
        kf[(size_t) i].timeStamp = ts[i];
        for (int a = 0; a < 4; ++a) kf[(size_t) i].unitQwb[a] = q[4 * i + a];
        for (int a = 0; a < 3; ++a) kf[(size_t) i].twb[a] = t[3 * i + a];
    }
    try {
        auto seg = EventCalibSpline::segmentsFromKeyframes(kf, step, false);
        if (which >= 0 && which < (int) seg.size()) {
            const auto &s = seg[(size_t) which];
            *n_cp_out = (int) s.rot_cp.size() / 4;
            std::copy(s.knots.begin(), s.knots.end(), knots);
            std::copy(s.rot_cp.begin(), s.rot_cp.end(), rot);
            std::copy(s.trans_cp.begin(), s.trans_cp.end(), trans);
        }
        return (int) seg.size();
    } catch (const std::logic_error &) {
        return -1;
    }
}
// EventCalibSpline::fitSpline (BsplineReal constructor fit): knots (cpNum + 4) and control points (cpNum x dim)
extern "C" void fh_fit_spline(const double *us, const double *data, int n, int dim, int cpNum, double *knots, double *cp) {
    std::vector<double> u(us, us + n), d(data, data + (size_t) n * dim), k, c;
    EventCalibSpline::fitSpline(u, d, dim, cpNum, k, c);
    std::copy(k.begin(), k.end(), knots);
    std::copy(c.begin(), c.end(), cp);
}
// opengv2::EventCalibIni::replayKeep: the cvCalibration loop replayed over precomputed poses (q = x y z w) and rectify verdicts
extern "C" void fh_replay_keep(const double *ts, const double *q, const double *t, const int *verdict, int n, double step,
                               char *keep, int *counters) {
    std::vector<EventCalibIni::KeyFrame> kf((size_t) n);
    std::vector<const EventCalibIni::KeyFrame *> seq;
    std::vector<int32_t> v(verdict, verdict + n);
    for (int i = 0; i < n; ++i) {
        kf[(size_t) i].timeStamp = ts[i];
        for (int a = 0; a < 4; ++a) kf[(size_t) i].unitQwb[a] = q[4 * i + a];
        for (int a = 0; a < 3; ++a) kf[(size_t) i].twb[a] = t[3 * i + a];
        seq.push_back(&kf[(size_t) i]);
    }
    counters[0] = counters[1] = 0;
    const std::vector<char> k = EventCalibIni::replayKeep(seq, v, step, counters[0], counters[1]);
    for (int i = 0; i < n; ++i) keep[i] = k[(size_t) i];
}

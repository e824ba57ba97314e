// unit_test_eventCameraCalib settingFilePath binFilePath SavePath
//
// Drop-in for the reference CLI (ECC/test/eventCameraCalib.cpp:104-233): same 3 positional arguments, same usage / exit
// codes, same `parameter/event_calibration/example.yaml` keys, same stdout lines ("Events from ... loaded.", "N frames in
// Map.", "Frame t contain n events.", the calibrateCamera report, "N frames in Map after Initialization.", the intrinsics
// before / after the optimisation) and the same SavePath/TrajectoryByEvent.txt — with the window loop running as ONE batched
// GPU launch per lattice instead of hardware_concurrency()-2 CPU threads, rectifyFeatures batched over all key frames, and the
// spline optimisation's residuals / Jacobians / normal equations on the GPU.  Headless (the reference opens a Pangolin
// viewer, :129).  SavePath/image/<timestamp>.png holds every key frame's debug image like :214-227 (cluster colours, medians,
// candidate and rectified circles; include/ecb/image_lite.hpp — ECB_NO_IMAGES=1 skips them).
//
// The window loop is the reference's adaptive one (accept -> jump by length + 5 steps, else grow by one step until the
// window holds FrameEventNumThreshold events or exceeds 3 lengths, then slide; :49-81) over the reference's time pieces.
// A window becomes a key frame when the candidate circles are found, ordered as the rows x cols grid, and pass the tracking
// gate (EventCalibIni::track: row directions vs the neighbouring key frame; the reference's worker threads race on the
// shared map, here the pieces of one round are gated in piece order).  The initialisation (EventCalibIni::cvCalibration)
// runs without OpenCV (include/ecb/calib_init.hpp); the ordered circles of the key frames are also written to
// SavePath/candidates.txt.
//
// Several GPUs (ECB_DEVICES, default: all visible): the reference's time pieces are dealt out to the GPUs in contiguous blocks,
// every GPU holds the records of its block, one host thread + context per GPU (include/ecb/multi_gpu.hpp); the spline
// optimisation shards the residuals the same way and sums the normal equations over NVLink inside the kernels.  The frames,
// candidates and key-frame map do not depend on the number of GPUs.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <sys/stat.h>
#include <thread>
#include <vector>

#include "../../include/ecb/multi_gpu.hpp"

using namespace opengv2;

// the subset of OpenCV FileStorage YAML 1.0 the reference's config uses: `key: scalar` and `key: [ a, b, ... ]`
struct Settings {
    std::map<std::string, std::string> kv;
    bool open(const std::string &path) {
        std::ifstream is(path);
        if (!is.is_open()) return false;
        std::string line;
        while (std::getline(is, line)) {
            const size_t h = line.find('#');
            if (h != std::string::npos) line = line.substr(0, h);
            if (line.rfind("%YAML", 0) == 0 || line.rfind("---", 0) == 0) continue;
            const size_t c = line.find(':');
            if (c == std::string::npos) continue;
            auto trim = [](std::string s) {
                const char *ws = " \t\r\n\"";
                const size_t b = s.find_first_not_of(ws);
                if (b == std::string::npos) return std::string();
                return s.substr(b, s.find_last_not_of(ws) - b + 1);
            };
            const std::string k = trim(line.substr(0, c)), v = trim(line.substr(c + 1));
            if (!k.empty()) kv[k] = v;
        }
        return true;
    }
    bool has(const std::string &k) const { return kv.count(k) && !kv.at(k).empty(); }
    double num(const std::string &k, double dflt = 0) const { return has(k) ? atof(kv.at(k).c_str()) : dflt; }
};

int main(int argc, char **argv) {
    if (argc != 4) {
        std::cerr << std::endl << "Usage: ./unit_test_eventCameraCalib settingFilePath binFilePath SavePath" << std::endl;
        return 1;
    }
    std::cout.precision(8);
    Settings fs;
    if (!fs.open(argv[1])) {
        std::cerr << "Failed to open settings file at: " << argv[1] << std::endl;
        exit(-1);
    }
    auto pattern = std::make_shared<CirclePatternParameters>();
    pattern->cols = (int) fs.num("BoardSize_Cols", 4);
    pattern->rows = (int) fs.num("BoardSize_Rows", 9);
    pattern->squareSize = fs.num("Square_Size", 5.5);
    pattern->isAsymmetric = fs.num("Is_Pattern_Asymmetric", 1) != 0;
    pattern->circleRadius = fs.num("Circles_Radius", 1.75);
    CalibrationSetting setting;  // parameters.hpp:32-46
    setting.circlePatternParameters = pattern;
    setting.aspectRatio = (float) fs.num("Calibrate_FixAspectRatio", 1);
    setting.calibZeroTangentDist = fs.num("Calibrate_AssumeZeroTangentialDistortion", 1) != 0;
    setting.calibFixPrincipalPoint = fs.num("Calibrate_FixPrincipalPointAtTheCenter", 1) != 0;
    setting.useFisheye = fs.num("Calibrate_UseFisheyeModel", 0) != 0;
    setting.fixK1 = fs.num("Fix_K1", 0) != 0;
    setting.fixK2 = fs.num("Fix_K2", 0) != 0;
    setting.fixK3 = fs.num("Fix_K3", 0) != 0;
    setting.fixK4 = fs.num("Fix_K4", 1) != 0;
    setting.fixK5 = fs.num("Fix_K5", 1) != 0;
    setting.NumOfFrameToUse = (int) fs.num("Calibrate_NrOfFrameToUse", 200);
    const double motionTimeStep = fs.num("MotionTimeStep");
    const int width = (int) fs.num("Camera.width"), height = (int) fs.num("Camera.height");
    if (!(motionTimeStep > 0) || width <= 0 || height <= 0) {
        std::cerr << "settings: MotionTimeStep / Camera.width / Camera.height missing" << std::endl;
        exit(-1);
    }

    // load: keep t >= StartTime, stop at the first t >= EndTime (eventCameraCalib.cpp:154-163)
    std::ifstream is(argv[2], std::ifstream::binary | std::ifstream::in);
    if (!is.is_open()) {
        std::cerr << "No such file: " << argv[2] << std::endl;  // EventStream ctor throws invalid_argument (EventStream.cpp:15)
        return 1;
    }
    is.seekg(0, std::ios::end);
    const int64_t n_all = (int64_t) is.tellg() / 25;
    is.seekg(0);
    std::vector<EventRecord> rec((size_t) n_all);
    is.read((char *) rec.data(), n_all * 25);
    const bool customEnd = fs.has("EndTime");
    const double startTime = fs.num("StartTime");
    double endTime = fs.num("EndTime");
    int64_t b = 0, e = n_all;
    while (b < n_all && rec[(size_t) b].t < startTime) ++b;
    if (customEnd)
        for (e = b; e < n_all && rec[(size_t) e].t < endTime; ++e) {}
    if (e <= b) {
        std::cerr << "no events in [StartTime, EndTime)" << std::endl;
        return 1;
    }
    // the reference sets endTime to the last loaded stamp (:165) and cuts [startTime, endTime] into its time pieces (:172-180)
    endTime = rec[(size_t) e - 1].t;

    FrontEnd::Params params;
    params.dbscan_eps = fs.num("dbscan_eps", 4);
    params.dbscan_startMinSample = (int) fs.num("dbscan_startMinSample", 2);
    params.clusterMinSample = (int) fs.num("clusterMinSample", 5);
    params.knn_num = (int) fs.num("knn_num", 3);
    params.fitCircle = fs.num("fitCircle", 0) != 0;
    // The reference's adaptive window loop (MultiProcess::process, eventCameraCalib.cpp:34-97) over its time pieces (:172-180),
    // run as a wavefront: every round evaluates the CURRENT window of every unfinished piece in one batched GPU call, then
    // each piece applies the reference's accept / grow / slide rule to its own result.  The pieces are independent, so the
    // frames found are those of the reference run with one worker per piece.  pieceNum = 5 * (hardware_concurrency() - 2)
    // like the reference (:172-174), or ECB_PIECES.
    const double len = 3 * motionTimeStep, frameGap = 5 * motionTimeStep;
    const int frameEventNumThreshold = (int) fs.num("FrameEventNumThreshold", 4000);
    int pieceNum = 5 * std::max(1, (int) std::thread::hardware_concurrency() - 2);
    if (getenv("ECB_PIECES")) pieceNum = std::max(1, atoi(getenv("ECB_PIECES")));
    const double pstep = (endTime - startTime) / pieceNum;
    struct Piece {
        double lo, hi, first, second;
        bool done;
    };
    std::vector<Piece> pieces;
    for (int k = 0; k < pieceNum; ++k) {
        Piece pc{endTime - pstep * (k + 1), endTime - pstep * k, 0, 0, false};
        pc.first = pc.lo;
        pc.second = pc.lo + len;
        pc.done = !(pc.second < pc.hi);
        pieces.push_back(pc);
    }
    typedef EventCalibIni::KeyFrame Frame;
    std::map<double, Frame> frames;  // MapBase keeps key frames ordered by time stamp
    // GPUs: contiguous blocks of pieces (piece k covers [endTime - (k+1) pstep, endTime - k pstep): ascending time = descending k)
    std::vector<int> devices = ecbDevices();
    if ((int) devices.size() > pieceNum) devices.resize((size_t) pieceNum);
    const int G = (int) devices.size();
    std::vector<double> cuts;
    for (int g = 1; g < G; ++g) cuts.push_back(pieces[(size_t) (pieceNum - (int) ((int64_t) g * pieceNum / G))].hi);
    ShardedEventContainer::Ptr container;
    try {
        container = std::make_shared<ShardedEventContainer>(devices, width, height);
        container->load(rec.data() + b, e - b, cuts);
    } catch (const std::exception &ex) {
        std::cerr << "ecb: " << ex.what() << std::endl;
        return 1;
    }
    std::cout << "Events from " << startTime << " second to " << endTime << " second loaded." << std::endl;
    ShardedFrontEnd fe(container, pattern, params);
    TrackingGate gate(pattern->rows, pattern->cols, motionTimeStep);  // EventCalibIni::track; pieces of one round in piece order
    size_t rounds = 0, evaluated = 0;
    for (;;) {
        std::vector<std::pair<double, double>> windows;
        std::vector<int> owner;
        for (int k = 0; k < pieceNum; ++k)
            if (!pieces[(size_t) k].done) {
                windows.emplace_back(pieces[(size_t) k].first, pieces[(size_t) k].second);
                owner.push_back(k);
            }
        if (windows.empty()) break;
        try {
            fe.run(windows);
        } catch (const std::exception &ex) {
            std::cerr << "ecb: " << ex.what() << std::endl;
            return 1;
        }
        ++rounds;
        evaluated += windows.size();
        for (size_t w = 0; w < windows.size(); ++w) {
            Piece &pc = pieces[(size_t) owner[w]];
            const int events_num = fe.eventsNum(w);
            // extractFeatures() (:56): candidate circles found AND ordered by the grid finder (findCirclesGrid, :332-356), then
            // the tracking gate (tracking->process, :60 -> EventCalibIni::track)
            std::vector<CalibCircleLite> c;
            const double ts0 = (pc.first + pc.second) / 2;
            if (CirclesEventFrame::orderFeatures(fe.candidates(w), *pattern, c) && gate.process(ts0, c)) {  // :56-60
                const double ts = (pc.first + pc.second) / 2;  // Bodyframe time stamp (:57)
                Frame f;
                f.timeStamp = ts;
                f.duration = {pc.first, pc.second};
                f.eventsNum = events_num;
                f.features = c;
                frames.emplace(ts, f);  // MapBase::addFrame through TrackingBase (an existing stamp is kept)
                pc.first = pc.second + frameGap;  // :60-62
                pc.second = pc.first + len;
            } else if (events_num > frameEventNumThreshold || (pc.second - pc.first) > 3 * len) {  // :66-68,75-77
                pc.first += motionTimeStep;
                pc.second = pc.first + len;
            } else {
                pc.second += motionTimeStep;  // :70,79
            }
            pc.done = !(pc.second < pc.hi);  // :50
        }
    }
    mkdir(argv[3], 0755);
    std::ofstream out(std::string(argv[3]) + "/candidates.txt");
    out.precision(17);
    std::cout << frames.size() << " frames in Map." << std::endl;
    for (const auto &kv : frames) {
        const Frame &f = kv.second;
        std::cout << "Frame " << f.timeStamp << " contain " << f.eventsNum << " events." << std::endl;
        for (size_t k = 0; k < f.features.size(); ++k)
            out << f.timeStamp << " " << k << " " << f.features[k].center[0] << " " << f.features[k].center[1] << " "
                << f.features[k].radius << "\n";
    }
    out.close();
    std::cerr << evaluated << " windows evaluated on " << G << " GPU(s) in " << rounds << " batched rounds over " << pieceNum
              << " time pieces" << std::endl;

    // EventCalibIni::cvCalibration (eventCameraCalib.cpp:199): intrinsics, frame poses, checkPose, rectifyFeatures
    EventCalibIni ini(setting, motionTimeStep, width, height);
    bool ok = false;
    try {
        ok = ini.cvCalibration(fe, frames, std::cout);
    } catch (const std::exception &ex) {
        std::cerr << "ecb: " << ex.what() << std::endl;
        return 1;
    }
    std::cout << frames.size() << " frames in Map after Initialization." << std::endl;

    // EventCalibSpline (eventCameraCalib.cpp:203-210): the constructor throws std::logic_error like the reference's
    const bool useSO3 = fs.num("useSO3", 0) != 0;
    if (!ok) {
        std::cerr << "initialisation failed" << std::endl;
        return 1;
    }
    std::vector<EventCalibSpline::KeyPose> poses;
    std::vector<EventCalibSpline::KeyFrame> kfs;
    std::vector<double> stamps;
    for (const auto &kv : frames) {
        EventCalibSpline::KeyPose kp;
        kp.timeStamp = kv.second.timeStamp;
        std::copy(kv.second.unitQwb, kv.second.unitQwb + 4, kp.unitQwb);
        std::copy(kv.second.twb, kv.second.twb + 3, kp.twb);
        poses.push_back(kp);
        kfs.push_back(EventCalibSpline::KeyFrame{kv.second.timeStamp, kv.second.circles});
        stamps.push_back(kv.second.timeStamp);
    }
    for (size_t i = 1; i < poses.size(); ++i) {  // one sign per quaternion so the real-valued spline fit sees a continuous curve
        double d = 0;
        for (int a = 0; a < 4; ++a) d += poses[i].unitQwb[a] * poses[i - 1].unitQwb[a];
        if (d < 0)
            for (int a = 0; a < 4; ++a) poses[i].unitQwb[a] = -poses[i].unitQwb[a];
    }
    try {
        std::vector<EventCalibSpline::Segment> segments = EventCalibSpline::segmentsFromKeyframes(poses, motionTimeStep, useSO3);
        // EventCalibSpline.cpp:94-108: radial part of the OpenCV model -> 5-term inverse polynomial
        const ecb::CameraModel &cam = ini.camera;
        const std::array<double, 4> radial = {cam.dist[0], cam.dist[1], cam.dist[4], 0.0};
        const std::array<double, 5> inv = inverseRadialDistortion(radial);
        const std::array<double, 9> intrinsics = {cam.fx, cam.fy, cam.cx, cam.cy, inv[0], inv[1], inv[2], inv[3], inv[4]};
        if (!getenv("ECB_REFERENCE_SIGNATURES")) {  // (the all-in-one constructor below prints them itself, like the reference's)
            std::cout << "OpenCV Distortion before optimization:";
            printEigenLike(std::cout, radial.data(), 1, 4);
            std::cout << std::endl << "Intrinsics before optimization:";
            printEigenLike(std::cout, intrinsics.data(), 1, 9);
            std::cout << std::endl;
        }
        // distinct devices: residuals sharded like the events; the same device listed several times (tests of the sharding
        // logic on a one-GPU box): the optimisation runs on one context holding the whole stream
        bool distinct = true;
        for (int g = 0; g < G; ++g)
            for (int h = 0; h < g; ++h) distinct = distinct && devices[(size_t) g] != devices[(size_t) h];
        ShardedEventContainer::Ptr opt_events = container;
        if (!distinct) {
            opt_events = std::make_shared<ShardedEventContainer>(std::vector<int>(1, devices[0]), width, height);
            opt_events->load(rec.data() + b, e - b, {});
        }
        if (getenv("ECB_REFERENCE_SIGNATURES")) {
            // The same back half through the reference's OWN argument lists (EventCalibSpline.hpp:19, CirclesEventFrame.hpp:43-65;
            // compat types of include/ecb/compat/): a MapBase of Bodyframes + landmarks + camera, the all-in-one constructor, and
            // for the first key frame CirclesEventFrame::rectifyFeatures(outlierIdxs, Rcw, tcw) / findCenter(Eigen::Vector2d).
            EventContainer::Ptr one = opt_events->shards() == 1 ? opt_events->shard[0] : nullptr;
            if (!one) {
                one = std::make_shared<EventContainer>(devices[0], width, height);
                one->load(rec.data() + b, e - b);
            }
            auto map = std::make_shared<MapBase>();
            map->camera = ini.camera;
            std::vector<LandmarkBase::Ptr> lms;
            const std::vector<double> bp = ini.boardPoints();
            for (size_t k = 0; k + 2 < bp.size(); k += 3) {
                lms.push_back(std::make_shared<LandmarkBase>((int) (k / 3), Eigen::Vector3d(bp[k], bp[k + 1], bp[k + 2])));
                map->addLandmark(lms.back());
            }
            for (const auto &kv : frames) {
                const Frame &f = kv.second;
                auto bf = std::make_shared<Bodyframe>(f.timeStamp, Eigen::Quaterniond(f.unitQwb[3], f.unitQwb[0], f.unitQwb[1], f.unitQwb[2]),
                                                      Eigen::Vector3d(f.twb[0], f.twb[1], f.twb[2]));
                bf->circles = f.circles;
                map->addFrame(bf);
            }
            {   // per-frame class with the reference's signatures on the first key frame
                const Frame &f = frames.begin()->second;
                CirclesEventFrame cf(one, f.duration, pattern, params);
                const bool found = cf.extractFeatures();
                cf.setSensor(ini.camera);
                cf.setLandmarks(lms);
                const Eigen::Quaterniond Qwb(f.unitQwb[3], f.unitQwb[0], f.unitQwb[1], f.unitQwb[2]);
                const Eigen::Matrix3d Rwb = Qwb.toRotationMatrix();
                Eigen::Matrix3d Rcw;
                Eigen::Vector3d tcw;
                for (int r = 0; r < 3; ++r) {
                    for (int c = 0; c < 3; ++c) Rcw(r, c) = Rwb(c, r);
                }
                for (int r = 0; r < 3; ++r) tcw[r] = -(Rcw(r, 0) * f.twb[0] + Rcw(r, 1) * f.twb[1] + Rcw(r, 2) * f.twb[2]);
                const bool ok2 = found && cf.rectifyFeatures(std::unordered_set<int>(), Rcw, tcw);
                double worst = 0;
                size_t alive = 0, fi = 0;
                for (const auto &c : f.circles) {
                    if (c[2] < 0) continue;
                    ++alive;
                    if (fi < cf.features().size()) {
                        const CalibCircleLite &g = cf.features()[fi++];
                        worst = std::max(worst, std::max(std::abs(g.center[0] - c[0]), std::max(std::abs(g.center[1] - c[1]), std::abs(g.radius - c[2]))));
                    }
                }
                int hits = 0;
                for (const auto &g : cf.features()) {
                    const LandmarkBase::Ptr lm = cf.findCenter(Eigen::Vector2d(g.center[0] + g.radius, g.center[1]));
                    hits += lm != nullptr;
                }
                std::cerr << "reference signatures: rectifyFeatures(outlierIdxs, Rcw, tcw) " << (ok2 ? "true" : "false") << ", "
                          << cf.features().size() << " of " << alive << " features, max |diff| to the batched result " << worst
                          << ", findCenter(Eigen::Vector2d) -> landmark for " << hits << " rim points" << std::endl;
            }
            EventCalibSpline spline(map, one, useSO3, fs.num("reduceMap", 0) != 0, motionTimeStep, pattern->circleRadius);
            std::ofstream tum(std::string(argv[3]) + "/TrajectoryByEvent.txt");
            tum << std::fixed;
            for (const auto &kv : map->keyframes()) {  // SystemBase::saveKeyFrameTrajectoryTUM on the updated map
                const Eigen::Quaterniond q = kv.second->unitQwb();
                const Eigen::Vector3d t = kv.second->twb();
                tum << std::setprecision(10) << kv.first << " " << t[0] << " " << t[1] << " " << t[2] << " " << q.x() << " " << q.y() << " "
                    << q.z() << " " << q.w() << std::endl;
            }
            std::cout << "press Enter to exit..." << std::endl;
            std::cin.ignore();
            return 0;
        }
        ShardedCalibSpline spline(opt_events, segments, intrinsics, motionTimeStep, pattern->circleRadius, useSO3);
        std::vector<std::array<double, 3>> landmarks;
        const std::vector<double> board = ini.boardPoints();
        for (size_t k = 0; k + 2 < board.size(); k += 3) landmarks.push_back({board[k], board[k + 1], board[k + 2]});
        const int64_t n_res = spline.associate(kfs, landmarks);
        ecb_lm_summary sum;
        if (!spline.optimize(&sum)) {
            std::cerr << "ecb: " << spline.lastError() << std::endl;
            return 1;
        }
        // in place of ceres::Solver::Summary::FullReport() (EventCalibSpline.cpp:248)
        std::cout << "Solver Summary: residuals " << n_res << ", splines " << segments.size() << ", iterations " << sum.iterations
                  << " (successful " << sum.successful_steps << "), cost " << sum.initial_cost << " -> " << sum.final_cost
                  << ", termination " << sum.termination << std::endl;
        std::cout.precision(12);  // updateMap(), EventCalibSpline.cpp:263-264
        std::cout << "Intrinsics after optimization:";
        printEigenLike(std::cout, spline.intrinsics().data(), 1, 9);
        std::cout << std::endl;
        spline.saveKeyFrameTrajectoryTUM(std::string(argv[3]) + "/TrajectoryByEvent.txt", stamps);  // eventCameraCalib.cpp:212
        if (!getenv("ECB_NO_IMAGES")) {  // :214-227: SavePath/image/<timestamp>.png = cf->image() of every key frame
            const std::string dir = std::string(argv[3]) + "/image/";
            mkdir(dir.c_str(), 0755);
            std::vector<std::pair<double, double>> kw;
            for (const auto &kv : frames) kw.push_back(kv.second.duration);
            fe.run(kw);
            size_t w = 0;
            for (const auto &kv : frames) {
                const ecb::Image8UC3 img = renderFrameImage(fe, w++, &kv.second.circles);
                ecb::write_png(dir + std::to_string(kv.second.timeStamp) + ".png", img);
            }
        }
    } catch (const std::logic_error &ex) {
        std::cerr << "terminate called after throwing an instance of 'std::logic_error'\n  what():  " << ex.what() << std::endl;
        return 134;  // the reference aborts on the uncaught exception
    } catch (const std::exception &ex) {
        std::cerr << "ecb: " << ex.what() << std::endl;
        return 1;
    }
    std::cout << "press Enter to exit..." << std::endl;
    std::cin.ignore();
    return 0;
}

// Drop-in for the reference's `dbscan.h` (DB/include/dbscan.h:42-113): same class name, template signature, `Run`
// arguments, return codes and public result members — the work is done by the CUDA library through the C ABI.
//
//   reference:  template<typename T, typename Float> class DBSCAN final;            dbscan.h:42-43
//               int Run(TVector* V, const uint dim, const Float eps, const uint min, const DistanceFunc& = ...)   :67-68
//               returns 0 SUCCESS / 1 FAILED when V->size()<1 || dim<1 || min<1, never throws                       :121-123
//               public: std::vector<std::vector<uint>> Clusters; std::vector<uint> Noise;                            :92-93
//
// Any T with operator[] convertible to double (dbscan.h:40-41,192), dim = 1 .. 4, any finite coordinates (duplicates
// included) and any eps are clustered on the device: distinct integer pixels with 1 <= eps <= 15 — what
// CirclesEventFrame.cpp:66-70 hands over — on the sensor-plane bitmap kernel, everything else on the general grid-hash path
// (DESIGN.md §2).  `Clusters` (discovery order, members in the reference's BFS pop order) and `Noise` are identical to the
// reference's as ordered lists.  FAILED beyond the reference's own conditions only for dim > 4, NaN / infinite input or
// a missing CUDA device (`last_error()` says which).  `disfunc` is accepted and ignored exactly like the reference's kd-tree
// build (dbscan.h:16,64).
// Thread model: one lazily created context per host thread (the reference runs Run() on hardware_concurrency()-2 threads).
#ifndef ECB_DBSCAN_H
#define ECB_DBSCAN_H

#include <functional>
#include <memory>
#include <string>
#include <vector>

#include "../eventcalib_b200.h"

#if __has_include(<Eigen/StdVector>)
#include <Eigen/Eigen>
#include <Eigen/StdVector>
#define ECB_ALLOC(T) Eigen::aligned_allocator<T>
#else
#define ECB_ALLOC(T) std::allocator<T>
#endif

typedef unsigned int uint;

namespace ecb {
struct CtxDeleter {
    void operator()(ecb_ctx *c) const { ecb_ctx_destroy(c); }
};
// per-thread context on device ECB_DEVICE (default 0)
inline ecb_ctx *thread_context() {
    static thread_local std::unique_ptr<ecb_ctx, CtxDeleter> ctx;
    if (!ctx) {
        ecb_ctx *c = nullptr;
        const char *d = getenv("ECB_DEVICE");
        if (ecb_ctx_create(d ? atoi(d) : 0, nullptr, &c) == ECB_OK) ctx.reset(c);
    }
    return ctx.get();
}
}  // namespace ecb

template <typename T, typename Float>
class DBSCAN final {
    enum ERROR_TYPE { SUCCESS = 0, FAILED, COUNT };
    using TVector = std::vector<T, ECB_ALLOC(T)>;
    using DistanceFunc = std::function<Float(const T &, const T &)>;

public:
    DBSCAN() {}
    ~DBSCAN() {}

    int Run(TVector *V, const uint dim, const Float eps, const uint min,
            const DistanceFunc &disfunc = [](const T &, const T &) -> Float { return 0; }) {
        (void) disfunc;
        if (V->size() < 1) return ERROR_TYPE::FAILED;
        if (dim < 1) return ERROR_TYPE::FAILED;
        if (min < 1) return ERROR_TYPE::FAILED;
        Clusters.clear();
        Noise.clear();
        ecb_ctx *ctx = ecb::thread_context();
        if (!ctx) {
            err_ = "no CUDA device (there is no CPU fallback)";
            return ERROR_TYPE::FAILED;
        }
        const int n = (int) V->size();
        std::vector<double> xy((size_t) dim * (size_t) n);  // kdtree only supports double (dbscan.h:190-193)
        for (int i = 0; i < n; ++i)
            for (uint c = 0; c < dim; ++c) xy[(size_t) dim * i + c] = (double) (*V)[i][c];
        std::vector<int32_t> labels((size_t) n), sizes((size_t) n);
        std::vector<uint32_t> members((size_t) n);
        int32_t nc = 0;
        const int rc = ecb_dbscan_run_nd(ctx, xy.data(), n, (int) dim, (double) eps, min, labels.data(), &nc, sizes.data(),
                                         members.data());
        if (rc != ECB_OK) {
            err_ = ecb_last_error(ctx);
            return ERROR_TYPE::FAILED;
        }
        Clusters.assign((size_t) nc, std::vector<uint>());
        size_t at = 0;
        for (int c = 0; c < nc; ++c) {  // members already in the reference's order (dbscan.h:229-259)
            Clusters[(size_t) c].assign(members.begin() + at, members.begin() + at + (size_t) sizes[(size_t) c]);
            at += (size_t) sizes[(size_t) c];
        }
        for (int i = 0; i < n; ++i)
            if (labels[i] < 0) Noise.push_back((uint) i);
        return ERROR_TYPE::SUCCESS;
    }

    const std::string &last_error() const { return err_; }

public:
    std::vector<std::vector<uint>> Clusters;
    std::vector<uint> Noise;

private:
    std::string err_;
};

#endif  // ECB_DBSCAN_H

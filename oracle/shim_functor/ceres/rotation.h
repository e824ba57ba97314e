// TEST INFRASTRUCTURE.  Stand-in for <ceres/rotation.h> (Ceres is external and absent): only the names the reference header
// mentions outside its templates — the `Create` factories return a ceres::AutoDiffCostFunction, never called here.
#ifndef ECB_ORACLE_CERES_SHIM
#define ECB_ORACLE_CERES_SHIM
namespace ceres {
class CostFunction {
public:
    virtual ~CostFunction() {}
};
template <class Functor, int... Ns>
class AutoDiffCostFunction : public CostFunction {
public:
    explicit AutoDiffCostFunction(Functor *f) : f_(f) {}
    ~AutoDiffCostFunction() override { delete f_; }

private:
    Functor *f_;
};
}  // namespace ceres
#endif

// TEST INFRASTRUCTURE.  Stand-in for <ceres/rotation.h> / <ceres/ceres.h> (Ceres is external and absent).  The reference's
// EventCalibSpline.cpp assembles a ceres::Problem and calls Solve(); here the Problem RECORDS what is added (parameter blocks
// with their local parameterisation, residual blocks with their loss and parameter pointers) so that a test can read the
// assembled structure, and Solve() is a no-op: nothing of Ceres' numerics is restated in this file.
#ifndef ECB_ORACLE_CERES_SHIM
#define ECB_ORACLE_CERES_SHIM
#include <string>
#include <vector>
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
    const Functor *functor() const { return f_; }

private:
    Functor *f_;
};
class LossFunction {
public:
    virtual ~LossFunction() {}
};
class HuberLoss : public LossFunction {
public:
    explicit HuberLoss(double a) : a_(a) {}
    double a_;
};
class LocalParameterization {
public:
    virtual ~LocalParameterization() {}
    // the interface the reference's LocalParameterizationSO3 overrides (BsplineSO3.hpp:190-221)
    virtual bool Plus(const double *, const double *, double *) const { return false; }
    virtual bool ComputeJacobian(const double *, double *) const { return false; }
    virtual int GlobalSize() const { return 0; }
    virtual int LocalSize() const { return 0; }
};
class EigenQuaternionParameterization : public LocalParameterization {};
enum LinearSolverType { DENSE_QR, SPARSE_NORMAL_CHOLESKY };
struct RecordedResidual {
    const CostFunction *cost;
    const LossFunction *loss;
    std::vector<double *> params;
};
struct RecordedParameter {
    double *ptr;
    int size;
    const LocalParameterization *local;
};
class Problem {
public:
    ~Problem() {
        for (auto &r : residuals) delete r.cost;
    }
    void AddParameterBlock(double *values, int size, LocalParameterization *local = nullptr) {
        parameters.push_back(RecordedParameter{values, size, local});
    }
    void SetParameterBlockConstant(double *values) { constant.push_back(values); }
    std::vector<double *> constant;
    template <class... Ps> void AddResidualBlock(CostFunction *cost, LossFunction *loss, Ps... ps) {
        residuals.push_back(RecordedResidual{cost, loss, std::vector<double *>{ps...}});
    }
    std::vector<RecordedResidual> residuals;
    std::vector<RecordedParameter> parameters;
};
namespace Solver {
struct Options {
    double gradient_tolerance = 1e-10, function_tolerance = 1e-6, parameter_tolerance = 1e-8;
    LinearSolverType linear_solver_type = DENSE_QR;
    int num_threads = 1, num_linear_solver_threads = 1, max_num_iterations = 50;
};
struct Summary {
    std::string BriefReport() const { return "ceres stand-in: Solve() is a no-op"; }
    std::string FullReport() const { return "ceres stand-in: Solve() is a no-op (oracle/shim_functor/ceres/rotation.h)"; }
};
}  // namespace Solver
// HOOK: defined by the test wrapper (it copies the recorded problem before the Problem is destroyed)
void Solve(const Solver::Options &options, Problem *problem, Solver::Summary *summary);
}  // namespace ceres
#endif

// TEST STUB (tests/stubs): VertexSE3Expmap / EdgeSE3Expmap / SE3Quat as far as the BA adapter reads them (types_six_dof_expmap.h:59-142).
#pragma once
#include "../core/block_solver.h"
namespace g2o {
class SE3Quat {
public:
    SE3Quat() { for (int i = 0; i < 7; i++) v_(i) = (i == 6) ? 1.0 : 0.0; }
    explicit SE3Quat(const Vector7d& v) : v_(v) {}
    Vector7d toVector() const { return v_; }  // x y z qx qy qz qw (se3quat.h:140-151)
private:
    Vector7d v_;
};
template <int D, class T>
class BaseVertex : public OptimizableGraph::Vertex {
public:
    int dimension() const { return D; }
    const T& estimate() const { return _estimate; }
    void setEstimate(const T& e) { _estimate = e; }
protected:
    T _estimate;
};
template <int D, class E, class VI, class VJ>
class BaseBinaryEdge : public OptimizableGraph::Edge {
public:
    typedef Matrix<double, D, D> InformationType;
    BaseBinaryEdge() { _vertices.resize(2); for (int i = 0; i < D; i++) _information(i, i) = 1.0; }
    const E& measurement() const { return _measurement; }
    void setMeasurement(const E& m) { _measurement = m; }
    const InformationType& information() const { return _information; }
    InformationType& information() { return _information; }
    virtual void mapHessianMemory(double* d, int, int, bool rowMajor) { _hessian = d; _hessianRowMajor = rowMajor; }  // base_binary_edge.hpp:207-218
protected:
    E _measurement;
    InformationType _information;
    double* _hessian = nullptr;      // protected in g2o as well: the adapter must not need it
    bool _hessianRowMajor = false;
};
class VertexSE3Expmap : public BaseVertex<6, SE3Quat> {};
class EdgeSE3Expmap : public BaseBinaryEdge<6, SE3Quat, VertexSE3Expmap, VertexSE3Expmap> {};
}  // namespace g2o

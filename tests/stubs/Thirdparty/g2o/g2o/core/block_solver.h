// TEST STUB (tests/stubs): a miniature of the g2o interfaces the BA adapter (integration/object_slam/cuboid_block_solver_b200.h) touches --
// OptimizableGraph::Vertex / Edge, SparseOptimizer, SparseBlockMatrix, BlockSolverX with its protected _Hpp / _Hll / _Hpl -- with the same
// names, signatures and block-mapping rules as object_slam/Thirdparty/g2o/g2o/core/{optimizable_graph.h, sparse_block_matrix.h,
// block_solver.h, block_solver.hpp:142-295}, so that the adapter compiles and its block copies can be run against the C ABI.  NOT g2o.
#pragma once
#include <Eigen/Core>
#include <algorithm>
#include <map>
#include <utility>
#include <vector>
namespace g2o {
using namespace Eigen;
typedef Matrix<double, 7, 1> Vector7d;
typedef Matrix<double, 10, 1> Vector10d;

class OptimizableGraph {
public:
    class Vertex {
    public:
        virtual ~Vertex() {}
        virtual int dimension() const = 0;
        int hessianIndex() const { return _hessianIndex; }
        void setHessianIndex(int i) { _hessianIndex = i; }
        bool fixed() const { return _fixed; }
        void setFixed(bool f) { _fixed = f; }
        bool marginalized() const { return _marginalized; }
        void setMarginalized(bool m) { _marginalized = m; }
        int colInHessian() const { return _colInHessian; }
        void setColInHessian(int c) { _colInHessian = c; }
        void mapHessianMemory(double* d) { _hessian = d; }
        double* hessianData() { return _hessian; }
        double* bData() { return _b.data(); }
        int copyB(double* b) const { std::copy(_b.begin(), _b.end(), b); return (int)_b.size(); }
        void allocB() { _b.assign(dimension(), 0.0); }
    protected:
        int _hessianIndex = -1, _colInHessian = -1;
        bool _fixed = false, _marginalized = false;
        double* _hessian = nullptr;
        std::vector<double> _b;
    };
    class Edge {
    public:
        virtual ~Edge() {}
        Vertex* vertex(size_t i) const { return _vertices[i]; }
        const std::vector<Vertex*>& vertices() const { return _vertices; }
        void setVertex(size_t i, Vertex* v) { if (_vertices.size() <= i) _vertices.resize(i + 1); _vertices[i] = v; }
        virtual void mapHessianMemory(double* d, int i, int j, bool rowMajor) = 0;
    protected:
        std::vector<Vertex*> _vertices;
    };
};

class SparseOptimizer {
public:
    typedef std::vector<OptimizableGraph::Edge*> EdgeContainer;
    typedef std::vector<OptimizableGraph::Vertex*> VertexContainer;
    const EdgeContainer& activeEdges() const { return _activeEdges; }
    const VertexContainer& indexMapping() const { return _ivMap; }
    EdgeContainer _activeEdges;
    VertexContainer _ivMap;
};

template <class MatrixType>
class SparseBlockMatrix {
public:
    ~SparseBlockMatrix() { for (auto& kv : _blocks) delete kv.second; }
    MatrixType* block(int r, int c, bool alloc = false) {
        auto it = _blocks.find(std::make_pair(r, c));
        if (it != _blocks.end()) return it->second;
        if (!alloc) return 0;
        MatrixType* m = new MatrixType(_rowDim[r], _colDim[c]);
        _blocks[std::make_pair(r, c)] = m;
        return m;
    }
    void clear() { for (auto& kv : _blocks) kv.second->setZero(); }
    std::vector<int> _rowDim, _colDim;
    std::map<std::pair<int, int>, MatrixType*> _blocks;
};

template <class M> class LinearSolver { public: virtual ~LinearSolver() {} };

class Solver {
public:
    virtual ~Solver() {}
    virtual bool buildStructure(bool zeroBlocks = false) = 0;
    virtual bool buildSystem() = 0;
    void setOptimizer(SparseOptimizer* o) { _optimizer = o; }
    double* b() { return _b; }
protected:
    SparseOptimizer* _optimizer = nullptr;
    double* _b = nullptr;
};

// BlockSolver< BlockSolverTraits<Dynamic, Dynamic> >: poses (not marginalized) first, then landmarks (marginalized)
class BlockSolverX : public Solver {
public:
    typedef MatrixXd PoseMatrixType;
    typedef MatrixXd LandmarkMatrixType;
    typedef MatrixXd PoseLandmarkMatrixType;
    typedef LinearSolver<PoseMatrixType> LinearSolverType;
    explicit BlockSolverX(LinearSolverType* ls) : _linearSolver(ls) {}
    ~BlockSolverX() { delete _Hpp; delete _Hll; delete _Hpl; delete _linearSolver; }
    // block_solver.hpp:142-295: hessian indices in indexMapping() order (poses, then landmarks), diagonal blocks mapped into the vertices,
    // one upper-triangle block per edge mapped into the edge (transposed when the edge's vertex order is the other way round)
    virtual bool buildStructure(bool zeroBlocks = false) {
        delete _Hpp; delete _Hll; delete _Hpl;
        _Hpp = new SparseBlockMatrix<PoseMatrixType>(); _Hll = new SparseBlockMatrix<LandmarkMatrixType>(); _Hpl = new SparseBlockMatrix<PoseLandmarkMatrixType>();
        _numPoses = 0; _numLandmarks = 0;
        int col = 0;
        for (auto* v : _optimizer->indexMapping()) {
            if (!v->marginalized()) { _Hpp->_rowDim.push_back(v->dimension()); _Hpp->_colDim.push_back(v->dimension()); _Hpl->_rowDim.push_back(v->dimension()); _numPoses++; }
            else { _Hll->_rowDim.push_back(v->dimension()); _Hll->_colDim.push_back(v->dimension()); _Hpl->_colDim.push_back(v->dimension()); _numLandmarks++; }
            v->setColInHessian(col); col += v->dimension(); v->allocB();
        }
        _bstore.assign(col, 0.0); _b = _bstore.data();
        int pi = 0, li = 0;
        for (auto* v : _optimizer->indexMapping()) {
            if (!v->marginalized()) { v->setHessianIndex(pi); v->mapHessianMemory(_Hpp->block(pi, pi, true)->data()); pi++; }
            else { v->setHessianIndex(_numPoses + li); v->mapHessianMemory(_Hll->block(li, li, true)->data()); li++; }
        }
        for (auto* e : _optimizer->activeEdges()) {
            auto* v1 = e->vertex(0); auto* v2 = e->vertex(1);
            int ind1 = v1->hessianIndex(), ind2 = v2->hessianIndex();
            if (ind1 == -1 || ind2 == -1) continue;
            bool transposedBlock = ind1 > ind2;
            if (transposedBlock) std::swap(ind1, ind2);
            if (!v1->marginalized() && !v2->marginalized()) e->mapHessianMemory(_Hpp->block(ind1, ind2, true)->data(), 0, 1, transposedBlock);
            else if (v1->marginalized() && v2->marginalized()) e->mapHessianMemory(_Hll->block(ind1 - _numPoses, ind2 - _numPoses, true)->data(), 0, 1, false);
            else if (v1->marginalized()) e->mapHessianMemory(_Hpl->block(v2->hessianIndex(), v1->hessianIndex() - _numPoses, true)->data(), 0, 1, true);
            else e->mapHessianMemory(_Hpl->block(v1->hessianIndex(), v2->hessianIndex() - _numPoses, true)->data(), 0, 1, false);
        }
        (void)zeroBlocks;
        return true;
    }
    virtual bool buildSystem() { return false; }
protected:
    SparseBlockMatrix<PoseMatrixType>* _Hpp = nullptr;
    SparseBlockMatrix<LandmarkMatrixType>* _Hll = nullptr;
    SparseBlockMatrix<PoseLandmarkMatrixType>* _Hpl = nullptr;
    LinearSolverType* _linearSolver = nullptr;
    int _numPoses = 0, _numLandmarks = 0;
    std::vector<double> _bstore;
};
}  // namespace g2o

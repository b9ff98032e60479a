// g2o batch hook for the BA half: a BlockSolver whose buildSystem() linearises every EdgeSE3Cuboid / EdgeSE3CuboidProj /
// EdgeSE3Expmap of the graph on the GPU in one call and writes the blocks into the Hessian memory g2o mapped in
// buildStructure().  The vertex / edge classes of object_slam/include/object_slam/g2o_Object.h stay untouched (they still
// define computeError()/oplusImpl(), which g2o uses for chi2 and the update step); only the per-edge virtual
// linearizeOplus()+constructQuadraticForm() loop of block_solver.hpp:501-560 is replaced.
//
// Usage in object_slam/src/main_obj.cpp:512-517 (one changed line):
//     g2o::BlockSolverX* solver_ptr = new g2o::CuboidBlockSolverB200(linearSolver);
//
// The vendored g2o stays UNPATCHED: the off-diagonal blocks are looked up in the solver's own _Hpp / _Hpl / _Hll (protected members of
// BlockSolver, the very blocks buildStructure() mapped into the edges, block_solver.hpp:222-249).
// Compiled by tests/test_adapters_compile.py against the interface stubs under tests/stubs/ (the build container has no Eigen / g2o build).
#pragma once

#include <algorithm>
#include <map>
#include <stdexcept>
#include <vector>

#include "Thirdparty/g2o/g2o/core/block_solver.h"
#include "Thirdparty/g2o/g2o/types/types_six_dof_expmap.h"
#include "object_slam/g2o_Object.h"

#include "cubeslam_b200.h"

namespace g2o {

class CuboidBlockSolverB200 : public BlockSolverX {
public:
    explicit CuboidBlockSolverB200(LinearSolverType* linearSolver) : BlockSolverX(linearSolver) {
        if (csb_create(&ctx_, 0) != CSB_OK) throw std::runtime_error("CuboidBlockSolverB200: no usable CUDA device (there is no CPU fallback)");
    }
    ~CuboidBlockSolverB200() { csb_destroy(ctx_); }

    // buildStructure() of the base class allocates Hpp blocks and maps them into vertices (diagonal) and edges (off-diagonal);
    // afterwards the graph topology, measurements and information matrices go to the device once.
    virtual bool buildStructure(bool zeroBlocks = false)
    {
        if (!BlockSolverX::buildStructure(zeroBlocks)) return false;
        cams_.clear(); cubes_.clear(); cam_fixed_.clear(); cube_fixed_.clear();
        ec_.clear(); ep_.clear(); eo_.clear();
        std::map<const OptimizableGraph::Vertex*, int> cam_id, cube_id;
        auto cam_of = [&](OptimizableGraph::Vertex* v) {
            auto it = cam_id.find(v);
            if (it != cam_id.end()) return it->second;
            int id = (int)cams_.size(); cam_id[v] = id; cams_.push_back(static_cast<VertexSE3Expmap*>(v)); cam_fixed_.push_back(v->fixed()); return id; };
        auto cube_of = [&](OptimizableGraph::Vertex* v) {
            auto it = cube_id.find(v);
            if (it != cube_id.end()) return it->second;
            int id = (int)cubes_.size(); cube_id[v] = id; cubes_.push_back(static_cast<VertexCuboid*>(v)); cube_fixed_.push_back(v->fixed()); return id; };
        // activeEdges() is sorted by edge id (sparse_optimizer.cpp:482-487); the C ABI accumulates ec, ep, eo in array order, which
        // equals that order for main_obj.cpp's id scheme (cuboid edges: frame index, odometry edges: N + frame index).
        std::vector<int32_t> ec_cam, ec_cube, ep_cam, ep_cube, eo_i, eo_j;
        std::vector<double> ec_meas, ec_info, ep_meas, ep_info, ep_K, eo_meas, eo_info;
        for (OptimizableGraph::Edge* e : _optimizer->activeEdges()) {
            auto* v0 = static_cast<OptimizableGraph::Vertex*>(e->vertex(0));
            auto* v1 = static_cast<OptimizableGraph::Vertex*>(e->vertex(1));
            if (auto* c = dynamic_cast<EdgeSE3Cuboid*>(e)) {
                ec_.push_back(c); ec_cam.push_back(cam_of(v0)); ec_cube.push_back(cube_of(v1));
                Vector10d m = c->measurement().toVector();
                ec_meas.insert(ec_meas.end(), m.data(), m.data() + 10);
                for (int r = 0; r < 9; r++) for (int k = 0; k < 9; k++) ec_info.push_back(c->information()(r, k));
            } else if (auto* p = dynamic_cast<EdgeSE3CuboidProj*>(e)) {
                ep_.push_back(p); ep_cam.push_back(cam_of(v0)); ep_cube.push_back(cube_of(v1));
                for (int k = 0; k < 4; k++) ep_meas.push_back(p->measurement()(k));
                for (int r = 0; r < 4; r++) for (int k = 0; k < 4; k++) ep_info.push_back(p->information()(r, k));
                for (int r = 0; r < 3; r++) for (int k = 0; k < 3; k++) ep_K.push_back(p->Kalib(r, k));
            } else if (auto* o = dynamic_cast<EdgeSE3Expmap*>(e)) {
                eo_.push_back(o); eo_i.push_back(cam_of(v0)); eo_j.push_back(cam_of(v1));
                Vector7d m = o->measurement().toVector();
                eo_meas.insert(eo_meas.end(), m.data(), m.data() + 7);
                for (int r = 0; r < 6; r++) for (int k = 0; k < 6; k++) eo_info.push_back(o->information()(r, k));
            } else
                return false;  // an edge type this solver does not know: fall back to a stock BlockSolverX in the caller
        }
        csb_ba_graph g;
        g.n_cam = (int)cams_.size(); g.n_cube = (int)cubes_.size(); g.cam_fixed = cam_fixed_.data(); g.cube_fixed = cube_fixed_.data();
        g.n_ec = (int)ec_.size(); g.ec_cam = ec_cam.data(); g.ec_cube = ec_cube.data(); g.ec_meas = ec_meas.data(); g.ec_info = ec_info.data();
        g.n_ep = (int)ep_.size(); g.ep_cam = ep_cam.data(); g.ep_cube = ep_cube.data(); g.ep_meas = ep_meas.data(); g.ep_info = ep_info.data(); g.ep_K = ep_K.data();
        g.n_eo = (int)eo_.size(); g.eo_cam_i = eo_i.data(); g.eo_cam_j = eo_j.data(); g.eo_meas = eo_meas.data(); g.eo_info = eo_info.data();
        // Online mode (main_obj.cpp:738-803 adds one keyframe per frame and optimises again): when the graph is the previous one plus ONE
        // camera with its edges (and landmarks first seen by it), only that frame goes to the device (csb_ba_add_frame); anything else --
        // also any change to an old measurement -- rebuilds the device graph.
        bool done = false;
        if (have_graph_ && grown_by_one_frame(ec_cam, ec_cube, ec_meas, ec_info, ep_cam, eo_i, eo_j, eo_meas, eo_info)) {
            const size_t oc = ec_cam_.size(), oo = eo_i_.size(), oq = n_cube_prev_;
            csb_ba_frame f = {};
            Vector7d cv = cams_.back()->estimate().toVector();
            f.cam7 = cv.data(); f.cam_fixed = cam_fixed_.back();
            std::vector<double> nq;
            for (size_t i = oq; i < cubes_.size(); i++) { Vector10d v = cubes_[i]->estimate().toVector(); nq.insert(nq.end(), v.data(), v.data() + 10); }
            f.n_new_cubes = (int)(cubes_.size() - oq); f.new_cubes10 = nq.data(); f.new_cube_fixed = cube_fixed_.data() + oq;
            f.n_ec = (int)(ec_cam.size() - oc); f.ec_cube = ec_cube.data() + oc; f.ec_meas = ec_meas.data() + 10 * oc; f.ec_info = ec_info.data() + 81 * oc;
            f.n_eo = (int)(eo_i.size() - oo); f.eo_cam_i = eo_i.data() + oo; f.eo_meas = eo_meas.data() + 7 * oo; f.eo_info = eo_info.data() + 36 * oo;
            int32_t idx = -1;
            if (csb_ba_add_frame(ctx_, &f, &idx) == CSB_OK && idx == (int)cams_.size() - 1) { done = true; n_incremental_++; }
        }
        if (!done && csb_ba_set_graph(ctx_, &g) != CSB_OK) { have_graph_ = false; return false; }
        ec_cam_ = ec_cam; ec_cube_ = ec_cube; ep_cam_ = ep_cam; ep_cube_ = ep_cube; eo_i_ = eo_i; eo_j_ = eo_j;
        ec_meas_ = ec_meas; ec_info_ = ec_info; eo_meas_ = eo_meas; eo_info_ = eo_info;
        n_cam_prev_ = cams_.size(); n_cube_prev_ = cubes_.size(); cam_fixed_prev_ = cam_fixed_; cube_fixed_prev_ = cube_fixed_;
        have_graph_ = true;
        return true;
    }

    // how many buildStructure() calls were served by csb_ba_add_frame
    int incrementalUpdates() const { return n_incremental_; }

    // One GPU linearisation instead of the per-edge loop; then copy blocks into the memory g2o mapped (Eigen blocks are
    // column-major, like the C ABI's) and gather b, exactly as block_solver.hpp:546-557 does.
    virtual bool buildSystem()
    {
        const size_t nc = cams_.size(), nq = cubes_.size();
        std::vector<double> cams7(7 * nc), cubes10(10 * nq);
        for (size_t i = 0; i < nc; i++) { Vector7d v = cams_[i]->estimate().toVector(); std::copy(v.data(), v.data() + 7, &cams7[7 * i]); }
        for (size_t i = 0; i < nq; i++) { Vector10d v = cubes_[i]->estimate().toVector(); std::copy(v.data(), v.data() + 10, &cubes10[10 * i]); }
        H_cam_.resize(36 * nc); b_cam_.resize(6 * nc); H_cube_.resize(81 * nq); b_cube_.resize(9 * nq);
        ec_Hij_.resize(54 * ec_.size()); ep_Hij_.resize(54 * ep_.size()); eo_Hij_.resize(36 * eo_.size());
        csb_ba_output out = {};
        out.H_cam = H_cam_.data(); out.b_cam = b_cam_.data(); out.H_cube = H_cube_.data(); out.b_cube = b_cube_.data();
        out.ec_Hij = ec_Hij_.data(); out.ep_Hij = ep_Hij_.data(); out.eo_Hij = eo_Hij_.data();
        if (csb_ba_linearize(ctx_, cams7.data(), cubes10.data(), &out) != CSB_OK) return false;
        _Hpp->clear();
        for (size_t i = 0; i < nc; i++) if (!cams_[i]->fixed()) {
            std::copy(&H_cam_[36 * i], &H_cam_[36 * i] + 36, cams_[i]->hessianData());
            std::copy(&b_cam_[6 * i], &b_cam_[6 * i] + 6, cams_[i]->bData());
        }
        for (size_t i = 0; i < nq; i++) if (!cubes_[i]->fixed()) {
            std::copy(&H_cube_[81 * i], &H_cube_[81 * i] + 81, cubes_[i]->hessianData());
            std::copy(&b_cube_[9 * i], &b_cube_[9 * i] + 9, cubes_[i]->bData());
        }
        // off-diagonal blocks: written where buildStructure() mapped them for the edges (block_solver.hpp:222-249)
        for (size_t e = 0; e < ec_.size(); e++) writeOffDiagonal(ec_[e], &ec_Hij_[54 * e], 6, 9);
        for (size_t e = 0; e < ep_.size(); e++) writeOffDiagonal(ep_[e], &ep_Hij_[54 * e], 6, 9);
        for (size_t e = 0; e < eo_.size(); e++) writeOffDiagonal(eo_[e], &eo_Hij_[36 * e], 6, 6);
        for (size_t i = 0; i < _optimizer->indexMapping().size(); ++i) {
            OptimizableGraph::Vertex* v = _optimizer->indexMapping()[i];
            v->copyB(_b + v->colInHessian());
        }
        return true;
    }

private:
    template <class T>
    static bool is_prefix(const std::vector<T>& a, const std::vector<T>& b) { return a.size() <= b.size() && std::equal(a.begin(), a.end(), b.begin()); }
    // the marshalled graph = the one on the device + exactly one camera, whose cuboid edges start at it and whose odometry edges end at it
    bool grown_by_one_frame(const std::vector<int32_t>& ec_cam, const std::vector<int32_t>& ec_cube, const std::vector<double>& ec_meas,
                            const std::vector<double>& ec_info, const std::vector<int32_t>& ep_cam, const std::vector<int32_t>& eo_i,
                            const std::vector<int32_t>& eo_j, const std::vector<double>& eo_meas, const std::vector<double>& eo_info) const
    {
        if (cams_.size() != n_cam_prev_ + 1 || cubes_.size() < n_cube_prev_ || !ep_cam.empty() || !ep_cam_.empty()) return false;
        if (!is_prefix(cam_fixed_prev_, cam_fixed_) || !is_prefix(cube_fixed_prev_, cube_fixed_)) return false;
        if (!is_prefix(ec_cam_, ec_cam) || !is_prefix(ec_cube_, ec_cube) || !is_prefix(ec_meas_, ec_meas) || !is_prefix(ec_info_, ec_info)) return false;
        if (!is_prefix(eo_i_, eo_i) || !is_prefix(eo_j_, eo_j) || !is_prefix(eo_meas_, eo_meas) || !is_prefix(eo_info_, eo_info)) return false;
        const int32_t nc = (int32_t)n_cam_prev_;
        for (size_t e = ec_cam_.size(); e < ec_cam.size(); e++) if (ec_cam[e] != nc) return false;
        for (size_t e = eo_i_.size(); e < eo_i.size(); e++) if (eo_j[e] != nc || eo_i[e] >= nc) return false;
        return true;
    }

    // blk = A^T Omega B of the edge (di x dj, column-major; A belongs to vertex 0).  The block g2o accumulates it into is found the way
    // buildStructure() allocated it: upper block triangle of _Hpp for two poses (stored transposed when vertex 0 has the larger
    // hessian index), _Hll for two marginalised vertices, _Hpl (pose row, landmark column) for a mixed pair.
    void writeOffDiagonal(OptimizableGraph::Edge* e, const double* blk, int di, int dj)
    {
        OptimizableGraph::Vertex* v1 = static_cast<OptimizableGraph::Vertex*>(e->vertex(0));
        OptimizableGraph::Vertex* v2 = static_cast<OptimizableGraph::Vertex*>(e->vertex(1));
        int ind1 = v1->hessianIndex(), ind2 = v2->hessianIndex();
        if (ind1 == -1 || ind2 == -1) return;  // a vertex is fixed: no block
        bool transposed = ind1 > ind2;
        if (transposed) std::swap(ind1, ind2);
        double* dst = 0;
        if (!v1->marginalized() && !v2->marginalized()) dst = _Hpp->block(ind1, ind2, false)->data();
        else if (v1->marginalized() && v2->marginalized()) { dst = _Hll->block(ind1 - _numPoses, ind2 - _numPoses, false)->data(); transposed = false; }
        else if (v1->marginalized()) { dst = _Hpl->block(v2->hessianIndex(), v1->hessianIndex() - _numPoses, false)->data(); transposed = true; }
        else { dst = _Hpl->block(v1->hessianIndex(), v2->hessianIndex() - _numPoses, false)->data(); transposed = false; }
        if (!transposed) std::copy(blk, blk + di * dj, dst);                                                  // di x dj, column-major
        else for (int r = 0; r < di; r++) for (int c = 0; c < dj; c++) dst[r * dj + c] = blk[c * di + r];       // the dj x di transpose, column-major
    }

    csb_context* ctx_ = nullptr;
    std::vector<VertexSE3Expmap*> cams_;
    std::vector<VertexCuboid*> cubes_;
    std::vector<int32_t> cam_fixed_, cube_fixed_, ec_cam_, ec_cube_, ep_cam_, ep_cube_, eo_i_, eo_j_;
    std::vector<EdgeSE3Cuboid*> ec_;
    std::vector<EdgeSE3CuboidProj*> ep_;
    std::vector<EdgeSE3Expmap*> eo_;
    std::vector<double> H_cam_, b_cam_, H_cube_, b_cube_, ec_Hij_, ep_Hij_, eo_Hij_;
    // what the device graph was built from (the incremental path compares against it)
    std::vector<double> ec_meas_, ec_info_, eo_meas_, eo_info_;
    std::vector<int32_t> cam_fixed_prev_, cube_fixed_prev_;
    size_t n_cam_prev_ = 0, n_cube_prev_ = 0;
    bool have_graph_ = false;
    int n_incremental_ = 0;
};

}  // namespace g2o

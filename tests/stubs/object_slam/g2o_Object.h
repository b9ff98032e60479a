// TEST STUB (tests/stubs): g2o::cuboid, VertexCuboid, EdgeSE3Cuboid, EdgeSE3CuboidProj as far as the BA adapter reads them
// (object_slam/include/object_slam/g2o_Object.h:23-292).
#pragma once
#include "Thirdparty/g2o/g2o/types/types_six_dof_expmap.h"
namespace g2o {
class cuboid {
public:
    cuboid() { for (int i = 0; i < 10; i++) v_(i) = (i == 6 || i >= 7) ? 1.0 : 0.0; }
    explicit cuboid(const Vector10d& v) : v_(v) {}
    Vector10d toVector() const { return v_; }  // x y z qx qy qz qw sx sy sz (g2o_Object.h:143-150)
private:
    Vector10d v_;
};
class VertexCuboid : public BaseVertex<9, cuboid> {};
class EdgeSE3Cuboid : public BaseBinaryEdge<9, cuboid, VertexSE3Expmap, VertexCuboid> {};
class EdgeSE3CuboidProj : public BaseBinaryEdge<4, Vector4d, VertexSE3Expmap, VertexCuboid> {
public:
    Matrix3d Kalib;
};
}  // namespace g2o

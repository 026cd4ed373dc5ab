/* TEST INFRASTRUCTURE.  A small functional Eigen stand-in (3x3 / 4x4 float matrices, 3- and 6-vectors, block views) for
 * compiling the reference's OWN SE(3) exponential -- SkewMatrix, WedgeSkew, ExpSO3 (lib/transforms/xregRotUtils.cpp) and
 * ExpSE3 (lib/transforms/xregRigidUtils.cpp) -- which the product's xrc_exp_se3 (caller-side glue)
 * mirrors.  Conventions: element-wise f32, products as left-to-right sums, no FMA (compiled with -ffp-contract=off);
 * `a * B * C` associates left to right as in C++.  The comparison made with it is a tolerance (1e-6), not bit equality. */
#ifndef XREG_REF_PIN_SE3_PRELUDE_H
#define XREG_REF_PIN_SE3_PRELUDE_H

#include <cassert>
#include <cmath>
#include <cstddef>


#include "ref_pin_eigen_small.h"

namespace xreg
{
using CoordScalar = float;
using Pt3 = Eigen::M<3, 1>;
using Pt6 = Eigen::M<6, 1>;
using Mat3x3 = Eigen::M<3, 3>;
using Mat4x4 = Eigen::M<4, 4>;

Mat3x3 SkewMatrix(const Pt3& v);
Pt3 WedgeSkew(const Mat3x3& W);
Mat3x3 ExpSO3(const Pt3& x);
Mat3x3 ExpSO3(const Mat3x3& W);
Mat4x4 ExpSE3(const Mat4x4& M);
Mat4x4 ExpSE3(const Pt6& x);
}  // namespace xreg

#endif

// idocp_b200_compat.hpp -- lets a program written against idocp's headers compile UNCHANGED against idocp_b200:
//   g++ -std=c++17 -I<repo>/include/idocp_b200/compat -I<repo>/include  examples/anymal/anymal_trotting.cpp  -lidocp_b200
// The directory holds forwarding headers with the reference's include paths ("idocp/robot/robot.hpp", "idocp/ocp/ocp_solver.hpp",
// "idocp/cost/...", "idocp/constraints/...", "idocp/utils/...", "idocp/unocp/...") and a minimal "Eigen/Core"; all of them
// include this file, which maps the namespaces: `idocp::X` is `idocp_b200::X`, `Eigen::VectorXd / Vector3d / Matrix3d` are the
// dense stand-ins of idocp_b200.hpp (comma initialiser, Zero, Constant, coeffRef).  With a real Eigen on the include path
// define IDOCP_B200_NO_EIGEN_SHIM and convert at the call sites instead.
// tests/test_cpp_host.py::test_reference_examples_compile_unchanged builds the reference's own example sources this way.
#ifndef IDOCP_B200_COMPAT_HPP_
#define IDOCP_B200_COMPAT_HPP_

#include "idocp_b200/ocp_solver.hpp"

namespace idocp = idocp_b200;

#ifndef IDOCP_B200_NO_EIGEN_SHIM
namespace Eigen {
using VectorXd = idocp_b200::VectorXd;
using Vector3d = idocp_b200::Vector3d;
using Matrix3d = idocp_b200::Matrix3d;
}  // namespace Eigen
// the user's TimeVaryingTaskSpace6DRefBase::compute_q_6d_ref(t, pinocchio::SE3&) (examples/iiwa14/task_space_ocp.cpp:21-46)
namespace pinocchio {
using SE3 = idocp_b200::SE3;
}  // namespace pinocchio
#endif

#endif  // IDOCP_B200_COMPAT_HPP_

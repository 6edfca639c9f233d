// anymal_ocp_benchmark.cpp -- the problem of the reference's examples/anymal/ocp_benchmark.cpp on the batched GPU engine:
// ANYmal standing on four feet, ConfigurationSpaceCost + ContactForceCost, the six joint limits and the NONLINEAR
// FrictionCone (src/constraints/friction_cone.cpp), T = 0.5, N = 20; `batch` copies of the example's state.  An optional
// third argument adds JointAcceleration{Lower,Upper}Limit with that bound.
//   g++ -std=c++17 -Iinclude examples/anymal_ocp_benchmark.cpp -Lidocp_b200 -lidocp_b200 -Wl,-rpath,$PWD/idocp_b200 -o build/anymal_ocp_benchmark
// (the reference's own source also compiles unchanged: -Iinclude/idocp_b200/compat)
#include <cstdlib>
#include <iomanip>
#include <iostream>
#include <memory>
#include <string>

#include "idocp_b200/ocp_solver.hpp"

namespace idocp = idocp_b200;

int main(int argc, char* argv[]) {
  const int batch = argc > 1 ? std::atoi(argv[1]) : 1;
  const int num_iteration = argc > 2 ? std::atoi(argv[2]) : 10;
  const double a_limit = argc > 3 ? std::atof(argv[3]) : 0.0;
  const char* urdf_env = std::getenv("IDOCP_B200_ANYMAL_URDF");
  idocp::Robot robot(urdf_env ? urdf_env : "", {14, 24, 34, 44});   // LF, LH, RF, RH

  idocp::VectorXd q_ref = {0, 0, 0.4792, 0, 0, 0, 1, -0.1, 0.7, -1.0, -0.1, -0.7, 1.0, 0.1, 0.7, -1.0, 0.1, -0.7, 1.0};
  auto cost = std::make_shared<idocp::CostFunction>();
  auto config_cost = std::make_shared<idocp::ConfigurationSpaceCost>(robot);
  config_cost->set_q_ref(q_ref);
  config_cost->set_v_ref(idocp::VectorXd::Zero(robot.dimv()));
  config_cost->set_q_weight(idocp::VectorXd::Constant(robot.dimv(), 10));
  config_cost->set_qf_weight(idocp::VectorXd::Constant(robot.dimv(), 10));
  config_cost->set_v_weight(idocp::VectorXd::Constant(robot.dimv(), 1));
  config_cost->set_vf_weight(idocp::VectorXd::Constant(robot.dimv(), 1));
  config_cost->set_a_weight(idocp::VectorXd::Constant(robot.dimv(), 0.01));
  auto contact_cost = std::make_shared<idocp::ContactForceCost>(robot);
  contact_cost->set_f_weight(std::vector<idocp::Vector3d>(4, idocp::Vector3d(0.001, 0.001, 0.001)));
  contact_cost->set_f_ref(std::vector<idocp::Vector3d>(4, idocp::Vector3d(0, 0, 70)));
  cost->push_back(config_cost);
  cost->push_back(contact_cost);

  auto constraints = std::make_shared<idocp::Constraints>();
  constraints->push_back(std::make_shared<idocp::JointPositionLowerLimit>(robot));
  constraints->push_back(std::make_shared<idocp::JointPositionUpperLimit>(robot));
  constraints->push_back(std::make_shared<idocp::JointVelocityLowerLimit>(robot));
  constraints->push_back(std::make_shared<idocp::JointVelocityUpperLimit>(robot));
  constraints->push_back(std::make_shared<idocp::JointTorquesLowerLimit>(robot));
  constraints->push_back(std::make_shared<idocp::JointTorquesUpperLimit>(robot));
  constraints->push_back(std::make_shared<idocp::FrictionCone>(robot, 0.7));
  if (a_limit > 0.0) {
    constraints->push_back(std::make_shared<idocp::JointAccelerationLowerLimit>(robot, idocp::VectorXd::Constant(12, -a_limit)));
    constraints->push_back(std::make_shared<idocp::JointAccelerationUpperLimit>(robot, idocp::VectorXd::Constant(12, a_limit)));
  }

  idocp::OCPSolver ocp_solver(robot, cost, constraints, 0.5, 20, 4, 4, batch);

  const double t = 0;
  idocp::VectorXd q = q_ref;
  idocp::VectorXd v = idocp::VectorXd::Zero(robot.dimv());
  auto contact_status = robot.createContactStatus();
  contact_status.activateContacts({0, 1, 2, 3});
  robot.updateFrameKinematics(q);
  robot.setContactPoints(contact_status);
  ocp_solver.setContactStatusUniformly(contact_status);
  ocp_solver.setSolution("q", q);
  ocp_solver.setSolution("v", v);
  ocp_solver.setSolution("f", idocp::Vector3d(0, 0, 0.25 * robot.totalWeight()));
  ocp_solver.initConstraints(t);

  std::cout << std::setprecision(17);
  idocp::ocpbenchmarker::Convergence(ocp_solver, t, q, v, num_iteration, false);
  idocp::ocpbenchmarker::CPUTime(ocp_solver, t, q, v, 50, false);
  return 0;
}

// Twin of the reference's examples/iiwa14/unocp_benchmark.cpp on the B200 engine: same problem,
// same calls, namespace idocp -> idocp_b200.  Optional argument: batch size (default 1).
//   g++ -std=c++17 -Iinclude examples/unocp_benchmark.cpp -Lidocp_b200 -lidocp_b200 -Wl,-rpath,$PWD/idocp_b200
#include <iomanip>
#include <iostream>
#include <memory>
#include <string>

#include "idocp_b200/idocp_b200.hpp"

int main(int argc, char* argv[]) {
  namespace idocp = idocp_b200;
  const int batch = argc > 1 ? std::atoi(argv[1]) : 1;
  const int num_iteration_time = argc > 2 ? std::atoi(argv[2]) : 1000;
  // Create a robot.
  const std::string path_to_urdf = "../iiwa_description/urdf/iiwa14.urdf";
  idocp::Robot robot(path_to_urdf);

  // Create a cost function.
  robot.setJointEffortLimit(idocp::VectorXd::Constant(robot.dimu(), 200));
  auto cost = std::make_shared<idocp::CostFunction>();
  auto config_cost = std::make_shared<idocp::ConfigurationSpaceCost>(robot);
  config_cost->set_q_ref(idocp::VectorXd::Constant(robot.dimv(), -5));
  config_cost->set_v_ref(idocp::VectorXd::Constant(robot.dimv(), -9));
  config_cost->set_q_weight(idocp::VectorXd::Constant(robot.dimv(), 10));
  config_cost->set_qf_weight(idocp::VectorXd::Constant(robot.dimv(), 10));
  config_cost->set_v_weight(idocp::VectorXd::Constant(robot.dimv(), 0.1));
  config_cost->set_vf_weight(idocp::VectorXd::Constant(robot.dimv(), 0.1));
  config_cost->set_a_weight(idocp::VectorXd::Constant(robot.dimv(), 0.01));
  config_cost->set_u_weight(idocp::VectorXd::Constant(robot.dimv(), 0.0));
  cost->push_back(config_cost);

  // Create joint constraints.
  idocp::JointConstraintsFactory constraints_factory(robot);
  auto constraints = constraints_factory.create();

  // Create the OCP solver for unconstrained rigid-body systems.
  const double T = 1;
  const int N = 20;
  const int nthreads = 4;
  const double t = 0;
  const idocp::VectorXd q = idocp::VectorXd::Constant(robot.dimq(), 2);
  const idocp::VectorXd v = idocp::VectorXd::Zero(robot.dimv());
  idocp::UnOCPSolver ocp_solver(robot, cost, constraints, T, N, nthreads, batch);

  // Solves the OCP.
  ocp_solver.setSolution("q", q);
  ocp_solver.setSolution("v", v);
  const int num_iteration = 50;
  const bool line_search = false;
  std::cout << std::setprecision(17);
  idocp::ocpbenchmarker::Convergence(ocp_solver, t, q, v, num_iteration, line_search);
  idocp::ocpbenchmarker::CPUTime(ocp_solver, t, q, v, num_iteration_time, line_search);
  return 0;
}

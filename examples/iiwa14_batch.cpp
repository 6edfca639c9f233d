// iiwa14_batch.cpp -- drives the batched engine through the C++ host layer (the reference's class API):
//
//   iiwa14_batch <problem> <solver> [batch] [iterations] [timing-iterations] [devices, e.g. 0,1]
//     problem : benchmark | config | task | task6d   (cost / limits of the reference's three iiwa14 set-ups;
//               task6d = the task set-up with the constant-reference TaskSpace6DCost)
//     solver  : unocp | unparnmpc
//
// Prints the KKT-error history of instance 0 (ocpbenchmarker::Convergence) and the time per batched
// updateSolution (ocpbenchmarker::CPUTime).  Build:
//   g++ -std=c++17 -Iinclude examples/iiwa14_batch.cpp -Lidocp_b200 -lidocp_b200 -Wl,-rpath,$PWD/idocp_b200
#include <cmath>
#include <cstring>
#include <iomanip>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <type_traits>
#include <vector>

#include "idocp_b200/ocp_solver.hpp"   // idocp_b200.hpp + the constraint component tag classes (JointAcceleration*Limit)

namespace ob = idocp_b200;

namespace {

struct Setup {
  double T;
  int N;
  ob::VectorXd q0;
  std::shared_ptr<ob::CostFunction> cost;
};

// circle in the y-z plane traced by the end effector, orientation fixed
class CircleReference final : public ob::TimeVaryingTaskSpace6DRefBase {
 public:
  void compute_q_6d_ref(const double t, ob::SE3& ref) const override {
    ob::Matrix3d R;
    R(0, 0) = 0; R(0, 1) = 0; R(0, 2) = 1;
    R(1, 0) = 0; R(1, 1) = 1; R(1, 2) = 0;
    R(2, 0) = -1; R(2, 1) = 0; R(2, 2) = 0;
    ref = ob::SE3(R, ob::Vector3d(0.546, 0.1 * std::sin(M_PI * t), 0.76 + 0.1 * std::cos(M_PI * t)));
  }
};

ob::VectorXd bent_arm(bool start_with_bend) {
  ob::VectorXd q(7);
  for (int j = 0; j < 7; ++j) q[j] = ((j % 2 == 0) == start_with_bend) ? M_PI_2 : 0.0;
  return q;
}

Setup make_setup(const std::string& problem, ob::Robot& robot) {
  Setup s;
  s.cost = std::make_shared<ob::CostFunction>();
  auto joint_cost = std::make_shared<ob::ConfigurationSpaceCost>(robot);
  const int n = robot.dimv();
  if (problem == "benchmark") {            // unreachable reference far outside the limits: every constraint becomes active
    robot.setJointEffortLimit(ob::VectorXd::Constant(n, 200));
    joint_cost->set_q_ref(ob::VectorXd::Constant(n, -5));
    joint_cost->set_v_ref(ob::VectorXd::Constant(n, -9));
    joint_cost->set_q_weight(ob::VectorXd::Constant(n, 10));
    joint_cost->set_qf_weight(ob::VectorXd::Constant(n, 10));
    joint_cost->set_v_weight(ob::VectorXd::Constant(n, 0.1));
    joint_cost->set_vf_weight(ob::VectorXd::Constant(n, 0.1));
    joint_cost->set_a_weight(ob::VectorXd::Constant(n, 0.01));
    s.cost->push_back(joint_cost);
    s.T = 1; s.N = 20;
    s.q0 = ob::VectorXd::Constant(n, 2);
  } else if (problem == "config") {        // swing between two bent postures
    robot.setJointEffortLimit(ob::VectorXd::Constant(n, 50));
    robot.setJointVelocityLimit(ob::VectorXd::Constant(n, M_PI_2));
    joint_cost->set_q_ref(bent_arm(false));
    joint_cost->set_q_weight(ob::VectorXd::Constant(n, 10));
    joint_cost->set_qf_weight(ob::VectorXd::Constant(n, 10));
    joint_cost->set_v_weight(ob::VectorXd::Constant(n, 0.01));
    joint_cost->set_vf_weight(ob::VectorXd::Constant(n, 0.01));
    joint_cost->set_a_weight(ob::VectorXd::Constant(n, 0.01));
    s.cost->push_back(joint_cost);
    s.T = 3; s.N = 60;
    s.q0 = bent_arm(true);
  } else if (problem == "task") {          // end-effector circle, frame 22
    robot.setJointEffortLimit(ob::VectorXd::Constant(n, 50));
    robot.setJointVelocityLimit(ob::VectorXd::Constant(n, M_PI_2));
    joint_cost->set_v_weight(ob::VectorXd::Constant(n, 0.01));
    joint_cost->set_vf_weight(ob::VectorXd::Constant(n, 0.01));
    joint_cost->set_a_weight(ob::VectorXd::Constant(n, 0.01));
    s.cost->push_back(joint_cost);
    auto task = std::make_shared<ob::TimeVaryingTaskSpace6DCost>(robot, 22, std::make_shared<CircleReference>());
    task->set_q_6d_weight(ob::Vector3d::Constant(1000), ob::Vector3d::Constant(1000));
    task->set_qf_6d_weight(ob::Vector3d::Constant(1000), ob::Vector3d::Constant(1000));
    s.cost->push_back(task);
    s.T = 6; s.N = 120;
    s.q0 = bent_arm(false);
  } else if (problem == "task6d") {        // the same set-up with TaskSpace6DCost: a fixed end-effector placement
    robot.setJointEffortLimit(ob::VectorXd::Constant(n, 50));
    robot.setJointVelocityLimit(ob::VectorXd::Constant(n, M_PI_2));
    joint_cost->set_v_weight(ob::VectorXd::Constant(n, 0.01));
    joint_cost->set_vf_weight(ob::VectorXd::Constant(n, 0.01));
    joint_cost->set_a_weight(ob::VectorXd::Constant(n, 0.01));
    s.cost->push_back(joint_cost);
    auto task = std::make_shared<ob::TaskSpace6DCost>(robot, 22);
    ob::SE3 ref;
    CircleReference().compute_q_6d_ref(0.0, ref);   // sin(0), cos(0): no libm rounding between this file and the oracle
    task->set_q_6d_ref(ref.translation, ref.rotation);
    task->set_q_6d_weight(ob::Vector3d::Constant(1000), ob::Vector3d::Constant(1000));
    task->set_qf_6d_weight(ob::Vector3d::Constant(1000), ob::Vector3d::Constant(1000));
    s.cost->push_back(task);
    s.T = 6; s.N = 120;
    s.q0 = bent_arm(false);
  } else if (problem == "task3d") {        // reach a fixed end-effector POSITION: TaskSpace3DCost
    robot.setJointEffortLimit(ob::VectorXd::Constant(n, 50));
    robot.setJointVelocityLimit(ob::VectorXd::Constant(n, M_PI_2));
    joint_cost->set_v_weight(ob::VectorXd::Constant(n, 0.01));
    joint_cost->set_vf_weight(ob::VectorXd::Constant(n, 0.01));
    joint_cost->set_a_weight(ob::VectorXd::Constant(n, 0.01));
    s.cost->push_back(joint_cost);
    auto task = std::make_shared<ob::TaskSpace3DCost>(robot, 22);
    task->set_q_3d_ref(ob::Vector3d(0.546, 0.1, 0.76));
    task->set_q_3d_weight(ob::Vector3d::Constant(1000));
    task->set_qf_3d_weight(ob::Vector3d::Constant(1000));
    s.cost->push_back(task);
    s.T = 1.5; s.N = 30;
    s.q0 = bent_arm(false);
  } else {
    std::cerr << "unknown problem '" << problem << "' (benchmark | config | task | task6d | task3d)\n";
    std::exit(EXIT_FAILURE);
  }
  return s;
}

template <typename Solver>
void run(Solver& solver, const Setup& s, int iterations, int timing_iterations) {
  const ob::VectorXd v0 = ob::VectorXd::Zero(7);
  const double t = 0;
  solver.setSolution("q", s.q0);
  solver.setSolution("v", v0);
  std::cout << std::setprecision(17);
  ob::ocpbenchmarker::Convergence(solver, t, s.q0, v0, iterations, false);
  if (timing_iterations > 0) ob::ocpbenchmarker::CPUTime(solver, t, s.q0, v0, timing_iterations, false);
  if constexpr (std::is_base_of<ob::detail::SolverBase, Solver>::value) {
    if (const char* dir = std::getenv("IDOCP_B200_SAVE_DIR")) {   // trajectory export in the reference's text format
      for (const char* name : {"q", "v", "a", "u"}) solver.saveSolution(std::string(dir) + "/" + name + ".dat", name);
      solver.printSolution("u");
    }
  }
}

}  // namespace

int main(int argc, char* argv[]) {
  if (argc < 3) {
    std::cerr << "usage: iiwa14_batch <benchmark|config|task|task6d> <unocp|unparnmpc> [batch] [iterations] [timing-iterations]\n";
    return EXIT_FAILURE;
  }
  const std::string problem = argv[1], kind = argv[2];
  const int batch = argc > 3 ? std::atoi(argv[3]) : 1;
  const int iterations = argc > 4 ? std::atoi(argv[4]) : 30;
  const int timing = argc > 5 ? std::atoi(argv[5]) : 0;
  std::vector<int> devices;   // e.g. "0,1,2,3": the batch sharded over several GPUs of this node (default: device 0)
  if (argc > 6) {
    std::stringstream list(argv[6]);
    for (std::string item; std::getline(list, item, ',');) devices.push_back(std::atoi(item.c_str()));
  }
  if (devices.empty()) devices.push_back(0);
  // a URDF path (the reference: "../iiwa_description/urdf/iiwa14.urdf") is verified against the compiled-in model
  const char* urdf_env = std::getenv("IDOCP_B200_IIWA14_URDF");
  ob::Robot robot(urdf_env ? urdf_env : "");
  Setup s = make_setup(problem, robot);
  ob::JointConstraintsFactory factory(robot);
  auto constraints = factory.create();
  // optional: JointAccelerationLowerLimit / JointAccelerationUpperLimit (src/constraints/joint_acceleration_*_limit.cpp) with
  // amin = -limit, amax = +limit on every joint, pushed on top of the factory's six components
  if (const char* acc = std::getenv("IDOCP_B200_ACC_LIMIT")) {
    const double limit = std::atof(acc);
    constraints->push_back(std::make_shared<ob::JointAccelerationLowerLimit>(robot, ob::VectorXd::Constant(7, -limit)));
    constraints->push_back(std::make_shared<ob::JointAccelerationUpperLimit>(robot, ob::VectorXd::Constant(7, limit)));
  }
  const int nthreads = 1;   // accepted for source compatibility; the batch runs on the GPU
  if (kind == "unocp") {
    ob::UnOCPSolver solver(robot, s.cost, constraints, s.T, s.N, nthreads, batch, devices);
    run(solver, s, iterations, timing);
  } else if (kind == "unparnmpc") {
    ob::UnParNMPCSolver solver(robot, s.cost, constraints, s.T, s.N, nthreads, batch, devices);
    solver.setSolution("q", s.q0);
    solver.setSolution("v", ob::VectorXd::Zero(7));
    solver.initBackwardCorrection(0.0);
    run(solver, s, iterations, timing);
  } else if (kind == "ocp") {
    // the general class names on the fixed-base robot (examples/iiwa14/ocp_benchmark.cpp, parnmpc_benchmark.cpp): they forward
    // to the specialised solvers
    ob::OCPSolver solver(robot, s.cost, constraints, s.T, s.N, 0, nthreads, batch);
    run(solver, s, iterations, timing);
  } else if (kind == "parnmpc") {
    ob::ParNMPCSolver solver(robot, s.cost, constraints, s.T, s.N, 0, nthreads, batch);
    solver.setSolution("q", s.q0);
    solver.setSolution("v", ob::VectorXd::Zero(7));
    solver.initBackwardCorrection(0.0);
    run(solver, s, iterations, timing);
  } else {
    std::cerr << "unknown solver '" << kind << "' (unocp | unparnmpc | ocp | parnmpc)\n";
    return EXIT_FAILURE;
  }
  return 0;
}

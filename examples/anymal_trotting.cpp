// anymal_trotting.cpp -- the reference's examples/anymal/anymal_trotting.cpp on the batched GPU engine.  The body of
// main() is the reference's, with the namespace changed, QuadrupedRobot for Robot(path, contact_frames), the Hybrid*
// containers for CostFunction / Constraints, and a `batch` argument (all instances start from the same state here).
//   g++ -std=c++17 -Iinclude examples/anymal_trotting.cpp -Lidocp_b200 -lidocp_b200 -Wl,-rpath,$PWD/idocp_b200 -o build/anymal_trotting
#include <cstdlib>
#include <iomanip>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>

#include "idocp_b200/ocp_solver.hpp"

namespace idocp = idocp_b200;

int main(int argc, char* argv[]) {
  const int batch = argc > 1 ? std::atoi(argv[1]) : 1;
  std::vector<int> contact_frames = {14, 24, 34, 44};  // LF, LH, RF, RH
  // the reference passes "../anymal_b_simple_description/urdf/anymal.urdf"; a path given here is verified against the compiled-in model
  const char* urdf_env = std::getenv("IDOCP_B200_ANYMAL_URDF");
  const std::string path_to_urdf = urdf_env ? urdf_env : "";
  idocp::QuadrupedRobot robot(path_to_urdf, contact_frames);

  const double step_length = 0.15;
  const double t_start = 0.5;
  const double t_period = 0.5;

  auto cost = std::make_shared<idocp::HybridCostFunction>();
  idocp::VectorXd q_standing = {0, 0, 0.4792, 0, 0, 0, 1, -0.1, 0.7, -1.0, -0.1, -0.7, 1.0, 0.1, 0.7, -1.0, 0.1, -0.7, 1.0};
  idocp::VectorXd q_weight = idocp::VectorXd::Constant(robot.dimv(), 10);
  idocp::VectorXd v_weight = {1, 1, 1, 1, 1, 1, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1};
  idocp::VectorXd a_weight = {0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 0.01, 0.01, 0.01, 0.01, 0.01, 0.01, 0.01, 0.01, 0.01, 0.01, 0.01, 0.01};
  idocp::TrottingSwingAngles swing_angles;
  swing_angles.front_swing_knee = 1.7;
  swing_angles.hip_swing_knee = 1.7;
  auto config_cost = std::make_shared<idocp::TrottingConfigurationSpaceCost>(robot);
  config_cost->set_ref(t_start, t_period, q_standing, step_length, swing_angles);
  config_cost->set_q_weight(q_weight);
  config_cost->set_qf_weight(q_weight);
  config_cost->set_qi_weight(q_weight);
  config_cost->set_v_weight(v_weight);
  config_cost->set_vf_weight(v_weight);
  config_cost->set_vi_weight(v_weight);
  config_cost->set_a_weight(a_weight);
  config_cost->set_dvi_weight(a_weight);
  cost->push_back(config_cost);

  auto contact_cost = std::make_shared<idocp::ContactForceCost>(robot);
  std::vector<idocp::Vector3d> f_weight(contact_frames.size(), idocp::Vector3d(0.001, 0.001, 0.001));
  contact_cost->set_f_weight(f_weight);
  contact_cost->set_fi_weight(f_weight);
  contact_cost->set_f_ref(robot);
  cost->push_back(contact_cost);

  auto constraints = std::make_shared<idocp::HybridConstraints>();
  const double mu = 0.7;
  constraints->push_back(std::make_shared<idocp::JointPositionLowerLimit>(robot));
  constraints->push_back(std::make_shared<idocp::JointPositionUpperLimit>(robot));
  constraints->push_back(std::make_shared<idocp::JointVelocityLowerLimit>(robot));
  constraints->push_back(std::make_shared<idocp::JointVelocityUpperLimit>(robot));
  constraints->push_back(std::make_shared<idocp::JointTorquesLowerLimit>(robot));
  constraints->push_back(std::make_shared<idocp::JointTorquesUpperLimit>(robot));
  constraints->push_back(std::make_shared<idocp::LinearizedFrictionCone>(robot, mu));
  constraints->push_back(std::make_shared<idocp::LinearizedImpulseFrictionCone>(robot, mu));
  // optional (IDOCP_B200_CONTACT_DISTANCE=1 | 2): ContactDistance, literally as in the reference or with consistent derivatives
  if (const char* cd = std::getenv("IDOCP_B200_CONTACT_DISTANCE"))
    constraints->push_back(std::make_shared<idocp::ContactDistance>(robot, std::atoi(cd) == 2));

  // 2 steps
  const double T = 1.55;  // t_start + max_num_impulse_phase * t_period + 0.05
  const int N = 30;
  const int max_num_impulse_phase = 2;
  const int nthreads = 4;
  const double t = 0;
  // IDOCP_B200_DEVICES="0,1,...": the batch sharded over several GPUs by the one solver object (default: device 0)
  std::vector<int> devices;
  if (const char* dl = std::getenv("IDOCP_B200_DEVICES")) {
    std::stringstream list(dl);
    for (std::string item; std::getline(list, item, ',');) devices.push_back(std::atoi(item.c_str()));
  }
  idocp::OCPSolver ocp_solver = devices.empty()
      ? idocp::OCPSolver(robot, cost, constraints, T, N, max_num_impulse_phase + 1, nthreads, batch)
      : idocp::OCPSolver(robot, cost, constraints, T, N, max_num_impulse_phase + 1, nthreads, batch, devices);

  robot.updateFrameKinematics(q_standing);
  std::vector<idocp::Vector3d> contact_points(robot.maxPointContacts());
  robot.getContactPoints(contact_points);
  auto contact_status_initial = robot.createContactStatus();
  contact_status_initial.activateContacts({0, 1, 2, 3});
  contact_status_initial.setContactPoints(idocp::toPoints(contact_points));
  ocp_solver.setContactStatusUniformly(contact_status_initial);

  auto contact_status_even = robot.createContactStatus();
  contact_status_even.activateContacts({1, 2});
  contact_status_even.setContactPoints(idocp::toPoints(contact_points));
  ocp_solver.pushBackContactStatus(contact_status_even, t_start);

  auto contact_status_odd = robot.createContactStatus();
  contact_points[0].coeffRef(0) += 0.5 * step_length;
  contact_points[3].coeffRef(0) += 0.5 * step_length;
  contact_status_odd.activateContacts({0, 3});
  contact_status_odd.setContactPoints(idocp::toPoints(contact_points));
  ocp_solver.pushBackContactStatus(contact_status_odd, t_start + t_period);

  for (int i = 2; i <= max_num_impulse_phase; ++i) {
    if (i % 2 == 0) {
      contact_points[1].coeffRef(0) += step_length;
      contact_points[2].coeffRef(0) += step_length;
      contact_status_even.setContactPoints(idocp::toPoints(contact_points));
      ocp_solver.pushBackContactStatus(contact_status_even, t_start + i * t_period);
    } else {
      contact_points[0].coeffRef(0) += step_length;
      contact_points[3].coeffRef(0) += step_length;
      contact_status_odd.setContactPoints(idocp::toPoints(contact_points));
      ocp_solver.pushBackContactStatus(contact_status_odd, t_start + i * t_period);
    }
  }

  idocp::VectorXd q = q_standing;
  idocp::VectorXd v = idocp::VectorXd::Zero(robot.dimv());
  ocp_solver.setSolution("q", q);
  ocp_solver.setSolution("v", v);
  idocp::Vector3d f_init(0, 0, 0.25 * robot.totalWeight());
  ocp_solver.setSolution("f", f_init);

  ocp_solver.initConstraints(t);

  const bool line_search = false;
  std::cout << std::setprecision(17);
  idocp::ocpbenchmarker::Convergence(ocp_solver, t, q, v, 25, line_search);
  idocp::ocpbenchmarker::CPUTime(ocp_solver, t, q, v, 20, line_search);
  const auto qs = ocp_solver.getSolution("q");
  std::cout << "base x at the end of the horizon: " << qs.back()[0] << std::endl;
  if (argc > 2 && std::string(argv[2]) == "mpc") {
    // a few control ticks of the batch of MPC loops (idocp_b200::BatchedMPC): the measured state stays at the standing
    // posture here; the first phase is dropped once its switching time (0.5) has passed
    idocp::BatchedMPC mpc(ocp_solver, 2);
    std::vector<double> qb(static_cast<size_t>(batch) * robot.dimq()), vb(static_cast<size_t>(batch) * robot.dimv(), 0.0),
        u0(static_cast<size_t>(batch) * robot.dimu());
    for (int b = 0; b < batch; ++b)
      for (int j = 0; j < robot.dimq(); ++j) qb[static_cast<size_t>(b) * robot.dimq() + j] = q[j];
    for (const double tick : {0.0, 0.2, 0.45, 0.52, 0.6}) {
      mpc.tick(tick, qb.data(), vb.data(), u0.data());
      std::cout << "MPC tick t = " << tick << ": u0[0] = " << u0[0] << ", popped phases = " << mpc.numPoppedPhases() << std::endl;
    }
  }
  return 0;
}

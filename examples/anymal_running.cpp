// anymal_running.cpp -- the reference's examples/anymal/anymal_running.cpp (running gait with flight phases: T = 7,
// N = 240, 26 touch-downs, 14 lift-offs, TimeVaryingConfigurationSpaceCost) on the batched GPU engine.
//   g++ -std=c++17 -Iinclude examples/anymal_running.cpp -Lidocp_b200 -lidocp_b200 -Wl,-rpath,$PWD/idocp_b200 -o build/anymal_running
//   build/anymal_running [batch] [iterations] [line_search 0|1]
#include <cstdlib>
#include <iomanip>
#include <iostream>
#include <memory>
#include <string>

#include "idocp_b200/ocp_solver.hpp"

namespace idocp = idocp_b200;

int main(int argc, char* argv[]) {
  const int batch = argc > 1 ? std::atoi(argv[1]) : 1;
  const int iterations = argc > 2 ? std::atoi(argv[2]) : 350;
  const bool line_search = argc > 3 ? std::atoi(argv[3]) != 0 : false;
  std::vector<int> contact_frames = {14, 24, 34, 44};  // LF, LH, RF, RH
  // the reference passes "../anymal_b_simple_description/urdf/anymal.urdf"; a path given here is verified against the compiled-in model
  const char* urdf_env = std::getenv("IDOCP_B200_ANYMAL_URDF");
  idocp::QuadrupedRobot robot(urdf_env ? urdf_env : "", contact_frames);

  const double stride = 0.4;
  const double additive_stride_hip = 0.2;
  const double t_start = 1.0;
  const double t_front_swing = 0.135;
  const double t_front_hip_swing = 0.05;
  const double t_hip_swing = 0.165;
  const double t_period = t_front_swing + t_front_hip_swing + t_hip_swing;
  const int steps = 10;

  auto cost = std::make_shared<idocp::HybridCostFunction>();
  idocp::VectorXd q_standing = {-3, 0, 0.4792, 0, 0, 0, 1, -0.1, 0.7, -1.0, -0.1, -0.7, 1.0, 0.1, 0.7, -1.0, 0.1, -0.7, 1.0};
  idocp::VectorXd q_weight = {1, 1, 1, 10, 10, 10, 10, 10, 10, 10, 10, 10, 10, 10, 10, 10, 10, 10};
  idocp::VectorXd v_weight = {0.01, 0.01, 0.01, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1};
  idocp::VectorXd a_weight = idocp::VectorXd::Constant(robot.dimv(), 0.01);
  auto config_cost = std::make_shared<idocp::TimeVaryingConfigurationSpaceCost>(robot);
  idocp::VectorXd v_ref = idocp::VectorXd::Zero(robot.dimv());
  v_ref[0] = stride / t_period;
  config_cost->set_ref(robot, t_start, t_start + (0.5 + steps) * t_period, q_standing, v_ref);
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
  std::vector<idocp::Vector3d> f_weight(4, idocp::Vector3d(1e-01, 1e-01, 1.0e-07)), f_ref(4, idocp::Vector3d(0, 0, 70));
  contact_cost->set_f_weight(f_weight);
  contact_cost->set_fi_weight(f_weight);
  contact_cost->set_f_ref(f_ref);
  cost->push_back(contact_cost);

  auto constraints = std::make_shared<idocp::HybridConstraints>();
  const double mu = 0.8;
  constraints->push_back(std::make_shared<idocp::JointPositionLowerLimit>(robot));
  constraints->push_back(std::make_shared<idocp::JointPositionUpperLimit>(robot));
  constraints->push_back(std::make_shared<idocp::JointVelocityLowerLimit>(robot));
  constraints->push_back(std::make_shared<idocp::JointVelocityUpperLimit>(robot));
  constraints->push_back(std::make_shared<idocp::JointTorquesLowerLimit>(robot));
  constraints->push_back(std::make_shared<idocp::JointTorquesUpperLimit>(robot));
  constraints->push_back(std::make_shared<idocp::LinearizedFrictionCone>(robot, mu));
  constraints->push_back(std::make_shared<idocp::LinearizedImpulseFrictionCone>(robot, mu));

  const double T = 7;
  const int N = 240;
  const int max_num_impulse_phase = (steps + 3) * 2;
  const int nthreads = 4;
  const double t = 0;
  idocp::OCPSolver ocp_solver(robot, cost, constraints, T, N, max_num_impulse_phase, nthreads, batch);

  robot.updateFrameKinematics(q_standing);
  std::vector<idocp::Vector3d> contact_points(robot.maxPointContacts());
  robot.getContactPoints(contact_points);
  auto contact_status_initial = robot.createContactStatus();
  contact_status_initial.activateContacts({0, 1, 2, 3});
  auto contact_status_front_swing = robot.createContactStatus();
  contact_status_front_swing.activateContacts({1, 3});
  auto contact_status_hip_swing = robot.createContactStatus();
  contact_status_hip_swing.activateContacts({0, 2});
  auto contact_status_front_hip_swing = robot.createContactStatus();

  contact_status_initial.setContactPoints(idocp::toPoints(contact_points));
  ocp_solver.setContactStatusUniformly(contact_status_initial);

  const double t_initial_front_swing = 0.125, t_initial_front_hip_swing = 0.05, t_initial_hip_swing = 0.125;
  const double t_initial = t_initial_front_swing + t_initial_front_hip_swing + t_initial_hip_swing;
  const double t_initial_front_swing2 = 0.135, t_initial_front_hip_swing2 = 0.055, t_initial_hip_swing2 = 0.15;
  const double t_initial2 = t_initial_front_swing2 + t_initial_front_hip_swing2 + t_initial_hip_swing2;

  contact_status_front_swing.setContactPoints(idocp::toPoints(contact_points));
  ocp_solver.pushBackContactStatus(contact_status_front_swing, t_start);
  contact_status_front_hip_swing.setContactPoints(idocp::toPoints(contact_points));
  ocp_solver.pushBackContactStatus(contact_status_front_hip_swing, t_start + t_initial_front_swing);

  contact_points[0].coeffRef(0) += 0.25 * stride;
  contact_points[1].coeffRef(0) += 0.25 * stride + 0.5 * additive_stride_hip;
  contact_points[2].coeffRef(0) += 0.25 * stride;
  contact_points[3].coeffRef(0) += 0.25 * stride + 0.5 * additive_stride_hip;
  contact_status_hip_swing.setContactPoints(idocp::toPoints(contact_points));
  ocp_solver.pushBackContactStatus(contact_status_hip_swing, t_start + t_initial_front_swing + t_initial_front_hip_swing);

  contact_status_front_swing.setContactPoints(idocp::toPoints(contact_points));
  ocp_solver.pushBackContactStatus(contact_status_front_swing, t_start + t_initial);
  contact_status_front_hip_swing.setContactPoints(idocp::toPoints(contact_points));
  ocp_solver.pushBackContactStatus(contact_status_front_hip_swing, t_start + t_initial + t_initial_front_swing2);

  contact_points[0].coeffRef(0) += 0.5 * stride;
  contact_points[1].coeffRef(0) += 0.5 * stride + 0.5 * additive_stride_hip;
  contact_points[2].coeffRef(0) += 0.5 * stride;
  contact_points[3].coeffRef(0) += 0.5 * stride + 0.5 * additive_stride_hip;
  contact_status_hip_swing.setContactPoints(idocp::toPoints(contact_points));
  ocp_solver.pushBackContactStatus(contact_status_hip_swing, t_start + t_initial + t_initial_front_swing2 + t_initial_front_hip_swing2);
  const double t_end_init = t_start + t_initial + t_initial2;

  for (int i = 0; i < steps; ++i) {
    contact_status_front_swing.setContactPoints(idocp::toPoints(contact_points));
    ocp_solver.pushBackContactStatus(contact_status_front_swing, t_end_init + i * t_period);
    ocp_solver.pushBackContactStatus(contact_status_front_hip_swing, t_end_init + i * t_period + t_front_swing);
    for (int k = 0; k < 4; ++k) contact_points[k].coeffRef(0) += stride;
    contact_status_hip_swing.setContactPoints(idocp::toPoints(contact_points));
    ocp_solver.pushBackContactStatus(contact_status_hip_swing, t_end_init + i * t_period + t_front_swing + t_front_hip_swing);
  }
  contact_status_front_swing.setContactPoints(idocp::toPoints(contact_points));
  ocp_solver.pushBackContactStatus(contact_status_front_swing, t_end_init + steps * t_period);

  const double t_end_front_swing = 0.15, t_end_front_hip_swing = 0.05, t_end_hip_swing = 0.15;
  const double t_end = t_end_front_swing + t_end_front_hip_swing + t_end_hip_swing;
  ocp_solver.pushBackContactStatus(contact_status_front_hip_swing, t_end_init + steps * t_period + t_end_front_swing);
  contact_points[0].coeffRef(0) += 0.5 * stride;
  contact_points[2].coeffRef(0) += 0.5 * stride;
  contact_points[1].coeffRef(0) += 0.5 * stride - additive_stride_hip;
  contact_points[3].coeffRef(0) += 0.5 * stride - additive_stride_hip;
  contact_status_hip_swing.setContactPoints(idocp::toPoints(contact_points));
  ocp_solver.pushBackContactStatus(contact_status_hip_swing, t_end_init + steps * t_period + t_end_front_swing + t_end_front_hip_swing);
  contact_status_initial.setContactPoints(idocp::toPoints(contact_points));
  ocp_solver.pushBackContactStatus(contact_status_initial, t_end_init + steps * t_period + t_end);

  idocp::VectorXd q = q_standing;
  idocp::VectorXd v = idocp::VectorXd::Zero(robot.dimv());
  ocp_solver.setSolution("q", q);
  ocp_solver.setSolution("v", v);
  idocp::Vector3d f_init(0, 0, 0.25 * robot.totalWeight());
  ocp_solver.setSolution("f", f_init);
  ocp_solver.initConstraints(t);

  std::cout << std::setprecision(17);
  idocp::ocpbenchmarker::Convergence(ocp_solver, t, q, v, iterations, line_search);
  idocp::ocpbenchmarker::CPUTime(ocp_solver, t, q, v, 5, line_search);
  return 0;
}

// pin_dump.cpp -- dumps known-answer vectors from REAL upstream idocp (mayataka/idocp linked against pinocchio / Eigen) in
// the JSON schema of tests/golden/, so that the CPU oracle of this repository (oracle/*.c, a restatement) can be pinned to
// the upstream arithmetic instead of only to itself.  tests/test_upstream_pins.py consumes the files when they exist.
//
// This program is written against the PUBLIC API of idocp as its own examples use it (examples/iiwa14/*.cpp,
// examples/anymal/anymal_trotting.cpp); it is part of this repository's test tooling, not a copy of reference code.
//
//   pin_dump <path to idocp/examples> <output directory>
//
// Files:
//   upstream_robot.json    Robot::RNEA / RNEADerivatives of iiwa14 and ANYmal (with contact forces) and
//                          Robot::computeMJtJinv at seeded pseudo-random (q, v, a[, f]); foot-frame positions of ANYmal;
//                          the end-effector placement and frame Jacobian of iiwa14 (frame 22)
//   upstream_solvers.json  per problem: q0, v0, the KKT error before and after every iteration and the iterate (q, v, a, u,
//                          lmd, gmm, beta of every stage) after iterations 1, 2 and the last one, for
//                            unocp_benchmark (UnOCPSolver, 50 it.), config_space_ocp (UnOCPSolver, 30 it.),
//                            task_space_ocp (UnOCPSolver and UnParNMPCSolver, 30 it.), unparnmpc_benchmark (20 it.),
//                            anymal_trotting (OCPSolver, 25 it.; iterate at stages 0, 11, 20, N)
// The inputs are part of the dump, so the consumer needs no random-number agreement with this program.
#include <cmath>
#include <cstdint>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "Eigen/Core"
#include "idocp/constraints/constraints.hpp"
#include "idocp/constraints/joint_position_lower_limit.hpp"
#include "idocp/constraints/joint_position_upper_limit.hpp"
#include "idocp/constraints/joint_torques_lower_limit.hpp"
#include "idocp/constraints/joint_torques_upper_limit.hpp"
#include "idocp/constraints/joint_velocity_lower_limit.hpp"
#include "idocp/constraints/joint_velocity_upper_limit.hpp"
#include "idocp/constraints/linearized_friction_cone.hpp"
#include "idocp/constraints/linearized_impulse_friction_cone.hpp"
#include "idocp/cost/configuration_space_cost.hpp"
#include "idocp/cost/contact_force_cost.hpp"
#include "idocp/cost/cost_function.hpp"
#include "idocp/cost/time_varying_task_space_6d_cost.hpp"
#include "idocp/cost/trotting_configuration_space_cost.hpp"
#include "idocp/ocp/ocp_solver.hpp"
#include "idocp/robot/robot.hpp"
#include "idocp/unocp/unocp_solver.hpp"
#include "idocp/unocp/unparnmpc_solver.hpp"
#include "idocp/utils/joint_constraints_factory.hpp"

namespace {

// counter-based splitmix64 -> [0, 1): the generator of SURVEY.md section 8d (any generator would do: inputs are dumped)
double uniform(uint64_t seed, uint64_t index) {
  uint64_t z = seed + (index + 1) * 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  z = z ^ (z >> 31);
  return static_cast<double>(z >> 11) * (1.0 / 9007199254740992.0);
}

struct Json {
  std::ostringstream o;
  Json() { o << std::setprecision(17); }
  static std::string vec(const Eigen::VectorXd& v) {
    std::ostringstream s;
    s << std::setprecision(17) << "[";
    for (int i = 0; i < v.size(); ++i) s << (i ? ", " : "") << v[i];
    s << "]";
    return s.str();
  }
  // row-major nested list
  static std::string mat(const Eigen::MatrixXd& m) {
    std::ostringstream s;
    s << "[";
    for (int r = 0; r < m.rows(); ++r) s << (r ? ", " : "") << vec(m.row(r).transpose());
    s << "]";
    return s.str();
  }
  static std::string list(const std::vector<double>& v) {
    std::ostringstream s;
    s << std::setprecision(17) << "[";
    for (size_t i = 0; i < v.size(); ++i) s << (i ? ", " : "") << v[i];
    s << "]";
    return s.str();
  }
};

std::string split_solution(const idocp::SplitSolution& s, bool with_u) {
  std::ostringstream o;
  o << "{\"q\": " << Json::vec(s.q) << ", \"v\": " << Json::vec(s.v) << ", \"lmd\": " << Json::vec(s.lmd) << ", \"gmm\": "
    << Json::vec(s.gmm);
  if (with_u) o << ", \"a\": " << Json::vec(s.a) << ", \"u\": " << Json::vec(s.u) << ", \"beta\": " << Json::vec(s.beta);
  o << "}";
  return o.str();
}

// KKT history + iterates of one fixed-base solver run (ocpbenchmarker::Convergence pattern, utils/ocp_benchmarker.hxx:37-51)
template <typename Solver>
std::string run_fixed_base(Solver& solver, const Eigen::VectorXd& q0, const Eigen::VectorXd& v0, int iterations, int first_stage,
                           int last_stage, int last_stage_with_u) {
  const double t = 0;
  std::vector<double> kkt;
  solver.computeKKTResidual(t, q0, v0);
  kkt.push_back(solver.KKTError());
  std::ostringstream iterates;
  iterates << "{";
  bool first = true;
  for (int it = 1; it <= iterations; ++it) {
    solver.updateSolution(t, q0, v0, false);
    solver.computeKKTResidual(t, q0, v0);
    kkt.push_back(solver.KKTError());
    if (it == 1 || it == 2 || it == iterations) {
      iterates << (first ? "" : ", ") << "\"" << it << "\": [";
      for (int i = first_stage; i <= last_stage; ++i)
        iterates << (i > first_stage ? ", " : "") << split_solution(solver.getSolution(i), i <= last_stage_with_u);
      iterates << "]";
      first = false;
    }
  }
  iterates << "}";
  std::ostringstream o;
  o << "{\"q0\": " << Json::vec(q0) << ", \"v0\": " << Json::vec(v0) << ", \"kkt\": " << Json::list(kkt) << ", \"iterates\": "
    << iterates.str() << "}";
  return o.str();
}

class CircleRef final : public idocp::TimeVaryingTaskSpace6DRefBase {
 public:
  CircleRef() {
    rot_ << 0, 0, 1, 0, 1, 0, -1, 0, 0;
    center_ << 0.546, 0, 0.76;
  }
  void compute_q_6d_ref(const double t, pinocchio::SE3& ref) const override {
    Eigen::Vector3d pos(center_);
    pos.coeffRef(1) += 0.1 * std::sin(M_PI * t);
    pos.coeffRef(2) += 0.1 * std::cos(M_PI * t);
    ref = pinocchio::SE3(rot_, pos);
  }
  bool isActive(const double) const override { return true; }
 private:
  Eigen::Matrix3d rot_;
  Eigen::Vector3d center_;
};

std::shared_ptr<idocp::CostFunction> config_cost(idocp::Robot& robot, const Eigen::VectorXd& q_ref, const Eigen::VectorXd& v_ref,
                                                 double qw, double vw, double aw) {
  auto cost = std::make_shared<idocp::CostFunction>();
  auto c = std::make_shared<idocp::ConfigurationSpaceCost>(robot);
  const int n = robot.dimv();
  c->set_q_ref(q_ref);
  c->set_v_ref(v_ref);
  c->set_q_weight(Eigen::VectorXd::Constant(n, qw));
  c->set_qf_weight(Eigen::VectorXd::Constant(n, qw));
  c->set_v_weight(Eigen::VectorXd::Constant(n, vw));
  c->set_vf_weight(Eigen::VectorXd::Constant(n, vw));
  c->set_a_weight(Eigen::VectorXd::Constant(n, aw));
  cost->push_back(c);
  return cost;
}

}  // namespace

int main(int argc, char* argv[]) {
  if (argc < 3) {
    std::cerr << "usage: pin_dump <path to idocp/examples> <output directory>\n";
    return 1;
  }
  const std::string examples = argv[1], out_dir = argv[2];
  const std::string iiwa_urdf = examples + "/iiwa14/iiwa_description/urdf/iiwa14.urdf";
  const std::string anymal_urdf = examples + "/anymal/anymal_b_simple_description/urdf/anymal.urdf";
  const std::vector<int> feet = {14, 24, 34, 44};

  // ------------------------------------------------------------------------------------------------ robot level
  {
    std::ofstream f(out_dir + "/upstream_robot.json");
    f << "{\n\"source\": \"mayataka/idocp + pinocchio (tools/pin_against_idocp/pin_dump.cpp)\",\n";
    idocp::Robot iiwa(iiwa_urdf);
    f << "\"iiwa14\": [";
    for (int s = 0; s < 8; ++s) {
      Eigen::VectorXd q(7), v(7), a(7), tau(7);
      for (int j = 0; j < 7; ++j) {
        q[j] = 2.0 * (2 * uniform(11, 21 * s + j) - 1);
        v[j] = 1.5 * (2 * uniform(11, 21 * s + 7 + j) - 1);
        a[j] = 3.0 * (2 * uniform(11, 21 * s + 14 + j) - 1);
      }
      Eigen::MatrixXd dq = Eigen::MatrixXd::Zero(7, 7), dv = dq, da = dq, J = Eigen::MatrixXd::Zero(6, 7);
      iiwa.RNEA(q, v, a, tau);
      iiwa.RNEADerivatives(q, v, a, dq, dv, da);
      iiwa.updateKinematics(q, v, a);
      iiwa.getFrameJacobian(22, J);
      f << (s ? ",\n" : "\n") << "{\"q\": " << Json::vec(q) << ", \"v\": " << Json::vec(v) << ", \"a\": " << Json::vec(a)
        << ", \"tau\": " << Json::vec(tau) << ", \"dtau_dq\": " << Json::mat(dq) << ", \"dtau_dv\": " << Json::mat(dv)
        << ", \"dtau_da\": " << Json::mat(da) << ", \"frame22_position\": " << Json::vec(iiwa.framePosition(22))
        << ", \"frame22_rotation\": " << Json::mat(iiwa.frameRotation(22)) << ", \"frame22_jacobian_local\": " << Json::mat(J) << "}";
    }
    f << "],\n";
    idocp::Robot anymal(anymal_urdf, feet);
    f << "\"anymal\": [";
    for (int s = 0; s < 8; ++s) {
      Eigen::VectorXd q(19), v(18), a(18), tau(18);
      Eigen::Vector4d quat;
      for (int j = 0; j < 4; ++j) quat[j] = 2 * uniform(12, 100 * s + j) - 1;
      quat.normalize();
      q << 0.3 * (2 * uniform(12, 100 * s + 4) - 1), 0.3 * (2 * uniform(12, 100 * s + 5) - 1), 0.48 + 0.1 * uniform(12, 100 * s + 6),
          quat, Eigen::VectorXd::Zero(12);
      for (int j = 0; j < 12; ++j) q[7 + j] = 1.0 * (2 * uniform(12, 100 * s + 10 + j) - 1);
      for (int j = 0; j < 18; ++j) {
        v[j] = 1.0 * (2 * uniform(12, 100 * s + 30 + j) - 1);
        a[j] = 2.0 * (2 * uniform(12, 100 * s + 50 + j) - 1);
      }
      auto status = anymal.createContactStatus();
      std::vector<Eigen::Vector3d> forces;
      for (int c = 0; c < 4; ++c) {
        if ((s >> c) & 1) status.activateContact(c);
        forces.push_back(Eigen::Vector3d(20 * (2 * uniform(12, 100 * s + 70 + 3 * c) - 1), 20 * (2 * uniform(12, 100 * s + 71 + 3 * c) - 1),
                                         60 * uniform(12, 100 * s + 72 + 3 * c)));
      }
      if (s == 7) for (int c = 0; c < 4; ++c) status.activateContact(c);
      anymal.updateKinematics(q, v, a);
      anymal.setContactForces(status, forces);
      Eigen::MatrixXd dq = Eigen::MatrixXd::Zero(18, 18), dv = dq, da = dq;
      anymal.RNEA(q, v, a, tau);
      anymal.RNEADerivatives(q, v, a, dq, dv, da);
      const int dimf = status.dimf();
      Eigen::MatrixXd Jc = Eigen::MatrixXd::Zero(dimf, 18), MJtJinv = Eigen::MatrixXd::Zero(18 + dimf, 18 + dimf);
      std::vector<double> active;
      Eigen::VectorXd fstack(12);
      for (int c = 0; c < 4; ++c) {
        active.push_back(status.isContactActive(c) ? 1 : 0);
        fstack.segment<3>(3 * c) = forces[c];
      }
      f << (s ? ",\n" : "\n") << "{\"q\": " << Json::vec(q) << ", \"v\": " << Json::vec(v) << ", \"a\": " << Json::vec(a)
        << ", \"active\": " << Json::list(active) << ", \"f\": " << Json::vec(fstack) << ", \"tau\": " << Json::vec(tau)
        << ", \"dtau_dq\": " << Json::mat(dq) << ", \"dtau_dv\": " << Json::mat(dv) << ", \"dtau_da\": " << Json::mat(da);
      if (dimf > 0) {
        // contact Jacobian rows through the Baumgarte derivative with respect to a (point_contact.hxx:100-144): dC/da = J
        Eigen::MatrixXd dCdq = Jc, dCdv = Jc;
        anymal.computeBaumgarteDerivatives(status, 0.05, dCdq, dCdv, Jc);
        anymal.computeMJtJinv(da, Jc, MJtJinv);
        f << ", \"contact_jacobian\": " << Json::mat(Jc) << ", \"MJtJinv\": " << Json::mat(MJtJinv);
      }
      f << ", \"foot_positions\": [";
      for (int c = 0; c < 4; ++c) f << (c ? ", " : "") << Json::vec(anymal.framePosition(feet[c]));
      f << "]}";
    }
    f << "]\n}\n";
  }

  // ------------------------------------------------------------------------------------------------ solver level
  {
    std::ofstream f(out_dir + "/upstream_solvers.json");
    f << "{\n\"source\": \"mayataka/idocp + pinocchio (tools/pin_against_idocp/pin_dump.cpp)\"";
    const int nthreads = 1;
    {  // examples/iiwa14/unocp_benchmark.cpp
      idocp::Robot robot(iiwa_urdf);
      robot.setJointEffortLimit(Eigen::VectorXd::Constant(7, 200));
      auto cost = config_cost(robot, Eigen::VectorXd::Constant(7, -5), Eigen::VectorXd::Constant(7, -9), 10, 0.1, 0.01);
      auto constraints = idocp::JointConstraintsFactory(robot).create();
      const Eigen::VectorXd q0 = Eigen::VectorXd::Constant(7, 2), v0 = Eigen::VectorXd::Zero(7);
      idocp::UnOCPSolver solver(robot, cost, constraints, 1.0, 20, nthreads);
      solver.setSolution("q", q0);
      solver.setSolution("v", v0);
      f << ",\n\"unocp_benchmark_reference_instance\": " << run_fixed_base(solver, q0, v0, 50, 0, 20, 19);
      idocp::UnParNMPCSolver par(robot, cost, constraints, 1.0, 20, nthreads);
      par.setSolution("q", q0);
      par.setSolution("v", v0);
      par.initBackwardCorrection(0.0);
      f << ",\n\"unparnmpc_benchmark_reference_instance\": " << run_fixed_base(par, q0, v0, 20, 0, 19, 19);
    }
    {  // examples/iiwa14/config_space_ocp.cpp
      idocp::Robot robot(iiwa_urdf);
      robot.setJointEffortLimit(Eigen::VectorXd::Constant(7, 50));
      robot.setJointVelocityLimit(Eigen::VectorXd::Constant(7, M_PI_2));
      Eigen::VectorXd q_ref(7), q0(7);
      q_ref << 0, M_PI_2, 0, M_PI_2, 0, M_PI_2, 0;
      q0 << M_PI_2, 0, M_PI_2, 0, M_PI_2, 0, M_PI_2;
      auto cost = config_cost(robot, q_ref, Eigen::VectorXd::Zero(7), 10, 0.01, 0.01);
      auto constraints = idocp::JointConstraintsFactory(robot).create();
      const Eigen::VectorXd v0 = Eigen::VectorXd::Zero(7);
      idocp::UnOCPSolver solver(robot, cost, constraints, 3.0, 60, nthreads);
      solver.setSolution("q", q0);
      solver.setSolution("v", v0);
      f << ",\n\"config_space_ocp\": " << run_fixed_base(solver, q0, v0, 30, 0, 60, 59);
    }
    {  // examples/iiwa14/task_space_ocp.cpp through both solvers
      idocp::Robot robot(iiwa_urdf);
      robot.setJointEffortLimit(Eigen::VectorXd::Constant(7, 50));
      robot.setJointVelocityLimit(Eigen::VectorXd::Constant(7, M_PI_2));
      auto cost = std::make_shared<idocp::CostFunction>();
      auto c = std::make_shared<idocp::ConfigurationSpaceCost>(robot);
      c->set_v_weight(Eigen::VectorXd::Constant(7, 0.01));
      c->set_vf_weight(Eigen::VectorXd::Constant(7, 0.01));
      c->set_a_weight(Eigen::VectorXd::Constant(7, 0.01));
      cost->push_back(c);
      auto task = std::make_shared<idocp::TimeVaryingTaskSpace6DCost>(robot, 22, std::make_shared<CircleRef>());
      task->set_q_6d_weight(Eigen::Vector3d::Constant(1000), Eigen::Vector3d::Constant(1000));
      task->set_qf_6d_weight(Eigen::Vector3d::Constant(1000), Eigen::Vector3d::Constant(1000));
      cost->push_back(task);
      auto constraints = idocp::JointConstraintsFactory(robot).create();
      Eigen::VectorXd q0(7);
      q0 << 0, M_PI_2, 0, M_PI_2, 0, M_PI_2, 0;
      const Eigen::VectorXd v0 = Eigen::VectorXd::Zero(7);
      idocp::UnOCPSolver solver(robot, cost, constraints, 6.0, 120, nthreads);
      solver.setSolution("q", q0);
      solver.setSolution("v", v0);
      f << ",\n\"task_space_ocp_unocp\": " << run_fixed_base(solver, q0, v0, 30, 0, 120, 119);
      idocp::UnParNMPCSolver par(robot, cost, constraints, 6.0, 120, nthreads);
      par.setSolution("q", q0);
      par.setSolution("v", v0);
      par.initBackwardCorrection(0.0);
      f << ",\n\"task_space_ocp_unparnmpc\": " << run_fixed_base(par, q0, v0, 30, 0, 119, 119);
    }
    {  // examples/anymal/anymal_trotting.cpp: same problem, same schedule, same calls
      idocp::Robot robot(anymal_urdf, feet);
      const double step_length = 0.15, t_start = 0.5, t_period = 0.5;
      auto cost = std::make_shared<idocp::CostFunction>();
      Eigen::VectorXd q_standing(19), q_weight(18), v_weight(18), a_weight(18);
      q_standing << 0, 0, 0.4792, 0, 0, 0, 1, -0.1, 0.7, -1.0, -0.1, -0.7, 1.0, 0.1, 0.7, -1.0, 0.1, -0.7, 1.0;
      q_weight.setConstant(10);
      v_weight << 1, 1, 1, 1, 1, 1, Eigen::VectorXd::Constant(12, 0.1);
      a_weight << 0.1, 0.1, 0.1, 0.1, 0.1, 0.1, Eigen::VectorXd::Constant(12, 0.01);
      idocp::TrottingSwingAngles angles;
      angles.front_swing_knee = 1.7;
      angles.hip_swing_knee = 1.7;
      auto config = std::make_shared<idocp::TrottingConfigurationSpaceCost>(robot);
      config->set_ref(t_start, t_period, q_standing, step_length, angles);
      config->set_q_weight(q_weight); config->set_qf_weight(q_weight); config->set_qi_weight(q_weight);
      config->set_v_weight(v_weight); config->set_vf_weight(v_weight); config->set_vi_weight(v_weight);
      config->set_a_weight(a_weight); config->set_dvi_weight(a_weight);
      cost->push_back(config);
      auto force = std::make_shared<idocp::ContactForceCost>(robot);
      const std::vector<Eigen::Vector3d> f_weight(4, Eigen::Vector3d::Constant(0.001));
      force->set_f_weight(f_weight);
      force->set_fi_weight(f_weight);
      force->set_f_ref(robot);
      cost->push_back(force);
      auto constraints = std::make_shared<idocp::Constraints>();
      constraints->push_back(std::make_shared<idocp::JointPositionLowerLimit>(robot));
      constraints->push_back(std::make_shared<idocp::JointPositionUpperLimit>(robot));
      constraints->push_back(std::make_shared<idocp::JointVelocityLowerLimit>(robot));
      constraints->push_back(std::make_shared<idocp::JointVelocityUpperLimit>(robot));
      constraints->push_back(std::make_shared<idocp::JointTorquesLowerLimit>(robot));
      constraints->push_back(std::make_shared<idocp::JointTorquesUpperLimit>(robot));
      const double mu = 0.7;
      constraints->push_back(std::make_shared<idocp::LinearizedFrictionCone>(robot, mu));
      constraints->push_back(std::make_shared<idocp::LinearizedImpulseFrictionCone>(robot, mu));
      const double T = 1.55;
      const int N = 30, max_num_impulse_phase = 2;
      idocp::OCPSolver solver(robot, cost, constraints, T, N, max_num_impulse_phase + 1, nthreads);
      robot.updateFrameKinematics(q_standing);
      std::vector<Eigen::Vector3d> points(robot.maxPointContacts(), Eigen::Vector3d::Zero());
      robot.getContactPoints(points);
      auto initial = robot.createContactStatus();
      initial.activateContacts({0, 1, 2, 3});
      initial.setContactPoints(points);
      solver.setContactStatusUniformly(initial);
      auto even = robot.createContactStatus();
      even.activateContacts({1, 2});
      even.setContactPoints(points);
      solver.pushBackContactStatus(even, t_start);
      auto odd = robot.createContactStatus();
      points[0].coeffRef(0) += 0.5 * step_length;
      points[3].coeffRef(0) += 0.5 * step_length;
      odd.activateContacts({0, 3});
      odd.setContactPoints(points);
      solver.pushBackContactStatus(odd, t_start + t_period);
      for (int k = 2; k <= max_num_impulse_phase; ++k) {
        if (k % 2 == 0) {
          points[1].coeffRef(0) += step_length;
          points[2].coeffRef(0) += step_length;
          even.setContactPoints(points);
          solver.pushBackContactStatus(even, t_start + k * t_period);
        } else {
          points[0].coeffRef(0) += step_length;
          points[3].coeffRef(0) += step_length;
          odd.setContactPoints(points);
          solver.pushBackContactStatus(odd, t_start + k * t_period);
        }
      }
      const Eigen::VectorXd q0 = q_standing, v0 = Eigen::VectorXd::Zero(18);
      solver.setSolution("q", q0);
      solver.setSolution("v", v0);
      Eigen::Vector3d f_init;
      f_init << 0, 0, 0.25 * robot.totalWeight();
      solver.setSolution("f", f_init);
      solver.initConstraints(0.0);
      std::vector<double> kkt;
      solver.computeKKTResidual(0.0, q0, v0);
      kkt.push_back(solver.KKTError());
      for (int it = 0; it < 25; ++it) {
        solver.updateSolution(0.0, q0, v0, false);
        solver.computeKKTResidual(0.0, q0, v0);
        kkt.push_back(solver.KKTError());
      }
      f << ",\n\"anymal_trotting\": {\"kkt\": " << Json::list(kkt) << ", \"final\": {";
      const int stages[] = {0, 11, 20, N};
      for (int k = 0; k < 4; ++k) {
        const idocp::SplitSolution& s = solver.getSolution(stages[k]);
        f << (k ? ", " : "") << "\"" << stages[k] << "\": {\"q\": " << Json::vec(s.q) << ", \"v\": " << Json::vec(s.v);
        if (stages[k] < N) f << ", \"u\": " << Json::vec(s.u) << ", \"f_stack\": " << Json::vec(s.f_stack());
        f << "}";
      }
      f << "}}";
    }
    f << "\n}\n";
  }
  std::cout << "wrote " << out_dir << "/upstream_robot.json and upstream_solvers.json\n";
  return 0;
}

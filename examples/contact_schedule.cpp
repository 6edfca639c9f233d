// contact_schedule.cpp -- builds a quadruped trotting contact schedule with the host-side hybrid classes
// (ContactStatus, ContactSequence, OCPDiscretizer; include/idocp_b200/hybrid.hpp) and prints the flattened stage
// table a device-side OCPSolver consumes.  Pure host code:
//   g++ -std=c++17 -Iinclude examples/contact_schedule.cpp -o build/contact_schedule
#include <cstdio>
#include <vector>

#include "idocp_b200/hybrid.hpp"

namespace ob = idocp_b200;

int main() {
  const int feet = 4;                    // LF, LH, RF, RH
  const double T = 1.55, t0 = 0.0;
  const int N = 30, max_events = 3;
  ob::ContactStatus standing(feet), lf_rh_swing(feet), rf_lh_swing(feet);
  standing.activateContacts();
  lf_rh_swing.activateContacts({1, 2});  // LH, RF on the ground
  rf_lh_swing.activateContacts({0, 3});  // LF, RH on the ground
  ob::ContactSequence sequence(feet, max_events);
  sequence.setContactStatusUniformly(standing);
  sequence.push_back(lf_rh_swing, 0.5);  // two feet leave the ground: a lift event
  sequence.push_back(rf_lh_swing, 1.0);  // two feet land (the others leave): an impulse event
  sequence.push_back(lf_rh_swing, 1.5);
  ob::OCPDiscretizer grid(T, N, max_events);
  const bool ok = grid.discretizeOCP(sequence, t0);
  std::printf("well defined: %d  N = %d  impulses = %d  lifts = %d  stages = %d\n", ok ? 1 : 0, grid.N(),
              grid.N_impulse(), grid.N_lift(), grid.N_all());
  static const char* names[] = {"grid", "impulse", "aux", "lift", "terminal"};
  for (const ob::ScheduledStage& s : ob::flattenSchedule(grid)) {
    const ob::ContactStatus& cs = sequence.contactStatus(s.contact_phase);
    std::printf("%-8s %3d  t = %.4f  dt = %.4f  feet = %d%d%d%d  constraints(%d)%s\n", names[s.kind], s.index, s.t, s.dt,
                cs.isContactActive(0), cs.isContactActive(1), cs.isContactActive(2), cs.isContactActive(3),
                s.constraint_stage, s.before_impulse ? "  + switching constraint" : "");
  }
  return ok ? 0 : 1;
}

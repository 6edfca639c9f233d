// hybrid.hpp -- host-side contact schedule of the hybrid (contact / impulse / lift) optimal control problem.
//
// Row a13 of SURVEY.md section 8: the schedule types consumed by idocp's OCPSolver stay on the host (integer /
// time bookkeeping, negligible cost) and are flattened into one stage table per updateSolution(t) call that a
// device-side OCPSolver uploads when t or the sequence changes.  Same class names, methods, argument meaning and
// error behaviour as the reference:
//
//   ContactStatus     include/idocp/robot/contact_status.hpp / .hxx
//   ImpulseStatus     include/idocp/robot/impulse_status.hpp / .hxx
//   DiscreteEvent     include/idocp/hybrid/discrete_event.hpp / .hxx        (impulse vs lift classification :78-104)
//   ContactSequence   include/idocp/hybrid/contact_sequence.hpp / .hxx      (push_back :56-103, pop :112-154, ...)
//   OCPDiscretizer    include/idocp/hybrid/ocp_discretizer.hpp / .hxx       (event -> stage mapping :246-377)
//   StageSchedule     the flattened per-stage table = what hybrid_container.hpp's OCP container indexes
//                     (N + 1 grid stages, N_impulse impulse + aux stages, N_lift lift stages; ocp_linearizer.hxx:113-228)
//
// Internals are organised differently from the reference (one phase list instead of eight parallel deques; the
// discretiser walks one merged, time-ordered event list), and three places where the reference reads stale or
// mis-indexed state are defined instead (documented at the spot).  Header-only, no CUDA, no Eigen.
#ifndef IDOCP_B200_HYBRID_HPP_
#define IDOCP_B200_HYBRID_HPP_

#include <array>
#include <cmath>
#include <cstdlib>
#include <deque>
#include <iostream>
#include <limits>
#include <stdexcept>
#include <string>
#include <vector>

namespace idocp_b200 {

using Point3 = std::array<double, 3>;

class ContactStatus {
 public:
  ContactStatus() {}
  explicit ContactStatus(const int max_point_contacts)
      : on_(max_point_contacts, 0), points_(max_point_contacts, Point3{0.0, 0.0, 0.0}) {}

  int maxPointContacts() const { return static_cast<int>(on_.size()); }
  bool isContactActive(const int contact_index) const { return on_.at(contact_index) != 0; }
  std::vector<bool> isContactActive() const { return std::vector<bool>(on_.begin(), on_.end()); }
  int dimf() const {
    int n = 0;
    for (const char e : on_) n += e ? 3 : 0;
    return n;
  }
  bool hasActiveContacts() const { return dimf() > 0; }

  void setActivity(const std::vector<bool>& is_contact_active) {
    require(static_cast<int>(is_contact_active.size()) == maxPointContacts(), "is_contact_active.size()");
    for (int i = 0; i < maxPointContacts(); ++i) on_[i] = is_contact_active[i] ? 1 : 0;
  }
  void activateContact(const int contact_index) { on_.at(contact_index) = 1; }
  void deactivateContact(const int contact_index) { on_.at(contact_index) = 0; }
  void activateContacts(const std::vector<int>& contact_indices) { for (const int i : contact_indices) activateContact(i); }
  void deactivateContacts(const std::vector<int>& contact_indices) { for (const int i : contact_indices) deactivateContact(i); }
  void activateContacts() { on_.assign(on_.size(), 1); }
  void deactivateContacts() { on_.assign(on_.size(), 0); }

  void setContactPoint(const int contact_index, const Point3& contact_point) { points_.at(contact_index) = contact_point; }
  void setContactPoints(const std::vector<Point3>& contact_points) {
    require(contact_points.size() == points_.size(), "contact_points.size()");
    points_ = contact_points;
  }
  // the reference's signature takes std::vector<Eigen::Vector3d>: any 3-vector type with operator[]
  template <typename Vector3>
  void setContactPoints(const std::vector<Vector3>& contact_points) {
    require(contact_points.size() == points_.size(), "contact_points.size()");
    for (size_t i = 0; i < points_.size(); ++i) points_[i] = Point3{contact_points[i][0], contact_points[i][1], contact_points[i][2]};
  }
  const Point3& contactPoint(const int contact_index) const { return points_.at(contact_index); }
  const std::vector<Point3>& contactPoints() const { return points_; }

  // activity and contact points (isApprox with Eigen's default precision 1e-12, contact_status.hxx:34-46)
  bool operator==(const ContactStatus& other) const {
    if (other.maxPointContacts() != maxPointContacts()) return false;
    for (int i = 0; i < maxPointContacts(); ++i) {
      if (other.on_[i] != on_[i]) return false;
      if (!approx(other.points_[i], points_[i])) return false;
    }
    return true;
  }
  bool operator!=(const ContactStatus& other) const { return !(*this == other); }

 private:
  static void require(bool ok, const char* what) {
    if (!ok) {
      std::cerr << "invalid argument: " << what << " must equal maxPointContacts()!" << '\n';
      std::exit(EXIT_FAILURE);
    }
  }
  // Eigen::DenseBase::isApprox: |a - b|^2 <= prec^2 min(|a|^2, |b|^2)
  static bool approx(const Point3& a, const Point3& b) {
    double d2 = 0, a2 = 0, b2 = 0;
    for (int k = 0; k < 3; ++k) { d2 += (a[k] - b[k]) * (a[k] - b[k]); a2 += a[k] * a[k]; b2 += b[k] * b[k]; }
    const double prec = 1e-12;
    return d2 <= prec * prec * std::min(a2, b2);
  }
  std::vector<char> on_;
  std::vector<Point3> points_;
};

// which contacts become active at an event (impulse_status.hxx:67-85)
class ImpulseStatus {
 public:
  ImpulseStatus() {}
  explicit ImpulseStatus(const int max_point_contacts) : s_(max_point_contacts) {}
  int maxPointContacts() const { return s_.maxPointContacts(); }
  bool isImpulseActive(const int contact_index) const { return s_.isContactActive(contact_index); }
  std::vector<bool> isImpulseActive() const { return s_.isContactActive(); }
  bool hasActiveImpulse() const { return s_.hasActiveContacts(); }
  int dimf() const { return s_.dimf(); }
  void setActivity(const ContactStatus& pre_contact_status, const ContactStatus& post_contact_status) {
    for (int i = 0; i < maxPointContacts(); ++i) {
      if (!pre_contact_status.isContactActive(i) && post_contact_status.isContactActive(i)) s_.activateContact(i);
      else s_.deactivateContact(i);
    }
  }
  void setActivity(const std::vector<bool>& is_impulse_active) { s_.setActivity(is_impulse_active); }
  void activateImpulse(const int contact_index) { s_.activateContact(contact_index); }
  void deactivateImpulse(const int contact_index) { s_.deactivateContact(contact_index); }
  void activateImpulse() { s_.activateContacts(); }
  void deactivateImpulse() { s_.deactivateContacts(); }
  void setContactPoint(const int contact_index, const Point3& p) { s_.setContactPoint(contact_index, p); }
  void setContactPoints(const std::vector<Point3>& p) { s_.setContactPoints(p); }
  const std::vector<Point3>& contactPoints() const { return s_.contactPoints(); }
  bool operator==(const ImpulseStatus& o) const { return s_ == o.s_; }
  bool operator!=(const ImpulseStatus& o) const { return !(*this == o); }
 private:
  ContactStatus s_;
};

// transition between two contact statuses: an IMPULSE if any contact becomes active (it may lift others at the
// same time), otherwise a LIFT if any contact becomes inactive (discrete_event.hxx:78-104)
class DiscreteEvent {
 public:
  DiscreteEvent() {}
  explicit DiscreteEvent(const int max_point_contacts)
      : pre_(max_point_contacts), post_(max_point_contacts), impulse_(max_point_contacts) {}
  DiscreteEvent(const ContactStatus& pre_contact_status, const ContactStatus& post_contact_status)
      : DiscreteEvent(pre_contact_status.maxPointContacts()) {
    setDiscreteEvent(pre_contact_status, post_contact_status);
  }
  void setDiscreteEvent(const ContactStatus& pre_contact_status, const ContactStatus& post_contact_status) {
    touch_down_ = lift_off_ = false;
    for (int i = 0; i < pre_contact_status.maxPointContacts(); ++i) {
      const bool before = pre_contact_status.isContactActive(i), after = post_contact_status.isContactActive(i);
      touch_down_ = touch_down_ || (!before && after);
      lift_off_ = lift_off_ || (before && !after);
    }
    impulse_.setActivity(pre_contact_status, post_contact_status);
    impulse_.setContactPoints(post_contact_status.contactPoints());
    pre_ = pre_contact_status;
    post_ = post_contact_status;
  }
  bool existDiscreteEvent() const { return touch_down_ || lift_off_; }
  bool existImpulse() const { return touch_down_; }
  bool existLift() const { return lift_off_; }   // note: true also for an impulse event that lifts other feet
  const ImpulseStatus& impulseStatus() const { return impulse_; }
  const ContactStatus& preContactStatus() const { return pre_; }
  const ContactStatus& postContactStatus() const { return post_; }
  void setContactPoint(const int contact_index, const Point3& p) { impulse_.setContactPoint(contact_index, p); }
  void setContactPoints(const std::vector<Point3>& p) { impulse_.setContactPoints(p); }
  int maxPointContacts() const { return pre_.maxPointContacts(); }
 private:
  ContactStatus pre_, post_;
  ImpulseStatus impulse_;
  bool touch_down_ = false, lift_off_ = false;
};

namespace detail {
[[noreturn]] inline void schedule_die(const std::string& what) {
  std::cerr << what << '\n';
  std::exit(EXIT_FAILURE);
}
}  // namespace detail

// Sequence of contact phases separated by discrete events.  Phase k > 0 starts with event k - 1.
class ContactSequence {
 public:
  ContactSequence() {}
  // the reference takes (const Robot&, max_num_events) and asks the robot for its number of point contacts
  ContactSequence(const int max_point_contacts, const int max_num_events)
      : max_events_(max_num_events), default_(max_point_contacts) {
    if (max_num_events <= 0) detail::schedule_die("invalid argument: max_num_events must be positive!");
    phases_.push_back(Phase{default_, 0.0, false, ImpulseStatus(max_point_contacts)});
  }
  template <typename RobotType>
  ContactSequence(const RobotType& robot, const int max_num_events)
      : ContactSequence(robot.maxPointContacts(), max_num_events) {}

  void setContactStatusUniformly(const ContactStatus& contact_status) {
    phases_.clear();
    phases_.push_back(Phase{contact_status, 0.0, false, ImpulseStatus(contact_status.maxPointContacts())});
  }
  // 0 on success, otherwise the reference's error text in `why` (the throwing overloads below print + exit like
  // the reference, contact_sequence.hxx:56-103)
  int try_push_back(const DiscreteEvent& discrete_event, const double event_time, std::string* why) {
    if (numContactPhases() == 0) return err(why, "Call setContactStatusUniformly() before calling push_back()!");
    if (!discrete_event.existDiscreteEvent()) return err(why, "discrete_event.existDiscreteEvent() must be true!");
    if (discrete_event.preContactStatus() != phases_.back().status)
      return err(why, "discrete_event.preContactStatus() is not consistent with the last contact status!");
    if (numDiscreteEvents() + 1 > max_events_)
      return err(why, "Number of discrete events=" + std::to_string(numDiscreteEvents() + 1) +
                          " exceeds predefined max_num_events=" + std::to_string(max_events_) + "!");
    if (numDiscreteEvents() > 0 && event_time <= phases_.back().start)
      return err(why, "event_time=" + std::to_string(event_time) + " must be larger than the last event time=" +
                          std::to_string(phases_.back().start) + "!");
    phases_.push_back(Phase{discrete_event.postContactStatus(), event_time, discrete_event.existImpulse(),
                            discrete_event.impulseStatus()});
    return 0;
  }
  void push_back(const DiscreteEvent& discrete_event, const double event_time) {
    std::string why;
    if (try_push_back(discrete_event, event_time, &why) != 0) detail::schedule_die(why);
  }
  void push_back(const ContactStatus& contact_status, const double event_time) {
    push_back(DiscreteEvent(phases_.back().status, contact_status), event_time);
  }
  void pop_back() {
    if (numDiscreteEvents() > 0) phases_.pop_back();
    else if (numContactPhases() > 0) phases_.back().status = default_;
  }
  void pop_front() {
    if (numDiscreteEvents() > 0) {
      phases_.pop_front();
      phases_.front().impulse_start = false;   // the first phase has no starting event any more
      phases_.front().start = 0.0;
    } else if (numContactPhases() > 0) {
      phases_.back().status = default_;
    }
  }
  // contact_sequence.hxx:157-197 / :200-240.  The admissible interval is checked as the reference does:
  // against the previous event when there is one, otherwise against the next one.  (The reference keeps the
  // event positions in deques that pop_front does not renumber; positions here are always current.)
  int try_update_event_time(const bool impulse, const int index, const double time, std::string* why) {
    const char* kind = impulse ? "impulse" : "lift";
    const int count = impulse ? numImpulseEvents() : numLiftEvents();
    if (count <= 0)
      return err(why, std::string(impulse ? "numImpulseEvents()" : "numLiftEvents()") + " must be positive when calling this method!");
    if (index < 0) return err(why, std::string(kind) + "_index must be non-negative!");
    if (index >= count)
      return err(why, std::string(kind) + "_index=" + std::to_string(index) + " must be less than " +
                          (impulse ? "numImpulseEvents()=" : "numLiftEvents()=") + std::to_string(count) + "!");
    const int e = eventPosition(impulse, index);   // event e starts phase e + 1
    if (e > 0) {
      if (phases_[e].start >= time)
        return err(why, std::string(kind) + "_time=" + std::to_string(time) + " must be larger than event_time_[event_index-1]=" +
                            std::to_string(phases_[e].start) + "!");
    } else if (e + 1 < numDiscreteEvents()) {
      if (phases_[e + 2].start <= time)
        return err(why, std::string(kind) + "_time=" + std::to_string(time) + " must be smaller than event_time_[event_index+1]=" +
                            std::to_string(phases_[e + 2].start) + "!");
    }
    phases_[e + 1].start = time;
    return 0;
  }
  void updateImpulseTime(const int impulse_index, const double impulse_time) {
    std::string why;
    if (try_update_event_time(true, impulse_index, impulse_time, &why) != 0) detail::schedule_die(why);
  }
  void updateLiftTime(const int lift_index, const double lift_time) {
    std::string why;
    if (try_update_event_time(false, lift_index, lift_time, &why) != 0) detail::schedule_die(why);
  }
  // contact points of a phase and of the impulse that starts it (contact_sequence.hxx:243-262; the reference
  // indexes its impulse list with contact_phase - 1, i.e. by EVENT position, which is only right while every
  // earlier event is an impulse; here the phase owns its impulse)
  void setContactPoints(const int contact_phase, const std::vector<Point3>& contact_points) {
    if (contact_phase >= numContactPhases())
      detail::schedule_die("contact_phase=" + std::to_string(contact_phase) + " must be smaller than numContactPhases()" +
                           std::to_string(numContactPhases()) + "!");
    phases_[contact_phase].status.setContactPoints(contact_points);
    if (contact_phase > 0 && phases_[contact_phase].impulse_start) phases_[contact_phase].impulse.setContactPoints(contact_points);
  }

  int numContactPhases() const { return static_cast<int>(phases_.size()); }
  int numDiscreteEvents() const { return numContactPhases() - 1; }
  int numImpulseEvents() const {
    int n = 0;
    for (int k = 1; k < numContactPhases(); ++k) n += phases_[k].impulse_start ? 1 : 0;
    return n;
  }
  int numLiftEvents() const { return numDiscreteEvents() - numImpulseEvents(); }
  const ContactStatus& contactStatus(const int contact_phase) const { return phases_.at(contact_phase).status; }
  const ImpulseStatus& impulseStatus(const int impulse_index) const { return phases_[eventPosition(true, impulse_index) + 1].impulse; }
  double impulseTime(const int impulse_index) const { return phases_[eventPosition(true, impulse_index) + 1].start; }
  double liftTime(const int lift_index) const { return phases_[eventPosition(false, lift_index) + 1].start; }
  // event e (0-based, time ordered): its time and kind
  double eventTime(const int e) const { return phases_.at(e + 1).start; }
  bool isImpulseEvent(const int e) const { return phases_.at(e + 1).impulse_start; }
  int maxNumEvents() const { return max_events_; }

 private:
  struct Phase {
    ContactStatus status;
    double start;          // time of the event that starts the phase (unused for the first phase)
    bool impulse_start;    // that event is an impulse (else a lift)
    ImpulseStatus impulse;
  };
  static int err(std::string* why, const std::string& text) {
    if (why) *why = text;
    return -1;
  }
  // position in the event list of the index-th impulse (or lift) event
  int eventPosition(const bool impulse, const int index) const {
    int seen = 0;
    for (int k = 1; k < numContactPhases(); ++k)
      if (phases_[k].impulse_start == impulse && seen++ == index) return k - 1;
    throw std::out_of_range("event index");
  }
  int max_events_ = 0;
  ContactStatus default_;
  std::deque<Phase> phases_;
};

// Maps the events of a contact sequence onto the time grid of the horizon (ocp_discretizer.hxx:246-377).
//   * an event at time te inside grid interval i = floor((te - t) / dt_ideal) shortens stage i to dt(i) = te - t_i and
//     is followed by an impulse + aux stage (dt_aux = dt_ideal - dt(i)) or a lift stage (dt_lift likewise);
//   * an event within min_dt = sqrt(eps) AFTER a grid point merges that grid stage away (N shrinks by one, the event
//     follows the previous stage, its aux / lift stage gets the full dt_ideal);
//   * an event within min_dt BEFORE the next grid point is moved onto that grid point (same merge, one stage later).
class OCPDiscretizer {
 public:
  static constexpr double kMinDt = 1.4901161193847656e-08;   // sqrt(DBL_EPSILON), ocp_discretizer.hpp:108-109

  OCPDiscretizer() {}
  OCPDiscretizer(const double T, const int N, const int max_events)
      : T_(T), dt_ideal_(T / N), N_ideal_(N), N_(N), max_events_(max_events) { reset(); }

  // returns false when the schedule is not well defined (an event before t, two events in one grid interval, or
  // impulses after two consecutive stages): the reference only asserts this in Debug builds
  // (ocp_discretizer.hxx:62-72,205-217)
  bool discretizeOCP(const ContactSequence& contact_sequence, const double t) {
    reset();
    const int ne = contact_sequence.numDiscreteEvents();
    std::vector<int> cell(ne);       // grid interval of every event, later the stage it follows
    for (int e = 0; e < ne; ++e) cell[e] = static_cast<int>(std::floor((contact_sequence.eventTime(e) - t) / dt_ideal_));
    int next = 0, merged = 0;
    bool ok = true;
    for (int i = 0; i < N_ideal_; ++i) {
      const int stage = i - merged;
      if (next < ne && cell[next] == i) {
        const double te = contact_sequence.eventTime(next);
        const double d = te - i * dt_ideal_ - t;
        dt_[stage] = d;
        if (d <= kMinDt) {                       // on the grid point: the grid stage disappears
          t_[stage] = t + (i - 1) * dt_ideal_;
          record(contact_sequence.isImpulseEvent(next), te, stage - 1, dt_ideal_);
          ++merged;
          ++next;
        } else if (d >= dt_ideal_ - kMinDt) {    // (numerically) on the next grid point: handled there
          t_[stage] = t + i * dt_ideal_;
          cell[next] = i + 1;
        } else {
          t_[stage] = t + i * dt_ideal_;
          record(contact_sequence.isImpulseEvent(next), te, stage, dt_ideal_ - d);
          ++next;
        }
        if (next < ne && cell[next] == i) ok = false;   // a second event in the same interval is never visited
      } else {
        dt_[stage] = dt_ideal_;
        t_[stage] = t + i * dt_ideal_;
      }
    }
    N_ = N_ideal_ - merged;
    t_[N_] = t + T_;
    // stage flags and contact phases
    int phase = 0;
    for (int i = 0; i <= N_; ++i) {
      phase_[i] = phase;
      if (i == N_) break;
      for (size_t k = 0; k < stage_before_impulse_.size(); ++k)
        if (stage_before_impulse_[k] == i) { impulse_after_[i] = static_cast<int>(k); ++phase; }
      for (size_t k = 0; k < stage_before_lift_.size(); ++k)
        if (stage_before_lift_[k] == i) { lift_after_[i] = static_cast<int>(k); ++phase; }
    }
    // events that found no stage: fine when they lie beyond the horizon (cell >= N_ideal; the reference would
    // count them in N_impulse() / N_lift() with undefined stage data), an error when they are in the past or
    // share a grid interval with another event
    for (int e = next; e < ne; ++e)
      if (cell[e] < N_ideal_) ok = false;
    return ok && isWellDefined();
  }

  int N() const { return N_; }
  int N_impulse() const { return static_cast<int>(stage_before_impulse_.size()); }
  int N_lift() const { return static_cast<int>(stage_before_lift_.size()); }
  int N_all() const { return N() + 1 + 2 * N_impulse() + N_lift(); }
  int N_ideal() const { return N_ideal_; }
  int contactPhase(const int time_stage) const { return phase_.at(time_stage); }
  int contactPhaseAfterImpulse(const int impulse_index) const { return contactPhase(timeStageAfterImpulse(impulse_index)); }
  int contactPhaseAfterLift(const int lift_index) const { return contactPhase(timeStageAfterLift(lift_index)); }
  int impulseIndexAfterTimeStage(const int time_stage) const { return impulse_after_.at(time_stage); }
  int liftIndexAfterTimeStage(const int time_stage) const { return lift_after_.at(time_stage); }
  int timeStageBeforeImpulse(const int impulse_index) const { return stage_before_impulse_.at(impulse_index); }
  int timeStageAfterImpulse(const int impulse_index) const { return timeStageBeforeImpulse(impulse_index) + 1; }
  int timeStageBeforeLift(const int lift_index) const { return stage_before_lift_.at(lift_index); }
  int timeStageAfterLift(const int lift_index) const { return timeStageBeforeLift(lift_index) + 1; }
  bool isTimeStageBeforeImpulse(const int time_stage) const { return time_stage < N_ && impulse_after_.at(time_stage) >= 0; }
  bool isTimeStageAfterImpulse(const int time_stage) const { return isTimeStageBeforeImpulse(time_stage - 1); }
  bool isTimeStageBeforeLift(const int time_stage) const { return time_stage < N_ && lift_after_.at(time_stage) >= 0; }
  bool isTimeStageAfterLift(const int time_stage) const { return isTimeStageBeforeLift(time_stage - 1); }
  double t(const int time_stage) const { return t_.at(time_stage); }
  double t_impulse(const int impulse_index) const { return t_impulse_.at(impulse_index); }
  double t_lift(const int lift_index) const { return t_lift_.at(lift_index); }
  double dt(const int time_stage) const { return dt_.at(time_stage); }
  double dt_aux(const int impulse_index) const { return dt_aux_.at(impulse_index); }
  double dt_lift(const int lift_index) const { return dt_lift_.at(lift_index); }
  bool isWellDefined() const {
    for (int i = 0; i < N_; ++i)
      if (isTimeStageBeforeImpulse(i) && isTimeStageBeforeLift(i)) return false;
    for (int i = 0; i + 1 < N_; ++i)
      if (isTimeStageBeforeImpulse(i) && isTimeStageBeforeImpulse(i + 1)) return false;
    return true;
  }

 private:
  void reset() {
    N_ = N_ideal_;
    t_.assign(N_ideal_ + 1, 0.0);
    dt_.assign(N_ideal_ + 1, dt_ideal_);
    phase_.assign(N_ideal_ + 1, 0);
    impulse_after_.assign(N_ideal_ + 1, -1);
    lift_after_.assign(N_ideal_ + 1, -1);
    stage_before_impulse_.clear(); stage_before_lift_.clear();
    t_impulse_.clear(); t_lift_.clear(); dt_aux_.clear(); dt_lift_.clear();
  }
  void record(const bool impulse, const double time, const int stage_before, const double dt_after) {
    if (impulse) { stage_before_impulse_.push_back(stage_before); t_impulse_.push_back(time); dt_aux_.push_back(dt_after); }
    else { stage_before_lift_.push_back(stage_before); t_lift_.push_back(time); dt_lift_.push_back(dt_after); }
  }
  double T_ = 0, dt_ideal_ = 0;
  int N_ideal_ = 0, N_ = 0, max_events_ = 0;
  std::vector<double> t_, dt_, t_impulse_, t_lift_, dt_aux_, dt_lift_;
  std::vector<int> phase_, impulse_after_, lift_after_, stage_before_impulse_, stage_before_lift_;
};

// One row of the flattened schedule, in the order the Riccati recursion visits the stages
// (riccati_recursion_solver.cpp:48-107): grid stage i, then the impulse + aux pair or the lift stage that follows it.
struct ScheduledStage {
  enum Kind { GRID = 0, IMPULSE = 1, AUX = 2, LIFT = 3, TERMINAL = 4 };
  int kind;
  int index;            // grid stage / impulse index / lift index
  double t, dt;         // start time and length (0 for IMPULSE and TERMINAL)
  int contact_phase;    // contact status of the stage = ContactSequence::contactStatus(contact_phase)
  int constraint_stage; // argument of createConstraintsData: grid index, 0 for aux / lift stages, -1 for impulses
                        // (ocp_linearizer.cpp:40-68)
  int before_impulse;   // 1: the switching constraint of the next impulse is imposed here (two stages ahead of the
                        // touch-down, ocp_linearizer.hxx:139-150), and its impulse index in switching_impulse
  int switching_impulse;
};

inline std::vector<ScheduledStage> flattenSchedule(const OCPDiscretizer& d) {
  std::vector<ScheduledStage> out;
  for (int i = 0; i < d.N(); ++i) {
    ScheduledStage s{ScheduledStage::GRID, i, d.t(i), d.dt(i), d.contactPhase(i), i, 0, -1};
    if (i + 1 < d.N() && d.isTimeStageBeforeImpulse(i + 1)) {
      s.before_impulse = 1;
      s.switching_impulse = d.impulseIndexAfterTimeStage(i + 1);
    }
    out.push_back(s);
    if (d.isTimeStageBeforeImpulse(i)) {
      const int k = d.impulseIndexAfterTimeStage(i);
      out.push_back(ScheduledStage{ScheduledStage::IMPULSE, k, d.t_impulse(k), 0.0, d.contactPhase(i + 1), -1, 0, -1});
      out.push_back(ScheduledStage{ScheduledStage::AUX, k, d.t_impulse(k), d.dt_aux(k), d.contactPhase(i + 1), 0, 0, -1});
    } else if (d.isTimeStageBeforeLift(i)) {
      const int k = d.liftIndexAfterTimeStage(i);
      out.push_back(ScheduledStage{ScheduledStage::LIFT, k, d.t_lift(k), d.dt_lift(k), d.contactPhase(i + 1), 0, 0, -1});
    }
  }
  out.push_back(ScheduledStage{ScheduledStage::TERMINAL, d.N(), d.t(d.N()), 0.0, d.contactPhase(d.N()), -1, 0, -1});
  return out;
}

}  // namespace idocp_b200
#endif  // IDOCP_B200_HYBRID_HPP_

/*
 * hybrid_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT THE PRODUCT) for the contact schedule of idocp's
 * hybrid optimal control problems.  Plain-C restatement, in the reference's own
 * data organisation (parallel event queues, two-cursor grid walk), of
 *
 *   ContactStatus / ImpulseStatus   include/idocp/robot/contact_status.hxx, impulse_status.hxx
 *   DiscreteEvent                   include/idocp/hybrid/discrete_event.hxx:78-104
 *   ContactSequence                 include/idocp/hybrid/contact_sequence.hxx:50-336
 *   OCPDiscretizer                  include/idocp/hybrid/ocp_discretizer.hxx:61-377
 *
 * The product (include/idocp_b200/hybrid.hpp) is organised differently (one phase list, merged event walk), so
 * agreement between the two on random and on the reference's own test scenarios (test/hybrid/, the *_test.cpp files) is a
 * real check.  This part of the reference has no dependency on pinocchio / Eigen beyond a 3-vector, so the
 * restatement is literal; "parity unpinned" only in the sense that the reference itself cannot be compiled here.
 *
 * Only tests/ may load this (through oracle/liboracle.so).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "hybrid_oracle.h"

typedef struct {
  int max_point_contacts;
  int active[HY_MAX_CONTACTS];
  double points[HY_MAX_CONTACTS][3];
} hy_status_t;

typedef struct {
  hy_status_t pre, post, impulse;
  int exist_impulse, exist_lift;
} hy_event_t;

/* DiscreteEvent::setDiscreteEvent (discrete_event.hxx:78-104) */
static void hy_event_set(hy_event_t* e, const hy_status_t* pre, const hy_status_t* post) {
  e->exist_impulse = 0;
  e->exist_lift = 0;
  e->impulse.max_point_contacts = pre->max_point_contacts;
  for (int i = 0; i < pre->max_point_contacts; ++i) {
    if (pre->active[i]) {
      e->impulse.active[i] = 0;
      if (!post->active[i]) e->exist_lift = 1;
    } else {
      if (post->active[i]) {
        e->impulse.active[i] = 1;
        e->exist_impulse = 1;
      } else {
        e->impulse.active[i] = 0;
      }
    }
  }
  memcpy(e->impulse.points, post->points, sizeof(post->points));
  e->pre = *pre;
  e->post = *post;
}

/* ContactStatus::operator== (contact_status.hxx:34-46): activity + isApprox contact points */
static int hy_status_equal(const hy_status_t* a, const hy_status_t* b) {
  for (int i = 0; i < a->max_point_contacts; ++i) {
    if (a->active[i] != b->active[i]) return 0;
    double d2 = 0, a2 = 0, b2 = 0;
    for (int k = 0; k < 3; ++k) {
      const double d = a->points[i][k] - b->points[i][k];
      d2 += d * d; a2 += a->points[i][k] * a->points[i][k]; b2 += b->points[i][k] * b->points[i][k];
    }
    if (!(d2 <= 1e-24 * (a2 < b2 ? a2 : b2))) return 0;
  }
  return 1;
}

/* ContactSequence: the reference's deques as arrays with explicit sizes */
struct oracle_contact_sequence {
  int max_point_contacts, max_num_events;
  hy_status_t default_status;
  int n_status;
  hy_status_t statuses[HY_MAX_EVENTS + 1];
  int n_impulse;
  hy_event_t impulse_events[HY_MAX_EVENTS];
  int event_index_impulse[HY_MAX_EVENTS];
  double impulse_time[HY_MAX_EVENTS];
  int n_lift;
  int event_index_lift[HY_MAX_EVENTS];
  double lift_time[HY_MAX_EVENTS];
  int n_event;
  double event_time[HY_MAX_EVENTS];
  int is_impulse_event[HY_MAX_EVENTS];
};

static void hy_status_init(hy_status_t* s, int n) {
  memset(s, 0, sizeof(*s));
  s->max_point_contacts = n;
}

oracle_contact_sequence_t* oracle_cs_create(int max_point_contacts, int max_num_events) {
  if (max_point_contacts > HY_MAX_CONTACTS || max_num_events <= 0 || max_num_events > HY_MAX_EVENTS) return NULL;
  oracle_contact_sequence_t* cs = (oracle_contact_sequence_t*)calloc(1, sizeof(*cs));
  cs->max_point_contacts = max_point_contacts;
  cs->max_num_events = max_num_events;
  hy_status_init(&cs->default_status, max_point_contacts);
  cs->statuses[0] = cs->default_status;   /* ctor: clear_all(); push_back(default) (:27-28) */
  cs->n_status = 1;
  return cs;
}
void oracle_cs_destroy(oracle_contact_sequence_t* cs) { free(cs); }

static void hy_fill(hy_status_t* s, int n, const int* active, const double* points) {
  hy_status_init(s, n);
  for (int i = 0; i < n; ++i) {
    s->active[i] = active[i] ? 1 : 0;
    if (points) for (int k = 0; k < 3; ++k) s->points[i][k] = points[3 * i + k];
  }
}

/* setContactStatusUniformly (:50-54) */
void oracle_cs_set_uniform(oracle_contact_sequence_t* cs, const int* active, const double* points) {
  cs->n_status = cs->n_impulse = cs->n_lift = cs->n_event = 0;
  hy_fill(&cs->statuses[0], cs->max_point_contacts, active, points);
  cs->n_status = 1;
}

/* push_back(ContactStatus, event_time) (:106-110 -> :56-103); returns 0 or the number of the failed check */
int oracle_cs_push_back(oracle_contact_sequence_t* cs, const int* active, const double* points, double event_time) {
  hy_status_t post;
  hy_fill(&post, cs->max_point_contacts, active, points);
  hy_event_t ev;
  if (cs->n_status == 0) return 1;
  hy_event_set(&ev, &cs->statuses[cs->n_status - 1], &post);
  if (!(ev.exist_impulse || ev.exist_lift)) return 2;
  if (!hy_status_equal(&ev.pre, &cs->statuses[cs->n_status - 1])) return 3;
  if (cs->n_event + 1 > cs->max_num_events) return 4;
  if (cs->n_impulse > 0 || cs->n_lift > 0)
    if (event_time <= cs->event_time[cs->n_event - 1]) return 5;
  cs->statuses[cs->n_status++] = ev.post;
  cs->event_time[cs->n_event] = event_time;
  if (ev.exist_impulse) {
    cs->impulse_events[cs->n_impulse] = ev;
    cs->event_index_impulse[cs->n_impulse] = cs->n_status - 2;
    cs->impulse_time[cs->n_impulse] = event_time;
    cs->n_impulse++;
    cs->is_impulse_event[cs->n_event] = 1;
  } else {
    cs->event_index_lift[cs->n_lift] = cs->n_status - 2;
    cs->lift_time[cs->n_lift] = event_time;
    cs->n_lift++;
    cs->is_impulse_event[cs->n_event] = 0;
  }
  cs->n_event++;
  return 0;
}

/* pop_back (:113-132) */
void oracle_cs_pop_back(oracle_contact_sequence_t* cs) {
  if (cs->n_event > 0) {
    if (cs->is_impulse_event[cs->n_event - 1]) cs->n_impulse--;
    else cs->n_lift--;
    cs->n_event--;
    cs->n_status--;
  } else if (cs->n_status > 0) {
    cs->statuses[0] = cs->default_status;
    cs->n_status = 1;
  }
}

/* pop_front (:135-154).  The reference leaves event_index_impulse_ / event_index_lift_ un-renumbered after a
 * pop_front (they then point one event too far); they are renumbered here, which is what every later use needs. */
void oracle_cs_pop_front(oracle_contact_sequence_t* cs) {
  if (cs->n_event > 0) {
    if (cs->is_impulse_event[0]) {
      memmove(&cs->impulse_events[0], &cs->impulse_events[1], sizeof(hy_event_t) * (size_t)(cs->n_impulse - 1));
      memmove(&cs->event_index_impulse[0], &cs->event_index_impulse[1], sizeof(int) * (size_t)(cs->n_impulse - 1));
      memmove(&cs->impulse_time[0], &cs->impulse_time[1], sizeof(double) * (size_t)(cs->n_impulse - 1));
      cs->n_impulse--;
    } else {
      memmove(&cs->event_index_lift[0], &cs->event_index_lift[1], sizeof(int) * (size_t)(cs->n_lift - 1));
      memmove(&cs->lift_time[0], &cs->lift_time[1], sizeof(double) * (size_t)(cs->n_lift - 1));
      cs->n_lift--;
    }
    memmove(&cs->event_time[0], &cs->event_time[1], sizeof(double) * (size_t)(cs->n_event - 1));
    memmove(&cs->is_impulse_event[0], &cs->is_impulse_event[1], sizeof(int) * (size_t)(cs->n_event - 1));
    cs->n_event--;
    memmove(&cs->statuses[0], &cs->statuses[1], sizeof(hy_status_t) * (size_t)(cs->n_status - 1));
    cs->n_status--;
    for (int i = 0; i < cs->n_impulse; ++i) cs->event_index_impulse[i]--;
    for (int i = 0; i < cs->n_lift; ++i) cs->event_index_lift[i]--;
  } else if (cs->n_status > 0) {
    cs->statuses[0] = cs->default_status;
    cs->n_status = 1;
  }
}

/* updateImpulseTime / updateLiftTime (:157-240) */
int oracle_cs_update_event_time(oracle_contact_sequence_t* cs, int impulse, int index, double time) {
  const int count = impulse ? cs->n_impulse : cs->n_lift;
  if (count <= 0) return 1;
  if (index < 0) return 2;
  if (index >= count) return 3;
  const int event_index = impulse ? cs->event_index_impulse[index] : cs->event_index_lift[index];
  if (event_index > 0) {
    if (cs->event_time[event_index - 1] >= time) return 4;
  } else if (event_index + 1 < cs->n_event) {
    if (cs->event_time[event_index + 1] <= time) return 5;
  }
  if (impulse) cs->impulse_time[index] = time;
  else cs->lift_time[index] = time;
  cs->event_time[event_index] = time;
  return 0;
}

/* setContactPoints (:243-262); the impulse of the event that starts the phase receives the points too (the
 * reference indexes impulse_events_ with contact_phase - 1, right only while all earlier events are impulses) */
int oracle_cs_set_contact_points(oracle_contact_sequence_t* cs, int phase, const double* points) {
  if (phase >= cs->n_status) return 1;
  for (int i = 0; i < cs->max_point_contacts; ++i)
    for (int k = 0; k < 3; ++k) cs->statuses[phase].points[i][k] = points[3 * i + k];
  if (phase > 0 && cs->is_impulse_event[phase - 1]) {
    for (int j = 0; j < cs->n_impulse; ++j)
      if (cs->event_index_impulse[j] == phase - 1)
        memcpy(cs->impulse_events[j].impulse.points, cs->statuses[phase].points, sizeof(cs->statuses[phase].points));
  }
  return 0;
}

void oracle_cs_counts(const oracle_contact_sequence_t* cs, int* phases, int* impulses, int* lifts) {
  *phases = cs->n_status; *impulses = cs->n_impulse; *lifts = cs->n_event - cs->n_impulse;
}
void oracle_cs_get_phase(const oracle_contact_sequence_t* cs, int phase, int* active, double* points) {
  for (int i = 0; i < cs->max_point_contacts; ++i) {
    active[i] = cs->statuses[phase].active[i];
    for (int k = 0; k < 3; ++k) points[3 * i + k] = cs->statuses[phase].points[i][k];
  }
}
void oracle_cs_get_impulse(const oracle_contact_sequence_t* cs, int impulse_index, int* active, double* points, double* time) {
  for (int i = 0; i < cs->max_point_contacts; ++i) {
    active[i] = cs->impulse_events[impulse_index].impulse.active[i];
    for (int k = 0; k < 3; ++k) points[3 * i + k] = cs->impulse_events[impulse_index].impulse.points[i][k];
  }
  *time = cs->impulse_time[impulse_index];
}
double oracle_cs_lift_time(const oracle_contact_sequence_t* cs, int lift_index) { return cs->lift_time[lift_index]; }

/* ------------------------------------------------------------------------------------------------------ */
/* OCPDiscretizer::discretizeOCP (ocp_discretizer.hxx:61-72): countDiscreteEvents, countTimeSteps,          */
/* countTimeStages, countContactPhase.  A fresh discretiser per call (the reference object keeps arrays of   */
/* earlier calls; with a shrinking event count it would read those stale entries at index N_impulse).       */
/* ------------------------------------------------------------------------------------------------------ */
/* oracle_discretization_t: hybrid_oracle.h */

static int hy_well_defined(const oracle_discretization_t* d) {
  for (int i = 0; i < d->N; ++i)
    if (d->before_impulse_flag[i] && d->before_lift_flag[i]) return 0;
  for (int i = 0; i < d->N - 1; ++i)
    if (d->before_impulse_flag[i] && d->before_impulse_flag[i + 1]) return 0;
  return 1;
}

int oracle_discretize_ocp(const oracle_contact_sequence_t* cs, double T, int N_ideal, double t,
                          oracle_discretization_t* d) {
  if (N_ideal > HY_MAX_N) return -1;
  const double min_dt = sqrt(2.2204460492503131e-16);   /* ocp_discretizer.hpp:108-109 */
  const double dt_ideal = T / N_ideal;
  const double max_dt = dt_ideal - min_dt;
  memset(d, 0, sizeof(*d));
  for (int i = 0; i <= N_ideal; ++i) { d->dt[i] = dt_ideal; d->impulse_after[i] = -1; d->lift_after[i] = -1; }
  for (int i = 0; i <= HY_MAX_EVENTS; ++i) { d->stage_before_impulse[i] = -1; d->stage_before_lift[i] = -1; }
  /* countDiscreteEvents (:220-237) */
  d->N_impulse = cs->n_impulse;
  for (int i = 0; i < d->N_impulse; ++i) {
    d->t_impulse[i] = cs->impulse_time[i];
    d->stage_before_impulse[i] = (int)floor((d->t_impulse[i] - t) / dt_ideal);
  }
  d->N_lift = cs->n_event - cs->n_impulse;
  for (int i = 0; i < d->N_lift; ++i) {
    d->t_lift[i] = cs->lift_time[i];
    d->stage_before_lift[i] = (int)floor((d->t_lift[i] - t) / dt_ideal);
  }
  d->N = N_ideal;
  /* countTimeSteps (:240-300) */
  int impulse_index = 0, lift_index = 0, num_events_on_grid = 0;
  for (int i = 0; i < N_ideal; ++i) {
    const int stage = i - num_events_on_grid;
    if (impulse_index < d->N_impulse && i == d->stage_before_impulse[impulse_index]) {
      d->dt[stage] = d->t_impulse[impulse_index] - i * dt_ideal - t;
      if (d->dt[stage] <= min_dt) {
        d->stage_before_impulse[impulse_index] = stage - 1;
        d->dt_aux[impulse_index] = dt_ideal;
        d->t[stage] = t + (i - 1) * dt_ideal;
        ++num_events_on_grid;
        ++impulse_index;
      } else if (d->dt[stage] >= max_dt) {
        d->stage_before_impulse[impulse_index] = i + 1;
        d->t[stage] = t + i * dt_ideal;
      } else {
        d->stage_before_impulse[impulse_index] = stage;
        d->dt_aux[impulse_index] = dt_ideal - d->dt[stage];
        d->t[stage] = t + i * dt_ideal;
        ++impulse_index;
      }
    } else if (lift_index < d->N_lift && i == d->stage_before_lift[lift_index]) {
      d->dt[stage] = d->t_lift[lift_index] - i * dt_ideal - t;
      if (d->dt[stage] <= min_dt) {
        d->stage_before_lift[lift_index] = stage - 1;
        d->dt_lift[lift_index] = dt_ideal;
        d->t[stage] = t + (i - 1) * dt_ideal;
        ++num_events_on_grid;
        ++lift_index;
      } else if (d->dt[stage] >= max_dt) {
        d->stage_before_lift[lift_index] = i + 1;
        d->t[stage] = t + i * dt_ideal;
      } else {
        d->stage_before_lift[lift_index] = stage;
        d->dt_lift[lift_index] = dt_ideal - d->dt[stage];
        d->t[stage] = t + i * dt_ideal;
        ++lift_index;
      }
    } else {
      d->dt[stage] = dt_ideal;
      d->t[stage] = t + i * dt_ideal;
    }
  }
  d->N = N_ideal - num_events_on_grid;
  d->t[d->N] = t + T;
  /* countTimeStages (:303-345) */
  impulse_index = 0; lift_index = 0;
  for (int i = 0; i < d->N; ++i) {
    if (impulse_index < d->N_impulse && i == d->stage_before_impulse[impulse_index]) {
      d->before_impulse_flag[i] = 1;
      d->impulse_after[i] = impulse_index++;
    } else {
      d->before_impulse_flag[i] = 0;
      d->impulse_after[i] = -1;
    }
    if (lift_index < d->N_lift && i == d->stage_before_lift[lift_index]) {
      d->before_lift_flag[i] = 1;
      d->lift_after[i] = lift_index++;
    } else {
      d->before_lift_flag[i] = 0;
      d->lift_after[i] = -1;
    }
  }
  d->before_impulse_flag[d->N] = 0;
  d->before_lift_flag[d->N] = 0;
  /* countContactPhase (:348-359) */
  int num_events = 0;
  for (int i = 0; i < d->N; ++i) {
    d->contact_phase[i] = num_events;
    if (d->before_impulse_flag[i] || d->before_lift_flag[i]) ++num_events;
  }
  d->contact_phase[d->N] = num_events;
  /* every event of the sequence must have found its stage; isWellDefined (:205-217)
   * -- except the events beyond the end of the horizon (cell >= N_ideal), which the reference simply never visits */
  int stray = 0;
  for (int i = impulse_index; i < d->N_impulse; ++i)
    if ((int)floor((d->t_impulse[i] - t) / dt_ideal) < N_ideal) stray = 1;
  for (int i = lift_index; i < d->N_lift; ++i)
    if ((int)floor((d->t_lift[i] - t) / dt_ideal) < N_ideal) stray = 1;
  d->well_defined = hy_well_defined(d) && !stray;
  return 0;
}

int oracle_discretization_size(void) { return (int)sizeof(oracle_discretization_t); }
int oracle_hybrid_limits(int* max_contacts, int* max_events, int* max_n) {
  *max_contacts = HY_MAX_CONTACTS; *max_events = HY_MAX_EVENTS; *max_n = HY_MAX_N;
  return 0;
}

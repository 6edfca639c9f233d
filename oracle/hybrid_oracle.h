/* oracle/hybrid_oracle.h -- TEST INFRASTRUCTURE: types shared by hybrid_oracle.c (contact schedule) and fb_ocp.c
 * (the floating-base OCP oracle that consumes the schedule). */
#ifndef ORACLE_HYBRID_ORACLE_H_
#define ORACLE_HYBRID_ORACLE_H_
#define HY_MAX_CONTACTS 8
#define HY_MAX_EVENTS 64
#define HY_MAX_N 1024

typedef struct oracle_contact_sequence oracle_contact_sequence_t;

typedef struct {
  int N, N_impulse, N_lift, well_defined;
  double t[HY_MAX_N + 1], dt[HY_MAX_N + 1];
  int contact_phase[HY_MAX_N + 1], impulse_after[HY_MAX_N + 1], lift_after[HY_MAX_N + 1];
  int before_impulse_flag[HY_MAX_N + 1], before_lift_flag[HY_MAX_N + 1];
  int stage_before_impulse[HY_MAX_EVENTS + 1], stage_before_lift[HY_MAX_EVENTS + 1];
  double t_impulse[HY_MAX_EVENTS + 1], t_lift[HY_MAX_EVENTS + 1], dt_aux[HY_MAX_EVENTS + 1], dt_lift[HY_MAX_EVENTS + 1];
} oracle_discretization_t;

void oracle_cs_counts(const oracle_contact_sequence_t* cs, int* phases, int* impulses, int* lifts);
void oracle_cs_get_phase(const oracle_contact_sequence_t* cs, int phase, int* active, double* points);
void oracle_cs_get_impulse(const oracle_contact_sequence_t* cs, int impulse_index, int* active, double* points, double* time);
int oracle_discretize_ocp(const oracle_contact_sequence_t* cs, double T, int N_ideal, double t, oracle_discretization_t* d);
#endif

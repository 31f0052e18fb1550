/* TEST INFRASTRUCTURE (see go1_physics_oracle.c). fp64 restatement of the rigid-body substep. */
#pragma once
#ifdef __cplusplus
extern "C" {
#endif
typedef struct {
    double dt, gravity[3];
    double erp, cfm, max_depen_vel, contact_margin, bounce_threshold;
    int pgs_iters, _pad;
    double terrain_friction, terrain_restitution;
    double pen_k[4], pen_c[4], pen_mt;   /* penalty contacts: base corners, hips, knees, calf-mid */
    double limit_k, limit_c;             /* implicit joint-limit spring/damper */
    const short* hf; int hf_rows, hf_cols; double hf_hscale, hf_vscale, hf_border;
} Go1PhysParams;
typedef struct { double pos[3], quat[4], linvel[3], angvel[3], q[12], qd[12]; } Go1PhysState; /* world-frame root, xyzw */
typedef struct { double friction, restitution, payload, com_disp[3]; } Go1PhysDR;
typedef struct { double contact_force[17][3]; } Go1PhysOut;  /* Isaac Gym body order: base, then per leg hip,thigh,calf,foot */

void go1_oracle_default_params(Go1PhysParams* P);
void go1_oracle_substep(const Go1PhysParams* P, const Go1PhysDR* dr, Go1PhysState* s, const double tau[12], Go1PhysOut* out);
void go1_oracle_substep_batch(const Go1PhysParams* P, int n, const Go1PhysDR* dr, Go1PhysState* s, const double* tau, Go1PhysOut* out);
/* cap on the worker threads of the *_batch calls (0 = all online processors) */
void go1_oracle_set_threads(int t);
void go1_oracle_feet_batch(int n, const Go1PhysState* s, double* foot_pos, double* foot_vel);
void go1_oracle_feet(const Go1PhysState* s, double foot_pos[4][3], double foot_vel[4][3]);
void go1_oracle_aba(const Go1PhysParams* P, const Go1PhysDR* dr, const Go1PhysState* s, const double tau[12], double a0_out[6], double qdd_out[12]);
#ifdef __cplusplus
}
#endif

/* orc_model.h -- data structures of the CPU ORACLE (test infrastructure, NOT product code).
 *
 * The oracle is a float64, one-environment, scalar restatement of the reference's HandMimic hot path:
 * MuJoCo 2.1 `mj_step` semantics (MuJoCo itself is an un-vendored third-party
 * binary -- mujoco210 / mujoco_py==2.1.2.14, /root/reference/requirements.txt:16) plus the env logic of
 * /root/reference/uhc/envs/ho_im4.py and ho_reward.py.  PARITY UNPINNED: the reference ships no golden
 * vectors and MuJoCo cannot be run here; the oracle is pinned only by analytic known-answer tests.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
 */
#ifndef ORC_MODEL_H
#define ORC_MODEL_H

#include <stddef.h>

#define ORC_NB 32      /* max bodies */
#define ORC_NV 32      /* max dofs */
#define ORC_NQ 40      /* max qpos */
#define ORC_NG 32      /* max geoms */
#define ORC_NJ 32      /* max joints */
#define ORC_NPAIR 64   /* max explicit pairs */
#define ORC_NCON 100   /* nconmax of the hand MJCF (sphere_mesh_hand_add_geom.xml:8) */
#define ORC_NEFC 500   /* njmax */
#define ORC_OBS 617
#define ORC_FRAME 212  /* floats per expert frame */
/* record layouts, identical to include/hoic.h (HOIC_REC_*, HOIC_SURF_*) */
#define ORC_REC_MAXCON 32
#define ORC_REC_CON 116
#define ORC_REC_STRIDE (ORC_REC_CON + 19 * ORC_REC_MAXCON)
#define ORC_SURF_MAXPT 95
#define ORC_SURF_PT 8
#define ORC_SURF_STRIDE (ORC_SURF_PT + 15 * ORC_SURF_MAXPT + 3)
#define ORC_NHG 24     /* max hand geoms tracked for contact averaging */

enum { GEOM_PLANE = 0, GEOM_SPHERE = 2, GEOM_CAPSULE = 3, GEOM_BOX = 6, GEOM_MESH = 7 };
enum { JNT_FREE = 0, JNT_SLIDE = 2, JNT_HINGE = 3 };
enum { EFC_FRICTION = 0, EFC_LIMIT = 1, EFC_CONTACT_FRICTIONLESS = 2, EFC_CONTACT_PYRAMIDAL = 3 };
enum { ST_SATISFIED = 0, ST_QUADRATIC = 1, ST_LINEARNEG = 2, ST_LINEARPOS = 3 };

typedef struct {
  int nq, nv, nbody, ngeom, njnt, nu, npair, nmeshvert;
  double timestep, gravity[3], tolerance, impratio, mpr_tolerance, meaninertia;
  int iterations, mpr_iterations, nconmax, njmax;
  int body_parentid[ORC_NB], body_rootid[ORC_NB], body_weldid[ORC_NB], body_dofadr[ORC_NB], body_dofnum[ORC_NB];
  int body_jntadr[ORC_NB], body_jntnum[ORC_NB], body_geomadr[ORC_NB], body_geomnum[ORC_NB];
  double body_pos[ORC_NB][3], body_quat[ORC_NB][4], body_ipos[ORC_NB][3], body_iquat[ORC_NB][4];
  double body_mass[ORC_NB], body_inertia[ORC_NB][3], body_subtreemass[ORC_NB], body_invweight0[ORC_NB][2];
  int jnt_type[ORC_NJ], jnt_bodyid[ORC_NJ], jnt_qposadr[ORC_NJ], jnt_dofadr[ORC_NJ], jnt_limited[ORC_NJ];
  double jnt_pos[ORC_NJ][3], jnt_axis[ORC_NJ][3], jnt_range[ORC_NJ][2], jnt_margin[ORC_NJ];
  double jnt_solref[ORC_NJ][2], jnt_solimp[ORC_NJ][5];
  int dof_bodyid[ORC_NV], dof_jntid[ORC_NV], dof_parentid[ORC_NV];
  double dof_armature[ORC_NV], dof_damping[ORC_NV], dof_frictionloss[ORC_NV], dof_invweight0[ORC_NV];
  double dof_solref[ORC_NV][2], dof_solimp[ORC_NV][5];
  double qpos0[ORC_NQ];
  int geom_type[ORC_NG], geom_bodyid[ORC_NG], geom_contype[ORC_NG], geom_conaffinity[ORC_NG];
  int geom_condim[ORC_NG], geom_priority[ORC_NG], geom_vertadr[ORC_NG], geom_vertnum[ORC_NG];
  double geom_pos[ORC_NG][3], geom_quat[ORC_NG][4], geom_size[ORC_NG][3], geom_rbound[ORC_NG];
  double geom_friction[ORC_NG][3], geom_solref[ORC_NG][2], geom_solimp[ORC_NG][5], geom_solmix[ORC_NG];
  double geom_margin[ORC_NG], geom_gap[ORC_NG];
  double* mesh_vert; /* [nmeshvert][3] */
  int actuator_dofid[ORC_NV];
  int pair_geom1[ORC_NPAIR], pair_geom2[ORC_NPAIR], pair_dim[ORC_NPAIR], pair_signature[ORC_NPAIR];
  double pair_friction[ORC_NPAIR][5], pair_solref[ORC_NPAIR][2], pair_solimp[ORC_NPAIR][5];
  double pair_margin[ORC_NPAIR], pair_gap[ORC_NPAIR];
  /* env constants (ho_im4.py:72-107 + YAML) */
  int n_hand_body, hand_body_idx[ORC_NB], obj_body, palm_body, hand_geom0, hand_geom1, obj_geom0, obj_geom1;
  int sim_step, w_size, residual_force, surface_contact, explain_force;
  double base_pose[ORC_NV], ctrl_scale[ORC_NV], jkp[ORC_NV], jkd[ORC_NV], torque_lim[ORC_NV];
  double hand_mass, rf_scale, rt_scale, pos_thresh, rot_thresh, jpos_thresh, obj_pos_thresh, obj_rot_thresh;
} orc_model;

typedef struct {
  double dist, pos[3], frame[9], includemargin, friction[5], solref[2], solimp[5], mu;
  int dim, geom1, geom2, exclude, efc_address;
} orc_contact;

typedef struct {
  /* state */
  double qpos[ORC_NQ], qvel[ORC_NV], qacc_warmstart[ORC_NV], ctrl[ORC_NV], qfrc_applied[ORC_NV], time;
  /* position-dependent */
  double xpos[ORC_NB][3], xquat[ORC_NB][4], xmat[ORC_NB][9], xipos[ORC_NB][3], ximat[ORC_NB][9];
  double xanchor[ORC_NJ][3], xaxis[ORC_NJ][3], geom_xpos[ORC_NG][3], geom_xmat[ORC_NG][9];
  double subtree_com[ORC_NB][3], cinert[ORC_NB][10], crb[ORC_NB][10], cdof[ORC_NV][6];
  double qM[ORC_NV][ORC_NV], qL[ORC_NV][ORC_NV]; /* dense mass matrix and its Cholesky factor (lower) */
  int ncon, nefc;
  orc_contact contact[ORC_NCON];
  /* constraints */
  double efc_J[ORC_NEFC][ORC_NV];
  int efc_type[ORC_NEFC], efc_id[ORC_NEFC], efc_state[ORC_NEFC];
  double efc_pos[ORC_NEFC], efc_margin[ORC_NEFC], efc_frictionloss[ORC_NEFC], efc_diagApprox[ORC_NEFC];
  double efc_KBIP[ORC_NEFC][4], efc_D[ORC_NEFC], efc_R[ORC_NEFC], efc_vel[ORC_NEFC], efc_aref[ORC_NEFC];
  double efc_force[ORC_NEFC];
  /* velocity-dependent */
  double cvel[ORC_NB][6], cdof_dot[ORC_NV][6];
  double qfrc_bias[ORC_NV], qfrc_passive[ORC_NV], qfrc_actuator[ORC_NV], qfrc_smooth[ORC_NV];
  double qacc_smooth[ORC_NV], qacc[ORC_NV], qfrc_constraint[ORC_NV];
  /* diagnostics */
  int solver_iter, warn_contactfull, warn_cnstrfull, warn_bad, mpr_calls, mpr_iters, ls_evals;
  double solver_improvement, solver_gradient;
} orc_data;

typedef struct {
  const orc_model* m;
  orc_data d;
  /* expert motion: frames[T][212] (owned copy), already sliced at start_idx by the caller */
  double* motion;
  int expert_len;
  int cur_t, start_ind, mode_train;
  double reward_w[24];
  double end_reward;
  /* per-step scratch kept for inspection */
  double obj_vf[3], obj_vt[3], rfc_score;
  int sub_iters[32], sub_ncon[32];   /* per-substep solver iterations / contacts of the last step (analysis accessor) */
  double obj_avg_acc[6], geom_avg_vel[ORC_NG][3], geom_avg_ang_vel[ORC_NG][3];
  double contact_sum[ORC_NHG][12];
  int contact_count[ORC_NHG];
  int n_avg_cp, avg_cp_geom[ORC_NHG];
  double avg_cps[ORC_NHG][12], cp_ts[ORC_NHG];
  int qp_n, qp_iters;
  double rest_force[3], rest_torque[3];
  /* test-mode recording (MotionData of ho_im4.py:27-41, filled by compute_data_frame :599-609 and solve_rfc :1072-1081) */
  int record;
  double* rec_frames;   /* [sim_step][ORC_REC_STRIDE] */
  double* rec_surface;  /* [ORC_SURF_STRIDE] */
} orc_env;

/* physics (orc_physics.c) */
void orc_fwd_position(const orc_model* m, orc_data* d);
void orc_fwd_velocity(const orc_model* m, orc_data* d);
void orc_fwd_actuation(const orc_model* m, orc_data* d);
void orc_fwd_acceleration(const orc_model* m, orc_data* d);
void orc_fwd_constraint(const orc_model* m, orc_data* d);
void orc_forward(const orc_model* m, orc_data* d);
void orc_mj_step(const orc_model* m, orc_data* d);
void orc_reset_data(const orc_model* m, orc_data* d);
void orc_apply_ft(const orc_model* m, const orc_data* d, const double force[3], const double torque[3],
                  const double point[3], int body, double* qfrc);
void orc_jac(const orc_model* m, const orc_data* d, double* jacp, double* jacr, const double point[3], int body);
int orc_chol_factor(double* A, int n, int lda);
void orc_chol_solve(const double* L, int n, int lda, double* x);

/* collision (orc_collision.c) */
void orc_collision(const orc_model* m, orc_data* d);

#endif

/* orc_physics.c -- CPU ORACLE (test infrastructure): float64 restatement of MuJoCo 2.1 mj_step for the
 * HandMimic model class (hinge/slide/free joints, motors, dof friction loss, joint limits, pyramidal
 * contacts, Newton solver, Euler with implicit joint damping).  Follows SURVEY.md Appendix B.1-B.5, B.8-B.11;
 * MuJoCo is a third-party binary (mujoco210, un-vendored) so this restates its published algorithm.
 * Reference call sites: uhc/envs/ho_im4.py:545 (sim.step), uhc/khrylib/rl/envs/common/mujoco_env.py:96,113-114
 * (sim.reset / set_state / forward), ho_im4.py:398 (mj_fullM), :492,:527 (mj_applyFT).  PARITY UNPINNED.
 */
#include <stdio.h>
#include <stdlib.h>
#include "../include/hoic_blob.h"
#include "orc_math.h"
#include "orc_model.h"

/* ------------------------------------------------------------------------------------------ model load */

static int get_f64(const hoic_blob_view* v, const char* name, double* dst, size_t maxn) {
  int dt; size_t n;
  const void* p = hoic_blob_find(v, name, &dt, &n);
  if (!p || dt != HOIC_BLOB_F64 || n > maxn) { fprintf(stderr, "orc: bad blob field %s\n", name); return -1; }
  memcpy(dst, p, n * sizeof(double));
  return (int)n;
}
static int get_i32(const hoic_blob_view* v, const char* name, int* dst, size_t maxn) {
  int dt; size_t n;
  const void* p = hoic_blob_find(v, name, &dt, &n);
  if (!p || dt != HOIC_BLOB_I32 || n > maxn) { fprintf(stderr, "orc: bad blob field %s\n", name); return -1; }
  memcpy(dst, p, n * sizeof(int));
  return (int)n;
}
#define GF(name, field, maxn) if (get_f64(&v, name, (double*)(field), maxn) < 0) goto fail
#define GI(name, field, maxn) if (get_i32(&v, name, (int*)(field), maxn) < 0) goto fail

orc_model* orc_model_create(const void* blob, size_t nbytes) {
  hoic_blob_view v;
  orc_model* m;
  int ids[16], n;
  double sc[16];
  if (hoic_blob_open(&v, blob, nbytes) != 0) return NULL;
  m = (orc_model*)calloc(1, sizeof(orc_model));
  GI("nq", &m->nq, 1); GI("nv", &m->nv, 1); GI("nbody", &m->nbody, 1); GI("ngeom", &m->ngeom, 1);
  GI("njnt", &m->njnt, 1); GI("nu", &m->nu, 1);
  if (m->nq > ORC_NQ || m->nv > ORC_NV || m->nbody > ORC_NB || m->ngeom > ORC_NG || m->njnt > ORC_NJ) goto fail;
  GF("opt_timestep", &m->timestep, 1); GF("opt_gravity", m->gravity, 3); GF("opt_tolerance", &m->tolerance, 1);
  GF("opt_impratio", &m->impratio, 1); GF("opt_mpr_tolerance", &m->mpr_tolerance, 1);
  GF("stat_meaninertia", &m->meaninertia, 1);
  GI("opt_iterations", &m->iterations, 1); GI("opt_mpr_iterations", &m->mpr_iterations, 1);
  GI("size_nconmax", &m->nconmax, 1); GI("size_njmax", &m->njmax, 1);
  if (m->nconmax <= 0 || m->nconmax > ORC_NCON) m->nconmax = ORC_NCON;
  if (m->njmax <= 0 || m->njmax > ORC_NEFC) m->njmax = ORC_NEFC;
  GI("body_parentid", m->body_parentid, ORC_NB); GI("body_rootid", m->body_rootid, ORC_NB);
  GI("body_weldid", m->body_weldid, ORC_NB); GI("body_dofadr", m->body_dofadr, ORC_NB);
  GI("body_dofnum", m->body_dofnum, ORC_NB); GI("body_jntadr", m->body_jntadr, ORC_NB);
  GI("body_jntnum", m->body_jntnum, ORC_NB); GI("body_geomadr", m->body_geomadr, ORC_NB);
  GI("body_geomnum", m->body_geomnum, ORC_NB);
  GF("body_pos", m->body_pos, ORC_NB * 3); GF("body_quat", m->body_quat, ORC_NB * 4);
  GF("body_ipos", m->body_ipos, ORC_NB * 3); GF("body_iquat", m->body_iquat, ORC_NB * 4);
  GF("body_mass", m->body_mass, ORC_NB); GF("body_inertia", m->body_inertia, ORC_NB * 3);
  GF("body_subtreemass", m->body_subtreemass, ORC_NB); GF("body_invweight0", m->body_invweight0, ORC_NB * 2);
  GI("jnt_type", m->jnt_type, ORC_NJ); GI("jnt_bodyid", m->jnt_bodyid, ORC_NJ);
  GI("jnt_qposadr", m->jnt_qposadr, ORC_NJ); GI("jnt_dofadr", m->jnt_dofadr, ORC_NJ);
  GI("jnt_limited", m->jnt_limited, ORC_NJ);
  GF("jnt_pos", m->jnt_pos, ORC_NJ * 3); GF("jnt_axis", m->jnt_axis, ORC_NJ * 3);
  GF("jnt_range", m->jnt_range, ORC_NJ * 2); GF("jnt_margin", m->jnt_margin, ORC_NJ);
  GF("jnt_solref", m->jnt_solref, ORC_NJ * 2); GF("jnt_solimp", m->jnt_solimp, ORC_NJ * 5);
  GI("dof_bodyid", m->dof_bodyid, ORC_NV); GI("dof_jntid", m->dof_jntid, ORC_NV);
  GI("dof_parentid", m->dof_parentid, ORC_NV);
  GF("dof_armature", m->dof_armature, ORC_NV); GF("dof_damping", m->dof_damping, ORC_NV);
  GF("dof_frictionloss", m->dof_frictionloss, ORC_NV); GF("dof_invweight0", m->dof_invweight0, ORC_NV);
  GF("dof_solref", m->dof_solref, ORC_NV * 2); GF("dof_solimp", m->dof_solimp, ORC_NV * 5);
  GF("qpos0", m->qpos0, ORC_NQ);
  GI("geom_type", m->geom_type, ORC_NG); GI("geom_bodyid", m->geom_bodyid, ORC_NG);
  GI("geom_contype", m->geom_contype, ORC_NG); GI("geom_conaffinity", m->geom_conaffinity, ORC_NG);
  GI("geom_condim", m->geom_condim, ORC_NG); GI("geom_priority", m->geom_priority, ORC_NG);
  GI("geom_vertadr", m->geom_vertadr, ORC_NG); GI("geom_vertnum", m->geom_vertnum, ORC_NG);
  GF("geom_pos", m->geom_pos, ORC_NG * 3); GF("geom_quat", m->geom_quat, ORC_NG * 4);
  GF("geom_size", m->geom_size, ORC_NG * 3); GF("geom_rbound", m->geom_rbound, ORC_NG);
  GF("geom_friction", m->geom_friction, ORC_NG * 3); GF("geom_solref", m->geom_solref, ORC_NG * 2);
  GF("geom_solimp", m->geom_solimp, ORC_NG * 5); GF("geom_solmix", m->geom_solmix, ORC_NG);
  GF("geom_margin", m->geom_margin, ORC_NG); GF("geom_gap", m->geom_gap, ORC_NG);
  {
    int dt; size_t cnt;
    const void* p = hoic_blob_find(&v, "mesh_vert", &dt, &cnt);
    if (!p || dt != HOIC_BLOB_F64) goto fail;
    m->nmeshvert = (int)(cnt / 3);
    m->mesh_vert = (double*)malloc((cnt + 3) * sizeof(double));
    memcpy(m->mesh_vert, p, cnt * sizeof(double));
  }
  GI("actuator_dofid", m->actuator_dofid, ORC_NV);
  n = get_i32(&v, "pair_geom1", m->pair_geom1, ORC_NPAIR);
  if (n < 0) goto fail;
  m->npair = n;
  GI("pair_geom2", m->pair_geom2, ORC_NPAIR); GI("pair_dim", m->pair_dim, ORC_NPAIR);
  GI("pair_signature", m->pair_signature, ORC_NPAIR);
  GF("pair_friction", m->pair_friction, ORC_NPAIR * 5); GF("pair_solref", m->pair_solref, ORC_NPAIR * 2);
  GF("pair_solimp", m->pair_solimp, ORC_NPAIR * 5); GF("pair_margin", m->pair_margin, ORC_NPAIR);
  GF("pair_gap", m->pair_gap, ORC_NPAIR);
  n = get_i32(&v, "env_hand_body_idx", m->hand_body_idx, ORC_NB);
  if (n < 0) goto fail;
  m->n_hand_body = n;
  if (get_i32(&v, "env_ids", ids, 16) != 11) goto fail;
  m->obj_body = ids[0]; m->palm_body = ids[1]; m->hand_geom0 = ids[2]; m->hand_geom1 = ids[3];
  m->obj_geom0 = ids[4]; m->obj_geom1 = ids[5]; m->sim_step = ids[6]; m->w_size = ids[7];
  m->residual_force = ids[8]; m->surface_contact = ids[9]; m->explain_force = ids[10];
  GF("env_base_pose", m->base_pose, ORC_NV); GF("env_ctrl_scale", m->ctrl_scale, ORC_NV);
  GF("env_jkp", m->jkp, ORC_NV); GF("env_jkd", m->jkd, ORC_NV); GF("env_torque_lim", m->torque_lim, ORC_NV);
  if (get_f64(&v, "env_scalars", sc, 16) != 8) goto fail;
  m->hand_mass = sc[0]; m->rf_scale = sc[1]; m->rt_scale = sc[2]; m->pos_thresh = sc[3]; m->rot_thresh = sc[4];
  m->jpos_thresh = sc[5]; m->obj_pos_thresh = sc[6]; m->obj_rot_thresh = sc[7];
  if (m->hand_geom1 - m->hand_geom0 + 1 > ORC_NHG) goto fail;
  return m;
fail:
  if (m) { free(m->mesh_vert); free(m); }
  return NULL;
}

void orc_model_destroy(orc_model* m) {
  if (m) { free(m->mesh_vert); free(m); }
}

/* ------------------------------------------------------------------------------------------ linear algebra */

/* in-place lower Cholesky of the leading n x n block (row stride lda); returns 0 ok, -1 not SPD */
int orc_chol_factor(double* A, int n, int lda) {
  for (int j = 0; j < n; j++) {
    double s = A[j * lda + j];
    for (int k = 0; k < j; k++) s -= A[j * lda + k] * A[j * lda + k];
    if (!(s > ORC_MINVAL)) return -1;
    s = sqrt(s);
    A[j * lda + j] = s;
    for (int i = j + 1; i < n; i++) {
      double t = A[i * lda + j];
      for (int k = 0; k < j; k++) t -= A[i * lda + k] * A[j * lda + k];
      A[i * lda + j] = t / s;
    }
  }
  return 0;
}
void orc_chol_solve(const double* L, int n, int lda, double* x) {
  for (int i = 0; i < n; i++) {
    double s = x[i];
    for (int k = 0; k < i; k++) s -= L[i * lda + k] * x[k];
    x[i] = s / L[i * lda + i];
  }
  for (int i = n - 1; i >= 0; i--) {
    double s = x[i];
    for (int k = i + 1; k < n; k++) s -= L[k * lda + i] * x[k];
    x[i] = s / L[i * lda + i];
  }
}

/* ------------------------------------------------------------------------------------------ spatial algebra
 * 6-vectors are (rotational; translational), expressed at the subtree CoM of the kinematic tree root, world axes.
 * 10-vector inertia: Ixx Iyy Izz Ixy Ixz Iyz  m*dx m*dy m*dz  m                                             */

static void inert_com(double* res, const double* inert, const double* mat, const double* dif, double mass) {
  /* rotate diag(inert) by mat, shift by dif (Appendix B.2) */
  double I[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      I[3 * i + j] = mat[3 * i] * inert[0] * mat[3 * j] + mat[3 * i + 1] * inert[1] * mat[3 * j + 1] +
                     mat[3 * i + 2] * inert[2] * mat[3 * j + 2];
  res[0] = I[0] + mass * (dif[1] * dif[1] + dif[2] * dif[2]);
  res[1] = I[4] + mass * (dif[0] * dif[0] + dif[2] * dif[2]);
  res[2] = I[8] + mass * (dif[0] * dif[0] + dif[1] * dif[1]);
  res[3] = I[1] - mass * dif[0] * dif[1];
  res[4] = I[2] - mass * dif[0] * dif[2];
  res[5] = I[5] - mass * dif[1] * dif[2];
  res[6] = mass * dif[0]; res[7] = mass * dif[1]; res[8] = mass * dif[2];
  res[9] = mass;
}
/* res = I * v : momentum (angular; linear) of motion v */
static void mul_inert_vec(double* res, const double* i, const double* v) {
  res[0] = i[0] * v[0] + i[3] * v[1] + i[4] * v[2] - i[8] * v[4] + i[7] * v[5];
  res[1] = i[3] * v[0] + i[1] * v[1] + i[5] * v[2] + i[8] * v[3] - i[6] * v[5];
  res[2] = i[4] * v[0] + i[5] * v[1] + i[2] * v[2] - i[7] * v[3] + i[6] * v[4];
  res[3] = i[8] * v[1] - i[7] * v[2] + i[9] * v[3];
  res[4] = i[6] * v[2] - i[8] * v[0] + i[9] * v[4];
  res[5] = i[7] * v[0] - i[6] * v[1] + i[9] * v[5];
}
/* motion cross product: res = vel x v */
static void cross_motion(double* res, const double* vel, const double* v) {
  double a[3], b[3], c[3];
  cross3(a, vel, v);          /* w x v_rot */
  cross3(b, vel, v + 3);      /* w x v_lin */
  cross3(c, vel + 3, v);      /* u x v_rot */
  res[0] = a[0]; res[1] = a[1]; res[2] = a[2];
  res[3] = b[0] + c[0]; res[4] = b[1] + c[1]; res[5] = b[2] + c[2];
}
/* force cross product: res = vel x* f */
static void cross_force(double* res, const double* vel, const double* f) {
  double a[3], b[3], c[3];
  cross3(a, vel, f);          /* w x f_rot */
  cross3(b, vel + 3, f + 3);  /* u x f_lin */
  cross3(c, vel, f + 3);      /* w x f_lin */
  res[0] = a[0] + b[0]; res[1] = a[1] + b[1]; res[2] = a[2] + b[2];
  res[3] = c[0]; res[4] = c[1]; res[5] = c[2];
}

/* ------------------------------------------------------------------------------------------ position stage */

static void local2global(const orc_data* d, double* xpos, double* xmat, const double* pos, const double* quat, int body) {
  double q[4];
  mulmat3(xpos, d->xmat[body], pos);
  add3(xpos, xpos, d->xpos[body]);
  mulquat(q, d->xquat[body], quat);
  quat2mat(xmat, q);
}

static void kinematics(const orc_model* m, orc_data* d) {
  /* normalise free-joint quaternions in qpos */
  for (int j = 0; j < m->njnt; j++)
    if (m->jnt_type[j] == JNT_FREE) normalize4(d->qpos + m->jnt_qposadr[j] + 3);
  zero3(d->xpos[0]);
  d->xquat[0][0] = 1; d->xquat[0][1] = d->xquat[0][2] = d->xquat[0][3] = 0;
  quat2mat(d->xmat[0], d->xquat[0]);
  for (int b = 1; b < m->nbody; b++) {
    int ja = m->body_jntadr[b], jn = m->body_jntnum[b], p = m->body_parentid[b];
    double xpos[3], xquat[4];
    if (jn == 1 && m->jnt_type[ja] == JNT_FREE) {
      int a = m->jnt_qposadr[ja];
      copy3(xpos, d->qpos + a);
      memcpy(xquat, d->qpos + a + 3, 4 * sizeof(double));
      copy3(d->xanchor[ja], xpos);
      copy3(d->xaxis[ja], m->jnt_axis[ja]);
    } else {
      mulmat3(xpos, d->xmat[p], m->body_pos[b]);
      add3(xpos, xpos, d->xpos[p]);
      mulquat(xquat, d->xquat[p], m->body_quat[b]);
      for (int j = ja; j < ja + jn; j++) {
        double q = d->qpos[m->jnt_qposadr[j]] - m->qpos0[m->jnt_qposadr[j]];
        rotvecquat(d->xaxis[j], m->jnt_axis[j], xquat);
        rotvecquat(d->xanchor[j], m->jnt_pos[j], xquat);
        add3(d->xanchor[j], d->xanchor[j], xpos);
        if (m->jnt_type[j] == JNT_SLIDE) {
          addscl3(xpos, xpos, d->xaxis[j], q);
        } else if (m->jnt_type[j] == JNT_HINGE) {
          double ql[4], qn[4], vec[3];
          axisangle2quat(ql, m->jnt_axis[j], q);
          mulquat(qn, xquat, ql);
          memcpy(xquat, qn, sizeof qn);
          rotvecquat(vec, m->jnt_pos[j], xquat);
          sub3(xpos, d->xanchor[j], vec);
        }
      }
    }
    normalize4(xquat);
    copy3(d->xpos[b], xpos);
    memcpy(d->xquat[b], xquat, sizeof xquat);
    quat2mat(d->xmat[b], xquat);
  }
  for (int b = 0; b < m->nbody; b++) local2global(d, d->xipos[b], d->ximat[b], m->body_ipos[b], m->body_iquat[b], b);
  for (int g = 0; g < m->ngeom; g++)
    local2global(d, d->geom_xpos[g], d->geom_xmat[g], m->geom_pos[g], m->geom_quat[g], m->geom_bodyid[g]);
}

static void com_pos(const orc_model* m, orc_data* d) {
  memset(d->subtree_com, 0, sizeof d->subtree_com);
  for (int b = m->nbody - 1; b >= 0; b--) {
    addscl3(d->subtree_com[b], d->subtree_com[b], d->xipos[b], m->body_mass[b]);
    if (b) add3(d->subtree_com[m->body_parentid[b]], d->subtree_com[m->body_parentid[b]], d->subtree_com[b]);
    if (m->body_subtreemass[b] < ORC_MINVAL) copy3(d->subtree_com[b], d->xipos[b]);
    else scl3(d->subtree_com[b], d->subtree_com[b], 1.0 / m->body_subtreemass[b]);
  }
  memset(d->cinert[0], 0, sizeof d->cinert[0]);
  for (int b = 1; b < m->nbody; b++) {
    double off[3];
    sub3(off, d->xipos[b], d->subtree_com[m->body_rootid[b]]);
    inert_com(d->cinert[b], m->body_inertia[b], d->ximat[b], off, m->body_mass[b]);
  }
  for (int j = 0; j < m->njnt; j++) {
    int b = m->jnt_bodyid[j], da = m->jnt_dofadr[j];
    double off[3];
    sub3(off, d->subtree_com[m->body_rootid[b]], d->xanchor[j]);
    if (m->jnt_type[j] == JNT_FREE) {
      memset(d->cdof[da], 0, 18 * sizeof(double));
      d->cdof[da][3] = 1; d->cdof[da + 1][4] = 1; d->cdof[da + 2][5] = 1;
      for (int k = 0; k < 3; k++) {
        double ax[3] = {d->xmat[b][k], d->xmat[b][3 + k], d->xmat[b][6 + k]};
        copy3(d->cdof[da + 3 + k], ax);
        cross3(d->cdof[da + 3 + k] + 3, ax, off);
      }
    } else if (m->jnt_type[j] == JNT_SLIDE) {
      zero3(d->cdof[da]);
      copy3(d->cdof[da] + 3, d->xaxis[j]);
    } else {
      copy3(d->cdof[da], d->xaxis[j]);
      cross3(d->cdof[da] + 3, d->xaxis[j], off);
    }
  }
}

static void crb(const orc_model* m, orc_data* d) {
  int nv = m->nv;
  memcpy(d->crb, d->cinert, sizeof d->crb);
  for (int b = m->nbody - 1; b > 0; b--) {
    int p = m->body_parentid[b];
    if (p > 0) for (int k = 0; k < 10; k++) d->crb[p][k] += d->crb[b][k];
  }
  memset(d->qM, 0, sizeof d->qM);
  for (int i = 0; i < nv; i++) {
    double buf[6];
    mul_inert_vec(buf, d->crb[m->dof_bodyid[i]], d->cdof[i]);
    for (int j = i; j >= 0; j = m->dof_parentid[j]) {
      double s = 0;
      for (int k = 0; k < 6; k++) s += d->cdof[j][k] * buf[k];
      d->qM[i][j] = d->qM[j][i] = s;
    }
    d->qM[i][i] += m->dof_armature[i];
  }
}

static void factor_m(const orc_model* m, orc_data* d) {
  memcpy(d->qL, d->qM, sizeof d->qL);
  if (orc_chol_factor(&d->qL[0][0], m->nv, ORC_NV) != 0) d->warn_bad = 1;
}

/* mj_jac: 3 x nv translational / rotational Jacobian of `point` attached to `body` (Appendix B.5) */
void orc_jac(const orc_model* m, const orc_data* d, double* jacp, double* jacr, const double point[3], int body) {
  int nv = m->nv, i;
  double off[3];
  if (jacp) memset(jacp, 0, 3 * nv * sizeof(double));
  if (jacr) memset(jacr, 0, 3 * nv * sizeof(double));
  sub3(off, point, d->subtree_com[m->body_rootid[body]]);
  while (body && !m->body_dofnum[body]) body = m->body_parentid[body];
  if (!body) return;
  i = m->body_dofadr[body] + m->body_dofnum[body] - 1;
  while (i >= 0) {
    const double* c = d->cdof[i];
    if (jacr) { jacr[i] = c[0]; jacr[nv + i] = c[1]; jacr[2 * nv + i] = c[2]; }
    if (jacp) {
      double t[3];
      cross3(t, c, off);
      jacp[i] = c[3] + t[0]; jacp[nv + i] = c[4] + t[1]; jacp[2 * nv + i] = c[5] + t[2];
    }
    i = m->dof_parentid[i];
  }
}

void orc_apply_ft(const orc_model* m, const orc_data* d, const double force[3], const double torque[3],
                  const double point[3], int body, double* qfrc) {
  double jp[3 * ORC_NV], jr[3 * ORC_NV];
  int nv = m->nv;
  orc_jac(m, d, jp, jr, point, body);
  for (int i = 0; i < nv; i++)
    qfrc[i] += jp[i] * force[0] + jp[nv + i] * force[1] + jp[2 * nv + i] * force[2] +
               jr[i] * torque[0] + jr[nv + i] * torque[1] + jr[2 * nv + i] * torque[2];
}

/* ------------------------------------------------------------------------------------------ constraints */

static int add_row(const orc_model* m, orc_data* d, const double* J, double pos, double margin, double frictionloss,
                   int type, int id) {
  int r = d->nefc;
  if (r >= m->njmax) { d->warn_cnstrfull = 1; return -1; }
  memcpy(d->efc_J[r], J, m->nv * sizeof(double));
  d->efc_pos[r] = pos; d->efc_margin[r] = margin; d->efc_frictionloss[r] = frictionloss;
  d->efc_type[r] = type; d->efc_id[r] = id;
  d->nefc++;
  return r;
}

static void get_impedance(const double* solimp, double pos, double margin, double* imp) {
  double x, y;
  if (solimp[0] == solimp[1] || solimp[2] <= ORC_MINVAL) { *imp = 0.5 * (solimp[0] + solimp[1]); return; }
  x = fabs((pos - margin) / solimp[2]);
  if (x >= 1 || x <= 0) { *imp = x >= 1 ? solimp[1] : solimp[0]; return; }
  if (solimp[4] == 1) y = x;
  else if (x <= solimp[3]) y = pow(x, solimp[4]) / pow(solimp[3], solimp[4] - 1);
  else y = 1 - pow(1 - x, solimp[4]) / pow(1 - solimp[3], solimp[4] - 1);
  *imp = solimp[0] + y * (solimp[1] - solimp[0]);
}

static void make_constraint(const orc_model* m, orc_data* d) {
  int nv = m->nv;
  double J[2][ORC_NV];
  d->nefc = 0;
  /* (1) dof friction loss */
  for (int i = 0; i < nv; i++)
    if (m->dof_frictionloss[i] > 0) {
      memset(J[0], 0, sizeof J[0]);
      J[0][i] = 1;
      add_row(m, d, J[0], 0, 0, m->dof_frictionloss[i], EFC_FRICTION, i);
    }
  /* (2) joint limits */
  for (int j = 0; j < m->njnt; j++) {
    if (!m->jnt_limited[j] || m->jnt_type[j] == JNT_FREE) continue;
    double value = d->qpos[m->jnt_qposadr[j]], margin = m->jnt_margin[j];
    for (int side = -1; side <= 1; side += 2) {
      double dist = side * (m->jnt_range[j][(side + 1) / 2] - value);
      if (dist < margin) {
        memset(J[0], 0, sizeof J[0]);
        J[0][m->jnt_dofadr[j]] = -side;
        add_row(m, d, J[0], dist, margin, 0, EFC_LIMIT, j);
      }
    }
  }
  /* (3) contacts, pyramidal */
  for (int c = 0; c < d->ncon; c++) {
    orc_contact* con = &d->contact[c];
    double jp1[3 * ORC_NV], jr1[3 * ORC_NV], jp2[3 * ORC_NV], jr2[3 * ORC_NV], jac[6][ORC_NV];
    int b1 = m->geom_bodyid[con->geom1], b2 = m->geom_bodyid[con->geom2];
    con->efc_address = -1;
    if (con->exclude) continue;
    orc_jac(m, d, jp1, jr1, con->pos, b1);
    orc_jac(m, d, jp2, jr2, con->pos, b2);
    for (int i = 0; i < nv; i++) {
      double dp[3] = {jp2[i] - jp1[i], jp2[nv + i] - jp1[nv + i], jp2[2 * nv + i] - jp1[2 * nv + i]};
      double dr[3] = {jr2[i] - jr1[i], jr2[nv + i] - jr1[nv + i], jr2[2 * nv + i] - jr1[2 * nv + i]};
      for (int k = 0; k < 3; k++) {
        jac[k][i] = dot3(con->frame + 3 * k, dp);
        jac[3 + k][i] = dot3(con->frame + 3 * k, dr);
      }
    }
    con->efc_address = d->nefc;
    if (con->dim == 1) {
      add_row(m, d, jac[0], con->dist, con->includemargin, 0, EFC_CONTACT_FRICTIONLESS, c);
    } else {
      for (int k = 1; k < con->dim; k++) {
        for (int i = 0; i < nv; i++) {
          J[0][i] = jac[0][i] + con->friction[k - 1] * jac[k][i];
          J[1][i] = jac[0][i] - con->friction[k - 1] * jac[k][i];
        }
        add_row(m, d, J[0], con->dist, con->includemargin, 0, EFC_CONTACT_PYRAMIDAL, c);
        add_row(m, d, J[1], con->dist, con->includemargin, 0, EFC_CONTACT_PYRAMIDAL, c);
      }
    }
  }
  /* diagApprox (Appendix B.8) */
  for (int r = 0; r < d->nefc; r++) {
    int id = d->efc_id[r];
    switch (d->efc_type[r]) {
      case EFC_FRICTION: d->efc_diagApprox[r] = m->dof_invweight0[id]; break;
      case EFC_LIMIT: d->efc_diagApprox[r] = m->dof_invweight0[m->jnt_dofadr[id]]; break;
      default: {
        const orc_contact* con = &d->contact[id];
        int b1 = m->geom_bodyid[con->geom1], b2 = m->geom_bodyid[con->geom2];
        double tran = m->body_invweight0[b1][0] + m->body_invweight0[b2][0];
        double rot = m->body_invweight0[b1][1] + m->body_invweight0[b2][1];
        if (d->efc_type[r] == EFC_CONTACT_FRICTIONLESS) d->efc_diagApprox[r] = tran;
        else {
          int nr = 2 * (con->dim - 1);
          for (int j = 0; j < nr; j++) {
            double f = con->friction[j / 2];
            d->efc_diagApprox[r + j] = tran + f * f * (j / 2 < 2 ? tran : rot);
          }
          r += nr - 1;
        }
      }
    }
  }
  /* impedance, K, B, R, D */
  for (int r = 0; r < d->nefc; r++) {
    double solref[2], solimp[5], imp, K, B;
    int id = d->efc_id[r];
    switch (d->efc_type[r]) {
      case EFC_FRICTION: memcpy(solref, m->dof_solref[id], sizeof solref); memcpy(solimp, m->dof_solimp[id], sizeof solimp); break;
      case EFC_LIMIT: memcpy(solref, m->jnt_solref[id], sizeof solref); memcpy(solimp, m->jnt_solimp[id], sizeof solimp); break;
      default: memcpy(solref, d->contact[id].solref, sizeof solref); memcpy(solimp, d->contact[id].solimp, sizeof solimp);
    }
    solimp[0] = fmin(ORC_MAXIMP, fmax(ORC_MINIMP, solimp[0]));
    solimp[1] = fmin(ORC_MAXIMP, fmax(ORC_MINIMP, solimp[1]));
    solimp[2] = fmax(0, solimp[2]);
    solimp[3] = fmin(ORC_MAXIMP, fmax(ORC_MINIMP, solimp[3]));
    solimp[4] = fmax(1, solimp[4]);
    get_impedance(solimp, d->efc_pos[r], d->efc_margin[r], &imp);
    d->efc_R[r] = fmax(ORC_MINVAL, (1 - imp) * d->efc_diagApprox[r] / imp);
    if (solref[0] > 0) {
      double tc = fmax(solref[0], 2 * m->timestep), dr = solref[1];
      B = 2 / (solimp[1] * tc);
      K = 1 / (solimp[1] * solimp[1] * tc * tc * dr * dr);
    } else {
      K = -solref[0] / (solimp[1] * solimp[1]);
      B = -solref[1] / solimp[1];
    }
    if (d->efc_type[r] == EFC_FRICTION) K = 0;
    d->efc_KBIP[r][0] = K; d->efc_KBIP[r][1] = B; d->efc_KBIP[r][2] = imp; d->efc_KBIP[r][3] = 0;
  }
  /* pyramidal contacts: all edges share R = 2 mu^2 R_first, mu = friction[0]*sqrt(1/impratio) */
  for (int r = 0; r < d->nefc; r++)
    if (d->efc_type[r] == EFC_CONTACT_PYRAMIDAL) {
      orc_contact* con = &d->contact[d->efc_id[r]];
      int nr = 2 * (con->dim - 1);
      double R1 = d->efc_R[r] / fmax(ORC_MINVAL, m->impratio);
      con->mu = con->friction[0] * sqrt(R1 / d->efc_R[r]);
      double Rpy = 2 * con->mu * con->mu * d->efc_R[r];
      for (int j = 0; j < nr; j++) d->efc_R[r + j] = Rpy;
      r += nr - 1;
    }
  for (int r = 0; r < d->nefc; r++) d->efc_D[r] = 1.0 / d->efc_R[r];
}

/* ------------------------------------------------------------------------------------------ velocity stage */

static void com_vel(const orc_model* m, orc_data* d) {
  memset(d->cvel[0], 0, sizeof d->cvel[0]);
  for (int b = 1; b < m->nbody; b++) {
    double cvel[6];
    int da = m->body_dofadr[b], dn = m->body_dofnum[b];
    memcpy(cvel, d->cvel[m->body_parentid[b]], sizeof cvel);
    for (int j = 0; j < dn; j++) {
      int jt = m->jnt_type[m->dof_jntid[da + j]];
      if (jt == JNT_FREE) {
        /* translations: cdof_dot = 0, then 3 rotations using the velocity after translations */
        memset(d->cdof_dot[da + j], 0, 18 * sizeof(double));
        for (int k = 0; k < 3; k++)
          for (int c = 0; c < 6; c++) cvel[c] += d->cdof[da + j + k][c] * d->qvel[da + j + k];
        j += 3;
        for (int k = 0; k < 3; k++) cross_motion(d->cdof_dot[da + j + k], cvel, d->cdof[da + j + k]);
        for (int k = 0; k < 3; k++)
          for (int c = 0; c < 6; c++) cvel[c] += d->cdof[da + j + k][c] * d->qvel[da + j + k];
        j += 2;
      } else {
        cross_motion(d->cdof_dot[da + j], cvel, d->cdof[da + j]);
        for (int c = 0; c < 6; c++) cvel[c] += d->cdof[da + j][c] * d->qvel[da + j];
      }
    }
    memcpy(d->cvel[b], cvel, sizeof cvel);
  }
}

static void rne_bias(const orc_model* m, orc_data* d) {
  double cacc[ORC_NB][6], cfrc[ORC_NB][6];
  memset(cacc[0], 0, sizeof cacc[0]);
  for (int k = 0; k < 3; k++) cacc[0][3 + k] = -m->gravity[k];
  memset(cfrc[0], 0, sizeof cfrc[0]);
  for (int b = 1; b < m->nbody; b++) {
    int da = m->body_dofadr[b], dn = m->body_dofnum[b];
    double t1[6], t2[6], t3[6];
    memcpy(cacc[b], cacc[m->body_parentid[b]], sizeof cacc[b]);
    for (int j = 0; j < dn; j++)
      for (int c = 0; c < 6; c++) cacc[b][c] += d->cdof_dot[da + j][c] * d->qvel[da + j];
    mul_inert_vec(t1, d->cinert[b], cacc[b]);
    mul_inert_vec(t2, d->cinert[b], d->cvel[b]);
    cross_force(t3, d->cvel[b], t2);
    for (int c = 0; c < 6; c++) cfrc[b][c] = t1[c] + t3[c];
  }
  for (int b = m->nbody - 1; b > 0; b--) {
    int p = m->body_parentid[b];
    if (p) for (int c = 0; c < 6; c++) cfrc[p][c] += cfrc[b][c];
  }
  for (int i = 0; i < m->nv; i++) {
    double s = 0;
    for (int c = 0; c < 6; c++) s += d->cdof[i][c] * cfrc[m->dof_bodyid[i]][c];
    d->qfrc_bias[i] = s;
  }
}

void orc_fwd_position(const orc_model* m, orc_data* d) {
  kinematics(m, d);
  com_pos(m, d);
  crb(m, d);
  factor_m(m, d);
  orc_collision(m, d);
  make_constraint(m, d);
}

void orc_fwd_velocity(const orc_model* m, orc_data* d) {
  int nv = m->nv;
  for (int r = 0; r < d->nefc; r++) {
    double s = 0;
    for (int i = 0; i < nv; i++) s += d->efc_J[r][i] * d->qvel[i];
    d->efc_vel[r] = s;
  }
  com_vel(m, d);
  for (int i = 0; i < nv; i++) d->qfrc_passive[i] = -m->dof_damping[i] * d->qvel[i];
  for (int r = 0; r < d->nefc; r++)
    d->efc_aref[r] = -d->efc_KBIP[r][1] * d->efc_vel[r] - d->efc_KBIP[r][0] * d->efc_KBIP[r][2] * (d->efc_pos[r] - d->efc_margin[r]);
  rne_bias(m, d);
}

void orc_fwd_actuation(const orc_model* m, orc_data* d) {
  memset(d->qfrc_actuator, 0, sizeof d->qfrc_actuator);
  for (int u = 0; u < m->nu; u++) d->qfrc_actuator[m->actuator_dofid[u]] += d->ctrl[u];
}

void orc_fwd_acceleration(const orc_model* m, orc_data* d) {
  for (int i = 0; i < m->nv; i++) {
    d->qfrc_smooth[i] = d->qfrc_passive[i] - d->qfrc_bias[i] + d->qfrc_applied[i] + d->qfrc_actuator[i];
    d->qacc_smooth[i] = d->qfrc_smooth[i];
  }
  orc_chol_solve(&d->qL[0][0], m->nv, ORC_NV, d->qacc_smooth);
}

/* ------------------------------------------------------------------------------------------ Newton solver */

typedef struct {
  double jar[ORC_NEFC], Jv[ORC_NEFC], Ma[ORC_NV], Mv[ORC_NV], grad[ORC_NV], search[ORC_NV];
  double cost, gauss;
} solver_ctx;

/* per-row force/state/cost from jar (Appendix B.9); returns the constraint part of the cost */
static double constraint_update(const orc_model* m, orc_data* d, const double* jar, int write) {
  double cost = 0;
  (void)m;
  for (int r = 0; r < d->nefc; r++) {
    double x = jar[r], D = d->efc_D[r], f, c;
    int st;
    if (d->efc_type[r] == EFC_FRICTION) {
      double fl = d->efc_frictionloss[r], Rf = d->efc_R[r] * fl;
      if (x <= -Rf) { st = ST_LINEARNEG; f = fl; c = fl * (-0.5 * Rf - x); }
      else if (x >= Rf) { st = ST_LINEARPOS; f = -fl; c = fl * (-0.5 * Rf + x); }
      else { st = ST_QUADRATIC; f = -D * x; c = 0.5 * D * x * x; }
    } else {
      if (x < 0) { st = ST_QUADRATIC; f = -D * x; c = 0.5 * D * x * x; }
      else { st = ST_SATISFIED; f = 0; c = 0; }
    }
    cost += c;
    if (write) { d->efc_force[r] = f; d->efc_state[r] = st; }
  }
  return cost;
}

static void mul_M(const orc_model* m, const orc_data* d, double* res, const double* v) {
  for (int i = 0; i < m->nv; i++) {
    double s = 0;
    for (int j = 0; j < m->nv; j++) s += d->qM[i][j] * v[j];
    res[i] = s;
  }
}
static void mul_J(const orc_model* m, const orc_data* d, double* res, const double* v) {
  for (int r = 0; r < d->nefc; r++) {
    double s = 0;
    for (int j = 0; j < m->nv; j++) s += d->efc_J[r][j] * v[j];
    res[r] = s;
  }
}

static double total_cost(const orc_model* m, orc_data* d, solver_ctx* c, int write) {
  double g = 0;
  for (int i = 0; i < m->nv; i++) g += 0.5 * (c->Ma[i] - d->qfrc_smooth[i]) * (d->qacc[i] - d->qacc_smooth[i]);
  c->gauss = g;
  return g + constraint_update(m, d, c->jar, write);
}

static int newton_direction(const orc_model* m, orc_data* d, solver_ctx* c) {
  int nv = m->nv;
  double H[ORC_NV][ORC_NV];
  for (int i = 0; i < nv; i++) {
    double s = c->Ma[i] - d->qfrc_smooth[i];
    for (int r = 0; r < d->nefc; r++) s -= d->efc_J[r][i] * d->efc_force[r];
    c->grad[i] = s;
  }
  memcpy(H, d->qM, sizeof H);
  for (int r = 0; r < d->nefc; r++) {
    if (d->efc_state[r] != ST_QUADRATIC) continue;
    const double* J = d->efc_J[r];
    double D = d->efc_D[r];
    for (int i = 0; i < nv; i++) {
      if (J[i] == 0) continue;
      double a = D * J[i];
      for (int j = 0; j <= i; j++) H[i][j] += a * J[j];
    }
  }
  if (orc_chol_factor(&H[0][0], nv, ORC_NV) != 0) return -1;
  for (int i = 0; i < nv; i++) c->search[i] = -c->grad[i];
  orc_chol_solve(&H[0][0], nv, ORC_NV, c->search);
  return 0;
}

typedef struct { double alpha, dslope; } brk;
static int brk_cmp(const void* a, const void* b) {
  double x = ((const brk*)a)->alpha, y = ((const brk*)b)->alpha;
  return (x > y) - (x < y);
}

/* exact minimiser of the convex piecewise-quadratic cost along `search`.  MuJoCo's line search is a
 * bracketed Newton iteration to ls_tolerance; both locate the same unique minimiser, and the outer Newton
 * iteration converges to the same optimum, so the exact form is used here (documented in DESIGN.md). */
static double line_search(const orc_model* m, orc_data* d, solver_ctx* c) {
  int nv = m->nv, nb = 0;
  static __thread brk bk[2 * ORC_NEFC];
  double g0, slope, qg1 = 0, qg2 = 0, alpha = 0;
  mul_M(m, d, c->Mv, c->search);
  mul_J(m, d, c->Jv, c->search);
  for (int i = 0; i < nv; i++) {
    qg1 += c->search[i] * (c->Ma[i] - d->qfrc_smooth[i]);
    qg2 += 0.5 * c->search[i] * c->Mv[i];
  }
  g0 = qg1;
  slope = 2 * qg2;
  for (int r = 0; r < d->nefc; r++) {
    double x = c->jar[r], v = c->Jv[r], D = d->efc_D[r];
    if (d->efc_type[r] == EFC_FRICTION) {
      double fl = d->efc_frictionloss[r], Rf = d->efc_R[r] * fl;
      /* derivative wrt x: -fl (x<=-Rf), D x (|x|<Rf), +fl (x>=Rf) */
      if (x <= -Rf) g0 += -fl * v;
      else if (x >= Rf) g0 += fl * v;
      else { g0 += D * x * v; slope += D * v * v; }
      if (v != 0) {
        double a1 = (-Rf - x) / v, a2 = (Rf - x) / v; /* crossing -Rf and +Rf */
        /* moving with v>0: crossing -Rf enters quadratic (+), crossing +Rf leaves (-); v<0 mirrored */
        if (a1 > 0) { bk[nb].alpha = a1; bk[nb].dslope = (v > 0 ? 1 : -1) * D * v * v; nb++; }
        if (a2 > 0) { bk[nb].alpha = a2; bk[nb].dslope = (v > 0 ? -1 : 1) * D * v * v; nb++; }
      }
    } else {
      if (x < 0) { g0 += D * x * v; slope += D * v * v; }
      if (v != 0) {
        double a = -x / v;
        if (a > 0) { bk[nb].alpha = a; bk[nb].dslope = (v > 0 ? -1 : 1) * D * v * v; nb++; }
      }
    }
  }
  if (!(g0 < 0)) return 0;
  qsort(bk, nb, sizeof(brk), brk_cmp);
  {
    double g = g0, a_prev = 0;
    int k = 0;
    for (;;) {
      double a_next = k < nb ? bk[k].alpha : INFINITY;
      double g_next = g + slope * (a_next - a_prev);
      if (k >= nb || g_next >= 0) { alpha = slope > 0 ? a_prev - g / slope : a_prev; break; }
      g = g_next; a_prev = a_next; slope += bk[k].dslope; k++;
    }
  }
  d->ls_evals += nb;
  return alpha;
}

void orc_fwd_constraint(const orc_model* m, orc_data* d) {
  int nv = m->nv;
  static __thread solver_ctx ctx;
  solver_ctx* c = &ctx;
  double scale = 1.0 / (m->meaninertia * (nv > 1 ? nv : 1));
  d->solver_iter = 0;
  if (d->nefc == 0) {
    memcpy(d->qacc, d->qacc_smooth, nv * sizeof(double));
    memset(d->qfrc_constraint, 0, nv * sizeof(double));
    return;
  }
  /* warm start: take qacc_warmstart if its cost is lower than qacc_smooth's */
  {
    double cw, cs;
    memcpy(d->qacc, d->qacc_warmstart, nv * sizeof(double));
    mul_M(m, d, c->Ma, d->qacc);
    mul_J(m, d, c->jar, d->qacc);
    for (int r = 0; r < d->nefc; r++) c->jar[r] -= d->efc_aref[r];
    cw = total_cost(m, d, c, 0);
    memcpy(d->qacc, d->qacc_smooth, nv * sizeof(double));
    mul_J(m, d, c->jar, d->qacc);
    for (int r = 0; r < d->nefc; r++) c->jar[r] -= d->efc_aref[r];
    cs = constraint_update(m, d, c->jar, 0);
    if (cw <= cs) memcpy(d->qacc, d->qacc_warmstart, nv * sizeof(double));
  }
  mul_M(m, d, c->Ma, d->qacc);
  mul_J(m, d, c->jar, d->qacc);
  for (int r = 0; r < d->nefc; r++) c->jar[r] -= d->efc_aref[r];
  c->cost = total_cost(m, d, c, 1);
  if (newton_direction(m, d, c) != 0) { d->warn_bad = 1; return; }
  while (d->solver_iter < m->iterations) {
    double alpha = line_search(m, d, c), oldcost = c->cost, gn = 0;
    if (alpha == 0) break;
    for (int i = 0; i < nv; i++) { d->qacc[i] += alpha * c->search[i]; c->Ma[i] += alpha * c->Mv[i]; }
    for (int r = 0; r < d->nefc; r++) c->jar[r] += alpha * c->Jv[r];
    c->cost = total_cost(m, d, c, 1);
    if (newton_direction(m, d, c) != 0) { d->warn_bad = 1; break; }
    for (int i = 0; i < nv; i++) gn += c->grad[i] * c->grad[i];
    d->solver_improvement = scale * (oldcost - c->cost);
    d->solver_gradient = scale * sqrt(gn);
    d->solver_iter++;
    if (d->solver_improvement < m->tolerance || d->solver_gradient < m->tolerance) break;
  }
  for (int i = 0; i < nv; i++) {
    double s = 0;
    for (int r = 0; r < d->nefc; r++) s += d->efc_J[r][i] * d->efc_force[r];
    d->qfrc_constraint[i] = s;
  }
}

/* ------------------------------------------------------------------------------------------ step */

static int bad_vec(const double* v, int n) {
  for (int i = 0; i < n; i++)
    if (isnan(v[i]) || v[i] > ORC_MAXVAL || v[i] < -ORC_MAXVAL) return 1;
  return 0;
}

void orc_forward(const orc_model* m, orc_data* d) {
  orc_fwd_position(m, d);
  orc_fwd_velocity(m, d);
  orc_fwd_actuation(m, d);
  orc_fwd_acceleration(m, d);
  orc_fwd_constraint(m, d);
}

static void euler(const orc_model* m, orc_data* d) {
  int nv = m->nv, damp = 0;
  double h = m->timestep, qacc[ORC_NV];
  for (int i = 0; i < nv; i++) damp |= m->dof_damping[i] > 0;
  if (damp) {
    double A[ORC_NV][ORC_NV];
    memcpy(A, d->qM, sizeof A);
    for (int i = 0; i < nv; i++) {
      A[i][i] += h * m->dof_damping[i];
      qacc[i] = d->qfrc_smooth[i] + d->qfrc_constraint[i];
    }
    if (orc_chol_factor(&A[0][0], nv, ORC_NV) != 0) { d->warn_bad = 1; return; }
    orc_chol_solve(&A[0][0], nv, ORC_NV, qacc);
  } else {
    memcpy(qacc, d->qacc, nv * sizeof(double));
  }
  for (int i = 0; i < nv; i++) d->qvel[i] += h * qacc[i];
  for (int j = 0; j < m->njnt; j++) {
    int qa = m->jnt_qposadr[j], da = m->jnt_dofadr[j];
    if (m->jnt_type[j] == JNT_FREE) {
      double w[3], qr[4], qn[4], ang;
      for (int k = 0; k < 3; k++) d->qpos[qa + k] += h * d->qvel[da + k];
      copy3(w, d->qvel + da + 3);
      ang = h * normalize3(w);
      axisangle2quat(qr, w, ang);
      normalize4(d->qpos + qa + 3);
      mulquat(qn, d->qpos + qa + 3, qr);
      memcpy(d->qpos + qa + 3, qn, sizeof qn);
    } else {
      d->qpos[qa] += h * d->qvel[da];
    }
  }
  d->time += h;
  memcpy(d->qacc_warmstart, d->qacc, nv * sizeof(double));
}

void orc_mj_step(const orc_model* m, orc_data* d) {
  if (bad_vec(d->qpos, m->nq) || bad_vec(d->qvel, m->nv)) d->warn_bad = 1;
  orc_forward(m, d);
  if (bad_vec(d->qacc, m->nv)) d->warn_bad = 1;
  if (d->warn_bad) return;
  euler(m, d);
}

void orc_reset_data(const orc_model* m, orc_data* d) {
  memset(d, 0, sizeof *d);
  memcpy(d->qpos, m->qpos0, m->nq * sizeof(double));
}

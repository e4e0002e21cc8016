/* orc_env.c -- CPU ORACLE (test infrastructure): float64 restatement of the HandObjMimic4 env logic.
 * Follows /root/reference/uhc/envs/ho_im4.py: step :611-662, do_simulation :503-565, compute_torque :412-486,
 * compute_desired_accel :393-410, rfc_obj :488-501, record_contact :883-889, classify_contact :567-597,
 * solve_rfc :941-1083, calc_ho_diff :664-688, get_full_obs_v5 :280-356, reset_model :690-716, expert accessors
 * :739-792; uhc/khrylib/rl/envs/common/mujoco_env.py:95-114 (reset / set_state); uhc/envs/ho_reward.py:943-1046
 * (ho_mimic_reward_9); uhc/utils/math_utils.py:103-131,210-236; uhc/utils/transforms.py:99-158,414-490.
 * The QP of solve_rfc (qpsolvers 'daqp', third-party) is solved by an exact active-set method.  PARITY UNPINNED.
 */
#include <stdio.h>
#include <stdlib.h>
#include "orc_math.h"
#include "orc_model.h"

orc_model* orc_model_create(const void* blob, size_t nbytes);
void orc_model_destroy(orc_model* m);

/* expert frame offsets */
#define F_DOF 0
#define F_DOFVEL 26
#define F_OBJPOSE 52
#define F_OBJVEL 59
#define F_OBJANGVEL 62
#define F_BPOS 65
#define F_BQUAT 128

orc_env* orc_env_create(const orc_model* m) {
  orc_env* e = (orc_env*)calloc(1, sizeof(orc_env));
  e->m = m;
  e->mode_train = 1;
  orc_reset_data(m, &e->d);
  return e;
}
void orc_env_destroy(orc_env* e) {
  if (e) { free(e->motion); free(e->rec_frames); free(e->rec_surface); free(e); }
}
/* test-mode recording on/off (the reference records when mode == 'test', ho_im4.py:547-548, 1072) */
void orc_env_set_record(orc_env* e, int on) {
  e->record = on;
  if (on && !e->rec_frames) {
    e->rec_frames = (double*)calloc((size_t)e->m->sim_step * ORC_REC_STRIDE, sizeof(double));
    e->rec_surface = (double*)calloc(ORC_SURF_STRIDE, sizeof(double));
  }
}
int orc_get_record(const orc_env* e, double* frames, double* surface) {
  if (!e->rec_frames) return 0;
  memcpy(frames, e->rec_frames, (size_t)e->m->sim_step * ORC_REC_STRIDE * sizeof(double));
  memcpy(surface, e->rec_surface, ORC_SURF_STRIDE * sizeof(double));
  return e->m->sim_step;
}
void orc_env_set_motion(orc_env* e, const double* frames, int T) {
  free(e->motion);
  e->motion = (double*)malloc((size_t)T * ORC_FRAME * sizeof(double));
  memcpy(e->motion, frames, (size_t)T * ORC_FRAME * sizeof(double));
  e->expert_len = T;
}
void orc_env_set_reward_weights(orc_env* e, const double* w) { memcpy(e->reward_w, w, sizeof e->reward_w); }
void orc_env_set_mode(orc_env* e, int train) { e->mode_train = train; }

static const double* expert(const orc_env* e, int delta) {
  int ind = e->cur_t + delta + e->start_ind;
  if (ind > e->expert_len - 1) ind = e->expert_len - 1;
  return e->motion + (size_t)ind * ORC_FRAME;
}

/* ------------------------------------------------------------------------------------------ observation */

static void put_rot2(double* o, const double* R) { /* R[:, 0:2].flatten() */
  o[0] = R[0]; o[1] = R[1]; o[2] = R[3]; o[3] = R[4]; o[4] = R[6]; o[5] = R[7];
}

void orc_env_obs(const orc_env* e, double* obs) {
  const orc_model* m = e->m;
  const orc_data* d = &e->d;
  const double *qpos = d->qpos, *qvel = d->qvel;
  const double* rp = d->xpos[m->palm_body];
  const double* rq = d->xquat[m->palm_body];
  double R[9], rqi[4], t[3], q[4], q2[4], M[9];
  int k = 0, W = m->w_size, nh = m->nq - 7, nhv = m->nv - 6, nb = m->n_hand_body;
  quat2mat(R, rq);
  negquat(rqi, rq);
  put_rot2(obs + k, R); k += 6;
  for (int i = 6; i < nh; i++) obs[k++] = qpos[i];
  for (int w = 0; w < W; w++) {
    const double* f = expert(e, 1 + w);
    for (int i = 6; i < nh; i++) obs[k++] = f[F_DOF + i] - qpos[i];
  }
  for (int i = 6; i < nhv; i++) obs[k++] = qvel[i];
  mulmatT3(obs + k, R, qvel); k += 3;
  mulmatT3(obs + k, R, qvel + 3); k += 3;
  for (int w = 0; w < W; w++) {
    const double* f = expert(e, 1 + w);
    sub3(t, f + F_BPOS, rp);
    mulmatT3(obs + k, R, t); k += 3;
    mulquat(q, rqi, f + F_BQUAT);
    quat2mat(M, q);
    put_rot2(obs + k, M); k += 6;
  }
  /* joint positions relative to qpos[:3], component-major (transform_vec_batch returns (3,N)) */
  for (int c = 0; c < 3; c++)
    for (int b = 1; b < nb; b++) {
      sub3(t, d->xpos[m->hand_body_idx[b]], qpos);
      obs[k++] = R[c] * t[0] + R[3 + c] * t[1] + R[6 + c] * t[2];
    }
  for (int w = 0; w < W; w++) {
    const double* f = expert(e, 1 + w);
    for (int c = 0; c < 3; c++)
      for (int b = 1; b < nb; b++) {
        sub3(t, f + F_BPOS + 3 * b, d->xpos[m->hand_body_idx[b]]);
        obs[k++] = R[c] * t[0] + R[3 + c] * t[1] + R[6 + c] * t[2];
      }
  }
  {
    const double* op = qpos + nh;
    double oqi[4];
    sub3(t, op, rp);
    mulmatT3(obs + k, R, t); k += 3;
    mulquat(q, rqi, op + 3);
    quat2mat(M, q);
    put_rot2(obs + k, M); k += 6;
    mulmatT3(obs + k, R, qvel + nhv); k += 3;
    mulmatT3(obs + k, R, qvel + nhv + 3); k += 3;
    negquat(oqi, op + 3);
    {
      double n2 = op[3] * op[3] + op[4] * op[4] + op[5] * op[5] + op[6] * op[6];
      for (int i = 0; i < 4; i++) oqi[i] /= n2; /* quaternion_inverse divides by |q|^2 */
    }
    for (int w = 0; w < W; w++) {
      const double* f = expert(e, 1 + w);
      sub3(t, f + F_OBJPOSE, op);
      mulmatT3(obs + k, R, t); k += 3;
      mulquat(q, f + F_OBJPOSE + 3, oqi);
      mulquat(q2, rqi, q);
      {
        /* transforms3d quat2mat normalises by |q|^2 */
        double n2 = q2[0] * q2[0] + q2[1] * q2[1] + q2[2] * q2[2] + q2[3] * q2[3], s = 1.0 / sqrt(n2);
        for (int i = 0; i < 4; i++) q2[i] *= s;
      }
      quat2mat(M, q2);
      put_rot2(obs + k, M); k += 6;
    }
  }
  if (k != ORC_OBS) { fprintf(stderr, "orc: obs size %d\n", k); abort(); }
}

/* ------------------------------------------------------------------------------------------ reset / set_state */

void orc_env_set_state(orc_env* e, const double* qpos, const double* qvel) {
  memcpy(e->d.qpos, qpos, e->m->nq * sizeof(double));
  memcpy(e->d.qvel, qvel, e->m->nv * sizeof(double));
  orc_forward(e->m, &e->d);
}

void orc_env_reset(orc_env* e, double* obs) {
  const orc_model* m = e->m;
  double qpos[ORC_NQ], qvel[ORC_NV];
  const double* f = e->motion;
  int nh = m->nq - 7;
  orc_reset_data(m, &e->d);
  e->cur_t = 0;
  e->start_ind = 0;
  memcpy(qpos, f + F_DOF, nh * sizeof(double));
  memcpy(qpos + nh, f + F_OBJPOSE, 7 * sizeof(double));
  memcpy(qvel, f + F_DOFVEL, nh * sizeof(double));
  memcpy(qvel + nh, f + F_OBJVEL, 3 * sizeof(double));
  memcpy(qvel + nh + 3, f + F_OBJANGVEL, 3 * sizeof(double));
  orc_env_set_state(e, qpos, qvel);
  if (obs) orc_env_obs(e, obs);
}

/* ------------------------------------------------------------------------------------------ control */

/* stable PD (ho_im4.py:393-486) with the STALE qM / qfrc_bias of the last forward */
void orc_env_compute_torque(const orc_env* e, const double* ctrl, double* torque) {
  const orc_model* m = e->m;
  const orc_data* d = &e->d;
  int n = m->nv - 6;
  double dt = m->timestep, target[ORC_NV], err[ORC_NV], A[ORC_NV][ORC_NV], rhs[ORC_NV];
  const double* ref = expert(e, 0) + F_DOF;
  for (int i = 0; i < 3; i++) target[i] = ref[i] + 0.1 * ctrl[i];
  for (int i = 3; i < 6; i++) target[i] = ref[i] + 0.3 * ctrl[i];
  for (int i = 6; i < n; i++) target[i] = ref[i] + m->ctrl_scale[i] * ctrl[i];
  for (int i = 0; i < n; i++) {
    err[i] = d->qpos[i] + d->qvel[i] * dt - target[i];
    if (i >= 3) {
      while (err[i] > ORC_PI) err[i] -= 2 * ORC_PI;
      while (err[i] < -ORC_PI) err[i] += 2 * ORC_PI;
    }
  }
  for (int i = 0; i < n; i++) {
    for (int j = 0; j < n; j++) A[i][j] = d->qM[i][j];
    A[i][i] += m->jkd[i] * dt;
    rhs[i] = -d->qfrc_bias[i] - m->jkp[i] * err[i] - m->jkd[i] * d->qvel[i];
  }
  orc_chol_factor(&A[0][0], n, ORC_NV);
  orc_chol_solve(&A[0][0], n, ORC_NV, rhs);
  for (int i = 0; i < n; i++) {
    double t = -m->jkp[i] * err[i] - m->jkd[i] * (d->qvel[i] + rhs[i] * dt);
    if (t > m->torque_lim[i]) t = m->torque_lim[i];
    if (t < -m->torque_lim[i]) t = -m->torque_lim[i];
    torque[i] = t;
  }
}

static void record_contact(orc_env* e) {
  const orc_model* m = e->m;
  const orc_data* d = &e->d;
  for (int c = 0; c < d->ncon; c++) {
    int g1 = d->contact[c].geom1, g2 = d->contact[c].geom2;
    if (g1 >= m->hand_geom0 && g1 <= m->hand_geom1 && g2 >= m->obj_geom0 && g2 <= m->obj_geom1) {
      int k = g1 - m->hand_geom0;
      for (int i = 0; i < 3; i++) e->contact_sum[k][i] += d->contact[c].pos[i];
      for (int i = 0; i < 9; i++) e->contact_sum[k][3 + i] += d->contact[c].frame[i];
      e->contact_count[k]++;
    }
  }
}

/* pytorch3d-style matrix -> axis-angle (uhc/utils/transforms.py:99-158, 414-490) */
static void matrix_to_axis_angle(const double* R, double* aa) {
  double m00 = R[0], m01 = R[1], m02 = R[2], m10 = R[3], m11 = R[4], m12 = R[5], m20 = R[6], m21 = R[7], m22 = R[8];
  double qa[4] = {1 + m00 + m11 + m22, 1 + m00 - m11 - m22, 1 - m00 + m11 - m22, 1 - m00 - m11 + m22};
  double cand[4][4] = {{0, m21 - m12, m02 - m20, m10 - m01}, {m21 - m12, 0, m10 + m01, m02 + m20},
                       {m02 - m20, m10 + m01, 0, m12 + m21}, {m10 - m01, m20 + m02, m21 + m12, 0}};
  double q[4], nrm, half, ang, s;
  int best = 0;
  for (int i = 0; i < 4; i++) { qa[i] = qa[i] > 0 ? sqrt(qa[i]) : 0; cand[i][i] = qa[i] * qa[i]; }
  for (int i = 1; i < 4; i++) if (qa[i] > qa[best]) best = i;
  for (int i = 0; i < 4; i++) q[i] = cand[best][i] / (2.0 * (qa[best] > 0.1 ? qa[best] : 0.1));
  nrm = sqrt(q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  half = atan2(nrm, q[0]);
  ang = 2 * half;
  s = fabs(ang) < 1e-6 ? 0.5 - ang * ang / 48 : sin(half) / ang;
  aa[0] = q[1] / s; aa[1] = q[2] / s; aa[2] = q[3] / s;
}

/* ------------------------------------------------------------------------------------------ RFC QP */

/* min 0.5 x'Qx + p'x  s.t. x >= 0, Q SPD: primal active-set (Lawson-Hanson style).  Returns iterations. */
static int qp_nonneg(int n, const double* Q, const double* p, double* x) {
  int* inP = (int*)calloc(n, sizeof(int));
  int* idx = (int*)malloc(n * sizeof(int));
  double* w = (double*)malloc(n * sizeof(double));
  double* s = (double*)malloc(n * sizeof(double));
  double* A = NULL;
  int iters = 0, np = 0;
  double pscale = 0;
  for (int i = 0; i < n; i++) { x[i] = 0; if (fabs(p[i]) > pscale) pscale = fabs(p[i]); }
  A = (double*)malloc((size_t)n * n * sizeof(double));
  for (int outer = 0; outer < 3 * n + 10; outer++) {
    int j = -1;
    double wmax = 1e-11 * (1 + pscale);
    for (int i = 0; i < n; i++) {
      double g = p[i];
      for (int k = 0; k < n; k++) if (x[k] != 0) g += Q[(size_t)i * n + k] * x[k];
      w[i] = -g;
      if (!inP[i] && w[i] > wmax) { wmax = w[i]; j = i; }
    }
    if (j < 0) break;
    inP[j] = 1;
    for (int inner = 0; inner < 3 * n + 10; inner++) {
      double alpha = 1;
      int blocking = -1;
      iters++;
      np = 0;
      for (int i = 0; i < n; i++) if (inP[i]) idx[np++] = i;
      for (int a = 0; a < np; a++) {
        for (int b = 0; b < np; b++) A[a * np + b] = Q[(size_t)idx[a] * n + idx[b]];
        s[a] = -p[idx[a]];
      }
      if (orc_chol_factor(A, np, np) != 0) { inP[j] = 0; break; }
      orc_chol_solve(A, np, np, s);
      for (int a = 0; a < np; a++)
        if (s[a] <= 0) {
          double xi = x[idx[a]], al = xi / (xi - s[a]);
          if (al < alpha) { alpha = al; blocking = a; }
        }
      if (blocking < 0) {
        for (int a = 0; a < np; a++) x[idx[a]] = s[a];
        break;
      }
      for (int a = 0; a < np; a++) x[idx[a]] += alpha * (s[a] - x[idx[a]]);
      for (int a = 0; a < np; a++)
        if (a == blocking || x[idx[a]] <= 0) { x[idx[a]] = 0; inP[idx[a]] = 0; }
    }
  }
  free(inP); free(idx); free(w); free(s); free(A);
  return iters;
}

static void solve_rfc(orc_env* e) {
  const orc_model* m = e->m;
  const orc_data* d = &e->d;
  const double dx = 0.0025, mu = 0.75, w_t = 1e4;
  int nq = m->nq, nv = m->nv, lastg = m->ngeom - 1;
  const double *obj_p = d->qpos + nq - 7, *obj_quat = d->qpos + nq - 4;
  double R[9], I[9], tmp[9], F[3], T[3], Iw[3], c3[3];
  const double *obj_v = e->geom_avg_vel[lastg], *obj_w = e->geom_avg_ang_vel[lastg];
  const double *obj_a = e->obj_avg_acc, *obj_oa = e->obj_avg_acc + 3;
  double mass = m->body_mass[m->nbody - 1];
  const double* inb = m->body_inertia[m->nbody - 1];
  (void)nv;
  if (!m->explain_force) {
    e->rfc_score = norm3(e->obj_vf) + w_t * norm3(e->obj_vt);
    return;
  }
  {
    double qn[4] = {obj_quat[0], obj_quat[1], obj_quat[2], obj_quat[3]};
    double n2 = qn[0] * qn[0] + qn[1] * qn[1] + qn[2] * qn[2] + qn[3] * qn[3], s = 1 / sqrt(n2);
    for (int i = 0; i < 4; i++) qn[i] *= s;
    quat2mat(R, qn);
  }
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) tmp[3 * i + j] = R[3 * i + j] * inb[j];
  mulmat33T(I, tmp, R);
  for (int k = 0; k < 3; k++) F[k] = mass * (obj_a[k] + (k == 2 ? 9.8 : 0));
  mulmat3(Iw, I, obj_w);
  cross3(c3, obj_w, Iw);
  mulmat3(T, I, obj_oa);
  add3(T, T, c3);
  e->qp_n = 0; e->qp_iters = 0;
  if (e->n_avg_cp == 0) {
    copy3(e->rest_force, F); copy3(e->rest_torque, T);
    e->rfc_score = norm3(F) + w_t * norm3(T);
    return;
  }
  {
    int ec = m->surface_contact ? 5 : 1, nc = e->n_avg_cp * ec, n = 4 * nc;
    double* Jf = (double*)calloc((size_t)3 * n, sizeof(double));
    double* Jt = (double*)calloc((size_t)3 * n, sizeof(double));
    double* pl = (double*)calloc(n, sizeof(double));
    double* Q = (double*)malloc((size_t)n * n * sizeof(double));
    double* x = (double*)malloc(n * sizeof(double));
    double isq = 1.0 / sqrt(1 + mu * mu), rf[3], rt[3];
    for (int ci = 0; ci < e->n_avg_cp; ci++) {
      const double* cf = e->avg_cps[ci];
      const double *fr0 = cf + 3, *fr1 = cf + 6, *fr2 = cf + 9;
      int g1 = e->avg_cp_geom[ci];
      for (int j = 0; j < ec; j++) {
        int i = ci * ec + j;
        double pos[3], crh[3], cro[3], cvh[3], cvo[3], relv[3], vn[3], vt[3], t3[3], vnd, A[4][3], r[3], dirs[4], ts;
        copy3(pos, cf);
        if (j == 1) addscl3(pos, pos, fr1, dx);
        else if (j == 2) addscl3(pos, pos, fr1, -dx);
        else if (j == 3) addscl3(pos, pos, fr2, dx);
        else if (j == 4) addscl3(pos, pos, fr2, -dx);
        sub3(crh, pos, d->geom_xpos[g1]);
        sub3(cro, pos, obj_p);
        cross3(t3, e->geom_avg_ang_vel[g1], crh); add3(cvh, e->geom_avg_vel[g1], t3);
        cross3(t3, obj_w, cro); add3(cvo, obj_v, t3);
        sub3(relv, cvo, cvh);
        vnd = dot3(fr0, relv);
        scl3(vn, fr0, vnd);
        sub3(vt, relv, vn);
        /* cp_ts[i // 5] quirk (ho_im4.py:1012): only meaningful with surface_contact */
        ts = e->cp_ts[(i / 5) < e->n_avg_cp ? (i / 5) : e->n_avg_cp - 1] / m->sim_step;
        for (int k = 0; k < 3; k++) {
          A[0][k] = (fr0[k] + mu * fr1[k]) * isq; A[1][k] = (fr0[k] - mu * fr1[k]) * isq;
          A[2][k] = (fr0[k] + mu * fr2[k]) * isq; A[3][k] = (fr0[k] - mu * fr2[k]) * isq;
        }
        sub3(r, pos, obj_p);
        for (int a = 0; a < 4; a++) {
          double rxA[3];
          cross3(rxA, r, A[a]);
          for (int k = 0; k < 3; k++) { Jf[k * n + 4 * i + a] = ts * A[a][k]; Jt[k * n + 4 * i + a] = ts * rxA[k]; }
        }
        if (dot3(vn, fr0) <= 0) { double nn = norm3(vn); for (int a = 0; a < 4; a++) pl[4 * i + a] += nn; }
        dirs[0] = -dot3(vt, fr1); dirs[1] = dot3(vt, fr1); dirs[2] = -dot3(vt, fr2); dirs[3] = dot3(vt, fr2);
        {
          int mx = 0; double nt = norm3(vt);
          for (int a = 1; a < 4; a++) if (dirs[a] > dirs[mx]) mx = a;
          for (int a = 0; a < 4; a++) if (a != mx) pl[4 * i + a] += nt;
        }
      }
    }
    for (int a = 0; a < n; a++) {
      for (int b = 0; b < n; b++) {
        double s = 0;
        for (int k = 0; k < 3; k++) s += Jf[k * n + a] * Jf[k * n + b] + w_t * Jt[k * n + a] * Jt[k * n + b];
        Q[(size_t)a * n + b] = 2 * s + (a == b ? 1e-7 : 0);
      }
      {
        double s = 0;
        for (int k = 0; k < 3; k++) s += -2 * Jf[k * n + a] * F[k] - 2 * w_t * Jt[k * n + a] * T[k];
        pl[a] += s;
      }
    }
    e->qp_n = n;
    e->qp_iters = qp_nonneg(n, Q, pl, x);
    for (int k = 0; k < 3; k++) {
      double sf = 0, st = 0;
      for (int a = 0; a < n; a++) { sf += Jf[k * n + a] * x[a]; st += Jt[k * n + a] * x[a]; }
      rf[k] = F[k] - sf; rt[k] = T[k] - st;
    }
    copy3(e->rest_force, rf); copy3(e->rest_torque, rt);
    e->rfc_score = norm3(rf) + sqrt(w_t) * norm3(rt);
    if (e->record) {
      /* surface_contact_frames / surface_contact_force / rest_obj_ft (ho_im4.py:1072-1081): f_i = (ts A_i) x_i */
      double* o = e->rec_surface;
      o[0] = nc; memcpy(o + 1, rf, sizeof rf); memcpy(o + 4, rt, sizeof rt); o[7] = 1;
      for (int ci = 0; ci < e->n_avg_cp; ci++)
        for (int j = 0; j < ec; j++) {
          int i = ci * ec + j;
          const double* cf = e->avg_cps[ci];
          double* q = o + ORC_SURF_PT + 15 * i;
          if (i >= ORC_SURF_MAXPT) break;
          memcpy(q, cf, 12 * sizeof(double));
          if (j == 1) addscl3(q, q, cf + 6, dx); else if (j == 2) addscl3(q, q, cf + 6, -dx);
          else if (j == 3) addscl3(q, q, cf + 9, dx); else if (j == 4) addscl3(q, q, cf + 9, -dx);
          for (int k = 0; k < 3; k++) { q[12 + k] = 0; for (int a = 0; a < 4; a++) q[12 + k] += Jf[k * n + 4 * i + a] * x[4 * i + a]; }
        }
    }
    free(Jf); free(Jt); free(pl); free(Q); free(x);
  }
}

/* ------------------------------------------------------------------------------------------ step */

/* [MJ] mj_contactForce: force / torque of contact c in its own frame from the pyramid edge forces (mju_decodePyramid) */
void orc_contact_force(const orc_data* d, int c, double* out6) {
  const orc_contact* con = &d->contact[c];
  memset(out6, 0, 6 * sizeof(double));
  if (con->efc_address < 0) return;
  if (con->dim == 1) { out6[0] = d->efc_force[con->efc_address]; return; }
  for (int i = 0; i < 2 * (con->dim - 1); i++) out6[0] += d->efc_force[con->efc_address + i];
  for (int i = 0; i < con->dim - 1; i++)
    out6[i + 1] = (d->efc_force[con->efc_address + 2 * i] - d->efc_force[con->efc_address + 2 * i + 1]) * con->friction[i];
}

/* compute_data_frame (ho_im4.py:599-609) after substep `i`: qpos / qvel are the new state, body positions, contacts,
   contact forces and qacc belong to the forward that produced it. */
static void record_frame(orc_env* e, int i) {
  const orc_model* m = e->m;
  const orc_data* d = &e->d;
  double* o = e->rec_frames + (size_t)i * ORC_REC_STRIDE;
  int nh = m->nq - 7, n = 0;
  memset(o, 0, ORC_REC_STRIDE * sizeof(double));
  for (int b = 0; b < m->n_hand_body; b++) memcpy(o + 3 * b, d->xpos[m->hand_body_idx[b]], 3 * sizeof(double));
  memcpy(o + 63, d->qpos, nh * sizeof(double));
  memcpy(o + 89, e->obj_vf, 3 * sizeof(double)); memcpy(o + 92, e->obj_vt, 3 * sizeof(double));
  memcpy(o + 95, d->qpos + nh, 7 * sizeof(double));
  memcpy(o + 102, d->qvel + m->nv - 6, 6 * sizeof(double));
  memcpy(o + 108, d->qacc + m->nv - 6, 6 * sizeof(double));
  for (int c = 0; c < d->ncon && n < ORC_REC_MAXCON; c++) {
    int g1 = d->contact[c].geom1, g2 = d->contact[c].geom2;
    if (g1 >= m->hand_geom0 && g1 <= m->hand_geom1 && g2 >= m->obj_geom0 && g2 <= m->obj_geom1) {
      double* q = o + ORC_REC_CON + 19 * n++;
      memcpy(q, d->contact[c].pos, 3 * sizeof(double));
      memcpy(q + 3, d->contact[c].frame, 9 * sizeof(double));
      orc_contact_force(d, c, q + 12);
      q[18] = g1;
    }
  }
  o[114] = n; o[115] = 1;
}

static int warned(const orc_data* d) { return d->warn_bad || d->warn_contactfull || d->warn_cnstrfull; }

/* one substep exactly as the loop body of do_simulation (ho_im4.py:516-545); returns nonzero on failure */
int orc_env_substep(orc_env* e, const double* action) {
  const orc_model* m = e->m;
  orc_data* d = &e->d;
  double torque[ORC_NV], qfrc[ORC_NV], zero[3] = {0, 0, 0}, fg[3] = {0, 0, m->hand_mass * 9.8};
  orc_env_compute_torque(e, action, torque);
  memcpy(d->ctrl, torque, m->nu * sizeof(double));
  memset(qfrc, 0, sizeof qfrc);
  orc_apply_ft(m, d, fg, zero, d->geom_xpos[m->hand_geom0], m->palm_body, qfrc);
  if (m->residual_force) orc_apply_ft(m, d, e->obj_vf, e->obj_vt, d->qpos + m->nq - 7, m->obj_body, qfrc);
  memcpy(d->qfrc_applied, qfrc, m->nv * sizeof(double));
  record_contact(e);
  orc_mj_step(m, d);
  return warned(d);
}
int orc_env_ncontact_force(orc_env* e, double* out) {   /* test accessor: [ncon][6] */
  for (int c = 0; c < e->d.ncon; c++) orc_contact_force(&e->d, c, out + 6 * c);
  return e->d.ncon;
}

static void classify_contact(orc_env* e) {
  const orc_model* m = e->m;
  int nhg = m->hand_geom1 - m->hand_geom0 + 1;
  e->n_avg_cp = 0;
  for (int i = 0; i < nhg; i++) {
    double f[12], *n = f + 3, ex[3] = {1, 0, 0}, ey[3] = {0, 1, 0}, s;
    if (e->contact_count[i] == 0) continue;
    for (int k = 0; k < 12; k++) f[k] = e->contact_sum[i][k] / e->contact_count[i];
    s = norm3(n);
    scl3(n, n, 1.0 / s);
    if (fabs(n[0]) >= 1e-5) cross3(f + 6, n, ex); else cross3(f + 6, n, ey);
    s = norm3(f + 6); scl3(f + 6, f + 6, 1.0 / s);
    cross3(f + 9, n, f + 6);
    s = norm3(f + 9); scl3(f + 9, f + 9, 1.0 / s);
    memcpy(e->avg_cps[e->n_avg_cp], f, sizeof f);
    e->avg_cp_geom[e->n_avg_cp] = i + m->hand_geom0;
    e->cp_ts[e->n_avg_cp] = e->contact_count[i];
    e->n_avg_cp++;
  }
}

void orc_env_ho_diff(const orc_env* e, double* out5) {
  const orc_model* m = e->m;
  const orc_data* d = &e->d;
  const double* f = expert(e, 0);
  const double* rp = d->xpos[m->palm_body];
  const double* rq = d->xquat[m->palm_body];
  double t[3], qi[4], qd[4], s = 0, n;
  sub3(t, rp, f + F_BPOS);
  out5[0] = norm3(t);
  for (int b = 0; b < m->n_hand_body; b++) {
    sub3(t, d->xpos[m->hand_body_idx[b]], f + F_BPOS + 3 * b);
    s += norm3(t);
  }
  out5[2] = s / m->n_hand_body;
  negquat(qi, rq);
  mulquat(qd, f + F_BQUAT, qi);
  n = norm3(qd + 1);
  out5[1] = 2.0 * asin(n > 1 ? 1 : n);
  sub3(t, d->qpos + m->nq - 7, f + F_OBJPOSE);
  out5[3] = norm3(t);
  out5[4] = 0.0; /* ho_im4.py:685 compares the expert pose with itself */
}

void orc_env_reward(const orc_env* e, double* reward, double* info9) {
  const orc_model* m = e->m;
  const orc_data* d = &e->d;
  const double* w = e->reward_w;
  double w_p = w[0], w_wp = w[1], w_v = w[2], w_j = w[3], w_op = w[4], w_or = w[5], w_ov = w[6], w_orfc = w[7];
  double k_p = w[8], k_wp = w[9], k_v = w[10], k_j = w[11], k_op = w[12], k_or = w[13], k_ov = w[14], k_orfc = w[15];
  const double* f = expert(e, 0);
  int nh = m->nq - 7, nb = m->n_hand_body;
  double s, t[3], pose_r, wpose_r, vel_r, jpos_r, opos_r, orot_r, ovel_r, orfc_r, hand, obj;
  s = 0;
  for (int i = 6; i < nh; i++) s += fabs(d->qpos[i] - f[F_DOF + i]);
  pose_r = exp(-k_p * s / (nh - 6));
  s = 0;
  for (int b = 0; b < nb; b++) {
    double qi[4], qd[4], dd[4];
    const double* eq = f + F_BQUAT + 4 * b;
    double n2 = eq[0] * eq[0] + eq[1] * eq[1] + eq[2] * eq[2] + eq[3] * eq[3];
    negquat(qi, eq);
    for (int i = 0; i < 4; i++) qi[i] /= n2;
    mulquat(qd, d->xquat[m->hand_body_idx[b]], qi);
    dd[0] = fabs(qd[0]) - 1; dd[1] = qd[1]; dd[2] = qd[2]; dd[3] = qd[3];
    s += sqrt(dd[0] * dd[0] + dd[1] * dd[1] + dd[2] * dd[2] + dd[3] * dd[3]);
  }
  wpose_r = exp(-k_wp * s / nb);
  s = 0;
  for (int i = 0; i < nh; i++) s += fabs(d->qvel[i] - f[F_DOFVEL + i]);
  vel_r = exp(-k_v * s / nh);
  s = 0;
  for (int b = 0; b < nb; b++) { sub3(t, d->xpos[m->hand_body_idx[b]], f + F_BPOS + 3 * b); s += norm3(t); }
  jpos_r = exp(-k_j * s / nb);
  sub3(t, d->qpos + nh, f + F_OBJPOSE);
  opos_r = exp(-k_op * norm3(t));
  {
    double qi[4], qd[4], dd[4];
    const double* eq = f + F_OBJPOSE + 3;
    double n2 = eq[0] * eq[0] + eq[1] * eq[1] + eq[2] * eq[2] + eq[3] * eq[3];
    negquat(qi, eq);
    for (int i = 0; i < 4; i++) qi[i] /= n2;
    mulquat(qd, d->qpos + nh + 3, qi);
    dd[0] = fabs(qd[0]) - 1; dd[1] = qd[1]; dd[2] = qd[2]; dd[3] = qd[3];
    orot_r = exp(-k_or * sqrt(dd[0] * dd[0] + dd[1] * dd[1] + dd[2] * dd[2] + dd[3] * dd[3]));
  }
  s = 0;
  for (int i = 0; i < 3; i++) s += fabs(d->qvel[nh + i] - f[F_OBJVEL + i]) + fabs(d->qvel[nh + 3 + i] - f[F_OBJANGVEL + i]);
  ovel_r = exp(-k_ov * s / 6);
  orfc_r = m->residual_force ? exp(-k_orfc * e->rfc_score) : 1.0;
  hand = (w_p * pose_r + w_wp * wpose_r + w_j * jpos_r + w_v * vel_r) / (w_p + w_wp + w_j + w_v);
  obj = (w_op * opos_r + w_or * orot_r + w_ov * ovel_r + w_orfc * orfc_r) / (w_op + w_or + w_ov + w_orfc);
  *reward = hand * obj;
  info9[0] = pose_r; info9[1] = wpose_r; info9[2] = jpos_r; info9[3] = vel_r; info9[4] = opos_r; info9[5] = orot_r;
  info9[6] = ovel_r; info9[7] = orfc_r; info9[8] = 1.0;
}

/* HandObjMimic4.step + ho_mimic_reward_9.  flags: [done, fail, end]; scal: [reward, percent, rfc_score] */
void orc_env_step(orc_env* e, const double* action_in, double* obs, double* scal, double* info9, int* flags) {
  const orc_model* m = e->m;
  orc_data* d = &e->d;
  double a[ORC_NV + 8], old_obj_vel[6], old_gpos[ORC_NG][3], old_gmat[ORC_NG][9], diff[5];
  int nact = m->nu + (m->residual_force ? 6 : 0), fail = 0, end, body_fail;
  double delta_t = m->sim_step * m->timestep;
  for (int i = 0; i < nact; i++) a[i] = action_in[i] > 1 ? 1 : (action_in[i] < -1 ? -1 : action_in[i]);
  e->rfc_score = 0;
  for (int k = 0; k < 3; k++) {
    e->obj_vf[k] = m->residual_force ? m->rf_scale * a[m->nu + k] : 0;
    e->obj_vt[k] = m->residual_force ? m->rt_scale * a[m->nu + 3 + k] : 0;
  }
  /* do_simulation */
  memcpy(old_obj_vel, d->qvel + m->nv - 6, sizeof old_obj_vel);
  memcpy(old_gpos, d->geom_xpos, sizeof old_gpos);
  memcpy(old_gmat, d->geom_xmat, sizeof old_gmat);
  memset(e->contact_sum, 0, sizeof e->contact_sum);
  memset(e->contact_count, 0, sizeof e->contact_count);
  if (e->record) {
    memset(e->rec_frames, 0, (size_t)m->sim_step * ORC_REC_STRIDE * sizeof(double));
    memset(e->rec_surface, 0, ORC_SURF_STRIDE * sizeof(double));
  }
  for (int i = 0; i < m->sim_step && !fail; i++) {
    fail = orc_env_substep(e, a);
    e->sub_iters[i & 31] = d->solver_iter; e->sub_ncon[i & 31] = d->ncon;
    if (!fail && e->record) record_frame(e, i);   /* a MujocoException leaves do_simulation before compute_data_frame */
  }
  if (!fail) {
    for (int k = 0; k < 6; k++) e->obj_avg_acc[k] = (d->qvel[m->nv - 6 + k] - old_obj_vel[k]) / delta_t;
    for (int g = 0; g < m->ngeom; g++) {
      double Rd[9];
      for (int k = 0; k < 3; k++) e->geom_avg_vel[g][k] = (d->geom_xpos[g][k] - old_gpos[g][k]) / delta_t;
      mulmat33T(Rd, d->geom_xmat[g], old_gmat[g]);
      matrix_to_axis_angle(Rd, e->geom_avg_ang_vel[g]);
      scl3(e->geom_avg_ang_vel[g], e->geom_avg_ang_vel[g], 1.0 / delta_t);
    }
    classify_contact(e);
    if (m->residual_force) solve_rfc(e);
  }
  e->cur_t += 1;
  orc_env_ho_diff(e, diff);
  body_fail = diff[0] > m->pos_thresh || diff[1] > m->rot_thresh || diff[2] > m->jpos_thresh ||
              diff[3] > m->obj_pos_thresh || diff[4] > m->obj_rot_thresh;
  if (e->mode_train) fail = fail || body_fail;
  end = (e->cur_t + e->start_ind >= e->expert_len - m->w_size - 1);
  flags[0] = fail || end; flags[1] = fail; flags[2] = end;
  scal[1] = (double)e->cur_t / (e->expert_len - 1);
  scal[2] = e->rfc_score;
  orc_env_obs(e, obs);
  orc_env_reward(e, &scal[0], info9);
}

/* ------------------------------------------------------------------------------------------ accessors for tests */
orc_data* orc_env_data(orc_env* e) { return &e->d; }
size_t orc_sizeof_data(void) { return sizeof(orc_data); }
int orc_env_cur_t(const orc_env* e) { return e->cur_t; }
void orc_env_set_cur_t(orc_env* e, int t) { e->cur_t = t; }
void orc_env_clear_warnings(orc_env* e) { e->d.warn_bad = e->d.warn_contactfull = e->d.warn_cnstrfull = 0; }
#define GETTER(name, expr, n) int orc_get_##name(orc_env* e, double* out) { memcpy(out, expr, (n) * sizeof(double)); return (int)(n); }
GETTER(qpos, e->d.qpos, e->m->nq)
GETTER(qvel, e->d.qvel, e->m->nv)
GETTER(qacc, e->d.qacc, e->m->nv)
GETTER(qacc_smooth, e->d.qacc_smooth, e->m->nv)
GETTER(qacc_warmstart, e->d.qacc_warmstart, e->m->nv)
GETTER(qfrc_bias, e->d.qfrc_bias, e->m->nv)
GETTER(qfrc_applied, e->d.qfrc_applied, e->m->nv)
GETTER(qfrc_constraint, e->d.qfrc_constraint, e->m->nv)
GETTER(ctrl, e->d.ctrl, e->m->nu)
GETTER(xpos, e->d.xpos, e->m->nbody * 3)
GETTER(xquat, e->d.xquat, e->m->nbody * 4)
GETTER(geom_xpos, e->d.geom_xpos, e->m->ngeom * 3)
GETTER(geom_xmat, e->d.geom_xmat, e->m->ngeom * 9)
GETTER(subtree_com, e->d.subtree_com, e->m->nbody * 3)
GETTER(cdof, e->d.cdof, e->m->nv * 6)
GETTER(qM, e->d.qM, ORC_NV * ORC_NV)
GETTER(efc_force, e->d.efc_force, e->d.nefc)
GETTER(efc_aref, e->d.efc_aref, e->d.nefc)
GETTER(efc_D, e->d.efc_D, e->d.nefc)
GETTER(efc_pos, e->d.efc_pos, e->d.nefc)
void orc_set_warmstart(orc_env* e, const double* w) { memcpy(e->d.qacc_warmstart, w, e->m->nv * sizeof(double)); }
void orc_set_ctrl_applied(orc_env* e, const double* ctrl, const double* qfrc) {
  memcpy(e->d.ctrl, ctrl, e->m->nu * sizeof(double));
  memcpy(e->d.qfrc_applied, qfrc, e->m->nv * sizeof(double));
}
void orc_env_mj_step(orc_env* e) { orc_mj_step(e->m, &e->d); }
void orc_env_forward(orc_env* e) { orc_forward(e->m, &e->d); }
int orc_get_efc_J(orc_env* e, double* out) {
  for (int r = 0; r < e->d.nefc; r++) memcpy(out + r * e->m->nv, e->d.efc_J[r], e->m->nv * sizeof(double));
  return e->d.nefc;
}
/* contacts: per contact 16 doubles [dist, pos3, frame9, geom1, geom2, dim]; returns ncon */
int orc_get_contacts(orc_env* e, double* out) {
  for (int c = 0; c < e->d.ncon; c++) {
    const orc_contact* k = &e->d.contact[c];
    double* o = out + 16 * c;
    o[0] = k->dist; memcpy(o + 1, k->pos, 3 * sizeof(double)); memcpy(o + 4, k->frame, 9 * sizeof(double));
    o[13] = k->geom1; o[14] = k->geom2; o[15] = k->dim;
  }
  return e->d.ncon;
}
int orc_get_substep_log(const orc_env* e, int* iters, int* ncon) {
  memcpy(iters, e->sub_iters, sizeof e->sub_iters); memcpy(ncon, e->sub_ncon, sizeof e->sub_ncon); return e->m->sim_step;
}
/* ints: [ncon, nefc, solver_iter, warn_bad, warn_contactfull, warn_cnstrfull, mpr_calls, mpr_iters, qp_n, qp_iters, n_avg_cp] */
void orc_get_counters(orc_env* e, int* out) {
  out[0] = e->d.ncon; out[1] = e->d.nefc; out[2] = e->d.solver_iter; out[3] = e->d.warn_bad;
  out[4] = e->d.warn_contactfull; out[5] = e->d.warn_cnstrfull; out[6] = e->d.mpr_calls; out[7] = e->d.mpr_iters;
  out[8] = e->qp_n; out[9] = e->qp_iters; out[10] = e->n_avg_cp;
}
void orc_get_rfc(orc_env* e, double* out) { /* rest_force3, rest_torque3, rfc_score */
  memcpy(out, e->rest_force, 3 * sizeof(double)); memcpy(out + 3, e->rest_torque, 3 * sizeof(double)); out[6] = e->rfc_score;
}

/* test hook: the QP solver on caller-supplied data */
int qp_nonneg_test(int n, const double* Q, const double* p, double* x) { return qp_nonneg(n, Q, p, x); }

/* ------------------------------------------------------------------------------------------ dataset preprocessing
   DatasetSingleDepth.preprocess_seq (uhc/data_loaders/dataset_singledepth.py:78-142) of one raw sequence: clip the hand
   dofs to the joint ranges (:100), finite-difference velocities (compute_vel_from_seq :152-185), FK of the hand bodies
   (compute_body_pos_quat_from_seq :222-237) and the per-frame hand-object contact pre-pass (compute_contact_info :187-220,
   `contact_info[g1] = contact.pos`: the last contact of a hand geom wins).  frames_out[T][212] in the expert frame layout,
   cmask[T][ORC_NHG] / cpos[T][ORC_NHG][3] indexed by hand geom slot. */
int orc_preprocess_seq(const orc_model* m, const double* hand_pose_seq, const double* obj_pose_seq, int T, double motion_freq,
                       int clip, double* frames_out, int* cmask, double* cpos) {
  int nh = m->nq - 7, nhg = m->hand_geom1 - m->hand_geom0 + 1;
  orc_data* d = (orc_data*)calloc(1, sizeof(orc_data));
  double* hand = (double*)malloc((size_t)T * nh * sizeof(double));
  const double M_PI_ = 3.14159265358979323846;
  if (T < 2) { free(d); free(hand); return -1; }
  for (int t = 0; t < T; t++)
    for (int i = 0; i < nh; i++) {
      double v = hand_pose_seq[t * nh + i], lo = m->jnt_range[i][0], hi = m->jnt_range[i][1];
      hand[t * nh + i] = !clip ? v : (v < lo ? lo : (v > hi ? hi : v));
    }
  for (int t = 0; t < T; t++) {
    double* f = frames_out + (size_t)t * ORC_FRAME;
    int a = t == 0 ? 1 : t, b = a - 1;   /* vel[0] = vel[1] */
    const double *oa = obj_pose_seq + 7 * a, *ob = obj_pose_seq + 7 * b;
    double Ra[9], Rb[9], Rd[9], qa[4], qb[4], na = 0, nb = 0;
    memcpy(f + F_DOF, hand + t * nh, nh * sizeof(double));
    memcpy(f + F_OBJPOSE, obj_pose_seq + 7 * t, 7 * sizeof(double));
    for (int i = 0; i < nh; i++) {
      double v = hand[a * nh + i] - hand[b * nh + i];
      if (i >= 3 && i < 6) { while (v > M_PI_) v -= 2 * M_PI_; while (v < -M_PI_) v += 2 * M_PI_; }
      f[F_DOFVEL + i] = v * motion_freq;
    }
    for (int k = 0; k < 3; k++) f[F_OBJVEL + k] = (oa[k] - ob[k]) * motion_freq;
    for (int k = 0; k < 4; k++) { na += oa[3 + k] * oa[3 + k]; nb += ob[3 + k] * ob[3 + k]; }
    for (int k = 0; k < 4; k++) { qa[k] = oa[3 + k] / sqrt(na); qb[k] = ob[3 + k] / sqrt(nb); }
    quat2mat(Ra, qa); quat2mat(Rb, qb);
    mulmat33T(Rd, Ra, Rb);
    matrix_to_axis_angle(Rd, f + F_OBJANGVEL);
    for (int k = 0; k < 3; k++) f[F_OBJANGVEL + k] *= motion_freq;
    /* forward of (hand dofs, object pose) of this frame: body poses + contacts */
    orc_reset_data(m, d);
    memcpy(d->qpos, hand + t * nh, nh * sizeof(double));
    memcpy(d->qpos + nh, obj_pose_seq + 7 * t, 7 * sizeof(double));
    orc_fwd_position(m, d);
    for (int bi = 0; bi < m->n_hand_body; bi++) {
      memcpy(f + F_BPOS + 3 * bi, d->xpos[m->hand_body_idx[bi]], 3 * sizeof(double));
      memcpy(f + F_BQUAT + 4 * bi, d->xquat[m->hand_body_idx[bi]], 4 * sizeof(double));
    }
    for (int g = 0; g < ORC_NHG; g++) { cmask[t * ORC_NHG + g] = 0; cpos[(t * ORC_NHG + g) * 3] = cpos[(t * ORC_NHG + g) * 3 + 1] = cpos[(t * ORC_NHG + g) * 3 + 2] = 0; }
    for (int c = 0; c < d->ncon; c++) {
      int g1 = d->contact[c].geom1, g2 = d->contact[c].geom2;
      if (g1 >= m->hand_geom0 && g1 <= m->hand_geom1 && g2 >= m->obj_geom0 && g2 <= m->obj_geom1 && g1 - m->hand_geom0 < nhg) {
        cmask[t * ORC_NHG + g1 - m->hand_geom0] = 1;
        memcpy(cpos + (t * ORC_NHG + g1 - m->hand_geom0) * 3, d->contact[c].pos, 3 * sizeof(double));
      }
    }
  }
  free(d); free(hand);
  return 0;
}

/* ------------------------------------------------------------------------------------------ standalone wrench QP
   The QP of solve_rfc (ho_im4.py:1003-1070) / PhysMetrics.solve_force (scripts/metrics.py:156-196) on explicit inputs,
   solved like solve_rfc above: dense Q, p and the primal active-set.  Checker of hoic_solve_wrench_qp. */
int orc_solve_wrench_qp(int nc, int ec, const double* contacts, const double* weights, const double* penalties, const double* target,
                        const double* center, double mu, double w_t, double* rest, double* score, double* point_force) {
  const double dx = 0.0025, *F = target, *T = target + 3;
  int n = 4 * ec * nc, iters = 0;
  if (nc == 0) {
    memcpy(rest, target, 6 * sizeof(double));
    *score = norm3(F) + sqrt(w_t) * norm3(T);
    return 0;
  }
  double* Jf = (double*)calloc((size_t)3 * n, sizeof(double));
  double* Jt = (double*)calloc((size_t)3 * n, sizeof(double));
  double* pl = (double*)calloc(n, sizeof(double));
  double* Q = (double*)malloc((size_t)n * n * sizeof(double));
  double* x = (double*)malloc(n * sizeof(double));
  double isq = 1.0 / sqrt(1 + mu * mu);
  for (int ci = 0; ci < nc; ci++) {
    const double* cf = contacts + 12 * ci;
    const double *fr0 = cf + 3, *fr1 = cf + 6, *fr2 = cf + 9;
    double ts = weights ? weights[ci] : 1.0;
    for (int j = 0; j < ec; j++) {
      int i = ci * ec + j;
      double pos[3], r[3], A[4][3];
      copy3(pos, cf);
      if (j == 1) addscl3(pos, pos, fr1, dx); else if (j == 2) addscl3(pos, pos, fr1, -dx);
      else if (j == 3) addscl3(pos, pos, fr2, dx); else if (j == 4) addscl3(pos, pos, fr2, -dx);
      for (int k = 0; k < 3; k++) {
        A[0][k] = (fr0[k] + mu * fr1[k]) * isq; A[1][k] = (fr0[k] - mu * fr1[k]) * isq;
        A[2][k] = (fr0[k] + mu * fr2[k]) * isq; A[3][k] = (fr0[k] - mu * fr2[k]) * isq;
      }
      sub3(r, pos, center);
      for (int a = 0; a < 4; a++) {
        double rxA[3];
        cross3(rxA, r, A[a]);
        for (int k = 0; k < 3; k++) { Jf[k * n + 4 * i + a] = ts * A[a][k]; Jt[k * n + 4 * i + a] = ts * rxA[k]; }
        if (penalties) pl[4 * i + a] = penalties[4 * i + a];
      }
    }
  }
  for (int a = 0; a < n; a++) {
    double s = 0;
    for (int b = 0; b < n; b++) {
      double q = 0;
      for (int k = 0; k < 3; k++) q += Jf[k * n + a] * Jf[k * n + b] + w_t * Jt[k * n + a] * Jt[k * n + b];
      Q[(size_t)a * n + b] = 2 * q + (a == b ? 1e-7 : 0);
    }
    for (int k = 0; k < 3; k++) s += -2 * Jf[k * n + a] * F[k] - 2 * w_t * Jt[k * n + a] * T[k];
    pl[a] += s;
  }
  iters = qp_nonneg(n, Q, pl, x);
  for (int k = 0; k < 3; k++) {
    double sf = 0, st = 0;
    for (int a = 0; a < n; a++) { sf += Jf[k * n + a] * x[a]; st += Jt[k * n + a] * x[a]; }
    rest[k] = F[k] - sf; rest[3 + k] = T[k] - st;
  }
  *score = norm3(rest) + sqrt(w_t) * norm3(rest + 3);
  if (point_force)
    for (int i = 0; i < nc * ec; i++)
      for (int k = 0; k < 3; k++) { point_force[3 * i + k] = 0; for (int a = 0; a < 4; a++) point_force[3 * i + k] += Jf[k * n + 4 * i + a] * x[4 * i + a]; }
  free(Jf); free(Jt); free(pl); free(Q); free(x);
  return iters;
}

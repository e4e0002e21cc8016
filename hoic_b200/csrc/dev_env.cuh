// dev_env.cuh -- the HandObjMimic4 env logic around the physics, one warp per environment:
// stable PD + torque clipping (uhc/envs/ho_im4.py:393-486, :521), gravity compensation and residual wrench through
// J^T (:488-501, :526-536), contact recording / averaging (:883-889, :567-597), residual-force QP (:941-1083),
// termination (:648-660, :664-688), reward 9 (uhc/envs/ho_reward.py:943-1046) and observation v5 (:280-356).
#pragma once
#include "dev_math.cuh"

// expert frame offsets (hoic_b200/motion.py)
#define F_DOF 0
#define F_DOFVEL 26
#define F_OBJPOSE 52
#define F_OBJVEL 59
#define F_OBJANGVEL 62
#define F_BPOS 65
#define F_BQUAT 128

struct EnvCtx {
  const float* motion;   // frames of this env's expert (already offset by motion start + start_idx)
  int expert_len, cur_t, start_ind;
};
__device__ __forceinline__ const float* expert_frame(const EnvCtx& e, int delta) {
  int ind = e.cur_t + delta + e.start_ind;
  if (ind > e.expert_len - 1) ind = e.expert_len - 1;
  return e.motion + (size_t)ind * H_FRAME;
}

// ---- stable PD with the previous forward's M and bias.  Writes S_CTRL.
static __device__ __noinline__ bool pd_torque(float* s, const DevModel& m, int lane, const EnvCtx& e) {
  const int n = m.nhand_v;
  const float dt = m.timestep;
  const float* ref = expert_frame(e, 0) + F_DOF;
  float err = 0.f, rhs = 0.f, kp = 0.f, kd = 0.f, qv = 0.f;
  if (lane < n) {
    const float a = s[S_ACT + lane];
    const float r = __ldg(ref + lane);
    const float target = lane < 3 ? r + 0.1f * a : (lane < 6 ? r + 0.3f * a : r + m.ctrl_scale[lane] * a);
    qv = s[S_QVEL + lane];
    err = s[S_QPOS + lane] + qv * dt - target;
    if (lane >= 3) {
      // the reference's `while err > pi: err -= 2 pi` (ho_im4.py:455-460), bounded: a state whose error is beyond a few turns
      // has blown up and is failed by the state check of the forward that follows
#pragma unroll 1
      for (int it = 0; it < 8 && err > F_PI; it++) err -= 2.f * F_PI;
#pragma unroll 1
      for (int it = 0; it < 8 && err < -F_PI; it++) err += 2.f * F_PI;
    }
    kp = m.jkp[lane]; kd = m.jkd[lane];
    rhs = -s[S_BIAS + lane] - kp * err - kd * qv;
  }
  // (M_hand + Kd dt) acc = rhs: the object's dofs are a separate tree, their part of the solve is ignored
  float acc;
  const bool ok = tree_solve(s, m, lane, lane < n ? kd * dt : 0.f, rhs, &acc);
  float t = 0.f;
  if (lane < n) {
    t = -kp * err - kd * (qv + acc * dt);
    t = fminf(fmaxf(t, -m.torque_lim[lane]), m.torque_lim[lane]);
  }
  s[S_CTRL + lane] = t;
  __syncwarp();
  return ok;
}

// ---- record_contact: add hand-object contacts of the last forward to the per-hand-geom sums (world positions)
static __device__ __noinline__ void record_contacts(float* s, const DevModel& m, int lane, float* csum) {
  const int ncon = ((const int*)(s + S_MISC))[MI_NCON];
  const V3 org = ld3(s + S_ORG);
  for (int c = 0; c < ncon; c++) {
    const float* cr = s + U_CON + 16 * c;
    const int slot = m.cand[__float_as_int(cr[13])].hand_slot;
    if (slot < 0) continue;
    if (lane < 12) {
      float v = cr[lane];
      if (lane < 3) v += comp(org, lane);
      csum[12 * slot + lane] += v;
    } else if (lane == 12) {
      csum[240 + slot] += 1.f;
    }
  }
  __syncwarp();
}

// ---- test-mode recording (compute_data_frame, ho_im4.py:599-609; layout HOIC_REC_* in include/hoic.h).  Like the
// reference, which records after sim.step(), body positions / contacts / contact forces / qacc are those of the forward
// that produced the new state.  record_contact_frames runs before the solve (the Newton weights reuse the head of the
// contact records), record_forward after it (U_FORCE / U_CINFO live), record_state after Euler.
static __device__ __noinline__ void record_contact_frames(const float* s, const DevModel& m, int lane, float* o) {
  const int ncon = ((const int*)(s + S_MISC))[MI_NCON];
  const V3 org = ld3(s + S_ORG);
  int n = 0;
  for (int c = 0; c < ncon; c++) {
    const float* cr = s + U_CON + 16 * c;
    const DevCand& cd = m.cand[__float_as_int(cr[13])];
    if (cd.hand_slot < 0) continue;
    if (n < HOIC_REC_MAXCON) {
      float* q = o + HOIC_REC_CON + 19 * n;
      if (lane < 12) q[lane] = cr[lane] + (lane < 3 ? comp(org, lane) : 0.f);
      else if (lane == 18) q[18] = (float)cd.g1;
    }
    n++;
  }
  if (lane == 0) o[114] = (float)(n < HOIC_REC_MAXCON ? n : HOIC_REC_MAXCON);
}
// `con` / `force` / `cinfo`: the U_CON / U_FORCE / U_CINFO arrays of the constraint stage that produced qacc (the pipeline's
// kernels pass their copies in the global working set: the kernel that runs Euler does not hold the constraint stage).
static __device__ __noinline__ void record_forward(const float* s, const DevModel& m, int lane, float* o, const float* con, const float* force,
                                            const float* cinfo) {
  const int* misc = (const int*)(s + S_MISC);
  const V3 org = ld3(s + S_ORG);
  for (int i = lane; i < 3 * m.n_hand_body; i += 32) {
    const int b = i / 3, k = i - 3 * b;
    o[i] = s[S_XPOS + 3 * m.hand_body_idx[b] + k] + comp(org, k);
  }
  if (lane < 6) o[108 + lane] = s[S_QACC + m.nv - 6 + lane];
  const int ncon = misc[MI_NCON];
  const float* F = force + H_RCON;
  int n = 0;
  for (int c = 0; c < ncon; c++) {
    const DevCand& cd = m.cand[__float_as_int(con[16 * c + 13])];
    if (cd.hand_slot < 0) continue;
    if (n < HOIC_REC_MAXCON && lane < 6) {
      // mj_contactForce: mju_decodePyramid of the edge forces (frictionless contacts: the single row's force)
      const float* ci = cinfo + 8 * c;
      const int eoff = ((const int*)ci)[2], ne = ((const int*)ci)[3], k = lane;
      float f = 0.f;
      if (ne == 1) f = k == 0 ? F[eoff] : 0.f;
      else if (k == 0) { for (int e = 0; e < ne; e++) f += F[eoff + e]; }
      else if (2 * k <= ne) f = (F[eoff + 2 * (k - 1)] - F[eoff + 2 * (k - 1) + 1]) * (k <= 2 ? ci[4] : ci[5]);
      o[HOIC_REC_CON + 19 * n + 12 + lane] = f;
    }
    n++;
  }
}
static __device__ __noinline__ void record_state(const float* s, const DevModel& m, int lane, V3 vf, V3 vt, float* o) {
  if (lane < m.nhand_q) o[63 + lane] = s[S_QPOS + lane];
  if (lane < 7) o[95 + lane] = s[S_QPOS + m.nhand_q + lane];
  if (lane < 6) {
    o[102 + lane] = s[S_QVEL + m.nv - 6 + lane];
    o[89 + lane] = lane < 3 ? comp(vf, lane) : comp(vt, lane - 3);
  }
  if (lane == 0) o[115] = 1.f;
}

// pytorch3d-style matrix -> axis-angle (uhc/utils/transforms.py:99-158, 414-490)
static __device__ __noinline__ V3 matrix_to_axis_angle(const M3& R) {
  const float* r = R.m;
  float qa[4] = {1 + r[0] + r[4] + r[8], 1 + r[0] - r[4] - r[8], 1 - r[0] + r[4] - r[8], 1 - r[0] - r[4] + r[8]};
  float cand[4][4] = {{0, r[7] - r[5], r[2] - r[6], r[3] - r[1]}, {r[7] - r[5], 0, r[3] + r[1], r[2] + r[6]},
                      {r[2] - r[6], r[3] + r[1], 0, r[5] + r[7]}, {r[3] - r[1], r[6] + r[2], r[7] + r[5], 0}};
  int best = 0;
  for (int i = 0; i < 4; i++) { qa[i] = qa[i] > 0.f ? sqrtf(qa[i]) : 0.f; cand[i][i] = qa[i] * qa[i]; }
  for (int i = 1; i < 4; i++) if (qa[i] > qa[best]) best = i;
  float q[4];
  for (int i = 0; i < 4; i++) q[i] = cand[best][i] / (2.0f * fmaxf(qa[best], 0.1f));
  // small rotations: the vector part is exact to first order, take the angle from it (atan2 of tiny/1 in float32)
  const float nrm = sqrtf(q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  const float half = atan2f(nrm, q[0]), ang = 2.f * half;
  const float sc = fabsf(ang) < 1e-6f ? 0.5f - ang * ang / 48.f : sinf(half) / ang;
  return v3(q[1] / sc, q[2] / sc, q[3] / sc);
}

// ---- residual-force QP (solve_rfc).  Dual Newton in float64 on the 6-D wrench multiplier (DESIGN.md "RFC QP").
// Scratch (union region): QP_CF contact-geom frames [19][12], QP_TS [19], QP_G1 [19] ints, QP_C [380] penalties.
#define QP_CF (QP_BASE)
#define QP_TS (QP_CF + H_NHG * 12)
#define QP_G1 (QP_TS + 32)
#define QP_C (QP_G1 + 32)
#define QP_GV (QP_C + 384)       // geom_avg_vel / ang_vel of the contact geoms [19][6]
#define QP_END (QP_GV + H_NHG * 6)

struct QpVar { double g[6]; };
static __device__ __noinline__ void qp_var(const float* s, int i, int ec, V3 obj_p, float sqw, float mu, QpVar& out, V3* pos_out) {
  const float dx = 0.0025f;
  const int pt = i >> 2, e = i & 3, k = pt / ec, j = pt % ec;
  const float* cf = s + QP_CF + 12 * k;
  const V3 n = ld3(cf + 3), t = ld3(cf + 6), b = ld3(cf + 9);
  V3 pos = ld3(cf);
  if (j == 1) pos = pos + t * dx; else if (j == 2) pos = pos - t * dx; else if (j == 3) pos = pos + b * dx; else if (j == 4) pos = pos - b * dx;
  const double isq = 1.0 / sqrt(1.0 + (double)mu * mu);
  const V3 tt = e < 2 ? t : b;
  const double sg = (e & 1) ? -(double)mu : (double)mu;
  const double ts = (double)s[QP_TS + k];
  const double ax = ((double)n.x + sg * tt.x) * isq, ay = ((double)n.y + sg * tt.y) * isq, az = ((double)n.z + sg * tt.z) * isq;
  const double rx = (double)pos.x - obj_p.x, ry = (double)pos.y - obj_p.y, rz = (double)pos.z - obj_p.z;
  out.g[0] = ts * ax; out.g[1] = ts * ay; out.g[2] = ts * az;
  out.g[3] = ts * sqw * (ry * az - rz * ay); out.g[4] = ts * sqw * (rz * ax - rx * az); out.g[5] = ts * sqw * (rx * ay - ry * ax);
  if (pos_out) *pos_out = pos;
}

// in-register 6x6 SPD solve (all lanes redundantly).  One reciprocal per pivot: a float64 division or square root is a
// ~30-instruction sequence with a slow-path call, and this runs on the critical path of every QP Newton iteration.
__device__ __forceinline__ bool solve6(double H[6][6], const double* rhs, double* x) {
  double inv[6];
  for (int j = 0; j < 6; j++) {
    double sd = H[j][j];
    for (int k = 0; k < j; k++) sd -= H[j][k] * H[j][k];
    if (!(sd > 1e-300)) return false;
    const double r = rsqrt(sd);            // L_jj = sd * r, 1 / L_jj = r
    inv[j] = r;
    for (int i = j + 1; i < 6; i++) {
      double t = H[i][j];
      for (int k = 0; k < j; k++) t -= H[i][k] * H[j][k];
      H[i][j] = t * r;
    }
  }
  for (int i = 0; i < 6; i++) { double t = rhs[i]; for (int k = 0; k < i; k++) t -= H[i][k] * x[k]; x[i] = t * inv[i]; }
  for (int i = 5; i >= 0; i--) { double t = x[i]; for (int k = i + 1; k < 6; k++) t -= H[k][i] * x[k]; x[i] = t * inv[i]; }
  return true;
}

// sums v[0..31] across the warp by recursive halving: lane L ends with the total of element L (31 shuffles of doubles
// in 5 dependent steps instead of 5 per element)
__device__ __forceinline__ double warp_transpose_sum(double (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int j = 0; j < off; j++) {
      const double send = upper ? v[j] : v[j + off];
      const double keep = upper ? v[j + off] : v[j];
      v[j] = keep + __shfl_xor_sync(FULLMASK, send, off);
    }
  }
  return v[0];
}

#define QP_G (QP_END + 2)                 // [380][6] float: wrench column of each variable (scaled torque rows)
#define QP_RED (QP_G + 384 * 6)           // 32 doubles: reduced gradient / Hessian entries
#define QP_NPER 12                        // variables per lane (380 / 32 rounded up)
static_assert(((QP_RED + 1) & ~1) + 64 <= U_QP_END, "QP scratch exceeds the union region reserved for it");

// Dual Newton iterations of the residual-force QP on the per-variable data left in QP_C / QP_G by rfc_qp.  NPER = variables
// per lane: the loops over a lane's variables are unrolled (their values live in registers), so small problems (most
// grasps touch 2-5 hand geoms = 40-100 variables) run a 3- or 6-wide instance instead of the 12-wide one.
template <int NPER>
static __device__ __noinline__ float rfc_qp_core(float* s, int lane, int n, V3 F, V3 T, float* rest, unsigned& qp_iters, float* surf, double sqw) {
  const double eps = 1e-7, inv_eps = 1.0 / eps;
  double* red = reinterpret_cast<double*>(s + ((QP_RED + 1) & ~1));
  const double b6[6] = {F.x, F.y, F.z, sqw * T.x, sqw * T.y, sqw * T.z};
  double bmax = 0;
  for (int a = 0; a < 6; a++) bmax = fmax(bmax, fabs(b6[a]));
  // (warm-starting lam from the env's previous QP was measured: 7.8 vs 7.4 iterations, no gain -- the target wrench of
  // consecutive steps comes from averaged accelerations and moves too much)
  double lam[6] = {0, 0, 0, 0, 0, 0};
  double sv[NPER], gd[NPER];       // per owned variable: s_i = c_i + g_i.lam and g_i.dl
  for (int it = 0; it < 60; it++) {
    double acc[32];
#pragma unroll
    for (int k = 0; k < 32; k++) acc[k] = 0;
#pragma unroll
    for (int t = 0; t < NPER; t++) {
      const int i = lane + 32 * t;
      sv[t] = 1.0;
      if (i < n) {
        double g[6], si = (double)s[QP_C + i];
#pragma unroll
        for (int k = 0; k < 6; k++) { g[k] = (double)s[QP_G + 6 * i + k]; si += g[k] * lam[k]; }
        sv[t] = si;
        if (si < 0) {
          int q = 6;
#pragma unroll
          for (int a = 0; a < 6; a++) {
            acc[a] += si * g[a];
#pragma unroll
            for (int c = 0; c <= a; c++) acc[q++] += g[a] * g[c];
          }
        }
      }
    }
    const double mine = warp_transpose_sum(acc, lane);
    if (lane < 27) red[lane] = mine;
    __syncwarp();
    double H[6][6], grad[6], dl[6];
    {
      int q = 6;
#pragma unroll
      for (int a = 0; a < 6; a++) {
        grad[a] = 0.5 * lam[a] + b6[a] + red[a] * inv_eps;
#pragma unroll
        for (int c = 0; c <= a; c++) { const double hv = red[q++] * inv_eps; H[a][c] = hv; H[c][a] = hv; }
        H[a][a] += 0.5;
      }
    }
    __syncwarp();
    double gmax = 0;
    for (int a = 0; a < 6; a++) gmax = fmax(gmax, fabs(grad[a]));
    qp_iters++;
    if (gmax <= 1e-9 * (1.0 + bmax)) break;
    double ng[6];
    for (int a = 0; a < 6; a++) ng[a] = -grad[a];
    if (!solve6(H, ng, dl)) break;
    double ll = 0, dd = 0, bd = 0;
    for (int a = 0; a < 6; a++) { ll += lam[a] * dl[a]; dd += dl[a] * dl[a]; bd += b6[a] * dl[a]; }
#pragma unroll
    for (int t = 0; t < NPER; t++) {
      const int i = lane + 32 * t;
      gd[t] = 0;
      if (i < n) {
#pragma unroll
        for (int k = 0; k < 6; k++) gd[t] += (double)s[QP_G + 6 * i + k] * dl[k];
      }
    }
    // exact line search on the piecewise quadratic along dl (safeguarded Newton on the derivative)
    double lo = 0, hi = -1, alpha = 1.0, d1 = 0, d2 = 0;
    for (int ls = 0; ls < 40; ls++) {
      double a1 = 0, a2 = 0;
#pragma unroll
      for (int t = 0; t < NPER; t++) {
        const double sa = sv[t] + alpha * gd[t];
        if (sa < 0) { a1 += sa * gd[t]; a2 += gd[t] * gd[t]; }
      }
      d1 = 0.5 * (ll + alpha * dd) + bd + wsumd(a1) * inv_eps;
      d2 = 0.5 * dd + wsumd(a2) * inv_eps;
      if (fabs(d1) <= 1e-12 * (fabs(bd) + fabs(0.5 * ll) + 1e-30) || !(d2 > 0)) break;
      if (d1 < 0) lo = alpha; else hi = alpha;
      double nxt = alpha - d1 / d2;
      if (hi >= 0) {
        if (!(nxt > lo && nxt < hi)) nxt = 0.5 * (lo + hi);
        if (hi - lo <= 1e-15 * hi) { alpha = nxt; break; }
      } else if (!(nxt > alpha)) nxt = 2 * alpha;
      alpha = nxt;
    }
    double step = 0, lmax = 0;
    for (int a = 0; a < 6; a++) { lam[a] += alpha * dl[a]; step = fmax(step, fabs(alpha * dl[a])); lmax = fmax(lmax, fabs(lam[a])); }
    if (step <= 1e-14 * (1.0 + lmax)) break;
  }
  // primal recovery: x_i = max(0, -s_i/eps); residual wrench in the scaled space
  double gx[6] = {0, 0, 0, 0, 0, 0};
  for (int i = lane; i < ((n + 31) & ~31); i += 32) {
    double g[6] = {0, 0, 0, 0, 0, 0}, x = 0;
    if (i < n) {
      double si = (double)s[QP_C + i];
      for (int k = 0; k < 6; k++) { g[k] = (double)s[QP_G + 6 * i + k]; si += g[k] * lam[k]; }
      if (si < 0) { x = -si * inv_eps; for (int k = 0; k < 6; k++) gx[k] += g[k] * x; }
    }
    if (surf) {
      // test-mode record: surface_contact_force of a point = (ts A) x over its 4 pyramid edges = 4 adjacent lanes
      double f[3];
      for (int k = 0; k < 3; k++) {
        f[k] = g[k] * x;
        f[k] += __shfl_xor_sync(FULLMASK, f[k], 1);
        f[k] += __shfl_xor_sync(FULLMASK, f[k], 2);
      }
      if (i < n && (i & 3) == 0) { float* q = surf + HOIC_SURF_PT + 15 * (i >> 2) + 12; q[0] = (float)f[0]; q[1] = (float)f[1]; q[2] = (float)f[2]; }
    }
  }
  double r[6];
  for (int k = 0; k < 6; k++) r[k] = b6[k] - wsumd(gx[k]);
  for (int k = 0; k < 3; k++) { rest[k] = (float)r[k]; rest[3 + k] = (float)(r[3 + k] / sqw); }
  return (float)(sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]) + sqrt(r[3] * r[3] + r[4] * r[4] + r[5] * r[5]));
}

// Standalone wrench QP on caller-supplied contacts (PhysMetrics.solve_force, scripts/metrics.py:156-196, and the QP of
// solve_rfc with explicit inputs): contact frames in QP_CF, weights in QP_TS, optional penalties c (NULL = 0).
static __device__ __noinline__ float wrench_qp(float* s, int lane, int ncg, int ec, V3 F, V3 T, V3 center, float mu, float sqw, const float* pen,
                                        float* rest, unsigned& qp_iters, float* surf) {
  const int n = 4 * ec * ncg;
  for (int i = lane; i < n; i += 32) {
    QpVar v;
    V3 pos;
    qp_var(s, i, ec, center, sqw, mu, v, &pos);
    s[QP_C + i] = pen ? pen[i] : 0.f;
    for (int a = 0; a < 6; a++) s[QP_G + 6 * i + a] = (float)v.g[a];
    if (surf && (i & 3) == 0) {
      const float* cf = s + QP_CF + 12 * ((i >> 2) / ec);
      float* q = surf + HOIC_SURF_PT + 15 * (i >> 2);
      st3(q, pos); st3(q + 3, ld3(cf + 3)); st3(q + 6, ld3(cf + 6)); st3(q + 9, ld3(cf + 9));
    }
  }
  __syncwarp();
  float score;
  if (n <= 96) score = rfc_qp_core<3>(s, lane, n, F, T, rest, qp_iters, surf, (double)sqw);
  else if (n <= 192) score = rfc_qp_core<6>(s, lane, n, F, T, rest, qp_iters, surf, (double)sqw);
  else score = rfc_qp_core<QP_NPER>(s, lane, n, F, T, rest, qp_iters, surf, (double)sqw);
  if (surf && lane < 8) surf[lane] = lane == 0 ? (float)(n >> 2) : (lane == 7 ? 1.f : rest[lane - 1]);
  return score;
}

// returns rfc_score; fills rest[6]
static __device__ __noinline__ float rfc_qp(float* s, int lane, int ncg, int ec, V3 F, V3 T, V3 obj_p, V3 obj_v, V3 obj_w, const float* gpos_now,
                        V3 org_now, float* rest, unsigned& qp_iters, float* surf) {
  const double sqw = 100.0;
  const int n = 4 * ec * ncg;
  // per-variable data: penalty c_i and wrench column g_i
  for (int i = lane; i < n; i += 32) {
    QpVar v;
    V3 pos;
    qp_var(s, i, ec, obj_p, (float)sqw, 0.75f, v, &pos);
    const int pt = i >> 2, e = i & 3, k = pt / ec;
    const float* cf = s + QP_CF + 12 * k;
    const V3 nn = ld3(cf + 3), t = ld3(cf + 6), b = ld3(cf + 9);
    const int g1 = ((const int*)(s + QP_G1))[k];
    const V3 crh = pos - (ld3(gpos_now + 3 * g1) + org_now), cro = pos - obj_p;
    const V3 cvh = ld3(s + QP_GV + 6 * k) + cross(ld3(s + QP_GV + 6 * k + 3), crh);
    const V3 cvo = obj_v + cross(obj_w, cro);
    const V3 rel = cvo - cvh;
    const float vnd = dot(nn, rel);
    const V3 vn = nn * vnd, vt = rel - vn;
    float c = vnd <= 0.f ? norm(vn) : 0.f;
    const float dirs[4] = {-dot(vt, t), dot(vt, t), -dot(vt, b), dot(vt, b)};
    int mx = 0;
    for (int a = 1; a < 4; a++) if (dirs[a] > dirs[mx]) mx = a;
    if (e != mx) c += norm(vt);
    s[QP_C + i] = c;
    for (int a = 0; a < 6; a++) s[QP_G + 6 * i + a] = (float)v.g[a];
    if (surf && e == 0) {   // surface_contact_frames: the expanded point with the averaged frame (ho_im4.py:983-992)
      float* q = surf + HOIC_SURF_PT + 15 * pt;
      st3(q, pos); st3(q + 3, nn); st3(q + 6, t); st3(q + 9, b);
    }
  }
  __syncwarp();
  float score;
  if (n <= 96) score = rfc_qp_core<3>(s, lane, n, F, T, rest, qp_iters, surf, sqw);
  else if (n <= 192) score = rfc_qp_core<6>(s, lane, n, F, T, rest, qp_iters, surf, sqw);
  else score = rfc_qp_core<QP_NPER>(s, lane, n, F, T, rest, qp_iters, surf, sqw);
  if (surf && lane < 8) surf[lane] = lane == 0 ? (float)(n >> 2) : (lane == 7 ? 1.f : rest[lane - 1]);
  return score;
}


// ---- observation v5 (617 floats) from the last forward's poses and the current qpos / qvel
__device__ __forceinline__ void put_rot2(float* o, const M3& R) {
  o[0] = R.m[0]; o[1] = R.m[1]; o[2] = R.m[3]; o[3] = R.m[4]; o[4] = R.m[6]; o[5] = R.m[7];
}
static __device__ __noinline__ void write_obs(const float* s, const DevModel& m, int lane, const EnvCtx& e, float* obs) {
  const float* qpos = s + S_QPOS;
  const float* qvel = s + S_QVEL;
  const V3 org = ld3(s + S_ORG);
  const V3 rp = ld3(s + S_XPOS + 3 * m.palm_body);   // relative to org
  const Q4 rq = ldq(s + S_XQUAT + 4 * m.palm_body);
  const M3 R = q2m(rq);
  const Q4 rqi = qconj(rq);
  const int W = m.w_size, nh = m.nhand_q, nhv = m.nhand_v, nb = m.n_hand_body, nf = nh - 6;
  int k = 0;
  if (lane == 0) put_rot2(obs + k, R);
  k += 6;
  for (int i = lane; i < nf; i += 32) obs[k + i] = qpos[6 + i];
  k += nf;
  for (int w = 0; w < W; w++) {
    const float* f = expert_frame(e, 1 + w);
    for (int i = lane; i < nf; i += 32) obs[k + i] = __ldg(f + F_DOF + 6 + i) - qpos[6 + i];
    k += nf;
  }
  for (int i = lane; i < nhv - 6; i += 32) obs[k + i] = qvel[6 + i];
  k += nhv - 6;
  if (lane == 0) { st3(obs + k, mulTv(R, ld3(qvel))); st3(obs + k + 3, mulTv(R, ld3(qvel + 3))); }
  k += 6;
  if (lane < W) {
    const float* f = expert_frame(e, 1 + lane);
    const V3 tp = v3(__ldg(f + F_BPOS), __ldg(f + F_BPOS + 1), __ldg(f + F_BPOS + 2));
    st3(obs + k + 9 * lane, mulTv(R, (tp - org) - rp));
    const Q4 tq = q4(__ldg(f + F_BQUAT), __ldg(f + F_BQUAT + 1), __ldg(f + F_BQUAT + 2), __ldg(f + F_BQUAT + 3));
    put_rot2(obs + k + 9 * lane + 3, q2m(qmul(rqi, tq)));
  }
  k += 9 * W;
  // joint positions relative to qpos[:3], component-major
  const V3 q3 = ld3(qpos);
  const int nj = nb - 1;
  for (int i = lane; i < 3 * nj; i += 32) {
    const int c = i / nj, b = 1 + i % nj;
    const V3 t = (ld3(s + S_XPOS + 3 * m.hand_body_idx[b]) + org) - q3;
    obs[k + i] = R.m[c] * t.x + R.m[3 + c] * t.y + R.m[6 + c] * t.z;
  }
  k += 3 * nj;
  for (int w = 0; w < W; w++) {
    const float* f = expert_frame(e, 1 + w);
    for (int i = lane; i < 3 * nj; i += 32) {
      const int c = i / nj, b = 1 + i % nj;
      const V3 tp = v3(__ldg(f + F_BPOS + 3 * b), __ldg(f + F_BPOS + 3 * b + 1), __ldg(f + F_BPOS + 3 * b + 2));
      const V3 t = (tp - org) - ld3(s + S_XPOS + 3 * m.hand_body_idx[b]);
      obs[k + i] = R.m[c] * t.x + R.m[3 + c] * t.y + R.m[6 + c] * t.z;
    }
    k += 3 * nj;
  }
  const V3 op = ld3(qpos + nh);
  const Q4 oq = ldq(qpos + nh + 3);
  if (lane == 0) {
    st3(obs + k, mulTv(R, (op - org) - rp));
    put_rot2(obs + k + 3, q2m(qmul(rqi, oq)));
    st3(obs + k + 9, mulTv(R, ld3(qvel + nhv)));
    st3(obs + k + 12, mulTv(R, ld3(qvel + nhv + 3)));
  }
  k += 15;
  if (lane < W) {
    const float* f = expert_frame(e, 1 + lane);
    const V3 tp = v3(__ldg(f + F_OBJPOSE), __ldg(f + F_OBJPOSE + 1), __ldg(f + F_OBJPOSE + 2));
    st3(obs + k + 9 * lane, mulTv(R, tp - op));
    const Q4 tq = q4(__ldg(f + F_OBJPOSE + 3), __ldg(f + F_OBJPOSE + 4), __ldg(f + F_OBJPOSE + 5), __ldg(f + F_OBJPOSE + 6));
    const float n2 = oq.w * oq.w + oq.x * oq.x + oq.y * oq.y + oq.z * oq.z;
    Q4 oqi = qconj(oq);
    oqi = q4(oqi.w / n2, oqi.x / n2, oqi.y / n2, oqi.z / n2);
    put_rot2(obs + k + 9 * lane + 3, q2m(qnormalize(qmul(rqi, qmul(tq, oqi)))));
  }
}

// ---- calc_ho_diff + reward 9.  Returns body_fail; reward / info written by lane 0.
static __device__ __noinline__ bool diff_and_reward(const float* s, const DevModel& m, int lane, const EnvCtx& e, const float* rw, float rfc_score,
                                float* reward, float* info9) {
  const float* f = expert_frame(e, 0);
  const float* qpos = s + S_QPOS;
  const float* qvel = s + S_QVEL;
  const V3 org = ld3(s + S_ORG);
  const int nh = m.nhand_q, nb = m.n_hand_body;
  // per-body terms, lane = hand body
  float jd = 0.f, qd = 0.f;
  if (lane < nb) {
    const int b = m.hand_body_idx[lane];
    const V3 ep = v3(__ldg(f + F_BPOS + 3 * lane), __ldg(f + F_BPOS + 3 * lane + 1), __ldg(f + F_BPOS + 3 * lane + 2));
    jd = norm((ld3(s + S_XPOS + 3 * b) + org) - ep);
    const Q4 eq = q4(__ldg(f + F_BQUAT + 4 * lane), __ldg(f + F_BQUAT + 4 * lane + 1), __ldg(f + F_BQUAT + 4 * lane + 2), __ldg(f + F_BQUAT + 4 * lane + 3));
    const float n2 = eq.w * eq.w + eq.x * eq.x + eq.y * eq.y + eq.z * eq.z;
    Q4 qi = qconj(eq);
    qi = q4(qi.w / n2, qi.x / n2, qi.y / n2, qi.z / n2);
    const Q4 d = qmul(ldq(s + S_XQUAT + 4 * b), qi);
    const float d0 = fabsf(d.w) - 1.f;
    qd = sqrtf(d0 * d0 + d.x * d.x + d.y * d.y + d.z * d.z);
  }
  const float jpos_mean = wsum(jd) / nb, wq_mean = wsum(qd) / nb;
  float pd = 0.f, vd = 0.f;
  if (lane >= 6 && lane < nh) pd = fabsf(qpos[lane] - __ldg(f + F_DOF + lane));
  if (lane < nh) vd = fabsf(qvel[lane] - __ldg(f + F_DOFVEL + lane));
  const float pose_mean = wsum(pd) / (nh - 6), vel_mean = wsum(vd) / nh;
  // root / object
  const V3 rp = ld3(s + S_XPOS + 3 * m.palm_body) + org;
  const Q4 rq = ldq(s + S_XQUAT + 4 * m.palm_body);
  const V3 erp = v3(__ldg(f + F_BPOS), __ldg(f + F_BPOS + 1), __ldg(f + F_BPOS + 2));
  const Q4 erq = q4(__ldg(f + F_BQUAT), __ldg(f + F_BQUAT + 1), __ldg(f + F_BQUAT + 2), __ldg(f + F_BQUAT + 3));
  const float pos_diff = norm(rp - erp);
  const Q4 qdf = qmul(erq, qconj(rq));
  const float rot_diff = 2.0f * asinf(fminf(sqrtf(qdf.x * qdf.x + qdf.y * qdf.y + qdf.z * qdf.z), 1.f));
  const V3 op = ld3(qpos + nh);
  const V3 eop = v3(__ldg(f + F_OBJPOSE), __ldg(f + F_OBJPOSE + 1), __ldg(f + F_OBJPOSE + 2));
  const float obj_dist = norm(op - eop);
  const bool body_fail = pos_diff > m.pos_thresh || rot_diff > m.rot_thresh || jpos_mean > m.jpos_thresh ||
                         obj_dist > m.obj_pos_thresh || 0.f > m.obj_rot_thresh;   // obj_rot_diff is identically 0 (ho_im4.py:685)
  if (lane == 0 && (reward || info9)) {
    const Q4 eoq = q4(__ldg(f + F_OBJPOSE + 3), __ldg(f + F_OBJPOSE + 4), __ldg(f + F_OBJPOSE + 5), __ldg(f + F_OBJPOSE + 6));
    const float n2 = eoq.w * eoq.w + eoq.x * eoq.x + eoq.y * eoq.y + eoq.z * eoq.z;
    Q4 qi = qconj(eoq);
    qi = q4(qi.w / n2, qi.x / n2, qi.y / n2, qi.z / n2);
    const Q4 d = qmul(ldq(qpos + nh + 3), qi);
    const float d0 = fabsf(d.w) - 1.f;
    const float orot = sqrtf(d0 * d0 + d.x * d.x + d.y * d.y + d.z * d.z);
    float ov = 0.f;
    for (int i = 0; i < 3; i++) ov += fabsf(qvel[m.nhand_v + i] - __ldg(f + F_OBJVEL + i)) + fabsf(qvel[m.nhand_v + 3 + i] - __ldg(f + F_OBJANGVEL + i));
    const float pose_r = expf(-rw[8] * pose_mean), wpose_r = expf(-rw[9] * wq_mean), vel_r = expf(-rw[10] * vel_mean);
    const float jpos_r = expf(-rw[11] * jpos_mean), opos_r = expf(-rw[12] * obj_dist), orot_r = expf(-rw[13] * orot);
    const float ovel_r = expf(-rw[14] * ov / 6.f), orfc_r = m.residual_force ? expf(-rw[15] * rfc_score) : 1.f;
    const float hand = (rw[0] * pose_r + rw[1] * wpose_r + rw[3] * jpos_r + rw[2] * vel_r) / (rw[0] + rw[1] + rw[3] + rw[2]);
    const float obj = (rw[4] * opos_r + rw[5] * orot_r + rw[6] * ovel_r + rw[7] * orfc_r) / (rw[4] + rw[5] + rw[6] + rw[7]);
    if (reward) *reward = hand * obj;
    if (info9) {
      info9[0] = pose_r; info9[1] = wpose_r; info9[2] = jpos_r; info9[3] = vel_r; info9[4] = opos_r; info9[5] = orot_r;
      info9[6] = ovel_r; info9[7] = orfc_r; info9[8] = 1.f;
    }
  }
  return body_fail;
}

// dev_dynamics.cuh -- forward kinematics, CoM quantities, composite-rigid-body mass matrix, bias forces (RNE),
// Jacobian transposes and semi-implicit Euler.  One warp per environment: lane = body / dof / geom index.
// Restates MuJoCo 2.1 mj_kinematics / mj_comPos / mj_crb / mj_comVel / mj_rne / mj_jac / mj_Euler for this model
// class; reference call site uhc/envs/ho_im4.py:545 (sim.step()).
// Positions are stored relative to S_ORG (the hand root translation at forward time) to keep float32 cancellation
// out of contact depths and relative poses.
#pragma once
#include "dev_math.cuh"

// ---------------------------------------------------------------------------------------------- position stage
// Scratch aliases inside the union (dead before the constraint stage): per-joint local rotation / slide and per-body
// local transform are computed lane-parallel first, so that the serial walk down the tree is one quaternion product
// and one rotation per level.
// `gpos` / `gmat`: where the geom poses go ([32][3], [32][9]); nothing in the position / velocity stage reads them back, so
// the kernel that only produces them (k_p1) points these at the global working set instead of shared memory.
static __device__ __noinline__ void pos_stage(float* s, const DevModel& m, int lane, float* gpos, float* gmat) {
  const float* q = s + S_QPOS;
  const int nbody = m.nbody;
  if (lane < 3) s[S_ORG + lane] = q[m.org_qadr + lane];
  // (1) lane = joint: local joint motion
  if (lane < m.njnt) {
    const int t = m.jnt_type[lane];
    const float qq = t == J_FREE ? 0.f : q[m.jnt_qposadr[lane]] - m.qpos0[m.jnt_qposadr[lane]];
    Q4 dq = q4(qq, 0.f, 0.f, 0.f);
    if (t == J_HINGE) dq = aa2q(ld3(m.jnt_axis[lane]), qq);
    stq(s + U_JQ + 4 * lane, dq);
  }
  __syncwarp();
  const V3 org = ld3(s + S_ORG);
  // (2) lane = body: transform of the body relative to its parent's frame, joint axes / anchors in the parent frame
  const int b = lane;
  const int lvl = b < nbody ? m.body_level[b] : -1;
  V3 lp = v3(0, 0, 0);
  Q4 lq = q4(1.f, 0.f, 0.f, 0.f);
  bool isfree = false;
  if (b > 0 && b < nbody) {
    const int ja = m.body_jntadr[b], jn = m.body_jntnum[b];
    isfree = jn == 1 && m.jnt_type[ja] == J_FREE;
    if (!isfree) {
      lp = ld3(m.body_pos[b]);
      lq = ldq(m.body_quat[b]);
      for (int j = ja; j < ja + jn; j++) {
        const M3 R = q2m(lq);
        const V3 ax = mulv(R, ld3(m.jnt_axis[j]));
        const V3 jp = ld3(m.jnt_pos[j]);
        const V3 anchor = lp + mulv(R, jp);
        st3(s + U_XAXIS + 3 * j, ax);
        st3(s + U_XANCH + 3 * j, anchor);
        const Q4 dq = ldq(s + U_JQ + 4 * j);
        if (m.jnt_type[j] == J_SLIDE) {
          lp = lp + ax * dq.w;
        } else {
          lq = qmul(lq, dq);
          lp = anchor - rotq(lq, jp);
        }
      }
    }
  }
  if (lane == 0) {
    st3(s + S_XPOS, -org);
    stq(s + S_XQUAT, q4(1.f, 0.f, 0.f, 0.f));
    stm(s + U_XMAT, q2m(q4(1.f, 0.f, 0.f, 0.f)));
  }
  __syncwarp();
  // (3) walk down the tree, one level at a time
  for (int L = 1; L <= m.nlevel; L++) {
    if (lvl == L) {
      V3 pos;
      Q4 quat;
      if (isfree) {
        const int a = m.jnt_qposadr[m.body_jntadr[b]];
        pos = ld3(q + a) - org;
        quat = qnormalize(ldq(q + a + 3));
        stq(s + S_QPOS + a + 3, quat);  // mj_kinematics normalises qpos in place
      } else {
        const int p = m.body_parent[b];
        pos = ld3(s + S_XPOS + 3 * p) + mulv(ldm(s + U_XMAT + 9 * p), lp);
        quat = qnormalize(qmul(ldq(s + S_XQUAT + 4 * p), lq));
      }
      st3(s + S_XPOS + 3 * b, pos);
      stq(s + S_XQUAT + 4 * b, quat);
      stm(s + U_XMAT + 9 * b, q2m(quat));
    }
    __syncwarp();
  }
  // (4) lane = joint: axis / anchor to world
  if (lane < m.njnt) {
    const int jb = m.jnt_body[lane];
    if (m.jnt_type[lane] == J_FREE) {
      st3(s + U_XANCH + 3 * lane, ld3(s + S_XPOS + 3 * jb));
      st3(s + U_XAXIS + 3 * lane, ld3(m.jnt_axis[lane]));
    } else {
      const int p = m.body_parent[jb];
      const M3 Rp = ldm(s + U_XMAT + 9 * p);
      const V3 ax = mulv(Rp, ld3(s + U_XAXIS + 3 * lane));
      const V3 an = ld3(s + S_XPOS + 3 * p) + mulv(Rp, ld3(s + U_XANCH + 3 * lane));
      st3(s + U_XAXIS + 3 * lane, ax);
      st3(s + U_XANCH + 3 * lane, an);
    }
  }
  // geoms
  if (lane < m.ngeom) {
    const int gb = m.geom_body[lane];
    const M3 R = ldm(s + U_XMAT + 9 * gb);
    st3(gpos + 3 * lane, ld3(s + S_XPOS + 3 * gb) + mulv(R, ld3(m.geom_pos[lane])));
    stm(gmat + 9 * lane, mulm(R, ldm(m.geom_mat[lane])));
  }
  // inertial frames, subtree CoM of each moving tree, cinert
  V3 xipos = v3(0, 0, 0);
  M3 ximat;
  float mass = 0.f;
  int slot = -1;
  if (b < nbody && b > 0) {
    const M3 R = ldm(s + U_XMAT + 9 * b);
    xipos = ld3(s + S_XPOS + 3 * b) + mulv(R, ld3(m.body_ipos[b]));
    ximat = mulm(R, ldm(m.body_imat[b]));
    mass = m.body_mass[b];
    slot = m.body_rootslot[b];
  }
  for (int r = 0; r < m.nroot; r++) {
    const float w = slot == r ? mass : 0.f;
    const float cx = wsum(w * xipos.x), cy = wsum(w * xipos.y), cz = wsum(w * xipos.z);
    if (lane == 0) st3(s + S_COM + 3 * r, v3(cx, cy, cz) * m.root_invmass[r]);
  }
  __syncwarp();
  if (b < nbody) {
    float ci[10];
    if (b > 0 && slot >= 0) inert_com(ci, m.body_inertia[b], ximat, xipos - ld3(s + S_COM + 3 * slot), mass);
    else
      for (int k = 0; k < 10; k++) ci[k] = 0.f;
    for (int k = 0; k < 10; k++) { s[U_CINERT + 10 * b + k] = ci[k]; s[U_CRB + 10 * b + k] = ci[k]; }
  }
  // cdof, lane = dof
  if (lane < m.nv) {
    const int j = m.dof_jnt[lane], jb = m.jnt_body[j], t = m.jnt_type[j];
    const V3 off = ld3(s + S_COM + 3 * m.dof_rootslot[lane]) - ld3(s + U_XANCH + 3 * j);
    V3 rot = v3(0, 0, 0), lin = v3(0, 0, 0);
    if (t == J_FREE) {
      const int k = lane - m.jnt_dofadr[j];
      if (k < 3) { setcomp(lin, k, 1.f); }
      else { rot = col(ldm(s + U_XMAT + 9 * jb), k - 3); lin = cross(rot, off); }
    } else if (t == J_SLIDE) {
      lin = ld3(s + U_XAXIS + 3 * j);
    } else {
      rot = ld3(s + U_XAXIS + 3 * j);
      lin = cross(rot, off);
    }
    st3(s + S_CDOF + 6 * lane, rot);
    st3(s + S_CDOF + 6 * lane + 3, lin);
  }
  __syncwarp();
  // composite inertias: parents gather their children, deepest level first
  for (int L = m.nlevel - 1; L >= 1; L--) {
    if (lvl == L) {
      float acc[10];
#pragma unroll
      for (int k = 0; k < 10; k++) acc[k] = s[U_CRB + 10 * b + k];
      for (int c = m.body_firstchild[b]; c >= 0; c = m.body_nextsib[c]) {
#pragma unroll
        for (int k = 0; k < 10; k++) acc[k] += s[U_CRB + 10 * c + k];
      }
#pragma unroll
      for (int k = 0; k < 10; k++) s[U_CRB + 10 * b + k] = acc[k];
    }
    __syncwarp();
  }
  // mass matrix, lane = dof i: entries of row i for its ancestors (by depth) and itself; the rest of the row stays zero
  {
    float4* z = reinterpret_cast<float4*>(s + S_MS + lane * MS_STRIDE);
    z[0] = z[1] = z[2] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (lane < m.nv) {
    float buf[6];
    mul_inert_vec(buf, s + U_CRB + 10 * m.dof_body[lane], s + S_CDOF + 6 * lane);
    int d = 0;   // ancestors come in increasing index = increasing depth
    for (unsigned mk = m.dof_anc[lane] | (1u << lane); mk; mk &= mk - 1, d++) {
      const int j = __ffs(mk) - 1;
      const float* c = s + S_CDOF + 6 * j;
      float v = c[0] * buf[0] + c[1] * buf[1] + c[2] * buf[2] + c[3] * buf[3] + c[4] * buf[4] + c[5] * buf[5];
      if (j == lane) v += m.dof_armature[lane];
      s[S_MS + lane * MS_STRIDE + d] = v;
    }
  }
  __syncwarp();
}

// ---------------------------------------------------------------------------------------------- velocity stage
// Bias forces by recursive Newton-Euler, written as lane-parallel sums over ancestor-dof masks instead of a serial walk:
//   cvel_b = sum_{e in chain(b)} cdof_e qvel_e ;  cdof_dot_d = cvel(before d) x cdof_d ;  cacc_b = -g + sum cdof_dot_e qvel_e
// (mj_comVel / mj_rne).  Needs U_CINERT from pos_stage (same substep).
static __device__ __noinline__ void vel_stage(float* s, const DevModel& m, int lane) {
  const int b = lane;
  const int lvl = b < m.nbody ? m.body_level[b] : -1;
  const float qv = lane < m.nv ? s[S_QVEL + lane] : 0.f;
  float cd[6];
#pragma unroll
  for (int k = 0; k < 6; k++) { cd[k] = lane < m.nv ? s[S_CDOF + 6 * lane + k] : 0.f; s[U_DV + 6 * lane + k] = cd[k] * qv; }
  __syncwarp();
  {
    float vb[6] = {0, 0, 0, 0, 0, 0}, t1[6];
    if (lane < m.nv) {
      for (unsigned mk = m.dof_premask[lane]; mk; mk &= mk - 1) {
        const int e = __ffs(mk) - 1;
#pragma unroll
        for (int k = 0; k < 6; k++) vb[k] += s[U_DV + 6 * e + k];
      }
    }
    cross_motion(t1, vb, cd);
    const float sc = (lane < m.nv && !m.dof_dotzero[lane]) ? qv : 0.f;
#pragma unroll
    for (int k = 0; k < 6; k++) s[U_DA + 6 * lane + k] = t1[k] * sc;
  }
  __syncwarp();
  if (b < m.nbody) {
    float cv[6] = {0, 0, 0, 0, 0, 0}, ca[6] = {0, 0, 0, -m.gravity[0], -m.gravity[1], -m.gravity[2]}, t1[6], t2[6], t3[6];
    for (unsigned mk = m.body_chain[b]; mk; mk &= mk - 1) {
      const int e = __ffs(mk) - 1;
#pragma unroll
      for (int k = 0; k < 6; k++) { cv[k] += s[U_DV + 6 * e + k]; ca[k] += s[U_DA + 6 * e + k]; }
    }
    mul_inert_vec(t1, s + U_CINERT + 10 * b, ca);
    mul_inert_vec(t2, s + U_CINERT + 10 * b, cv);
    cross_force(t3, cv, t2);
#pragma unroll
    for (int k = 0; k < 6; k++) s[U_CFRC + 6 * b + k] = b > 0 ? t1[k] + t3[k] : 0.f;
  }
  __syncwarp();
  for (int L = m.nlevel - 1; L >= 1; L--) {
    if (lvl == L) {
      float acc[6];
#pragma unroll
      for (int k = 0; k < 6; k++) acc[k] = s[U_CFRC + 6 * b + k];
      for (int c = m.body_firstchild[b]; c >= 0; c = m.body_nextsib[c]) {
#pragma unroll
        for (int k = 0; k < 6; k++) acc[k] += s[U_CFRC + 6 * c + k];
      }
#pragma unroll
      for (int k = 0; k < 6; k++) s[U_CFRC + 6 * b + k] = acc[k];
    }
    __syncwarp();
  }
  if (lane < m.nv) {
    const float* f = s + U_CFRC + 6 * m.dof_body[lane];
    s[S_BIAS + lane] = cd[0] * f[0] + cd[1] * f[1] + cd[2] * f[2] + cd[3] * f[3] + cd[4] * f[4] + cd[5] * f[5];
  } else {
    s[S_BIAS + lane] = 0.f;
  }
  __syncwarp();
}

// ---------------------------------------------------------------------------------------------- J^T wrench
// mj_applyFT: qfrc[lane] contribution of (force, torque) at `point` (relative to S_ORG of the stored kinematics)
__device__ __forceinline__ float apply_ft(const float* s, const DevModel& m, int lane, V3 force, V3 torque, V3 point, int body) {
  if (lane >= m.nv || !((m.body_chain[body] >> lane) & 1u)) return 0.f;
  const V3 off = point - ld3(s + S_COM + 3 * m.dof_rootslot[lane]);
  const V3 rot = ld3(s + S_CDOF + 6 * lane), lin = ld3(s + S_CDOF + 6 * lane + 3);
  const V3 jp = lin + cross(rot, off);
  return dot(jp, force) + dot(rot, torque);
}

// ---------------------------------------------------------------------------------------------- Euler
// mj_Euler with implicit joint damping: (M + h diag(damping)) qacc' = qfrc_smooth + qfrc_constraint.
// Returns false on a failed factorisation.
static __device__ __noinline__ bool euler_step(float* s, const DevModel& m, int lane, unsigned long long* chol_calls) {
  const float h = m.timestep;
  float qa;
  const float rhs = lane < m.nv ? s[S_SMOOTH + lane] + s[S_FCON + lane] : 0.f;
  const bool ok = tree_solve(s, m, lane, lane < m.nv ? h * m.dof_damping[lane] : 0.f, rhs, &qa);
  (void)chol_calls;
  __syncwarp();
  if (lane < m.nv) s[S_QVEL + lane] += h * qa;
  __syncwarp();
  if (lane < m.njnt) {
    const int qa_ = m.jnt_qposadr[lane], da = m.jnt_dofadr[lane];
    if (m.jnt_type[lane] == J_FREE) {
      for (int k = 0; k < 3; k++) s[S_QPOS + qa_ + k] += h * s[S_QVEL + da + k];
      V3 w = ld3(s + S_QVEL + da + 3);
      const float ang = h * normalize(w);
      const Q4 qn = qmul(qnormalize(ldq(s + S_QPOS + qa_ + 3)), aa2q(w, ang));
      stq(s + S_QPOS + qa_ + 3, qn);
    } else {
      s[S_QPOS + qa_] += h * s[S_QVEL + da];
    }
  }
  s[S_WARM + lane] = s[S_QACC + lane];
  __syncwarp();
  return ok;
}

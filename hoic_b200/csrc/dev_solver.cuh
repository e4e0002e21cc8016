// dev_solver.cuh -- constraint rows and the Newton solver of MuJoCo 2.1's soft-constraint model, one warp per env.
// Row classes: [0,32) dof friction loss (J = e_i, implicit), [32,64) joint limits (J = +-e_dof, implicit),
// [64, ...) pyramidal contact edges, stored as per-contact BASIS rows (normal, 2 tangents, torsion) so that
// edge row = basis_0 +- mu_k * basis_k.  Restates mj_makeConstraint / mj_makeImpedance / mj_referenceConstraint /
// mj_fwdConstraint(Newton) for this model class; reference call site ho_im4.py:545.
#pragma once
#include "dev_math.cuh"

#ifdef H_SOLVE_STATS
static __device__ unsigned long long g_solve_stats[64];   // [0..15] solves by iterations, [16..31] cycles, [32..47] ls evals, [48..] cycles by part
#define SST(i) do { const long long t_ = clock64(); if (lane == 0) atomicAdd(&g_solve_stats[48 + (i)], (unsigned long long)(t_ - sst_t)); sst_t = t_; } while (0)
#else
#define SST(i) do { } while (0)
#endif
#ifdef H_STATS
static __device__ unsigned long long g_stats[8];
static __device__ unsigned long long g_nst[64];   // [0..7] solves by iteration count, [8..15] their cycles, [16..] cycles by part
#define NST_T(i) do { const long long t_ = clock64(); if (lane == 0) atomicAdd(&g_nst[16 + (i)], (unsigned long long)(t_ - nst_t)); nst_t = t_; } while (0)
#else
#define NST_T(i) do { } while (0)
#endif
#ifndef NEWTON_ASM_UNROLL
#define NEWTON_ASM_UNROLL 1     // A/B knobs (tools/build_variant.sh): unroll factors of the Hessian-assembly column loop ...
#endif
#ifndef NEWTON_AUX_UNROLL
#define NEWTON_AUX_UNROLL 1     // ... and of the per-edge / per-row helper loops of newton_weights / newton_gradient
#endif
#define U_BINFO (U_MPRQ)   // basis row -> contact | k<<8 (the MPR queue is dead once contacts exist)
#define U_CW (U_CON)       // per-contact Hessian weights reuse the contact records' slots 14..15 + CINFO pad

// general power-law sigmoid of mj_makeImpedance; out of line: powf is large and the models here all use power 2
static __device__ __noinline__ float impedance_pow(float x, float s3, float s4) {
  if (x <= s3) return powf(x, s4) / powf(s3, s4 - 1.f);
  return 1.f - powf(1.f - x, s4) / powf(1.f - s3, s4 - 1.f);
}
__device__ __forceinline__ float impedance(const float* si, float pos, float margin) {
  float s0 = fminf(0.9999f, fmaxf(0.0001f, si[0])), s1 = fminf(0.9999f, fmaxf(0.0001f, si[1]));
  float s2 = fmaxf(0.f, si[2]), s3 = fminf(0.9999f, fmaxf(0.0001f, si[3])), s4 = fmaxf(1.f, si[4]);
  if (s0 == s1 || s2 <= F_MINVAL) return 0.5f * (s0 + s1);
  float x = fabsf((pos - margin) / s2), y;
  if (x >= 1.f || x <= 0.f) return x >= 1.f ? s1 : s0;
  if (s4 == 1.f) y = x;
  else if (s4 == 2.f) y = x <= s3 ? x * x / s3 : 1.f - (1.f - x) * (1.f - x) / (1.f - s3);
  else y = impedance_pow(x, s3, s4);
  return s0 + y * (s1 - s0);
}
__device__ __forceinline__ void kb_from_solref(const float* sr, const float* si, float timestep, float& K, float& B) {
  const float dmax = fminf(0.9999f, fmaxf(0.0001f, si[1]));
  if (sr[0] > 0.f) {
    const float tc = fmaxf(sr[0], 2.f * timestep), dr = sr[1];
    B = 2.f / (dmax * tc);
    K = 1.f / (dmax * dmax * tc * tc * dr * dr);
  } else {
    K = -sr[0] / (dmax * dmax);
    B = -sr[1] / dmax;
  }
}

// Builds all rows from the contacts in U_CON and the current qpos / qvel.  Sets MI_NLIM / MI_NBASIS / MI_NEDGE.
static __device__ __noinline__ void build_rows(float* s, const DevModel& m, int lane) {
  int* misc = (int*)(s + S_MISC);
  int* rinfo = (int*)(s + U_RINFO);
  int* binfo = (int*)(s + U_BINFO);
  const float* qv = s + S_QVEL;
  // friction-loss rows
  {
    const bool on = lane < m.nv && m.dof_frictionloss[lane] > 0.f;
    s[U_D + H_RFR + lane] = on ? m.fr_D[lane] : 0.f;
    s[U_AREF + H_RFR + lane] = on ? -m.fr_B[lane] * qv[lane] : 0.f;
  }
  // joint-limit rows: lower sides first, then upper sides (order does not affect the solution)
  int nlim = 0;
  s[U_D + H_RLIM + lane] = 0.f;
  s[U_AREF + H_RLIM + lane] = 0.f;
  rinfo[H_RLIM + lane] = 0;
  __syncwarp();
  for (int side = -1; side <= 1; side += 2) {
    bool on = false;
    float dist = 0.f;
    if (lane < m.njnt && m.jnt_limited[lane] && m.jnt_type[lane] != J_FREE) {
      const float value = s[S_QPOS + m.jnt_qposadr[lane]];
      dist = side * (m.jnt_range[lane][(side + 1) / 2] - value);
      on = dist < m.jnt_margin[lane];
    }
    const unsigned bal = __ballot_sync(FULLMASK, on);
    const int idx = nlim + __popc(bal & ((1u << lane) - 1u));
    if (on && idx < 32) {
      const int dof = m.jnt_dofadr[lane];
      const float margin = m.jnt_margin[lane];
      const float imp = impedance(m.jnt_solimp[lane], dist, margin);
      float K, B;
      kb_from_solref(m.jnt_solref[lane], m.jnt_solimp[lane], m.timestep, K, B);
      const float R = fmaxf(F_MINVAL, (1.f - imp) * m.jnt_invweight[lane] / imp);
      const float sgn = (float)(-side);
      s[U_D + H_RLIM + idx] = 1.0f / R;
      s[U_AREF + H_RLIM + idx] = -B * (sgn * qv[dof]) - K * imp * (dist - margin);
      rinfo[H_RLIM + idx] = dof | ((side < 0 ? 1 : 0) << 8);   // bit 8: J = +1 (lower side), else -1
    }
    nlim += __popc(bal);
  }
  if (nlim > 32) { nlim = 32; if (lane == 0) misc[MI_FLAGS] |= 2; }
  // contacts
  const int ncon = misc[MI_NCON];
  int nbasis = 0, nedge = 0;
  bool overflow = false;
  for (int c = 0; c < ncon; c++) {
    const float* cr = s + U_CON + 16 * c;
    const DevCand& cd = m.cand[__float_as_int(cr[13])];
    const int dim = cd.condim, nb = dim == 1 ? 1 : dim, ne = dim == 1 ? 1 : 2 * (dim - 1);
    float* ci = s + U_CINFO + 8 * c;
    if (nbasis + nb > H_NBASIS || nedge + ne > H_NEDGE) {
      overflow = true;
      if (lane == 0) { ((int*)ci)[0] = 0; ((int*)ci)[1] = 0; ((int*)ci)[2] = 0; ((int*)ci)[3] = 0; ci[6] = 0.f; }
      continue;
    }
    const V3 pos = ld3(cr), f0 = ld3(cr + 3), f1 = ld3(cr + 6), f2 = ld3(cr + 9);
    const float dist = cr[12];
    // jacobian difference column for dof = lane
    float sg = 0.f;
    if (lane < m.nv) sg = (float)((cd.chain2 >> lane) & 1u) - (float)((cd.chain1 >> lane) & 1u);
    const unsigned cmask = cd.chain1 ^ cd.chain2;                       // dofs with a non-zero column
    const int tcol = __popc(cmask & ((1u << lane) - 1u));               // compact column of this lane's dof
    V3 jp = v3(0, 0, 0), jr = v3(0, 0, 0);
    if (sg != 0.f) {
      const V3 off = pos - ld3(s + S_COM + 3 * m.dof_rootslot[lane]);
      const V3 rot = ld3(s + S_CDOF + 6 * lane), lin = ld3(s + S_CDOF + 6 * lane + 3);
      jp = (lin + cross(rot, off)) * sg;
      jr = rot * sg;
    }
    float bcol[4];
    bcol[0] = dot(f0, jp); bcol[1] = dot(f1, jp); bcol[2] = dot(f2, jp); bcol[3] = dot(f0, jr);
    float bvel[4];
    const float qvl = lane < m.nv ? qv[lane] : 0.f;
    for (int a = 0; a < nb; a++) {
      if (sg != 0.f) s[U_B + (nbasis + a) * B_STRIDE + tcol] = bcol[a];
      bvel[a] = wsum(bcol[a] * qvl);
      if (lane == 0) { binfo[nbasis + a] = c | (a << 8); ((unsigned*)(s + U_BMASK))[nbasis + a] = cmask; }
    }
    // impedance / reference (identical on all lanes)
    const float imp = impedance(cd.solimp, dist, cd.includemargin);
    float K, B;
    kb_from_solref(cd.solref, cd.solimp, m.timestep, K, B);
    const float mu0 = cd.friction[0], mu1 = cd.friction[1];
    float D;
    if (dim == 1) {
      D = 1.0f / fmaxf(F_MINVAL, (1.f - imp) * cd.dA_tran / imp);
    } else {
      const float dA = cd.dA_tran + mu0 * mu0 * cd.dA_tran;
      const float R0 = fmaxf(F_MINVAL, (1.f - imp) * dA / imp);
      const float mu = mu0 * sqrtf(1.0f / fmaxf(F_MINVAL, m.impratio));
      D = 1.0f / (2.f * mu * mu * R0);
    }
    const float posterm = K * imp * (dist - cd.includemargin);
    if (lane < ne) {
      float jv;
      if (dim == 1) jv = bvel[0];
      else {
        const int k = lane / 2 + 1;
        const float mu = k <= 2 ? mu0 : mu1;
        jv = bvel[0] + ((lane & 1) ? -mu : mu) * bvel[k];
      }
      const int r = H_RCON + nedge + lane;
      s[U_AREF + r] = -B * jv - posterm;
      s[U_D + r] = D;
      rinfo[r] = c | (lane << 8);
    }
    if (lane == 0) {
      ((int*)ci)[0] = nbasis; ((int*)ci)[1] = nb; ((int*)ci)[2] = nedge; ((int*)ci)[3] = ne;
      ci[4] = mu0; ci[5] = mu1; ci[6] = D;
      ((unsigned*)ci)[7] = cmask;
    }
    nbasis += nb;
    nedge += ne;
  }
  // clear unused edge rows of the last partially used 32-row group so that generic row loops can run unguarded
  for (int r = H_RCON + nedge + lane; r < H_RCON + ((nedge + 31) & ~31); r += 32) { s[U_D + r] = 0.f; s[U_AREF + r] = 0.f; rinfo[r] = 0; }
  if (lane == 0) {
    misc[MI_NLIM] = nlim; misc[MI_NBASIS] = nbasis; misc[MI_NEDGE] = nedge;
    if (overflow) misc[MI_FLAGS] |= 2;
  }
  __syncwarp();
}

// ---- row-wise products ----------------------------------------------------------------------------------------
// out[r] = (J v)[r] for all rows; v in shared (32 floats).  Uses U_BV as scratch.
static __device__ __noinline__ void mul_J(float* s, int lane, int nbasis, int nedge, const float* v, float* out) {
  const int* rinfo = (const int*)(s + U_RINFO);
#pragma unroll 1
  for (int a = lane; a < nbasis; a += 32) {
    const float* B = s + U_B + a * B_STRIDE;
    float acc0 = 0.f, acc1 = 0.f;
    unsigned mk = ((const unsigned*)(s + U_BMASK))[a];
#pragma unroll 1
    for (int t = 0; mk; t += 2) {
      const int j0 = __ffs(mk) - 1;
      mk &= mk - 1;
      acc0 = fmaf(B[t], v[j0], acc0);
      if (mk) {
        const int j1 = __ffs(mk) - 1;
        mk &= mk - 1;
        acc1 = fmaf(B[t + 1], v[j1], acc1);
      }
    }
    s[U_BV + a] = acc0 + acc1;
  }
  __syncwarp();
  out[H_RFR + lane] = v[lane];
  {
    const int ri = rinfo[H_RLIM + lane];
    out[H_RLIM + lane] = ((ri >> 8) & 1 ? 1.f : -1.f) * v[ri & 255];
  }
#pragma unroll 1
  for (int e = lane; e < nedge; e += 32) {
    const int ri = rinfo[H_RCON + e], c = ri & 255, ed = ri >> 8;
    const float* ci = s + U_CINFO + 8 * c;
    const int bo = ((const int*)ci)[0], ne = ((const int*)ci)[3];
    float val = s[U_BV + bo];
    if (ne > 1) {
      const int k = ed / 2 + 1;
      const float mu = k <= 2 ? ci[4] : ci[5];
      val += ((ed & 1) ? -mu : mu) * s[U_BV + bo + k];
    }
    out[H_RCON + e] = val;
  }
  __syncwarp();
}

// per-row force / cost / curvature at constraint-space residual x
__device__ __forceinline__ void row_eval(const float* s, const DevModel& m, int r, float x, float& force, float& cost, float& curv) {
  const float D = s[U_D + r];
  if (r < H_RLIM) {
    const float fl = m.dof_frictionloss[r], Rf = m.fr_Rf[r];
    if (D == 0.f) { force = 0.f; cost = 0.f; curv = 0.f; }
    else if (x <= -Rf) { force = fl; cost = fl * (-0.5f * Rf - x); curv = 0.f; }
    else if (x >= Rf) { force = -fl; cost = fl * (-0.5f * Rf + x); curv = 0.f; }
    else { force = -D * x; cost = 0.5f * D * x * x; curv = D; }
  } else {
    if (x < 0.f) { force = -D * x; cost = 0.5f * D * x * x; curv = D; }
    else { force = 0.f; cost = 0.f; curv = 0.f; }
  }
}

// constraint cost at jar (shared array), optionally writing forces; all lanes get the total
static __device__ __noinline__ float rows_update(float* s, const DevModel& m, int lane, int nedge, const float* jar, bool write) {
  float cost = 0.f;
  const int nrow = H_RCON + nedge;
#pragma unroll 1
  for (int r = lane; r < nrow; r += 32) {
    float f, c, k;
    row_eval(s, m, r, jar[r], f, c, k);
    cost += c;
    if (write) s[U_FORCE + r] = f;
  }
  cost = wsum(cost);
  __syncwarp();
  return cost;
}

// (M v)[lane] with M stored by (dof, ancestor depth) (S_MS): own row over the ancestors, the column part from the one
// descendant per level inside the lane's chain, and for the trunk (hub) dofs the sum over all branch dofs by warp sums.
// The helpers of the Newton iteration below use rolled loops (code bytes cost instruction fetches, profiles/README.md);
// mul_M is the exception: its unrolled form measured 4 % faster on the whole step than a rolled one.
static __device__ __noinline__ float mul_M(const float* s, const DevModel& m, int lane, const float* v) {
  const bool on = lane < m.nv;
  const int dep = on ? m.dof_depth[lane] : 0;
  const bool trunk = on && m.dof_is_trunk[lane];
  const bool branch = (m.branch_mask >> lane) & 1u;
  const unsigned a0 = m.dof_anc_at[lane][0], a1 = m.dof_anc_at[lane][1], a2 = m.dof_anc_at[lane][2];
  const unsigned c0 = m.dof_desc_at[lane][0], c1 = m.dof_desc_at[lane][1], c2 = m.dof_desc_at[lane][2];
  float r[12];
  ld_row12(r, s + S_MS + lane * MS_STRIDE);
  const float vl = v[lane];
  float acc0 = on ? s[S_MS + lane * MS_STRIDE + dep] * vl : vl, acc1 = 0.f;
#ifdef MULM_ROLLED
  {
    const float* Mr = s + S_MS + lane * MS_STRIDE;
#pragma unroll 1
    for (int d = 0; d < T_MAXD - 1; d++) {
      const unsigned w = d < 4 ? a0 : (d < 8 ? a1 : a2);
      const int j = (int)((w >> (8 * (d & 3))) & 0xffu);
      if (j != 0xff) acc0 = fmaf(Mr[d], v[j], acc0);
    }
#pragma unroll 1
    for (int l = 1; l < T_MAXD; l++) {
      const unsigned w = l < 4 ? c0 : (l < 8 ? c1 : c2);
      const int k = (int)((w >> (8 * (l & 3))) & 0xffu);
      if (k != 0xff) acc1 = fmaf(s[S_MS + k * MS_STRIDE + dep], v[k], acc1);
    }
  }
#else
#pragma unroll
  for (int d = 0; d < T_MAXD - 1; d++) {
    const int j = tbyte(a0, a1, a2, d);
    if (j != 0xff) acc0 = fmaf(r[d], v[j], acc0);
  }
#pragma unroll
  for (int l = 1; l < T_MAXD; l++) {
    const int k = tbyte(c0, c1, c2, l);
    if (k != 0xff) acc1 = fmaf(s[S_MS + k * MS_STRIDE + dep], v[k], acc1);
  }
#endif
  {
    // unconditional (uniform control flow keeps the shuffles free of convergence barriers); no hub -> no branch lanes
    float corr = 0.f;
#pragma unroll
    for (int t = 0; t <= T_HUBD; t++) {
      const float q = wsum(branch ? r[t] * vl : 0.f);
      if (t == dep) corr = q;
    }
    if (trunk) acc1 += corr;
  }
  return acc0 + acc1;
}

// gradient: grad = Ma - qfrc_smooth - J^T f (forces from the last rows_update).  Writes S_GRAD and S_FCON = J^T f,
// and the per-basis-row force coefficients in U_BV.  Returns |grad|^2 (all lanes).
static __device__ __noinline__ float newton_gradient(float* s, const DevModel& m, int lane, int nlim, int nbasis) {
  const int* rinfo = (const int*)(s + U_RINFO);
  const int* binfo = (const int*)(s + U_BINFO);
  (void)m;
#pragma unroll 1
  for (int a = lane; a < nbasis; a += 32) {
    const int c = binfo[a] & 255, k = binfo[a] >> 8;
    const float* ci = s + U_CINFO + 8 * c;
    const int eo = H_RCON + ((const int*)ci)[2], ne = ((const int*)ci)[3];
    float g;
    if (k == 0) {
      g = 0.f;
#pragma unroll 1
      for (int e = 0; e < ne; e++) g += s[U_FORCE + eo + e];
    }
    else { const float mu = k <= 2 ? ci[4] : ci[5]; g = mu * (s[U_FORCE + eo + 2 * (k - 1)] - s[U_FORCE + eo + 2 * (k - 1) + 1]); }
    s[U_BV + a] = g;
  }
  __syncwarp();
  float jtf = s[U_FORCE + H_RFR + lane];
#pragma unroll 1
  for (int r = 0; r < nlim; r++) {
    const int ri = rinfo[H_RLIM + r];
    if ((ri & 255) == lane) jtf += ((ri >> 8) & 1 ? 1.f : -1.f) * s[U_FORCE + H_RLIM + r];
  }
H_UNROLL(NEWTON_AUX_UNROLL)
  for (int a = 0; a < nbasis; a++) {
    const unsigned mk = ((const unsigned*)(s + U_BMASK))[a];
    if ((mk >> lane) & 1u) jtf = fmaf(s[U_B + a * B_STRIDE + __popc(mk & ((1u << lane) - 1u))], s[U_BV + a], jtf);
  }
  const float grad = s[S_MA + lane] - s[S_SMOOTH + lane] - jtf;
  s[S_GRAD + lane] = grad;
  s[S_FCON + lane] = jtf;
  __syncwarp();
  return wsum(grad * grad);
}

// Per-contact Hessian weights of the currently quadratic edge rows (written over contact record slots 0..6) and the
// diagonal contribution of the friction-loss / limit rows (returned, lane = dof).
static __device__ __noinline__ float newton_weights(float* s, const DevModel& m, int lane, int nlim, int ncon) {
  const int* rinfo = (const int*)(s + U_RINFO);
  const float* jar = s + U_JAR;
#pragma unroll 1
  for (int c = lane; c < ncon; c += 32) {
    // w[0] = C00, w[1..3] = C0k, w[4..6] = Ckk
    float* ci = s + U_CINFO + 8 * c;
    const int eo = H_RCON + ((const int*)ci)[2], ne = ((const int*)ci)[3];
    const float D = ci[6];
    float* w = s + U_CON + 16 * c;
    w[0] = 0.f; w[1] = 0.f; w[2] = 0.f; w[3] = 0.f; w[4] = 0.f; w[5] = 0.f; w[6] = 0.f;
H_UNROLL(NEWTON_AUX_UNROLL)
    for (int e = 0; e < ne; e++) {
      const float act = jar[eo + e] < 0.f ? D : 0.f;
      w[0] += act;
      if (ne > 1) {
        const int k = e / 2;
        const float mu = k < 2 ? ci[4] : ci[5];
        w[1 + k] += ((e & 1) ? -mu : mu) * act;
        w[4 + k] += mu * mu * act;
      }
    }
  }
  float diag = 0.f;
  {
    float f, c, k;
    row_eval(s, m, H_RFR + lane, jar[H_RFR + lane], f, c, k);
    diag += k;
  }
#pragma unroll 1
  for (int r = 0; r < nlim; r++) {
    const int ri = rinfo[H_RLIM + r];
    if ((ri & 255) == lane && jar[H_RLIM + r] < 0.f) diag += s[U_D + H_RLIM + r];
  }
  __syncwarp();
  return diag;
}

// Newton direction: H = M + J^T D_active J assembled in shared memory (lane i owns row i, columns j <= i, blocked
// layout of warp_chol_solve), Cholesky, search = -H^-1 grad.  All loops are rolled: see warp_chol_solve.
static __device__ __noinline__ bool newton_direction(float* s, const DevModel& m, int lane, int ncon, float diag) {
  float* A = s + S_H;
#ifdef H_STATS
  long long nst_t = clock64();
#endif
#ifdef H_SOLVE_STATS
  const long long sst_a = clock64();
#endif
  {
    const int I = lane >> 2, r = lane & 3;
#pragma unroll 1
    for (int J = 0; J <= I; J++) *reinterpret_cast<float4*>(A + HB_OFF(I, J) + 4 * r) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  // M + diag: entry d of the (dof, depth) row is M[lane, ancestor at depth d]; ancestors come in increasing index
  if (lane < m.nv) {
    const float* Mr = s + S_MS + lane * MS_STRIDE;
    int d = 0;
H_UNROLL(NEWTON_ASM_UNROLL)
    for (unsigned mk = m.dof_anc[lane]; mk; mk &= mk - 1, d++) A[hidx(lane, __ffs(mk) - 1)] = Mr[d];
    A[hidx(lane, lane)] = Mr[d] + diag;
  } else {
    A[hidx(lane, lane)] = 1.f;
  }
#pragma unroll 1
  for (int c = 0; c < ncon; c++) {
    const float* ci = s + U_CINFO + 8 * c;
    const int bo = ((const int*)ci)[0], nb = ((const int*)ci)[1];
    const float* w = s + U_CON + 16 * c;
    if (nb == 0 || w[0] == 0.f) continue;   // overflowed contact / no active edge
    const unsigned mask = ((const unsigned*)ci)[7];
    if (!((mask >> lane) & 1u)) continue;
    const float* B0 = s + U_B + bo * B_STRIDE;
    const int tl = __popc(mask & ((1u << lane) - 1u));
    const float b0 = B0[tl];
    float al0 = w[0] * b0, al1 = 0.f, al2 = 0.f, al3 = 0.f;
    if (nb > 1) {
      const float b1 = B0[B_STRIDE + tl], b2 = B0[2 * B_STRIDE + tl];
      al0 = fmaf(w[1], b1, fmaf(w[2], b2, al0));
      al1 = fmaf(w[4], b1, w[1] * b0);
      al2 = fmaf(w[5], b2, w[2] * b0);
      if (nb > 3) {
        const float b3 = B0[3 * B_STRIDE + tl];
        al0 = fmaf(w[3], b3, al0);
        al3 = fmaf(w[6], b3, w[3] * b0);
      }
    }
    // columns j <= lane of the contact's dofs; the compact column index runs with the set bits
    int tj = 0;
H_UNROLL(NEWTON_ASM_UNROLL)
    for (unsigned mk = mask & (0xffffffffu >> (31 - lane)); mk; mk &= mk - 1, tj++) {
      float acc = al0 * B0[tj];
      if (nb > 1) { acc = fmaf(al1, B0[B_STRIDE + tj], acc); acc = fmaf(al2, B0[2 * B_STRIDE + tj], acc); }
      if (nb > 3) acc = fmaf(al3, B0[3 * B_STRIDE + tj], acc);
      A[hidx(lane, __ffs(mk) - 1)] += acc;
    }
  }
  __syncwarp();
  NST_T(8);    // Hessian assembly
#ifdef H_SOLVE_STATS
  { const long long t_ = clock64(); if (lane == 0) atomicAdd(&g_solve_stats[48 + 5], (unsigned long long)(t_ - sst_a)); }
#endif
  float dir;
  const bool ok = warp_chol_solve(A, -s[S_GRAD + lane], &dir, lane);
  s[S_SEARCH + lane] = dir;
  __syncwarp();
  NST_T(9);    // Cholesky + substitutions
  return ok;
}

// derivative and curvature of the total cost along the search direction at step alpha
static __device__ __noinline__ void ls_eval(const float* s, const DevModel& m, int lane, int nedge, float alpha, float qg1,
                                        float qg2, float& d1, float& d2) {
  float a1 = 0.f, a2 = 0.f;
  const int nrow = H_RCON + nedge;
#pragma unroll 1
  for (int r = lane; r < nrow; r += 32) {
    const float jv = s[U_JV + r];
    float f, c, k;
    row_eval(s, m, r, s[U_JAR + r] + alpha * jv, f, c, k);
    a1 -= f * jv;          // d cost / d alpha = -force * Jv
    a2 += k * jv * jv;
  }
  d1 = qg1 + 2.f * alpha * qg2 + wsum(a1);
  d2 = 2.f * qg2 + wsum(a2);
}

// exact minimiser of the piecewise-quadratic cost along `search` by safeguarded Newton on its derivative
static __device__ __noinline__ float line_search(float* s, const DevModel& m, int lane, int nbasis, int nedge, unsigned& evals) {
  const float sv = s[S_SEARCH + lane];
#ifdef H_SOLVE_STATS
  const long long sst_l = clock64();
#endif
  const float mv = mul_M(s, m, lane, s + S_SEARCH);
  s[S_MV + lane] = mv;
  mul_J(s, lane, nbasis, nedge, s + S_SEARCH, s + U_JV);
#ifdef H_SOLVE_STATS
  { const long long t_ = clock64(); if (lane == 0) atomicAdd(&g_solve_stats[48 + 6], (unsigned long long)(t_ - sst_l)); }
#endif
  const float qg1 = wsum(sv * (s[S_MA + lane] - s[S_SMOOTH + lane]));
  const float qg2 = 0.5f * wsum(sv * mv);
  float d1, d2, lo = 0.f, hi = -1.f, alpha = 1.f;
  // at alpha = 0 the derivative is grad.search and the curvature search'H search = -grad.search (H search = -grad),
  // so the first trial point is the full Newton step alpha = 1 without evaluating the rows at 0
  d1 = wsum(sv * s[S_GRAD + lane]);
  if (!(d1 < 0.f)) return 0.f;
  const float gtol = 1e-6f * fabsf(d1);
#pragma unroll 1
  for (int it = 0; it < 24; it++) {
    ls_eval(s, m, lane, nedge, alpha, qg1, qg2, d1, d2);
    evals++;
    if (fabsf(d1) <= gtol) break;
    if (d1 < 0.f) lo = alpha; else hi = alpha;
    float nxt = alpha - d1 / d2;
    if (hi >= 0.f) {
      if (!(nxt > lo && nxt < hi)) nxt = 0.5f * (lo + hi);
      if (hi - lo <= 1e-7f * hi) { alpha = nxt; break; }
    } else if (!(nxt > alpha)) {
      nxt = 2.f * alpha;
    }
    alpha = nxt;
  }
  return alpha;
}

// ---- mj_fwdConstraint with the Newton solver, split into sub-phases that every warp of a CTA runs in lockstep.
// Inputs: S_SMOOTH, S_ASMOOTH, S_WARM, rows.  Outputs: S_QACC, S_FCON.
// warm start selection and initial cost; leaves S_QACC, S_MA, U_JAR, U_FORCE consistent
static __device__ __noinline__ float newton_init(float* s, const DevModel& m, int lane, int nbasis, int nedge) {
  const float asm_ = s[S_ASMOOTH + lane], fsm = s[S_SMOOTH + lane];
  float qacc = s[S_WARM + lane];
  s[S_QACC + lane] = qacc;
  __syncwarp();
  float ma = mul_M(s, m, lane, s + S_QACC);
  mul_J(s, lane, nbasis, nedge, s + S_QACC, s + U_JAR);
  for (int r = lane; r < H_RCON + nedge; r += 32) s[U_JAR + r] -= s[U_AREF + r];
  __syncwarp();
#ifdef NEWTON_WARM_ONLY
  // A/B: take the warm start whenever it is not the zero vector of a fresh reset (the optimum does not depend on the start)
  const bool cmp = !__any_sync(FULLMASK, qacc != 0.f);
#else
  const bool cmp = true;
#endif
  float cost_warm = 0.f, cost_smooth = 0.f;
  if (cmp) {
    cost_warm = rows_update(s, m, lane, nedge, s + U_JAR, false) + wsum(0.5f * (ma - fsm) * (qacc - asm_));
    mul_J(s, lane, nbasis, nedge, s + S_ASMOOTH, s + U_JV);
    for (int r = lane; r < H_RCON + nedge; r += 32) s[U_JV + r] -= s[U_AREF + r];
    __syncwarp();
    cost_smooth = rows_update(s, m, lane, nedge, s + U_JV, false);
  }
  if (cost_warm > cost_smooth) {
    qacc = asm_;
    s[S_QACC + lane] = qacc;
    for (int r = lane; r < H_RCON + nedge; r += 32) s[U_JAR + r] = s[U_JV + r];
    __syncwarp();
    ma = mul_M(s, m, lane, s + S_QACC);
  }
  s[S_MA + lane] = ma;
  __syncwarp();
  return rows_update(s, m, lane, nedge, s + U_JAR, true) + wsum(0.5f * (ma - fsm) * (qacc - asm_));
}

// line search along S_SEARCH, move, new cost.  Returns 0: zero step (nothing moved), 1: moved and some row changed
// state (active set / friction branch), 2: moved and no row changed state.
static __device__ __noinline__ int newton_move(float* s, const DevModel& m, int lane, int nbasis, int nedge, float& cost, unsigned& ls_evals) {
  const float alpha = line_search(s, m, lane, nbasis, nedge, ls_evals);
  if (alpha == 0.f) return 0;
  const float qacc = s[S_QACC + lane] + alpha * s[S_SEARCH + lane];
  const float ma = s[S_MA + lane] + alpha * s[S_MV + lane];
  s[S_QACC + lane] = qacc;
  s[S_MA + lane] = ma;
  bool changed = false;
#pragma unroll 1
  for (int r = lane; r < H_RCON + nedge; r += 32) {
    const float x0 = s[U_JAR + r], x1 = x0 + alpha * s[U_JV + r];
    s[U_JAR + r] = x1;
    if (s[U_D + r] != 0.f) {
      if (r < H_RLIM) { const float Rf = m.fr_Rf[r]; changed = changed || ((x0 <= -Rf) != (x1 <= -Rf)) || ((x0 >= Rf) != (x1 >= Rf)); }
      else changed = changed || ((x0 < 0.f) != (x1 < 0.f));
    }
  }
  changed = __any_sync(FULLMASK, changed);
  __syncwarp();
  cost = rows_update(s, m, lane, nedge, s + U_JAR, true) + wsum(0.5f * (ma - s[S_SMOOTH + lane]) * (qacc - s[S_ASMOOTH + lane]));
  return changed ? 1 : 2;
}

// Driver.  With H_NEWTON_LOCKSTEP the sub-phases of an iteration are separated by CTA barriers (every warp of the CTA
// must then call it, warps without work pass active = false); measured slower than free-running warps (box 4096 envs:
// 6.98 vs 6.84 ms), so it is off: the imbalance of the solve is the iteration COUNT, not fetch sharing inside it.
#ifdef H_NEWTON_LOCKSTEP
#define NEWTON_ANY(run) __syncthreads_or(run)
#define NEWTON_SYNC() __syncthreads()
#else
#define NEWTON_ANY(run) (run)
#define NEWTON_SYNC() do { } while (0)
#endif
// `budget` > 0: at most that many iterations in this call; a solve that is still running then is left OPEN: S_QACC, S_MA, U_JAR and
// MISC[MI_SV_*] hold its state (the caller stores them) and `open_out` is set.  `resume`: continue such a solve (the caller loaded
// S_QACC, S_MA, U_JAR; the row forces are recomputed from U_JAR, which reproduces them bit for bit), so that a solve split over two
// launches performs exactly the arithmetic of an unsplit one.
// SPLIT = false compiles all of that out (the default kernels).
template <bool SPLIT = false>
__device__ __forceinline__ bool solve_newton(float* s, const DevModel& m, int lane, bool active, unsigned& iters_out, unsigned& ls_evals,
                                             int budget = 0, bool resume = false, bool* open_out = nullptr) {
  int* misc = (int*)(s + S_MISC);
  const int nlim = misc[MI_NLIM], ncon = misc[MI_NCON], nbasis = misc[MI_NBASIS], nedge = misc[MI_NEDGE];
  const float scale = 1.0f / (m.meaninertia * (float)(m.nv > 1 ? m.nv : 1));
  bool run = active, ok = true;
  unsigned iter = 0;
  float cost = 0.f, oldcost = 0.f;
#ifdef H_SOLVE_STATS
  long long sst_t = clock64();
#endif
  if (SPLIT && run && resume) {
    iter = (unsigned)misc[MI_SV_ITER];
    cost = __int_as_float(misc[MI_SV_COST]);
    oldcost = __int_as_float(misc[MI_SV_OLDCOST]);
    rows_update(s, m, lane, nedge, s + U_JAR, true);      // U_FORCE of the saved point
  } else if (run) { cost = newton_init(s, m, lane, nbasis, nedge); oldcost = cost; }
  const unsigned iter0 = iter;
  SST(0);   // init
  for (;;) {
    if (run) {
      const float g2 = newton_gradient(s, m, lane, nlim, nbasis);
      SST(1);   // gradient
      if (iter > 0) {
        const float improvement = scale * (oldcost - cost), gradn = scale * sqrtf(g2);
        // float32 noise floor on the cost difference replaces MuJoCo's absolute 1e-8 (see DESIGN.md "precision")
        const float tol = fmaxf(m.tolerance, 2e-7f * scale * fabsf(cost));
        if (improvement < tol || gradn < m.tolerance) run = false;
      }
      if (iter >= (unsigned)m.iterations) run = false;
      if (SPLIT && run && budget > 0 && iter - iter0 >= (unsigned)budget) {
        // out of budget with the solve still running: leave it open at the top of an iteration (gradient of this point is cheap
        // to redo; cost / oldcost / iter are what the termination test of the continuation needs)
        if (lane == 0) {
          misc[MI_SV_OPEN] = 1; misc[MI_SV_ITER] = (int)iter;
          misc[MI_SV_COST] = __float_as_int(cost); misc[MI_SV_OLDCOST] = __float_as_int(oldcost);
        }
        __syncwarp();
        if (open_out) *open_out = true;
        run = false;
      }
#ifdef NEWTON_CAP      // timing experiment only (what the long solves cost): stop after NEWTON_CAP iterations, converged or not
      if (iter >= NEWTON_CAP) run = false;
#endif
    }
    if (!NEWTON_ANY(run)) break;
    if (run) {
      // full refactorisation every iteration: the blocked Cholesky is cheaper than tracking rank-one changes
      const float diag = newton_weights(s, m, lane, nlim, ncon);
      SST(2);   // weights
      ok = newton_direction(s, m, lane, ncon, diag);
      SST(3);   // Hessian assembly + Cholesky + substitutions
      if (!ok) run = false;
    }
    NEWTON_SYNC();
    if (run) {
      oldcost = cost;
      const int mv = newton_move(s, m, lane, nbasis, nedge, cost, ls_evals);
      SST(4);   // line search + move + cost
      if (mv == 0) run = false;
      else {
        iter++;
        if (mv == 2) {
          // no row changed state along the step: the cost was exactly quadratic there and the Newton step with the
          // exact line search landed on its minimiser, the optimum of the convex problem.  Refresh J^T f and stop.
          newton_gradient(s, m, lane, nlim, nbasis);
          run = false;
        }
      }
    }
  }
  iters_out += iter - iter0;
  // S_FCON holds J^T f of the final forces (written by newton_gradient)
  return ok;
}

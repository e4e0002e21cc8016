// p1_body.cuh -- the work of pipeline kernel 1 (see k_p1.cu) for one env / warp, shared by k_p1 (stand-alone launch) and
// the fused kernel k_a (k_a.cu: Newton solve of substep k-1 followed at once by this body for substep k, so that an env
// whose solve ends early does not wait for the slowest solve of the batch before its smooth dynamics start).
// FUSED: the working set is already in shared memory (MISC, QPOS, QVEL, BIAS, COM, CDOF, ORG, MS loaded by the caller,
// QACC / SMOOTH / FCON left there by the solve, S_ACT set); k >= 1.
#pragma once
#include "hoic_step.cuh"

// gpos_out / gmat_out: where mj_kinematics writes the geom poses (k_p1 / k_a: straight into the global working set; k_p12: its
// shared-memory fields, which k_p2's half reads).  Returns true when the forward of this launch ran (the poses are valid).
template <bool FUSED>
__device__ __forceinline__ bool p1_body(const BatchDev& B, const EnvSlot es, const int env, float* s, float* g, float* st, const DevModel& m,
                                        int* misc, const float* __restrict__ actions, const int k, const int lane, float* gpos_out,
                                        float* gmat_out) {
  bool fwd = false;
  bool ok = es.live;
  float* recf = (B.rec_frames && es.live) ? B.rec_frames + (size_t)env * m.sim_step * HOIC_REC_STRIDE : nullptr;
  if (!FUSED && k < 0) {
    misc[lane] = 0;
    __syncwarp();      // lane 0 sets MI_CARRIED (lane 20's slot) below: racecheck, profiles/r02b_sanitizer.txt
    int* ints = (int*)(st + ST_INTS);
    // The stale quantities of SURVEY S1 / S2 are the last forward of the previous env.step.  If this env's working set still
    // holds that forward (ints[7]: no reset / set_state since, no failure), nothing has to be recomputed: take the geom-pose
    // snapshot of ho_im4.py:508-511 and the first contact recording from what the previous step left.
    const bool carried = es.live && ints[7] != 0;
    if (es.live) {
      float* csum = B.csum + (size_t)env * CSUM_STRIDE;
      const float* carry = B.csum_carry + (size_t)env * CSUM_STRIDE;
      for (int i = lane; i < CSUM_STRIDE; i += 32) csum[i] = carried ? carry[i] : 0.f;
      if (recf) for (int i = lane; i < m.sim_step * HOIC_REC_STRIDE; i += 32) recf[i] = 0.f;
      if (B.rec_surface) {
        float* surf = B.rec_surface + (size_t)env * HOIC_SURF_STRIDE;
        for (int i = lane; i < HOIC_SURF_STRIDE; i += 32) surf[i] = 0.f;
      }
    }
    if (carried) {
      float* og = B.old_geom + (size_t)env * (HN * 12 + 4);
      for (int i = lane; i < m.ngeom * 3; i += 32) og[i] = g[G_OFF(GPOS) + i];
      for (int i = lane; i < m.ngeom * 9; i += 32) og[HN * 3 + i] = g[G_OFF(GMAT) + i];
      if (lane < 3) og[HN * 12 + lane] = g[G_OFF(ORG) + lane];
      if (lane == 0) misc[MI_CARRIED] = 1;
      __syncwarp();
      ST_FIELD(MISC);
      for (int i = lane; i < H_NQ; i += 32) g[G_OFF(QPOS) + i] = i < 33 ? st[ST_QPOS + i] : 0.f;
      g[G_OFF(QVEL) + lane] = st[ST_QVEL + lane];
      g[G_OFF(WARM) + lane] = st[ST_WARM + lane];
      return false;
    }
    load_state(s, st, lane, ST_PQPOS, ST_PQVEL);
    __syncwarp();
  } else if (!FUSED) {
    LD_FIELD(MISC); LD_FIELD(QPOS); LD_FIELD(QVEL); LD_FIELD(BIAS); LD_FIELD(COM); LD_FIELD(CDOF); LD_FIELD(ORG);
    LD_FIELD(MS);
    if (k > 0) {
      LD_FIELD(QACC); LD_FIELD(SMOOTH); LD_FIELD(FCON);   // Euler of substep k-1 (sets S_WARM)
      if (recf) LD_FIELD(XPOS);
    } else {
      LD_FIELD(WARM);
    }
    {
      const float a = lane < m.nu + (m.residual_force ? 6 : 0) ? actions[(size_t)env * HOIC_ACT_DIM + lane] : 0.f;
      s[S_ACT + lane] = fminf(fmaxf(a, -1.f), 1.f);
    }
    ld_wait();
    if (!ok || env_dead(misc)) return false;    // failed earlier in this step: the working set stays as it was
  }
  const V3 vf = (k >= 0 && m.residual_force) ? ld3(s + S_ACT + m.nu) * m.rf_scale : v3(0, 0, 0);
  const V3 vt = (k >= 0 && m.residual_force) ? ld3(s + S_ACT + m.nu + 3) * m.rt_scale : v3(0, 0, 0);
  if (ok && k > 0) {
    // integrate substep k-1 with the M of its forward
    float* rf = recf ? recf + (size_t)(k - 1) * HOIC_REC_STRIDE : nullptr;
    if (rf) record_forward(s, m, lane, rf, g + U_G(CON), g + U_G(FORCE), g + U_G(CINFO));
    ok = euler_step(s, m, lane, nullptr);
    if (!ok && lane == 0) misc[MI_FLAGS] |= 1;
    if (rf && ok) record_state(s, m, lane, vf, vt, rf);
  }
  if (ok && k >= 0) {
    // mj_checkPos / mj_checkVel at the head of mj_step: a non-finite or exploding state fails the env before anything reads it
    if (__any_sync(FULLMASK, bad(s[S_QPOS + lane]) || bad(s[S_QVEL + lane]) || (lane == 0 && bad(s[S_QPOS + 32])))) {
      ok = false;
      if (lane == 0) misc[MI_FLAGS] |= 1;
      __syncwarp();
    }
  }
  if (ok && k >= 0) {
    const EnvCtx ctx = make_ctx(B, (const int*)(st + ST_INTS));
    ok = pd_torque(s, m, lane, ctx);
    // J^T wrenches with the previous forward's Jacobians (stale cdof / com / geom_xpos)
    const V3 org = ld3(s + S_ORG);
    float qf = apply_ft(s, m, lane, v3(0.f, 0.f, m.hand_mass * 9.8f), v3(0, 0, 0), ld3(g + G_OFF(GPOS) + 3 * m.hand_geom0), m.palm_body);
    if (m.residual_force) qf += apply_ft(s, m, lane, vf, vt, ld3(s + S_QPOS + m.nhand_q) - org, m.obj_body);
    __syncwarp();
    s[S_QAPP + lane] = qf;
    // the state this forward is computed from: what the next env.step recomputes its stale quantities from
    st[ST_PQPOS + lane] = s[S_QPOS + lane];
    st[ST_PQVEL + lane] = s[S_QVEL + lane];
    if (lane == 0) st[ST_PQPOS + 32] = s[S_QPOS + 32];
    __syncwarp();
    if (!ok && lane == 0) misc[MI_FLAGS] |= 1;
  }
  if (ok) { fwd = phase_forward(s, m, lane, gpos_out, gmat_out); ok = fwd || k < 0; }   // a bad stale state was reported by the previous step already
  if (ok) vel_stage(s, m, lane);
  if (ok && k >= 0) {
    ok = phase_smooth(s, m, lane);
    if (!ok && lane == 0) misc[MI_FLAGS] |= 1;
  }
  __syncwarp();
  if (!es.live) return false;
  ST_FIELD(MISC);
  if (k < 0) {
    // scratch carries the CURRENT state into substep 0; the stale forward's derived quantities stay
    for (int i = lane; i < H_NQ; i += 32) g[G_OFF(QPOS) + i] = i < 33 ? st[ST_QPOS + i] : 0.f;
    g[G_OFF(QVEL) + lane] = st[ST_QVEL + lane];
    g[G_OFF(WARM) + lane] = st[ST_WARM + lane];
  } else {
    ST_FIELD(QPOS); ST_FIELD(QVEL); ST_FIELD(WARM); ST_FIELD(SMOOTH); ST_FIELD(ASMOOTH);
  }
  ST_FIELD(ORG); ST_FIELD(COM); ST_FIELD(CDOF); ST_FIELD(BIAS); ST_FIELD(MS);
  ST_FIELD(XPOS); ST_FIELD(XQUAT);   // read by the step tail (poses of the last forward) and by the test-mode record
  return fwd;
}

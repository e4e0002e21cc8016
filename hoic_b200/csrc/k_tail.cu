// k_tail.cu -- last kernel of env.step: mj_Euler of the last substep, then everything HandObjMimic4.step does after its
// substep loop (ho_im4.py:553-565, 631-662): averaged object acceleration / geom velocities, contact classification,
// the residual-force QP (solve_rfc), cur_t += 1, calc_ho_diff thresholds, reward 9, observation v5, state write-back.
// One env per CTA: the QP's iteration count differs from env to env (and contact-free envs skip it).
#define H_KERNEL HK_TAIL
#include "hoic_step.cuh"

#ifndef PT_W
#define PT_W 1
#endif
#ifndef PT_WPS
#define PT_WPS 12
#endif

__global__ void __launch_bounds__(32 * PT_W, PT_WPS / PT_W) k_tail(BatchDev B, StepPart sp, const float* __restrict__ actions, StepOut out) {
  extern __shared__ __align__(16) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const EnvSlot es = env_of_slot(B, sp, blockIdx.x * PT_W + warp);
  const int env = es.env;
  if (!es.live) return;
  float* s = smem + (size_t)warp * S_TOTAL;
  float* g = B.scratch + (size_t)env * G_TOTAL;
  float* st = B.state + (size_t)env * ST_STRIDE;
  int* ints = (int*)(st + ST_INTS);
  const DevModel& m = *B.env_model[env];
  int* misc = (int*)(s + S_MISC);
  LD_FIELD(MISC); LD_FIELD(QPOS); LD_FIELD(QVEL); LD_FIELD(WARM); LD_FIELD(ORG); LD_FIELD(XPOS); LD_FIELD(XQUAT); LD_FIELD(GPOS);
  LD_FIELD(GMAT); LD_FIELD(QACC); LD_FIELD(SMOOTH); LD_FIELD(FCON); LD_FIELD(MS);
  ld_wait();
  bool ok = !env_dead(misc);
  float* recf = B.rec_frames ? B.rec_frames + ((size_t)env * m.sim_step + (m.sim_step - 1)) * HOIC_REC_STRIDE : nullptr;
  float* surf = B.rec_surface ? B.rec_surface + (size_t)env * HOIC_SURF_STRIDE : nullptr;
  V3 vf = v3(0, 0, 0), vt = v3(0, 0, 0);
  if (m.residual_force) {
    const float* a = actions + (size_t)env * HOIC_ACT_DIM + m.nu;
    auto clip1 = [](float x) { return fminf(fmaxf(x, -1.f), 1.f); };
    vf = v3(clip1(a[0]), clip1(a[1]), clip1(a[2])) * m.rf_scale;
    vt = v3(clip1(a[3]), clip1(a[4]), clip1(a[5])) * m.rt_scale;
  }
  const float old_obj_vel = (lane >= m.nv - 6 && lane < m.nv) ? st[ST_QVEL + lane] : 0.f;   // state at the start of the step
  if (ok) {
    if (recf) record_forward(s, m, lane, recf, g + U_G(CON), g + U_G(FORCE), g + U_G(CINFO));
    ok = euler_step(s, m, lane, nullptr);
    if (!ok && lane == 0) misc[MI_FLAGS] |= 1;
    if (recf && ok) record_state(s, m, lane, vf, vt, recf);
    __syncwarp();
  }
  bool fail = !ok;
  const int flags = misc[MI_FLAGS];
  const bool phys = (flags & 1) != 0;
  Cnt cnt;
  memset(&cnt, 0, sizeof cnt);
  cnt.substeps = misc[MI_C_SUBSTEPS]; cnt.ncon = misc[MI_C_NCON]; cnt.nefc = misc[MI_C_NEFC]; cnt.newton = misc[MI_C_NEWTON];
  cnt.ls = misc[MI_C_LS]; cnt.mpr_calls = misc[MI_C_MPRCALLS]; cnt.mpr_iters = misc[MI_C_MPRITERS];
  cnt.max_ncon = misc[MI_C_MAXNCON]; cnt.max_nbasis = misc[MI_C_MAXNBASIS]; cnt.max_nedge = misc[MI_C_MAXNEDGE];
  if (flags & 1) cnt.fail_phys++;
  if (flags & 2) cnt.fail_over++;
  float rfc_score = 0.f, rest[6] = {0, 0, 0, 0, 0, 0};
  int ncg = 0;
  if (!fail)
    rfc_score = step_tail(s, m, lane, B.old_geom + (size_t)env * (HN * 12 + 4), old_obj_vel, vf, vt, rest, ncg, cnt.qp, surf,
                          B.csum + (size_t)env * CSUM_STRIDE);
  // ---- bookkeeping, termination, reward, observation (ho_im4.py:641-662)
  EnvCtx ctx = make_ctx(B, ints);
  ctx.cur_t += 1;
  float reward = 0.f;
  float* info = out.info ? out.info + (size_t)env * HOIC_INFO_DIM : nullptr;
  if (!phys) {
    const bool body_fail = diff_and_reward(s, m, lane, ctx, B.reward_w, rfc_score, &reward, info);
    if (B.mode_train) fail = fail || body_fail;
    if (out.obs) write_obs(s, m, lane, ctx, out.obs + (size_t)env * HOIC_OBS_DIM);
  } else {
    // The simulation blew up (non-finite state / failed factorisation; MuJoCo: mjWARN_BADQACC -> MujocoException -> fail,
    // ho_im4.py:635-637).  The transition is terminal; its reward, info and observation are finite (zero) so that one bad
    // env cannot poison the batch statistics, and the env keeps the state it had when the step started.
    if (info && lane < HOIC_INFO_DIM) info[lane] = 0.f;
    if (out.obs) for (int i = lane; i < HOIC_OBS_DIM; i += 32) out.obs[(size_t)env * HOIC_OBS_DIM + i] = 0.f;
  }
  const bool end = ctx.cur_t + ctx.start_ind >= ctx.expert_len - m.w_size - 1;
  if (lane == 0) {
    if (out.reward) out.reward[env] = reward;
    if (out.done) out.done[env] = fail || end;
    if (B.reset_list && (fail || end)) B.reset_list[atomicAdd(B.reset_count, 1)] = env;
    if (out.fail) out.fail[env] = fail;
    if (out.end) out.end[env] = end;
    if (out.percent) out.percent[env] = (float)ctx.cur_t / (float)(ctx.expert_len - 1);
    ints[0] = ctx.cur_t;
    ints[7] = (flags & 3) || !ok ? 0 : 1;      // the working set holds this step's last forward: the next step starts from it
    ints[5] = (int)(cnt.newton * 8 + cnt.ncon * 2 + cnt.mpr_iters + (ncg ? 40 + 8 * ncg : 0));
    st[ST_RFC] = rfc_score;
    for (int k = 0; k < 6; k++) st[ST_RFC + 1 + k] = rest[k];
    st[ST_RFC + 7] = (float)ncg;
  }
  if (!phys) {
    st[ST_QPOS + lane] = s[S_QPOS + lane];
    st[ST_QVEL + lane] = s[S_QVEL + lane];
    st[ST_WARM + lane] = s[S_WARM + lane];
    if (lane == 0) st[ST_QPOS + 32] = s[S_QPOS + 32];
  } else {
    st[ST_PQPOS + lane] = st[ST_QPOS + lane];
    st[ST_PQVEL + lane] = st[ST_QVEL + lane];
    if (lane == 0) st[ST_PQPOS + 32] = st[ST_QPOS + 32];
  }
  flush_counters(B, cnt, lane, 1);
}

static int g_pt_pad = -1;
cudaError_t hoic_configure_tail() {
  g_pt_pad = hoic_env_int("HOIC_PT_PAD", 0);
  return cudaFuncSetAttribute(k_tail, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hoic_step_smem(PT_W, g_pt_pad));
}
cudaError_t hoic_launch_tail(const BatchDev& B, StepPart sp, const float* actions, const StepOut& out, cudaStream_t stream) {
  k_tail<<<hoic_step_grid(B, sp, PT_W), 32 * PT_W, hoic_step_smem(PT_W, g_pt_pad < 0 ? 0 : g_pt_pad), stream>>>(B, sp, actions, out);
  return cudaGetLastError();
}

// dev_batch.cuh -- what the kernels of libhoic_b200.so share: the batch descriptor, per-step counters and the phases of
// one mj_step / of the env-step tail as out-of-line device functions.  Included by every kernel translation unit.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include "../../include/hoic.h"
#include "dev_collide.cuh"
#include "dev_dynamics.cuh"
#include "dev_env.cuh"
#include "dev_model.h"
#include "dev_solver.cuh"

struct BatchDev {
  int n_envs;
  const DevModel* const* env_model;   // [N]
  float* state;                       // [N][ST_STRIDE]
  float* old_geom;                    // [N][HN*12 + 4] geom poses (+origin) of the forward that preceded the step
  const float* frames;                // all motions
  const int* motion_offset;
  const int* motion_len;
  int n_motions;
  const float* reward_w;              // [24]
  int mode_train;
  unsigned long long* counters;       // hoic_counters as 12 x u64
  const int* perm;                    // optional env order (load balancing), NULL = identity
  int model_smem;                     // bytes of shared memory behind the env regions that k_step may stage a model into
  float* csum;                        // [N][CSUM_STRIDE] per-hand-geom contact sums / counts of the step (record_contact)
  float* csum_carry;                  // [N][CSUM_STRIDE] record of the LAST forward's contacts: the first recording of the next step (S2)
  float* rec_frames;                  // test-mode recording (hoic_set_record), NULL = off
  float* rec_surface;
  float* scratch;                     // [N][G_TOTAL] working set of the running step between the kernels of the pipeline
  // on-device episode sampler (hoic_set_sampler): motions of model i = samp_ids[samp_off[i] .. samp_off[i+1])
  const int* env_model_idx;           // [N] model index of each env
  const int* samp_off;
  const int* samp_ids;
  int samp_random_start, samp_min_episode;
  unsigned long long samp_seed;
  int* order;                         // [N] library-owned env order (hoic_set_balance), what `perm` points to when it is on
  // envs of the running substep whose constraint stage exceeds the solver's small tier, appended by k_p2 and consumed by
  // k_p3 (full tier): the full-tier launch then covers only those envs instead of starting one 19 KB CTA per env of the batch
  int* full_list;                     // [nparts][N]
  int* full_count;                    // [nparts][FULL_COUNT_STRIDE], entry k = substep, zeroed at the start of a step
  // envs whose episode ended in the running step (appended by k_tail): the auto-reset launch covers only those
  // "carry" mode of hoic_step: a capped k_p3s leaves the few long solves open; k_p3c finishes them on another stream while k_p1 /
  // k_p2 of the next substep already run for everybody else, then k_p1 / k_p2 run for the carried envs alone (short launches)
  int* carry_list;                    // [nparts][N] envs whose solve of the running substep is open (appended by k_p3s)
  int* carry_count;                   // [nparts][FULL_COUNT_STRIDE], behind reset_count's block
  int* late_stamp;                    // [N] substep + 1 of the last solve this env was carried / full-tier in (0 after the step's memset)
  int solve_budget;                   // Newton iterations of the capped launch (0: carry mode off)
  // "tick" mode of hoic_step (env-asynchronous schedule): every env carries its own substep index; a tick runs k_p12 for the envs
  // that are due for their smooth dynamics + collision and a BUDGETED solver launch for the envs that are in their solve (fresh or
  // continued); the kernels of tick t read the lists of tick t and append to those of tick t + 1
  int* tk_k;                          // [N] substep index of each env
  int* tk_list;                       // [3][2][N]: 0 = k_p12, 1 = small solver tier, 2 = full tier; [parity of the tick]
  int* tk_count;                      // [ticks + 2][4]
  int tk_budget;                      // Newton iterations per tick
  int* reset_list;                    // [N]
  int* reset_count;                   // one counter behind full_count's block (same memset)
};
#define FULL_COUNT_STRIDE 32
#define HOIC_LIST_PARTS 8            // = HOIC_MAX_PARTS of hoic_capi.cu
#define TK_LIST(B, which, t) ((B).tk_list + ((size_t)(which) * 2 + ((t) & 1)) * (B).n_envs)
#define TK_COUNT(B, which, t) ((B).tk_count + (size_t)(t) * 4 + (which))

struct StepOut {
  float* obs; float* reward; float* info; uint8_t* done; uint8_t* fail; uint8_t* end; float* percent;
};

struct Cnt { unsigned substeps, ncon, nefc, newton, ls, mpr_calls, mpr_iters, qp, fail_phys, fail_over, max_ncon, max_nbasis, max_nedge; };

__device__ __forceinline__ bool bad(float v) { return !(fabsf(v) <= F_MAXVAL); }

// Phases of one substep.  Kept as separate noinline calls so that every warp of a CTA runs the same (large) piece of
// code between two CTA barriers: the step kernel is instruction-fetch bound otherwise (profiles/r01_v1_*).
static __device__ __noinline__ bool phase_forward(float* s, const DevModel& m, int lane, float* gpos, float* gmat) {
  int* misc = (int*)(s + S_MISC);
  const bool fail = __any_sync(FULLMASK, bad(s[S_QPOS + lane]) || bad(s[S_QVEL + lane]) || (lane == 0 && bad(s[S_QPOS + 32])));
  if (fail) { if (lane == 0) misc[MI_FLAGS] |= 1; __syncwarp(); return false; }
  pos_stage(s, m, lane, gpos, gmat);
  return true;
}
static __device__ __noinline__ bool phase_smooth(float* s, const DevModel& m, int lane) {
  const bool on = lane < m.nv;
  const float passive = on ? -m.dof_damping[lane] * s[S_QVEL + lane] : 0.f;
  const float act = lane < m.nu ? s[S_CTRL + lane] : 0.f;   // motors on dofs 0..nu-1, gear 1
  const float fs = on ? passive - s[S_BIAS + lane] + s[S_QAPP + lane] + act : 0.f;
  s[S_SMOOTH + lane] = fs;
  float as;
  const bool ok = tree_solve(s, m, lane, 0.f, fs, &as);
  s[S_ASMOOTH + lane] = as;
  __syncwarp();
  return ok;
}
// Every warp of the CTA calls this (active = false: only the CTA barriers inside the solver are taken).
// budget / resume / open_out: see solve_newton (a solve split over a capped launch and its continuation); the per-substep counters
// are taken by the launch that STARTS the solve.
template <bool SPLIT = false>
static __device__ __noinline__ bool phase_solve(float* s, const DevModel& m, int lane, Cnt& cnt, bool active, int budget = 0, bool resume = false,
                                                bool* open_out = nullptr) {
  int* misc = (int*)(s + S_MISC);
  unsigned it = 0;
  bool fail = !solve_newton<SPLIT>(s, m, lane, active, it, cnt.ls, budget, resume, open_out);
  if (!active) return false;
  cnt.newton += it;
  if (!(SPLIT && resume)) {
    cnt.substeps++;
    cnt.ncon += misc[MI_NCON];
    cnt.nefc += m.nv + misc[MI_NLIM] + misc[MI_NEDGE];
    cnt.max_ncon = max(cnt.max_ncon, (unsigned)misc[MI_NCON]);
    cnt.max_nbasis = max(cnt.max_nbasis, (unsigned)misc[MI_NBASIS]);
    cnt.max_nedge = max(cnt.max_nedge, (unsigned)misc[MI_NEDGE]);
  }
  if (lane == 0) misc[MI_ITER] = (int)it + (SPLIT && resume ? misc[MI_ITER] : 0);
  fail = __any_sync(FULLMASK, fail || bad(s[S_QACC + lane]));
  if (fail) { if (lane == 0) misc[MI_FLAGS] |= 1; __syncwarp(); return false; }
  return !(misc[MI_FLAGS] & 2);   // capacity overflow: treated like MuJoCo's nconmax / njmax warnings
}

// forward dynamics of the state in S_QPOS/S_QVEL with S_CTRL/S_QAPP/S_WARM, then Euler (single-warp callers)
static __device__ bool substep_core(float* s, const DevModel& m, int lane, bool record, Cnt& cnt) {
  int* misc = (int*)(s + S_MISC);
  if (!phase_forward(s, m, lane, s + S_GPOS, s + S_GMAT)) return false;
  vel_stage(s, m, lane);
  if (!phase_smooth(s, m, lane)) { if (lane == 0) misc[MI_FLAGS] |= 1; __syncwarp(); return false; }
  collide(s, m, lane, cnt.mpr_calls, cnt.mpr_iters);
  (void)record;   // single-substep entry: no contact averaging
  build_rows(s, m, lane);
  if (!phase_solve(s, m, lane, cnt, true)) return false;
  if (!euler_step(s, m, lane, nullptr)) { if (lane == 0) misc[MI_FLAGS] |= 1; __syncwarp(); return false; }
  return true;
}

__device__ __forceinline__ void load_state(float* s, const float* st, int lane, int qoff, int voff) {
  s[S_QPOS + lane] = st[qoff + lane];
  if (lane < 4) s[S_QPOS + 32 + lane] = lane == 0 ? st[qoff + 32] : 0.f;
  s[S_QVEL + lane] = st[voff + lane];
}

static __device__ void flush_counters(const BatchDev& B, const Cnt& c, int lane, unsigned env_steps) {
  if (lane != 0) return;
  unsigned long long* g = B.counters;
  if (env_steps) atomicAdd(g + 0, (unsigned long long)env_steps);
  atomicAdd(g + 1, (unsigned long long)c.substeps);
  atomicAdd(g + 2, (unsigned long long)c.ncon);
  atomicAdd(g + 3, (unsigned long long)c.nefc);
  atomicAdd(g + 4, (unsigned long long)c.newton);
  atomicAdd(g + 5, (unsigned long long)c.ls);
  atomicAdd(g + 6, (unsigned long long)c.mpr_calls);
  atomicAdd(g + 7, (unsigned long long)c.mpr_iters);
  atomicAdd(g + 8, (unsigned long long)c.qp);
  if (c.fail_phys) atomicAdd(g + 9, (unsigned long long)c.fail_phys);
  if (c.fail_over) atomicAdd(g + 10, (unsigned long long)c.fail_over);
  atomicMax(g + 12, (unsigned long long)c.max_ncon);
  atomicMax(g + 13, (unsigned long long)c.max_nbasis);
  atomicMax(g + 14, (unsigned long long)c.max_nedge);
}

__device__ __forceinline__ EnvCtx make_ctx(const BatchDev& B, const int* ints) {
  EnvCtx e;
  const int mid = ints[1], sidx = ints[2];
  e.motion = B.frames + (size_t)(B.motion_offset[mid] + sidx) * H_FRAME;
  e.expert_len = ints[3];
  e.cur_t = ints[0];
  e.start_ind = 0;
  return e;
}

// ---- tail of env.step after the substeps: averages, contact classification, residual-force QP (ho_im4.py:553-565, 631)
static __device__ __noinline__ float step_tail(float* s, const DevModel& m, int lane, const float* og, float old_obj_vel, V3 vf, V3 vt,
                                        float* rest, int& ncg_out, unsigned& qp_iters, float* surf, const float* csum) {
  float rfc_score = 0.f;
  const int nhg = m.hand_geom1 - m.hand_geom0 + 1;
  const float inv_dt = 1.0f / (m.sim_step * m.timestep);
  const V3 org = ld3(s + S_ORG), oorg = ld3(og + HN * 12);
  const float acc = (s[S_QVEL + lane] - old_obj_vel) * inv_dt;   // lanes nv-6..nv-1: obj_avg_acc
  const V3 obj_a = v3(__shfl_sync(FULLMASK, acc, m.nv - 6), __shfl_sync(FULLMASK, acc, m.nv - 5), __shfl_sync(FULLMASK, acc, m.nv - 4));
  const V3 obj_oa = v3(__shfl_sync(FULLMASK, acc, m.nv - 3), __shfl_sync(FULLMASK, acc, m.nv - 2), __shfl_sync(FULLMASK, acc, m.nv - 1));
  // classify_contact: lane = hand geom slot
  const float ccnt = lane < nhg ? csum[240 + lane] : 0.f;
  const bool has = ccnt > 0.f;
  float fr[12];
  if (has) {
    const float ic = 1.0f / ccnt;
    for (int k = 0; k < 12; k++) fr[k] = csum[12 * lane + k] * ic;
    V3 n = ld3(fr + 3);
    n = n * (1.0f / norm(n));
    V3 t = fabsf(n.x) >= 1e-5f ? cross(n, v3(1.f, 0.f, 0.f)) : cross(n, v3(0.f, 1.f, 0.f));
    t = t * (1.0f / norm(t));
    V3 b = cross(n, t);
    b = b * (1.0f / norm(b));
    st3(fr + 3, n); st3(fr + 6, t); st3(fr + 9, b);
  }
  const float ts = has ? ccnt / (float)m.sim_step : 0.f;
  const unsigned bal = __ballot_sync(FULLMASK, has);
  const int ncg = __popc(bal);
  ncg_out = ncg;
  __syncwarp();
  // geom average velocities: contact geoms (lane = slot) and the last geom (object visual mesh)
  auto geom_vel = [&](int g, V3& lin, V3& ang) {
    const V3 pn = ld3(s + S_GPOS + 3 * g), po = ld3(og + 3 * g);
    lin = ((pn - po) + (org - oorg)) * inv_dt;
    const M3 Rd = mulmT(ldm(s + S_GMAT + 9 * g), ldm(og + HN * 3 + 9 * g));
    ang = matrix_to_axis_angle(Rd) * inv_dt;
  };
  V3 obj_v, obj_w;
  geom_vel(m.last_geom, obj_v, obj_w);
  if (has) {
    const int idx = __popc(bal & ((1u << lane) - 1u));
    for (int k = 0; k < 12; k++) s[QP_CF + 12 * idx + k] = fr[k];
    s[QP_TS + idx] = ts;
    ((int*)(s + QP_G1))[idx] = m.hand_geom0 + lane;
    V3 gl, ga;
    geom_vel(m.hand_geom0 + lane, gl, ga);
    st3(s + QP_GV + 6 * idx, gl); st3(s + QP_GV + 6 * idx + 3, ga);
  }
  __syncwarp();
  if (m.residual_force) {
    if (!m.explain_force) {
      rfc_score = norm(vf) + 1e4f * norm(vt);
    } else {
      const V3 obj_p = ld3(s + S_QPOS + m.nhand_q);
      const M3 R = q2m(qnormalize(ldq(s + S_QPOS + m.nhand_q + 3)));
      M3 RI;
      for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) RI.m[3 * i + j] = R.m[3 * i + j] * m.obj_inertia[j];
      const M3 I = mulmT(RI, R);
      const V3 F = (obj_a + v3(0.f, 0.f, 9.8f)) * m.obj_mass;
      const V3 T = mulv(I, obj_oa) + cross(obj_w, mulv(I, obj_w));
      if (ncg == 0) {
        rfc_score = norm(F) + 1e4f * norm(T);
        rest[0] = F.x; rest[1] = F.y; rest[2] = F.z; rest[3] = T.x; rest[4] = T.y; rest[5] = T.z;
      } else {
#ifdef H_NOQP   // timing experiment only (tools/ab.sh): what the residual-force QP costs in kernel wall time
        if (qp_iters > 1000000u)
#endif
        rfc_score = rfc_qp(s, lane, ncg, m.surface_contact ? 5 : 1, F, T, obj_p, obj_v, obj_w, s + S_GPOS, org, rest, qp_iters, surf);
      }
    }
  }
  return rfc_score;
}


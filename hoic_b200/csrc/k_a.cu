// k_a.cu -- fused pipeline kernel of env.step: the Newton solve of substep k (k_p3.cu: mj_fwdConstraint, reference call site
// ho_im4.py:545) followed AT ONCE, in the same warp, by k_p1's work for substep k + 1 (mj_Euler of substep k, stable PD,
// J^T wrenches, mj_fwdPosition / mj_fwdVelocity / mj_fwdAcceleration of the new state; p1_body.cuh).
// Why: with separate launches a part's substep costs max_env(solve) + max_env(k_p1) + max_env(k_p2); the solve takes 1 Newton
// iteration for half of the envs and 6-8 for the slowest one of a 2 048-env part (tools/newton_iter_stats.py), so k_p1 of
// 99 % of the envs waited for that one solve.  Fused, the launch costs max_env(solve + k_p1): the early finishers' smooth
// dynamics run underneath the long solves, and the working set of the solve (qacc, qfrc_constraint, M) stays in shared memory
// for the Euler step instead of travelling through L2.
// MEASURED SLOWER than the separate launches and therefore OFF by default (HOIC_FUSE=1 turns it on; GPU-tested for equality
// with the separate kernels): alone, whole batch of 4 096 box envs, k_as takes 264 us against 160 us (k_p3s) + 83 us (k_p1);
// a step takes 5.37 ms against 4.32 ms.  ncu (profiles/r02b_k_as_raw.csv): instruction-cache hit rate 73 % (k_p3s 93 %,
// k_p1 91 %), no-instruction stall 1.9 cycles per issued instruction (0.5), 16 resident warps per SM (20 / 22).  The warps of
// an SM are spread over 107 KB of SASS instead of 46 KB or 61 KB; kept as the evidence for the kernel split of DESIGN.md 4.
// Compiled twice like k_p3.cu: the SMALL tier (-DP3_SMALL: <= 8 contacts / 24 basis rows / 32 edges, most envs) and the full
// tier; each env is handled by exactly one of the two launches of a substep (side by side on two streams).
// The last substep's solve has no k_p1 after it (k_tail integrates it): for k + 1 >= sim_step the kernel ends like k_p3.
#define H_KERNEL HK_A
#ifdef P3_SMALL
#ifndef P3S_NCON      // (A/B builds pass -DP3S_NCON=.. -DP3S_NBASIS=.. -DP3S_NEDGE=..; the default lives in dev_model.h)
#define P3S_NCON 8
#define P3S_NBASIS 24
#define P3S_NEDGE 32
#endif
#define H_NCON P3S_NCON
#define H_NBASIS P3S_NBASIS
#define H_NEDGE P3S_NEDGE
#define K_A k_as
#define P3_TIER 0
#else
#define K_A k_a
#define P3_TIER 1
#endif
#include "p1_body.cuh"

#ifndef KA_WPS
#ifdef P3_SMALL
#define KA_WPS 16
#else
#define KA_WPS 10
#endif
#endif

__global__ void __launch_bounds__(32, KA_WPS) K_A(BatchDev B, StepPart sp, const float* __restrict__ actions, int k) {
  extern __shared__ __align__(16) float smem[];
  const int lane = threadIdx.x & 31;
  const EnvSlot es = env_of_slot(B, sp, blockIdx.x);
  const int env = es.env;
  if (!es.live) return;
  float* s = smem;
  float* g = B.scratch + (size_t)env * G_TOTAL;
  float* st = B.state + (size_t)env * ST_STRIDE;
  const DevModel& m = *B.env_model[env];
  int* misc = (int*)(s + S_MISC);
  if (k >= m.sim_step) return;
  const bool last = k + 1 >= m.sim_step;
  const bool rec = B.rec_frames != nullptr;
  // one load phase for both halves: the solve's inputs ...
  LD_FIELD(MISC); LD_FIELD(WARM); LD_FIELD(SMOOTH); LD_FIELD(ASMOOTH); LD_FIELD(MS);
#ifdef P3_SMALL
  LD_U(B, H_NBASIS * B_STRIDE);
  LD_U(AREF, H_NROW); LD_U(D, H_NROW); LD_U(RINFO, H_NROW);
  LD_U(BMASK, H_NBASIS); LD_U(MPRQ, H_NBASIS); LD_U(CINFO, H_NCON * 8);
#endif
  // ... and what k_p1 of the next substep reads besides the solve's results
  if (!last) {
    LD_FIELD(QPOS); LD_FIELD(QVEL); LD_FIELD(BIAS); LD_FIELD(COM); LD_FIELD(CDOF); LD_FIELD(ORG);
    if (rec) LD_FIELD(XPOS);
    const float a = lane < m.nu + (m.residual_force ? 6 : 0) ? actions[(size_t)env * HOIC_ACT_DIM + lane] : 0.f;
    s[S_ACT + lane] = fminf(fmaxf(a, -1.f), 1.f);
  }
  ld_wait();
  if (env_dead(misc)) return;
  const int nc = misc[MI_NCON], nb = misc[MI_NBASIS], ne = misc[MI_NEDGE], nrow = H_RCON + ((ne + 31) & ~31);
  if ((nc <= P3S_NCON && nb <= P3S_NBASIS && ne <= P3S_NEDGE) != (P3_TIER == 0)) return;    // the other tier's env
#ifndef P3_SMALL
  LD_U(B, nb * B_STRIDE);
  LD_U(AREF, nrow); LD_U(D, nrow); LD_U(RINFO, nrow);
  LD_U(BMASK, nb); LD_U(MPRQ, nb); LD_U(CINFO, nc * 8);
  ld_wait();
#endif
  Cnt cnt;
  memset(&cnt, 0, sizeof cnt);
  const bool solved = phase_solve(s, m, lane, cnt, true);
  if (lane == 0) {
    misc[MI_C_SUBSTEPS] += (int)cnt.substeps; misc[MI_C_NCON] += (int)cnt.ncon; misc[MI_C_NEFC] += (int)cnt.nefc;
    misc[MI_C_NEWTON] += (int)cnt.newton; misc[MI_C_LS] += (int)cnt.ls;
    misc[MI_C_MAXNCON] = max(misc[MI_C_MAXNCON], (int)cnt.max_ncon);
    misc[MI_C_MAXNBASIS] = max(misc[MI_C_MAXNBASIS], (int)cnt.max_nbasis);
    misc[MI_C_MAXNEDGE] = max(misc[MI_C_MAXNEDGE], (int)cnt.max_nedge);
  }
  __syncwarp();
  if (rec) { ST_U(FORCE, nrow); __syncwarp(); }     // record_forward (head of the k_p1 half / of k_tail) reads the forces from the working set
  if (last || !solved || env_dead(misc)) {
    // the last solve is integrated by k_tail; a failed env is skipped by every later kernel of the step
    ST_FIELD(MISC); ST_FIELD(QACC); ST_FIELD(FCON);
    return;
  }
  p1_body<true>(B, es, env, s, g, st, m, misc, actions, k + 1, lane, g + G_OFF(GPOS), g + G_OFF(GMAT));
}

static int g_ka_pad = -1;
#ifdef P3_SMALL
cudaError_t hoic_configure_as() {
  g_ka_pad = hoic_env_int("HOIC_KAS_PAD", 0);
#else
cudaError_t hoic_configure_a() {
  g_ka_pad = hoic_env_int("HOIC_KA_PAD", 0);
#endif
  return cudaFuncSetAttribute(K_A, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hoic_step_smem(1, g_ka_pad));
}
#ifdef P3_SMALL
cudaError_t hoic_launch_as(const BatchDev& B, StepPart sp, const float* actions, int k, cudaStream_t stream) {
#else
cudaError_t hoic_launch_a(const BatchDev& B, StepPart sp, const float* actions, int k, cudaStream_t stream) {
#endif
  K_A<<<hoic_step_grid(B, sp, 1), 32, hoic_step_smem(1, g_ka_pad < 0 ? 0 : g_ka_pad), stream>>>(B, sp, actions, k);
  return cudaGetLastError();
}

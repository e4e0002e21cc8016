// k_p1.cu -- pipeline kernel 1 of env.step: everything of a substep that costs the same for every environment.
//   k >= 1 : mj_Euler of substep k-1 (ho_im4.py:545, the tail of sim.step())
//   k >= 0 : stable-PD torque + gravity compensation / residual wrench through J^T with the PREVIOUS forward's M, bias and
//            Jacobians (ho_im4.py:412-501, 526-536), then mj_kinematics / comPos / crb, comVel / rne and
//            qacc_smooth of the new state (MuJoCo mj_fwdPosition / mj_fwdVelocity / mj_fwdAcceleration)
//   k = -1 : forward of the state that preceded the step (the "stale" quantities of S1 are recomputed, not stored)
// One warp per env; all warps do the same work, so the warps of a CTA stay in step without barriers.
#define H_KERNEL HK_P1
#include "p1_body.cuh"

// Compiled twice: one env per CTA (k_p1) and, with -DP1_WIDE, four envs per CTA (k_p1w).  Every CTA pays 1 KB of reserved
// shared memory, so the wide form holds 24 instead of 22 envs per SM: faster where the batch keeps every SM full (box,
// 16 384 envs: 13.33 -> 12.88 ms per step), slower at 4 096 envs (4.26 -> 4.36 ms: a part's kernel is about one wave there
// and a four-env CTA starts only when 38 KB are free at once).  hoic_step picks by the batch size.
#ifdef P1_WIDE
#define P1_W 4
#define P1_WPS 24
#define K_P1 k_p1w
#else
#define K_P1 k_p1
#endif
#ifndef P1_W
#define P1_W 1          // warps (= envs) per CTA
#endif
#ifndef P1_WPS
#define P1_WPS 23       // resident warps per SM the register budget is sized for
#endif

// GROUP (StepPart::group, a template parameter so that the default kernel carries none of the list-mode code): 0 = every env of the
// part; 1 / 2 = hoic_step's carry mode (everybody but the late envs / the late envs from their lists)
template <int GROUP>
__global__ void __launch_bounds__(32 * P1_W, P1_WPS / P1_W) K_P1(BatchDev B, StepPart sp, const float* __restrict__ actions, int k) {
  extern __shared__ __align__(16) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* s = smem + (size_t)warp * S_TOTAL;
  int* misc = (int*)(s + S_MISC);
  if (GROUP == 2) {
    // carry mode: the late envs of substep k (their solve of substep k - 1 was finished by k_p3c / the full tier), warp-strided
    const LateList l = late_list(B, sp, k);
    for (int i = blockIdx.x * P1_W + warp; i < l.n; i += gridDim.x * P1_W) {
      EnvSlot es;
      es.live = true;
      es.env = late_env(l, i);
      float* g = B.scratch + (size_t)es.env * G_TOTAL;
      const DevModel& m = *B.env_model[es.env];
      if (k < m.sim_step)
        p1_body<false>(B, es, es.env, s, g, B.state + (size_t)es.env * ST_STRIDE, m, misc, actions, k, lane, g + G_OFF(GPOS), g + G_OFF(GMAT));
      __syncwarp();
    }
    return;
  }
  const EnvSlot es = env_of_slot(B, sp, blockIdx.x * P1_W + warp);
  const int env = es.env;
  if (GROUP == 1 && k > 0 && es.live && B.late_stamp[env] == k) return;     // carry mode: this env comes with group 2
  float* g = B.scratch + (size_t)env * G_TOTAL;
  float* st = B.state + (size_t)env * ST_STRIDE;
  const DevModel& m = *B.env_model[env];
  if (k >= m.sim_step) return;
  p1_body<false>(B, es, env, s, g, st, m, misc, actions, k, lane, g + G_OFF(GPOS), g + G_OFF(GMAT));
}

static int g_p1_pad = -1;
#ifdef P1_WIDE
cudaError_t hoic_configure_p1w() {
#else
cudaError_t hoic_configure_p1() {
#endif
  g_p1_pad = hoic_env_int("HOIC_P1_PAD", 0);
  const size_t smem = hoic_step_smem(P1_W, g_p1_pad);
  cudaError_t e = cudaFuncSetAttribute(K_P1<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(K_P1<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(K_P1<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  return e;
}
#ifdef P1_WIDE
cudaError_t hoic_launch_p1w(const BatchDev& B, StepPart sp, const float* actions, int k, cudaStream_t stream) {
#else
cudaError_t hoic_launch_p1(const BatchDev& B, StepPart sp, const float* actions, int k, cudaStream_t stream) {
#endif
  const int grid = hoic_step_grid(B, sp, P1_W);
  const size_t smem = hoic_step_smem(P1_W, g_p1_pad < 0 ? 0 : g_p1_pad);
  if (sp.group == 2) K_P1<2><<<grid < 592 ? grid : 592, 32 * P1_W, smem, stream>>>(B, sp, actions, k);
  else if (sp.group == 1) K_P1<1><<<grid, 32 * P1_W, smem, stream>>>(B, sp, actions, k);
  else K_P1<0><<<grid, 32 * P1_W, smem, stream>>>(B, sp, actions, k);
  return cudaGetLastError();
}

// k_p12.cu -- k_p1's and k_p2's work of a substep in ONE launch (opt-in experiment, HOIC_FUSE12=1): every warp runs p1_body
// (Euler of the previous substep, stable PD, J^T wrenches, mj_fwdPosition / Velocity / Acceleration) and then, without leaving
// the SM, p2_body (mj_collision, record_contact, mj_makeConstraint) on the poses it just computed.  Unlike the solve, k_p1 costs
// the same for every env, so the warps of a CTA reach the collision code together (p2_body's CTA barriers still hold them
// there).  One launch and one wave tail less per substep, no round trip of the poses through L2.
// Shared memory: the fields k_p2 reads + the union at k_p2's size; the fields only k_p1's half touches live inside the union
// behind k_p1's own scratch and are dead (stored to the global working set) before the collision half overwrites them
// (dev_model.h: HK_P12) -- 14.2 KB per env, the same three 5-warp CTAs per SM as k_p2.
// MEASURED SLOWER and therefore off by default (bit-identical results, tests/test_gpu_parity2.py): box, 4 096 envs 4.383 vs
// 4.247 ms per step, 16 384 envs 14.53 vs 13.05 ms -- k_p1's half now runs at k_p2's 15 resident warps and 128 registers
// instead of 22 warps / 74 registers, which costs more than the launch and the wave tail it saves.
#define H_KERNEL HK_P12
#include "p1_body.cuh"
#include "p2_body.cuh"

static_assert(P12_P1_FIELDS >= UN_POS && P12_P1_FIELDS >= UN_TS + UN_TSF, "k_p12: k_p1's stages must fit below its private fields");
static_assert(S_OFF(MS) + 384 <= S_OFF(UNION) + UN_SIZE && S_OFF(WARM) >= S_OFF(UNION) + P12_P1_FIELDS, "k_p12: private fields inside the union");

__global__ void __launch_bounds__(32 * P2_W, P2_WPS / P2_W) k_p12(BatchDev B, StepPart sp, const float* __restrict__ actions, int k) {
  extern __shared__ __align__(16) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (sp.group == 3) {
    // tick mode: `k` is the tick; the envs that are due come from the tick's list, each with its own substep index
    const int t = k, n = *TK_COUNT(B, 0, t);
    const int* list = TK_LIST(B, 0, t);
    float* s = smem + (size_t)warp * S_TOTAL;
    int* misc = (int*)(s + S_MISC);
    for (int base = blockIdx.x * P2_W; base < n; base += gridDim.x * P2_W) {
      EnvSlot es;
      es.live = base + warp < n;
      es.env = list[es.live ? base + warp : base];
      const int env = es.env, kk = B.tk_k[env];
      float* g = B.scratch + (size_t)env * G_TOTAL;
      const DevModel& m = *B.env_model[env];
      bool fwd = false;
      if (kk < m.sim_step) fwd = p1_body<false>(B, es, env, s, g, B.state + (size_t)env * ST_STRIDE, m, misc, actions, kk, lane, s + S_GPOS, s + S_GMAT);
      __syncwarp();
      if (fwd && es.live) { ST_FIELD(GPOS); ST_FIELD(GMAT); }
      if (kk >= m.sim_step) misc[lane] = 0;
      __syncwarp();
      p2_body<true>(B, sp, es, env, s, g, m, misc, kk, lane);
      __syncwarp();
      if (es.live && lane == 0 && fwd && kk < m.sim_step && !env_dead(misc)) {
        // into the solver's list of the next tick, by capacity tier
        const int tier = (misc[MI_NCON] <= P3S_NCON && misc[MI_NBASIS] <= P3S_NBASIS && misc[MI_NEDGE] <= P3S_NEDGE) ? 1 : 2;
        TK_LIST(B, tier, t + 1)[atomicAdd(TK_COUNT(B, tier, t + 1), 1)] = env;
      }
      __syncthreads();
    }
    return;
  }
  const EnvSlot es = env_of_slot(B, sp, blockIdx.x * P2_W + warp);
  const int env = es.env;
  float* s = smem + (size_t)warp * S_TOTAL;
  float* g = B.scratch + (size_t)env * G_TOTAL;
  float* st = B.state + (size_t)env * ST_STRIDE;
  const DevModel& m = *B.env_model[env];
  int* misc = (int*)(s + S_MISC);
  bool fwd = false;
  if (k < m.sim_step) fwd = p1_body<false>(B, es, env, s, g, st, m, misc, actions, k, lane, s + S_GPOS, s + S_GMAT);
  __syncwarp();
  if (fwd && es.live) { ST_FIELD(GPOS); ST_FIELD(GMAT); }      // the step tail and the next substep's J^T wrench read them
  // (when k_p1's half ends early -- carried working set at k = -1, env failed earlier, slot beyond the batch -- MISC in shared
  // memory is what it loaded / wrote, and p2_body's `work` test reads exactly that)
  if (k >= m.sim_step) misc[lane] = 0;
  __syncwarp();
  p2_body<true>(B, sp, es, env, s, g, m, misc, k, lane);
}

static int g_p12_pad = -1;
cudaError_t hoic_configure_p12() {
  g_p12_pad = hoic_env_int("HOIC_P12_PAD", 0);
  return cudaFuncSetAttribute(k_p12, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hoic_step_smem(P2_W, g_p12_pad));
}
cudaError_t hoic_launch_p12(const BatchDev& B, StepPart sp, const float* actions, int k, cudaStream_t stream) {
  int grid = hoic_step_grid(B, sp, P2_W);
  if (sp.group == 3) grid = grid < 444 ? grid : 444;      // tick mode: one resident wave, CTA-strided over the tick's list
  k_p12<<<grid, 32 * P2_W, hoic_step_smem(P2_W, g_p12_pad < 0 ? 0 : g_p12_pad), stream>>>(B, sp, actions, k);
  return cudaGetLastError();
}

// k_p3.cu -- pipeline kernel 3 of env.step: mj_fwdConstraint, the Newton solver of MuJoCo 2.1's soft-constraint model
// (dev_solver.cuh; reference call site ho_im4.py:545).  The number of Newton iterations differs from env to env
// (1: 54 %, 2: 18 %, 3: 18 %, 4+: 10 % of the solves), so every env is its own CTA and the CTA scheduler keeps the SMs
// full while long solves finish.
// Compiled twice: the SMALL tier (-DP3_SMALL) holds a constraint stage of 8 contacts / 24 basis rows / 32 pyramid edges
// in shared memory (10 KB per env: 21 resident warps per SM) and solves the envs that fit, which is most of them; the
// full tier (24 / 56 / 96, 16 KB per env) takes the rest.  Each env is solved by exactly one of the two launches.
#define H_KERNEL HK_P3
#ifdef P3_SMALL
#ifndef P3S_NCON      // (A/B builds pass -DP3S_NCON=.. -DP3S_NBASIS=.. -DP3S_NEDGE=..; the default lives in dev_model.h)
#define P3S_NCON 8
#define P3S_NBASIS 24
#define P3S_NEDGE 32
#endif
#define H_NCON P3S_NCON
#define H_NBASIS P3S_NBASIS
#define H_NEDGE P3S_NEDGE
#if defined(P3_CONT)   // third build: k_p3c continues the solves a capped launch left open (hoic_step "carry" mode) and is the
#define K_P3 k_p3c    // small tier's budgeted solver of the "tick" mode
#elif defined(P3_BUDGET)   // fifth build: k_p3sb, the capped launch of the carry mode (k_p3s itself carries no budget code)
#define K_P3 k_p3sb
#else
#define K_P3 k_p3s
#endif
#define P3_TIER 0
#else
#ifdef P3_CONT         // fourth build: the full tier's budgeted solver of the tick mode
#define K_P3 k_p3fc
#else
#define K_P3 k_p3
#endif
#define P3_TIER 1
#endif
#include "hoic_step.cuh"

#ifndef P3_W
#define P3_W 1
#endif
#ifndef P3_LIST_GRID
#define P3_LIST_GRID 592     // CTAs of a full-tier launch over k_p2's env list (4 per SM; more envs than that are grid-strided)
#endif
#ifndef P3_WPS
#ifdef P3_SMALL
#define P3_WPS 21
#else
#define P3_WPS 14
#endif
#endif

// the solve of one env (one warp); s = the warp's shared-memory region
__device__ __forceinline__ void p3_env(const BatchDev& B, const StepPart sp, const int env, const int k, float* s, const int lane) {
  float* g = B.scratch + (size_t)env * G_TOTAL;
  const DevModel& m = *B.env_model[env];
  int* misc = (int*)(s + S_MISC);
  if (k >= m.sim_step) return;
  LD_FIELD(MISC); LD_FIELD(WARM); LD_FIELD(SMOOTH); LD_FIELD(ASMOOTH); LD_FIELD(MS);
#ifdef P3_SMALL
  // the small tier's whole constraint stage is 3 KB: fetch it at capacity in the same round trip as the counts
  LD_U(B, H_NBASIS * B_STRIDE);
  LD_U(AREF, H_NROW); LD_U(D, H_NROW); LD_U(RINFO, H_NROW);
  LD_U(BMASK, H_NBASIS); LD_U(MPRQ, H_NBASIS); LD_U(CINFO, H_NCON * 8);
#endif
#ifdef P3_CONT
  LD_FIELD(QACC); LD_FIELD(MA); LD_U(JAR, H_NROW);      // the open solve's point
#endif
  ld_wait();
  if (env_dead(misc)) return;
  const int nc = misc[MI_NCON], nb = misc[MI_NBASIS], ne = misc[MI_NEDGE], nrow = H_RCON + ((ne + 31) & ~31);
  if ((nc <= P3S_NCON && nb <= P3S_NBASIS && ne <= P3S_NEDGE) != (P3_TIER == 0)) return;    // the other tier's env
#ifdef P3_CONT
  if (!misc[MI_SV_OPEN]) return;
  if (lane == 0) misc[MI_SV_OPEN] = 0;
  __syncwarp();
#endif
#ifndef P3_SMALL
  LD_U(B, nb * B_STRIDE);
  LD_U(AREF, nrow); LD_U(D, nrow); LD_U(RINFO, nrow);
  LD_U(BMASK, nb); LD_U(MPRQ, nb); LD_U(CINFO, nc * 8);
  ld_wait();
#endif
  Cnt cnt;
  memset(&cnt, 0, sizeof cnt);
#ifdef H_SOLVE_STATS
  const long long t_solve = clock64();
#endif
  bool open = false;
#if defined(P3_CONT)
  phase_solve<true>(s, m, lane, cnt, true, 0, true, nullptr);
#elif defined(P3_BUDGET)
  phase_solve<true>(s, m, lane, cnt, true, B.solve_budget, false, &open);
#else
  phase_solve(s, m, lane, cnt, true);
#endif
#ifdef H_SOLVE_STATS
  if (lane == 0) {     // analysis build (tools/solve_stats.py): solves, cycles and line-search evaluations by iteration count
    const int it = min((int)cnt.newton, 15);
    atomicAdd(&g_solve_stats[it], 1ull);
    atomicAdd(&g_solve_stats[16 + it], (unsigned long long)(clock64() - t_solve));
    atomicAdd(&g_solve_stats[32 + it], (unsigned long long)cnt.ls);
  }
#endif
  if (lane == 0) {
    misc[MI_C_SUBSTEPS] += (int)cnt.substeps; misc[MI_C_NCON] += (int)cnt.ncon; misc[MI_C_NEFC] += (int)cnt.nefc;
    misc[MI_C_NEWTON] += (int)cnt.newton; misc[MI_C_LS] += (int)cnt.ls;
    misc[MI_C_MAXNCON] = max(misc[MI_C_MAXNCON], (int)cnt.max_ncon);
    misc[MI_C_MAXNBASIS] = max(misc[MI_C_MAXNBASIS], (int)cnt.max_nbasis);
    misc[MI_C_MAXNEDGE] = max(misc[MI_C_MAXNEDGE], (int)cnt.max_nedge);
  }
  __syncwarp();
#if defined(P3_BUDGET) && !defined(P3_CONT)
  if (open && !env_dead(misc)) {
    // out of budget: the solve's point goes to the working set, the env to the carry list of this (part, substep); k_p1 / k_p2 of
    // the next substep skip it (late_stamp) until k_p3c has finished the solve
    ST_FIELD(MISC); ST_FIELD(QACC); ST_FIELD(MA); ST_U(JAR, nrow);
    if (lane == 0) {
      B.late_stamp[env] = k + 1;
      B.carry_list[(size_t)sp.part * B.n_envs + atomicAdd(B.carry_count + sp.part * FULL_COUNT_STRIDE + k, 1)] = env;
    }
    return;
  }
#endif
  ST_FIELD(MISC); ST_FIELD(QACC); ST_FIELD(FCON);
  if (B.rec_frames) ST_U(FORCE, nrow);
}

#ifdef P3_CONT
// tick mode: one budgeted round of the solve of env `env` (fresh, or the continuation of a solve an earlier tick left open);
// afterwards the env goes to the next tick's list of whoever comes next: this solver tier again, or k_p12 of the next substep
__device__ __forceinline__ void p3_env_tick(const BatchDev& B, const int env, const int t, float* s, const int lane) {
  float* g = B.scratch + (size_t)env * G_TOTAL;
  const DevModel& m = *B.env_model[env];
  int* misc = (int*)(s + S_MISC);
  const int k = B.tk_k[env];
  if (k >= m.sim_step) return;
  LD_FIELD(MISC); LD_FIELD(WARM); LD_FIELD(SMOOTH); LD_FIELD(ASMOOTH); LD_FIELD(MS); LD_FIELD(QACC); LD_FIELD(MA);
#ifdef P3_SMALL
  LD_U(B, H_NBASIS * B_STRIDE);
  LD_U(AREF, H_NROW); LD_U(D, H_NROW); LD_U(RINFO, H_NROW);
  LD_U(BMASK, H_NBASIS); LD_U(MPRQ, H_NBASIS); LD_U(CINFO, H_NCON * 8); LD_U(JAR, H_NROW);
#endif
  ld_wait();
  if (env_dead(misc)) return;
  const int nc = misc[MI_NCON], nb = misc[MI_NBASIS], ne = misc[MI_NEDGE], nrow = H_RCON + ((ne + 31) & ~31);
#ifndef P3_SMALL
  LD_U(B, nb * B_STRIDE);
  LD_U(AREF, nrow); LD_U(D, nrow); LD_U(RINFO, nrow);
  LD_U(BMASK, nb); LD_U(MPRQ, nb); LD_U(CINFO, nc * 8); LD_U(JAR, nrow);
  ld_wait();
#endif
  const bool resume = misc[MI_SV_OPEN] != 0;
  __syncwarp();
  if (lane == 0) misc[MI_SV_OPEN] = 0;
  __syncwarp();
  Cnt cnt;
  memset(&cnt, 0, sizeof cnt);
  bool open = false;
  phase_solve<true>(s, m, lane, cnt, true, B.tk_budget, resume, &open);
  if (lane == 0) {
    misc[MI_C_SUBSTEPS] += (int)cnt.substeps; misc[MI_C_NCON] += (int)cnt.ncon; misc[MI_C_NEFC] += (int)cnt.nefc;
    misc[MI_C_NEWTON] += (int)cnt.newton; misc[MI_C_LS] += (int)cnt.ls;
    misc[MI_C_MAXNCON] = max(misc[MI_C_MAXNCON], (int)cnt.max_ncon);
    misc[MI_C_MAXNBASIS] = max(misc[MI_C_MAXNBASIS], (int)cnt.max_nbasis);
    misc[MI_C_MAXNEDGE] = max(misc[MI_C_MAXNEDGE], (int)cnt.max_nedge);
  }
  __syncwarp();
  if (open && !env_dead(misc)) {
    ST_FIELD(MISC); ST_FIELD(QACC); ST_FIELD(MA); ST_U(JAR, nrow);
    if (lane == 0) TK_LIST(B, 1 + P3_TIER, t + 1)[atomicAdd(TK_COUNT(B, 1 + P3_TIER, t + 1), 1)] = env;
    return;
  }
  ST_FIELD(MISC); ST_FIELD(QACC); ST_FIELD(FCON);
  if (B.rec_frames) ST_U(FORCE, nrow);
  if (lane == 0) {
    B.tk_k[env] = k + 1;
    if (k + 1 < m.sim_step && !env_dead(misc)) TK_LIST(B, 0, t + 1)[atomicAdd(TK_COUNT(B, 0, t + 1), 1)] = env;
  }
}
#endif

__global__ void __launch_bounds__(32 * P3_W, P3_WPS / P3_W) K_P3(BatchDev B, StepPart sp, int k) {
  extern __shared__ __align__(16) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* s = smem + (size_t)warp * S_TOTAL;
#ifdef P3_CONT
  if (sp.group == 3) {
    const int t = k, n = *TK_COUNT(B, 1 + P3_TIER, t);
    const int* list = TK_LIST(B, 1 + P3_TIER, t);
    for (int i = blockIdx.x * P3_W + warp; i < n; i += gridDim.x * P3_W) {
      p3_env_tick(B, list[i], t, s, lane);
      __syncwarp();
    }
    return;
  }
#endif
#ifndef P3_SMALL
  if (B.full_list) {
    // full tier: only the envs k_p2 listed for this substep (a few per cent of the batch), grid-strided
    const int n = B.full_count[sp.part * FULL_COUNT_STRIDE + k];
    const int* list = full_list_of(B, sp.part, k);
    for (int i = blockIdx.x * P3_W + warp; i < n; i += gridDim.x * P3_W) {
      p3_env(B, sp, list[i], k, s, lane);
      __syncwarp();
    }
    return;
  }
#endif
#if defined(P3_CONT) && defined(P3_SMALL)
  {
    const int n = B.carry_count[sp.part * FULL_COUNT_STRIDE + k];
    const int* list = B.carry_list + (size_t)sp.part * B.n_envs;
    for (int i = blockIdx.x * P3_W + warp; i < n; i += gridDim.x * P3_W) {
      p3_env(B, sp, list[i], k, s, lane);
      __syncwarp();
    }
    return;
  }
#endif
  const EnvSlot es = env_of_slot(B, sp, blockIdx.x * P3_W + warp);
  if (!es.live) return;
  p3_env(B, sp, es.env, k, s, lane);
}

static int g_p3_pad = -1;
#if defined(P3_CONT) && !defined(P3_SMALL)
cudaError_t hoic_configure_p3fc() {
  g_p3_pad = hoic_env_int("HOIC_P3_PAD", 0);
#elif defined(P3_CONT)
cudaError_t hoic_configure_p3c() {
  g_p3_pad = hoic_env_int("HOIC_P3S_PAD", 0);
#elif defined(P3_BUDGET)
cudaError_t hoic_configure_p3sb() {
  g_p3_pad = hoic_env_int("HOIC_P3S_PAD", 0);
#elif defined(P3_SMALL)
cudaError_t hoic_configure_p3s() {
  g_p3_pad = hoic_env_int("HOIC_P3S_PAD", 0);
#else
cudaError_t hoic_configure_p3() {
  g_p3_pad = hoic_env_int("HOIC_P3_PAD", 0);
#endif
  return cudaFuncSetAttribute(K_P3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hoic_step_smem(P3_W, g_p3_pad));
}
#if defined(P3_CONT) && !defined(P3_SMALL)
cudaError_t hoic_launch_p3fc(const BatchDev& B, StepPart sp, int k, cudaStream_t stream) {
#elif defined(P3_CONT)
cudaError_t hoic_launch_p3c(const BatchDev& B, StepPart sp, int k, cudaStream_t stream) {
#elif defined(P3_BUDGET)
cudaError_t hoic_launch_p3sb(const BatchDev& B, StepPart sp, int k, cudaStream_t stream) {
#elif defined(P3_SMALL)
cudaError_t hoic_launch_p3s(const BatchDev& B, StepPart sp, int k, cudaStream_t stream) {
#else
cudaError_t hoic_launch_p3(const BatchDev& B, StepPart sp, int k, cudaStream_t stream) {
#endif
  int grid = hoic_step_grid(B, sp, P3_W);
#ifndef P3_SMALL
  if (B.full_list) grid = grid < P3_LIST_GRID ? grid : P3_LIST_GRID;
#endif
#ifdef P3_CONT
#ifdef P3_SMALL
  if (sp.group == 3) grid = grid < 2960 ? grid : 2960;      // tick mode: up to one resident wave, warp-strided over the tick's list
#else
  if (sp.group == 3) grid = grid < 148 ? grid : 148;        // (the full tier's list is a handful of envs)
#endif
  else grid = grid < P3_LIST_GRID ? grid : P3_LIST_GRID;
#endif
  K_P3<<<grid, 32 * P3_W, hoic_step_smem(P3_W, g_p3_pad < 0 ? 0 : g_p3_pad), stream>>>(B, sp, k);
  return cudaGetLastError();
}

#if defined(H_SOLVE_STATS) && !defined(P3_CONT) && !defined(P3_BUDGET)
#ifdef P3_SMALL
extern "C" int hoic_debug_solve_stats_small(unsigned long long* out64, int reset) {
#else
extern "C" int hoic_debug_solve_stats_full(unsigned long long* out64, int reset) {
#endif
  if (cudaMemcpyFromSymbol(out64, g_solve_stats, sizeof(unsigned long long) * 64) != cudaSuccess) return -1;
  if (reset) { unsigned long long z[64] = {0}; cudaMemcpyToSymbol(g_solve_stats, z, sizeof z); }
  return 0;
}
#endif

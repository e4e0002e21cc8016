// k_p2.cu -- pipeline kernel 2 of env.step: mj_collision (analytic narrow phase + warp-cooperative MPR for meshes),
// record_contact (ho_im4.py:883-889; the contacts of the forward BEFORE sim.step()) and mj_makeConstraint.
// k = -1: contacts of the forward that preceded the step + the geom-pose snapshot of ho_im4.py:508-511.
// The work per env depends on its contacts, so every env is its own CTA: the CTA scheduler balances the load.
#define H_KERNEL HK_P2
#ifdef P2_CAP_NCON      // experiment (tools/build_variant.sh): constraint stage of k_p2 at a smaller capacity than the working set's
#define H_NCON P2_CAP_NCON
#define H_NBASIS P2_CAP_NBASIS
#define H_NEDGE P2_CAP_NEDGE
#endif
#include "p2_body.cuh"

// LIST (a template parameter so that the default kernel carries none of the list-mode code): hoic_step's carry mode, the late
// envs of substep k, CTA-strided so that the warps of a CTA keep meeting at p2_body's barriers
template <bool LIST>
__global__ void __launch_bounds__(32 * P2_W, P2_WPS / P2_W) k_p2(BatchDev B, StepPart sp, int k) {
  extern __shared__ __align__(16) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* s = smem + (size_t)warp * S_TOTAL;
  int* misc = (int*)(s + S_MISC);
  if (LIST) {
    const LateList l = late_list(B, sp, k);
    for (int base = blockIdx.x * P2_W; base < l.n; base += gridDim.x * P2_W) {
      EnvSlot es;
      es.live = base + warp < l.n;
      es.env = late_env(l, es.live ? base + warp : base);
      p2_body<false>(B, sp, es, es.env, s, B.scratch + (size_t)es.env * G_TOTAL, *B.env_model[es.env], misc, k, lane);
      __syncthreads();
    }
    return;
  }
  const EnvSlot es = env_of_slot(B, sp, blockIdx.x * P2_W + warp);
  const int env = es.env;
  float* g = B.scratch + (size_t)env * G_TOTAL;
  const DevModel& m = *B.env_model[env];
  p2_body<false>(B, sp, es, env, s, g, m, misc, k, lane);
}

static int g_p2_pad = -1;
cudaError_t hoic_configure_p2() {
  g_p2_pad = hoic_env_int("HOIC_P2_PAD", 0);
  cudaError_t e = cudaFuncSetAttribute(k_p2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hoic_step_smem(P2_W, g_p2_pad));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_p2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hoic_step_smem(P2_W, g_p2_pad));
  return e;
}
cudaError_t hoic_launch_p2(const BatchDev& B, StepPart sp, int k, cudaStream_t stream) {
  const int grid = hoic_step_grid(B, sp, P2_W);
  const size_t smem = hoic_step_smem(P2_W, g_p2_pad < 0 ? 0 : g_p2_pad);
  // (list mode: a few per cent of the part, one 5-env CTA per SM, strided beyond that)
  if (sp.group == 2) k_p2<true><<<grid < 148 ? grid : 148, 32 * P2_W, smem, stream>>>(B, sp, k);
  else k_p2<false><<<grid, 32 * P2_W, smem, stream>>>(B, sp, k);
  return cudaGetLastError();
}

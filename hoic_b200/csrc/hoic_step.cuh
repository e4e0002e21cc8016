// hoic_step.cuh -- helpers shared by the kernels of the env-step pipeline.
// One env.step = k_p1(-1) k_p2(-1) [k_p1(k) k_p2(k) k_p3(k)] x sim_step  k_tail, one warp per environment in every kernel.
// The working set of an env travels between the kernels through BatchDev.scratch (global memory, L2-resident at the
// batch sizes of the benchmark) in the FULL layout; each kernel copies the fields it needs into its own compact
// shared-memory layout, so that every kernel runs at its own occupancy, holds only its own code in the instruction cache
// and is load-balanced by the CTA scheduler (the Newton solve and collision take a data-dependent time per env).
#pragma once
#include "dev_batch.cuh"
#include "hoic_step.h"

// n floats, n % 4 == 0, both sides 16-byte aligned
__device__ __forceinline__ void cpf(float* dst, const float* src, int n, int lane) {
  const float4* s4 = reinterpret_cast<const float4*>(src);
  float4* d4 = reinterpret_cast<float4*>(dst);
  for (int i = lane; i < (n >> 2); i += 32) d4[i] = s4[i];
}
// global -> shared without a round trip through registers (LDGSTS): every copy of a kernel's load phase is in flight at
// once; ld_wait() before the first read.  The working set was written by an earlier kernel, so bypassing L1 (.cg) is safe
// and keeps L1 for the model constants.
__device__ __forceinline__ void ldf(float* dst_smem, const float* src, int n, int lane) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
  for (int i = lane; i < (n >> 2); i += 32)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + 16u * i), "l"(src + 4 * i) : "memory");
}
__device__ __forceinline__ void ld_wait() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
  __syncwarp();
}
#define LD_FIELD(name) ldf(s + S_OFF(name), g + G_OFF(name), F_SIZE(name), lane)
#define ST_FIELD(name) cpf(g + G_OFF(name), s + S_OFF(name), F_SIZE(name), lane)
// first n floats of a constraint-stage array (name as in UL_*: B, AREF, D, ...)
#define LD_U(f, n) ldf(s + U_S(f), g + U_G(f), ((n) + 3) & ~3, lane)
#define ST_U(f, n) cpf(g + U_G(f), s + U_S(f), ((n) + 3) & ~3, lane)

struct EnvSlot {
  int env;
  bool live;
};
__device__ __forceinline__ EnvSlot env_of_slot(const BatchDev& B, StepPart sp, int idx) {
  const int slot = idx * sp.nparts + sp.part;
  EnvSlot e;
  e.live = slot < B.n_envs;
  e.env = e.live ? (B.perm ? B.perm[slot] : slot) : B.n_envs - 1;
  return e;
}
// full-tier list of (part, substep k): double-buffered by the parity of k (carry mode reads substep k - 1's list while k_p2 of
// substep k appends to the other one)
__device__ __forceinline__ int* full_list_of(const BatchDev& B, int part, int k) {
  return B.full_list + ((size_t)(k & 1) * HOIC_LIST_PARTS + part) * B.n_envs;
}
// the envs of group 2 at substep k (see StepPart): carried solves first, then the full tier's envs, both of substep k - 1
struct LateList { const int* carry; const int* full; int n1, n; };
__device__ __forceinline__ LateList late_list(const BatchDev& B, StepPart sp, int k) {
  LateList l;
  l.carry = B.carry_list + (size_t)sp.part * B.n_envs;
  l.full = full_list_of(B, sp.part, k - 1);
  l.n1 = B.carry_count[sp.part * FULL_COUNT_STRIDE + k - 1];
  l.n = l.n1 + B.full_count[sp.part * FULL_COUNT_STRIDE + k - 1];
  return l;
}
__device__ __forceinline__ int late_env(const LateList& l, int i) { return i < l.n1 ? l.carry[i] : l.full[i - l.n1]; }
// env failed earlier in this step (physics failure or capacity overflow): the remaining kernels skip it
__device__ __forceinline__ bool env_dead(const int* misc) { return (misc[MI_FLAGS] & 3) != 0; }

// dynamic shared memory per CTA: `pad` extra bytes lower the number of resident CTAs (leaves more L1)
inline int hoic_step_grid(const BatchDev& B, StepPart sp, int warps) {
  const int n = (B.n_envs - sp.part + sp.nparts - 1) / sp.nparts;     // slots of this part
  return (n + warps - 1) / warps;
}
inline size_t hoic_step_smem(int warps, int pad_bytes) { return (size_t)warps * S_TOTAL * sizeof(float) + (size_t)pad_bytes; }
inline int hoic_env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v && *v ? atoi(v) : dflt;
}

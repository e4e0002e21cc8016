// p2_body.cuh -- the work of pipeline kernel 2 (see k_p2.cu) for one env / warp of a P2_W-warp CTA, shared by k_p2 and the fused
// kernel k_p12 (k_p12.cu).  LOADED: MISC, ORG, GPOS, GMAT (k >= 0: QPOS, QVEL, COM, CDOF too) are already in shared memory.
// EVERY warp of the CTA must call it (CTA barriers inside).
#pragma once
#include "hoic_step.cuh"

#ifndef P2_W
#define P2_W 5
#endif
#ifndef P2_WPS
#define P2_WPS 15
#endif
// P2_W > 1: the warps of a CTA start collision together and meet again before the constraint rows, so that they fetch the
// same (large) code at the same time; the kernel is instruction-fetch bound otherwise (profiles/README.md, r02).
#if P2_W > 1
#define P2_SYNC() __syncthreads()
#else
#define P2_SYNC() do { } while (0)
#endif

template <bool LOADED>
__device__ __forceinline__ void p2_body(const BatchDev& B, const StepPart sp, const EnvSlot es, const int env, float* s, float* g,
                                        const DevModel& m, int* misc, const int k, const int lane) {
  if (!LOADED) {
    LD_FIELD(MISC); LD_FIELD(ORG); LD_FIELD(GPOS); LD_FIELD(GMAT);
    if (k >= 0) { LD_FIELD(QPOS); LD_FIELD(QVEL); LD_FIELD(COM); LD_FIELD(CDOF); }
    ld_wait();
  }
  // work = this slot holds an env that is still running this step (k = -1: whose previous state had a forward)
  // (k = -1 is skipped for an env whose working set was carried over from the previous step, see k_p1)
  const bool late = sp.group == 1 && k > 0 && es.live && B.late_stamp[env] == k;     // carry mode: this env comes with group 2
  const bool work = es.live && !late && k < m.sim_step && (k < 0 ? !(misc[MI_FLAGS] & 1) && !misc[MI_CARRIED] : !env_dead(misc));
  float* csum = B.csum + (size_t)env * CSUM_STRIDE;
  unsigned c0 = 0, c1 = 0;
  P2_SYNC();
  if (work) {
    collide(s, m, lane, c0, c1);
    if (k + 1 < m.sim_step) record_contacts(s, m, lane, csum);
    else {
      // the last forward's contacts are what the NEXT env.step records first (record_contact runs before sim.step(), S2)
      float* carry = B.csum_carry + (size_t)env * CSUM_STRIDE;
      for (int i = lane; i < CSUM_STRIDE; i += 32) carry[i] = 0.f;
      __syncwarp();
      record_contacts(s, m, lane, carry);
    }
    if (lane == 0) { misc[MI_C_MPRCALLS] += (int)c0; misc[MI_C_MPRITERS] += (int)c1; }
  }
  if (k < 0) {
    if (!es.live || k >= m.sim_step || misc[MI_CARRIED]) return;
    if (work) {
      // snapshot of the geom poses for the averaged velocities (ho_im4.py:508-511)
      float* og = B.old_geom + (size_t)env * (HN * 12 + 4);
      for (int i = lane; i < m.ngeom * 3; i += 32) og[i] = s[S_GPOS + i];
      for (int i = lane; i < m.ngeom * 9; i += 32) og[HN * 3 + i] = s[S_GMAT + i];
      if (lane < 3) og[HN * 12 + lane] = s[S_ORG + lane];
    }
    __syncwarp();
    if (lane == 0) misc[MI_FLAGS] = 0;     // a bad stale state was reported by the previous step already
    __syncwarp();
    ST_FIELD(MISC);
    return;
  }
#ifndef P2_NO_MIDSYNC
  P2_SYNC();
#endif
  if (!work) return;
  build_rows(s, m, lane);
  if (B.rec_frames) record_contact_frames(s, m, lane, B.rec_frames + ((size_t)env * m.sim_step + k) * HOIC_REC_STRIDE);
  __syncwarp();
  // the used part of the constraint stage -> global working set (input of the Newton solve)
  const int nc = misc[MI_NCON], nb = misc[MI_NBASIS], nrow = H_RCON + ((misc[MI_NEDGE] + 31) & ~31);
  if (lane == 0 && sp.group != 3 && B.full_list && !(nc <= P3S_NCON && nb <= P3S_NBASIS && misc[MI_NEDGE] <= P3S_NEDGE) && !env_dead(misc)) {
    full_list_of(B, sp.part, k)[atomicAdd(B.full_count + sp.part * FULL_COUNT_STRIDE + k, 1)] = env;
    if (B.solve_budget > 0) B.late_stamp[env] = k + 1;      // carry mode: the full tier's envs take k_p1 / k_p2 of the next substep late, too
  }
  ST_FIELD(MISC);
  ST_U(B, nb * B_STRIDE);
  ST_U(AREF, nrow); ST_U(D, nrow); ST_U(RINFO, nrow);
  ST_U(BMASK, nb); ST_U(MPRQ, nb); ST_U(CINFO, nc * 8);      // (U_BINFO lives in the MPR queue's slots)
  if (B.rec_frames) ST_U(CON, nc * 16);
}

// hoic_capi.cu -- the C ABI of libhoic_b200.so (include/hoic.h) and its single-warp kernels (reset, state access, the
// single-mj_step parity entries, dataset preprocessing, standalone wrench QP).  env.step itself is the kernel pipeline
// of hoic_step.cuh (k_p1.cu / k_p2.cu / k_p3.cu / k_tail.cu), one warp per environment throughout.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <algorithm>
#include <utility>
#include <vector>

#include "../../include/hoic.h"
#include "../../include/hoic_blob.h"
#include "dev_batch.cuh"
#include "hoic_step.h"

// ================================================================================================ device side

// ------------------------------------------------------------------------------------------------ reset kernel
// counter-based uniform in [0, 1): splitmix64 of (seed, env, episode counter, draw)
__device__ __forceinline__ float hoic_uniform(unsigned long long seed, unsigned env, unsigned episode, unsigned draw) {
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (((unsigned long long)env << 32 | episode) * 4ull + draw + 1ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (float)(z >> 40) * (1.0f / 16777216.0f);
}

// `sample` != 0: draw (motion, start_idx) on the device as the sampler does between episodes (agent_handmimic.py:437-448):
// a uniform motion of the env's model, start ~ U{0 .. len - min_episode - 1} (0 without random_start).
static __device__ void reset_env(const BatchDev& B, const int env, const int* __restrict__ motion_id, const int* __restrict__ start_idx,
                                 float* obs_out, const int sample, float* s, const int lane) {
  const DevModel& m = *B.env_model[env];
  float* st = B.state + (size_t)env * ST_STRIDE;
  int* ints = (int*)(st + ST_INTS);
  int* misc = (int*)(s + S_MISC);
  misc[lane] = 0;
  int mid = ints[1], sidx = ints[2];
  const int episode = ints[6];
  if (sample) {
    const int mi = B.env_model_idx[env], lo = B.samp_off[mi], cnt = B.samp_off[mi + 1] - lo;
    mid = B.samp_ids[lo + min((int)(hoic_uniform(B.samp_seed, env, episode, 0) * cnt), cnt - 1)];
    sidx = 0;
    if (B.samp_random_start) {
      const int span = max(B.motion_len[mid] - B.samp_min_episode, 1);
      sidx = min((int)(hoic_uniform(B.samp_seed, env, episode, 1) * span), span - 1);
    }
  } else {
    if (motion_id) mid = min(max(motion_id[env], 0), B.n_motions - 1);
    if (start_idx) sidx = start_idx[env];
  }
  const int mlen = B.motion_len[mid];
  sidx = min(max(sidx, 0), mlen - 1);
  __syncwarp();
  if (lane == 0) { ints[0] = 0; ints[1] = mid; ints[2] = sidx; ints[3] = mlen - sidx; ints[6] = episode + 1; ints[7] = 0; }
  EnvCtx ctx;
  ctx.motion = B.frames + (size_t)(B.motion_offset[mid] + sidx) * H_FRAME;
  ctx.expert_len = mlen - sidx; ctx.cur_t = 0; ctx.start_ind = 0;
  const float* f = ctx.motion;
  const int nh = m.nhand_q, nhv = m.nhand_v;
  // reset_model (ho_im4.py:690-716): frame 0
  if (lane < nh) s[S_QPOS + lane] = f[F_DOF + lane];
  else if (lane < 32) s[S_QPOS + lane] = f[F_OBJPOSE + lane - nh];
  if (lane == 0) s[S_QPOS + 32] = f[F_OBJPOSE + 32 - nh];
  if (lane < nhv) s[S_QVEL + lane] = f[F_DOFVEL + lane];
  else if (lane < nhv + 3) s[S_QVEL + lane] = f[F_OBJVEL + lane - nhv];
  else s[S_QVEL + lane] = f[F_OBJANGVEL + lane - nhv - 3];
  __syncwarp();
  pos_stage(s, m, lane, s + S_GPOS, s + S_GMAT);   // sim.forward() of set_state: poses for the observation (and normalised quaternion)
  if (obs_out) write_obs(s, m, lane, ctx, obs_out + (size_t)env * HOIC_OBS_DIM);
  st[ST_QPOS + lane] = s[S_QPOS + lane];
  st[ST_QVEL + lane] = s[S_QVEL + lane];
  st[ST_WARM + lane] = 0.f;   // mj_resetData
  st[ST_PQPOS + lane] = s[S_QPOS + lane];
  st[ST_PQVEL + lane] = s[S_QVEL + lane];
  if (lane == 0) { st[ST_QPOS + 32] = s[S_QPOS + 32]; st[ST_PQPOS + 32] = s[S_QPOS + 32]; }
  if (lane < 8) st[ST_RFC + lane] = 0.f;
}
// use_list != 0: the envs k_tail listed as finished (B.reset_list / reset_count), grid-strided; otherwise env = blockIdx.x under `mask`
__global__ void __launch_bounds__(32) k_reset(BatchDev B, const uint8_t* __restrict__ mask, const int* __restrict__ motion_id,
                                              const int* __restrict__ start_idx, float* obs_out, int sample, int use_list) {
  extern __shared__ float s[];
  const int lane = threadIdx.x;
  if (use_list) {
    const int n = *B.reset_count;
    for (int i = blockIdx.x; i < n; i += gridDim.x) {
      reset_env(B, B.reset_list[i], motion_id, start_idx, obs_out, sample, s, lane);
      __syncwarp();
    }
    return;
  }
  const int env = blockIdx.x;
  if (env >= B.n_envs) return;
  if (mask && !mask[env]) return;
  reset_env(B, env, motion_id, start_idx, obs_out, sample, s, lane);
}

// Longest-processing-time env order for the step pipeline: a counting sort of the envs by the cost proxy of their last
// step (ints[5]: solver iterations, contacts, MPR iterations, QP size), most expensive first, so that the last CTAs of
// every kernel are the cheap ones.  One CTA; the order inside a bucket is irrelevant (results do not depend on the order).
#define ORDER_BUCKETS 512
__global__ void __launch_bounds__(1024) k_order(BatchDev B, int* order) {
  __shared__ int hist[ORDER_BUCKETS];
  for (int i = threadIdx.x; i < ORDER_BUCKETS; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  auto bucket = [&](int env) {
    const int w = ((const int*)(B.state + (size_t)env * ST_STRIDE + ST_INTS))[5];
    return ORDER_BUCKETS - 1 - min(max(w, 0) >> 2, ORDER_BUCKETS - 1);
  };
  for (int i = threadIdx.x; i < B.n_envs; i += blockDim.x) atomicAdd(&hist[bucket(i)], 1);
  __syncthreads();
  if (threadIdx.x < 32) {      // exclusive scan of the histogram by one warp
    int carry = 0;
    for (int base = 0; base < ORDER_BUCKETS; base += 32) {
      const int v = hist[base + threadIdx.x];
      const int ex = wexscan(v, threadIdx.x);
      hist[base + threadIdx.x] = carry + ex;
      carry += __shfl_sync(FULLMASK, ex + v, 31);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < B.n_envs; i += blockDim.x) order[atomicAdd(&hist[bucket(i)], 1)] = i;
}

// ------------------------------------------------------------------------------------------------ small kernels
__global__ void k_get_state(BatchDev B, float* qpos, float* qvel, float* warm) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B.n_envs * 33) return;
  const int env = i / 33, k = i % 33;
  const float* st = B.state + (size_t)env * ST_STRIDE;
  if (qpos) qpos[i] = st[ST_QPOS + k];
  if (k < 32) {
    if (qvel) qvel[env * 32 + k] = st[ST_QVEL + k];
    if (warm) warm[env * 32 + k] = st[ST_WARM + k];
  }
}
__global__ void k_set_state(BatchDev B, const float* qpos, const float* qvel, const float* warm, const int* cur_t) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B.n_envs * 33) return;
  const int env = i / 33, k = i % 33;
  float* st = B.state + (size_t)env * ST_STRIDE;
  if (qpos) { st[ST_QPOS + k] = qpos[i]; st[ST_PQPOS + k] = qpos[i]; }
  if (k < 32) {
    if (qvel) { st[ST_QVEL + k] = qvel[env * 32 + k]; st[ST_PQVEL + k] = qvel[env * 32 + k]; }
    if (warm) st[ST_WARM + k] = warm[env * 32 + k];
  }
  if (cur_t && k == 0) ((int*)(st + ST_INTS))[0] = cur_t[env];
  if ((qpos || qvel) && k == 0) ((int*)(st + ST_INTS))[7] = 0;      // the step pipeline's working set no longer matches the state
}
// env.cur_t / the expert assignment without touching the physics state (the tracking env re-bases its window every step)
__global__ void k_set_episode(BatchDev B, const int* motion_id, const int* start_idx, const int* cur_t) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B.n_envs) return;
  int* ints = (int*)(B.state + (size_t)i * ST_STRIDE + ST_INTS);
  int mid = ints[1], sidx = ints[2];
  if (motion_id) mid = min(max(motion_id[i], 0), B.n_motions - 1);
  if (start_idx) sidx = start_idx[i];
  const int mlen = B.motion_len[mid];
  sidx = min(max(sidx, 0), mlen - 1);
  ints[1] = mid; ints[2] = sidx; ints[3] = mlen - sidx;
  if (cur_t) ints[0] = cur_t[i];
}
// Sliding expert window of the online tracking env (HandObjMimicTest.insert_new_frame, ho_im_test.py:34-38), on the device:
// every motion is a window kept as [w0, w0, w1, ..., wK, wK] (first and last frame duplicated); pushing a frame drops w0 and
// appends the new frame: [w1, w1, w2, ..., wK, f, f].  One warp per motion.
__global__ void __launch_bounds__(32) k_push_frame(float* frames, const int* __restrict__ motion_offset, const int* __restrict__ motion_len,
                                                   int n_motions, const float* __restrict__ frame, int per_motion) {
  const int mtn = blockIdx.x, lane = threadIdx.x;
  if (mtn >= n_motions) return;
  const int len = motion_len[mtn];
  float* f = frames + (size_t)motion_offset[mtn] * H_FRAME;
  const float* nf = frame + (size_t)(per_motion ? mtn : 0) * H_FRAME;
  // frames 2 .. len-2 move one slot down; a lane owns the same float columns in every frame, so the in-place shift is safe
  for (int c = lane; c < H_FRAME; c += 32) {
    for (int k = 1; k + 2 < len; k++) f[(size_t)k * H_FRAME + c] = f[(size_t)(k + 1) * H_FRAME + c];
    f[c] = f[H_FRAME + c];
    const float v = nf[c];
    f[(size_t)(len - 2) * H_FRAME + c] = v;
    f[(size_t)(len - 1) * H_FRAME + c] = v;
  }
}
__global__ void k_get_rfc(BatchDev B, float* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B.n_envs * 8) return;
  out[i] = B.state[(size_t)(i / 8) * ST_STRIDE + ST_RFC + i % 8];
}

// forward of the state the last forward was computed from; exports derived quantities / contacts
__global__ void __launch_bounds__(32) k_forward_debug(BatchDev B, float* xpos, float* xquat, float* qM, float* bias, int* ncon_out,
                                                      int* pairs, float* cdata) {
  extern __shared__ float s[];
  const int env = blockIdx.x, lane = threadIdx.x;
  if (env >= B.n_envs) return;
  const DevModel& m = *B.env_model[env];
  const float* st = B.state + (size_t)env * ST_STRIDE;
  int* misc = (int*)(s + S_MISC);
  misc[lane] = 0;
  load_state(s, st, lane, ST_PQPOS, ST_PQVEL);
  __syncwarp();
  pos_stage(s, m, lane, s + S_GPOS, s + S_GMAT);
  vel_stage(s, m, lane);
  const V3 org = ld3(s + S_ORG);
  if (xpos) for (int i = lane; i < m.nbody * 3; i += 32) xpos[(size_t)env * m.nbody * 3 + i] = s[S_XPOS + i] + comp(org, i % 3);
  if (xquat) for (int i = lane; i < m.nbody * 4; i += 32) xquat[(size_t)env * m.nbody * 4 + i] = s[S_XQUAT + i];
  if (qM) {
    float* out = qM + (size_t)env * 1024;
    for (int i = lane; i < 1024; i += 32) out[i] = (i / 32 == i % 32 && i / 32 >= m.nv) ? 1.f : 0.f;
    __syncwarp();
    if (lane < m.nv) {
      int d = 0;
      for (unsigned mk = m.dof_anc[lane] | (1u << lane); mk; mk &= mk - 1, d++) {
        const int j = __ffs(mk) - 1;
        out[lane * 32 + j] = s[S_MS + lane * MS_STRIDE + d];
        out[j * 32 + lane] = s[S_MS + lane * MS_STRIDE + d];
      }
    }
  }
  if (bias) bias[(size_t)env * 32 + lane] = s[S_BIAS + lane];
  if (ncon_out) {
    unsigned a = 0, b = 0;
    collide(s, m, lane, a, b);
    const int ncon = misc[MI_NCON];
    if (lane == 0) ncon_out[env] = ncon;
    for (int c = 0; c < HOIC_MAX_CONTACTS; c++) {
      const float* cr = s + U_CON + 16 * c;
      const bool on = c < ncon;
      if (pairs && lane < 2) {
        const DevCand& cd = m.cand[on ? __float_as_int(cr[13]) : 0];
        pairs[((size_t)env * HOIC_MAX_CONTACTS + c) * 2 + lane] = on ? (lane == 0 ? cd.g1 : cd.g2) : -1;
      }
      if (cdata && lane < 13) {
        float v = 0.f;
        if (on) { v = lane == 0 ? cr[12] : cr[lane - 1]; if (lane >= 1 && lane <= 3) v += comp(org, lane - 1); }
        cdata[((size_t)env * HOIC_MAX_CONTACTS + c) * 13 + lane] = v;
      }
    }
  }
}

// one mj_step with explicit ctrl / qfrc_applied (single-substep parity entry)
__global__ void __launch_bounds__(32) k_substep_debug(BatchDev B, const float* ctrl, const float* qfrc_applied, float* qacc,
                                                      float* fcon, int* stats) {
  extern __shared__ float s[];
  const int env = blockIdx.x, lane = threadIdx.x;
  if (env >= B.n_envs) return;
  const DevModel& m = *B.env_model[env];
  float* st = B.state + (size_t)env * ST_STRIDE;
  int* misc = (int*)(s + S_MISC);
  misc[lane] = 0;
  Cnt cnt;
  memset(&cnt, 0, sizeof cnt);
  load_state(s, st, lane, ST_QPOS, ST_QVEL);
  s[S_WARM + lane] = st[ST_WARM + lane];
  s[S_CTRL + lane] = lane < m.nu ? ctrl[(size_t)env * m.nu + lane] : 0.f;
  s[S_QAPP + lane] = qfrc_applied[(size_t)env * 32 + lane];
  const float pq = s[S_QPOS + lane], pv = s[S_QVEL + lane], pq32 = st[ST_QPOS + 32];
  __syncwarp();
  const bool ok = substep_core(s, m, lane, false, cnt);
  if (qacc) qacc[(size_t)env * 32 + lane] = s[S_QACC + lane];
  if (fcon) fcon[(size_t)env * 32 + lane] = s[S_FCON + lane];
  if (stats && lane == 0) {
    stats[env * 4 + 0] = m.nv + misc[MI_NLIM] + misc[MI_NEDGE];
    stats[env * 4 + 1] = misc[MI_NCON];
    stats[env * 4 + 2] = misc[MI_ITER];
    stats[env * 4 + 3] = ok ? 0 : misc[MI_FLAGS] | 4;
  }
  st[ST_QPOS + lane] = s[S_QPOS + lane];
  st[ST_QVEL + lane] = s[S_QVEL + lane];
  st[ST_WARM + lane] = s[S_WARM + lane];
  st[ST_PQPOS + lane] = pq;
  st[ST_PQVEL + lane] = pv;
  if (lane == 0) { st[ST_QPOS + 32] = s[S_QPOS + 32]; st[ST_PQPOS + 32] = pq32; ((int*)(st + ST_INTS))[7] = 0; }
  flush_counters(B, cnt, lane, 0);
}

// standalone batched wrench QP, one warp per problem (hoic_solve_wrench_qp)
__global__ void __launch_bounds__(32) k_wrench_qp(int B, int K, int ec, const float* __restrict__ contacts, const int* __restrict__ ncontact,
                                                  const float* __restrict__ weights, const float* __restrict__ penalties,
                                                  const float* __restrict__ target, const float* __restrict__ center, float mu, float w_t,
                                                  float* rest_out, float* score_out, float* surface, int* iters_out) {
  extern __shared__ float s[];
  const int b = blockIdx.x, lane = threadIdx.x;
  if (b >= B) return;
  const int nc = min(max(ncontact[b], 0), K);
  const V3 F = ld3(target + 6 * (size_t)b), T = ld3(target + 6 * (size_t)b + 3), c = ld3(center + 3 * (size_t)b);
  float* surf = surface ? surface + (size_t)b * HOIC_SURF_STRIDE : nullptr;
  if (surf) for (int i = lane; i < HOIC_SURF_STRIDE; i += 32) surf[i] = 0.f;
  float rest[6] = {F.x, F.y, F.z, T.x, T.y, T.z};
  unsigned it = 0;
  float score;
  if (nc == 0) {
    score = norm(F) + sqrtf(w_t) * norm(T);     // nothing can be explained: the whole target wrench is left
  } else {
    for (int i = lane; i < 12 * nc; i += 32) s[QP_CF + i] = contacts[(size_t)b * K * 12 + i];
    if (lane < nc) s[QP_TS + lane] = weights ? weights[(size_t)b * K + lane] : 1.f;
    __syncwarp();
    score = wrench_qp(s, lane, nc, ec, F, T, c, mu, sqrtf(w_t), penalties ? penalties + (size_t)b * K * ec * 4 : nullptr, rest, it, surf);
  }
  if (lane < 6 && rest_out) rest_out[(size_t)b * 6 + lane] = rest[lane];
  if (lane == 0) { if (score_out) score_out[b] = score; if (iters_out) iters_out[b] = (int)it; }
}

// dataset preprocessing of one raw sequence, one warp per frame (DatasetSingleDepth.preprocess_seq, dataset_singledepth.py:78-142)
__global__ void __launch_bounds__(32) k_preprocess(const DevModel* mp, const float* __restrict__ hand, const float* __restrict__ obj, int T,
                                                   float freq, int clip, float* frames, int* cmask, float* cpos) {
  extern __shared__ float s[];
  const int t = blockIdx.x, lane = threadIdx.x;
  if (t >= T) return;
  const DevModel& m = *mp;
  int* misc = (int*)(s + S_MISC);
  misc[lane] = 0;
  const int nh = m.nhand_q;
  const int a = t == 0 ? 1 : t, b = a - 1;           // vel[0] = vel[1] (:157-158)
  float* f = frames + (size_t)t * H_FRAME;
  float q = 0.f;
  if (lane < nh) {
    // np.clip(hand_dof_seq, lower, upper) (:100); the raw sequence is processed unclipped (:109)
    const float lo = clip ? m.jnt_range[lane][0] : -F_MAXVAL, hi = clip ? m.jnt_range[lane][1] : F_MAXVAL;
    q = fminf(fmaxf(hand[(size_t)t * nh + lane], lo), hi);
    float v = fminf(fmaxf(hand[(size_t)a * nh + lane], lo), hi) - fminf(fmaxf(hand[(size_t)b * nh + lane], lo), hi);
    if (lane >= 3 && lane < 6) {
      for (int it = 0; it < 8 && v > F_PI; it++) v -= 2.f * F_PI;
      for (int it = 0; it < 8 && v < -F_PI; it++) v += 2.f * F_PI;
    }
    f[F_DOF + lane] = q;
    f[F_DOFVEL + lane] = v * freq;
  }
  if (lane < 7) f[F_OBJPOSE + lane] = obj[(size_t)7 * t + lane];
  if (lane < 3) f[F_OBJVEL + lane] = (obj[(size_t)7 * a + lane] - obj[(size_t)7 * b + lane]) * freq;
  {
    const M3 Rd = mulmT(q2m(qnormalize(ldq(obj + (size_t)7 * a + 3))), q2m(qnormalize(ldq(obj + (size_t)7 * b + 3))));
    const V3 w = matrix_to_axis_angle(Rd) * freq;
    if (lane < 3) f[F_OBJANGVEL + lane] = comp(w, lane);
  }
  // forward of (hand dofs, object pose): body poses (:222-237) and the contact pre-pass (:187-220)
  s[S_QPOS + lane] = lane < nh ? q : obj[(size_t)7 * t + lane - nh];
  if (lane == 0) s[S_QPOS + 32] = obj[(size_t)7 * t + 32 - nh];
  s[S_QVEL + lane] = 0.f;
  __syncwarp();
  pos_stage(s, m, lane, s + S_GPOS, s + S_GMAT);
  const V3 org = ld3(s + S_ORG);
  for (int i = lane; i < 3 * m.n_hand_body; i += 32) f[F_BPOS + i] = s[S_XPOS + 3 * m.hand_body_idx[i / 3] + i % 3] + comp(org, i % 3);
  for (int i = lane; i < 4 * m.n_hand_body; i += 32) f[F_BQUAT + i] = s[S_XQUAT + 4 * m.hand_body_idx[i / 4] + i % 4];
  if (!cmask || !cpos) return;
  unsigned c0 = 0, c1 = 0;
  collide(s, m, lane, c0, c1);
  if (lane < H_NHG) cmask[(size_t)t * H_NHG + lane] = 0;
  for (int i = lane; i < 3 * H_NHG; i += 32) cpos[(size_t)t * H_NHG * 3 + i] = 0.f;
  __syncwarp();
  const int ncon = misc[MI_NCON];
  for (int c = 0; c < ncon; c++) {       // contact_info[g1] = contact.pos: the last contact of a hand geom wins
    const float* cr = s + U_CON + 16 * c;
    const int slot = m.cand[__float_as_int(cr[13])].hand_slot;
    if (slot < 0 || slot >= H_NHG) continue;
    if (lane < 3) cpos[((size_t)t * H_NHG + slot) * 3 + lane] = cr[lane] + comp(org, lane);
    else if (lane == 3) cmask[(size_t)t * H_NHG + slot] = 1;
  }
}

// ================================================================================================ host side

static thread_local std::string g_err;
static unsigned long long g_launches = 0;
static int set_err(const std::string& e) { g_err = e; return -1; }
#define CUDA_OK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return set_err(std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)
#define CUDA_OKP(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { set_err(std::string(#call) + ": " + cudaGetErrorString(e_)); return nullptr; } } while (0)

struct hoic_model {
  int device;
  DevModel* dev;
  float4* mesh_vert;
  DevModel host;
};

#define HOIC_MAX_PARTS 8
static_assert(HOIC_MAX_PARTS == HOIC_LIST_PARTS, "dev_batch.cuh: HOIC_LIST_PARTS");
#define HOIC_COUNT_INTS (2 * HOIC_MAX_PARTS * FULL_COUNT_STRIDE + 32)   // full-tier and carry counters per (part, substep) + the reset counter
#define HOIC_TIMING_CLASSES 8          // k_p1, k_p2, k_p3s, k_p3, k_tail, k_reset, k_order, whole call
static int capi_env_int(const char* name, int dflt) { const char* v = getenv(name); return v && *v ? atoi(v) : dflt; }
struct hoic_batch {
  int device, n_envs;
  BatchDev b;
  std::vector<const DevModel*> env_model_host;
  const DevModel** env_model_dev;
  float *frames, *reward_w;
  int *motion_offset, *motion_len;
  int *env_model_idx = nullptr, *samp_off = nullptr, *samp_ids = nullptr, *order = nullptr;
  // optional per-kernel timing of hoic_step (hoic_set_timing): event pairs around every launch, summed per kernel class
  bool timing = false;
  std::vector<cudaEvent_t> tev;          // pairs (before, after)
  std::vector<int> tev_class;
  size_t tev_used = 0;
  double tacc[HOIC_TIMING_CLASSES] = {};
  int tsteps = 0;
  int balance = 0, n_models = 0, n_samp_ids = 0, min_motion_len = 0;
  const int* user_perm = nullptr;
  size_t smem_bytes, frames_cap = 0;
  int motions_cap = 0, nsub = 0;
  // hoic_step runs the batch as `nparts` interleaved parts on internal streams (fork / join around the caller's stream)
  int nparts = 1;
  bool fuse = false;                     // solve + next substep's k_p1 in one launch (k_a.cu)
  int ticks = 0;                         // tick mode: number of ticks per step (HOIC_TICKS), 0 = off
  bool fuse12 = false;                   // k_p1 + k_p2 of a substep in one launch (k_p12.cu, HOIC_FUSE12=1)
  bool p1_wide = false;                  // k_p1 with four envs per CTA (large batches)
  int part_delay_ns = 0;                 // experiment: part 1 starts this much later than part 0 (HOIC_PART_DELAY_US)
  cudaStream_t part_stream[HOIC_MAX_PARTS] = {}, aux_stream[HOIC_MAX_PARTS] = {};   // aux: the solver's full-capacity tier
  cudaStream_t carry_stream[HOIC_MAX_PARTS] = {};                                    // carry mode: the late group of every substep
  cudaEvent_t ev_p3s[HOIC_MAX_PARTS] = {}, ev_carry[HOIC_MAX_PARTS] = {};
  bool order_valid = false;              // b->order holds the env order the previous hoic_step computed for this one
  cudaEvent_t ev_tail = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join[HOIC_MAX_PARTS] = {}, ev_rows[HOIC_MAX_PARTS] = {}, ev_aux[HOIC_MAX_PARTS] = {};
};

extern "C" const char* hoic_last_error(void) { return g_err.c_str(); }
extern "C" int hoic_version(void) { return 1; }
extern "C" unsigned long long hoic_launch_count(void) { return g_launches; }

namespace {

struct Blob {
  hoic_blob_view v;
  bool ok = true;
  std::vector<double> f(const char* name, size_t expect = 0) {
    int dt; size_t n;
    const void* p = hoic_blob_find(&v, name, &dt, &n);
    if (!p || dt != HOIC_BLOB_F64 || (expect && n != expect)) { ok = false; g_err = std::string("blob field ") + name; return std::vector<double>(expect ? expect : 1, 0.0); }
    std::vector<double> out(n);
    memcpy(out.data(), p, n * sizeof(double));
    return out;
  }
  std::vector<int> i(const char* name, size_t expect = 0) {
    int dt; size_t n;
    const void* p = hoic_blob_find(&v, name, &dt, &n);
    if (!p || dt != HOIC_BLOB_I32 || (expect && n != expect)) { ok = false; g_err = std::string("blob field ") + name; return std::vector<int>(expect ? expect : 1, 0); }
    std::vector<int> out(n);
    memcpy(out.data(), p, n * sizeof(int));
    return out;
  }
};

void quat2mat_h(const double* q, float* m) {
  double w = q[0], x = q[1], y = q[2], z = q[3];
  m[0] = (float)(w * w + x * x - y * y - z * z); m[1] = (float)(2 * (x * y - w * z)); m[2] = (float)(2 * (x * z + w * y));
  m[3] = (float)(2 * (x * y + w * z)); m[4] = (float)(w * w - x * x + y * y - z * z); m[5] = (float)(2 * (y * z - w * x));
  m[6] = (float)(2 * (x * z - w * y)); m[7] = (float)(2 * (y * z + w * x)); m[8] = (float)(w * w - x * x - y * y + z * z);
}

double impedance_h(const double* si, double pos, double margin) {
  auto clip = [](double x, double a, double b) { return x < a ? a : (x > b ? b : x); };
  double s0 = clip(si[0], 1e-4, 0.9999), s1 = clip(si[1], 1e-4, 0.9999), s2 = si[2] > 0 ? si[2] : 0, s3 = clip(si[3], 1e-4, 0.9999), s4 = si[4] > 1 ? si[4] : 1;
  if (s0 == s1 || s2 <= 1e-15) return 0.5 * (s0 + s1);
  double x = fabs((pos - margin) / s2), y;
  if (x >= 1 || x <= 0) return x >= 1 ? s1 : s0;
  if (s4 == 1) y = x; else if (x <= s3) y = pow(x, s4) / pow(s3, s4 - 1); else y = 1 - pow(1 - x, s4) / pow(1 - s3, s4 - 1);
  return s0 + y * (s1 - s0);
}

static inline float __int_as_float_h(int i) { float f; memcpy(&f, &i, 4); return f; }
bool build_dev_model(const void* blob, size_t nbytes, DevModel& D, std::vector<float4>& verts, size_t& cl_base) {
  Blob B;
  if (hoic_blob_open(&B.v, blob, nbytes) != 0) { g_err = "not a HOIC model blob"; return false; }
  memset(&D, 0, sizeof D);
  D.nq = B.i("nq", 1)[0]; D.nv = B.i("nv", 1)[0]; D.nbody = B.i("nbody", 1)[0]; D.ngeom = B.i("ngeom", 1)[0];
  D.njnt = B.i("njnt", 1)[0]; D.nu = B.i("nu", 1)[0];
  if (!B.ok) return false;
  if (D.nq > H_NQ - 3 || D.nv > HN || D.nbody > HN || D.ngeom > HN || D.njnt > HN) { g_err = "model exceeds kernel maxima (32 bodies/dofs/geoms/joints)"; return false; }
  const int nb = D.nbody, nv = D.nv, ng = D.ngeom, nj = D.njnt;
  D.timestep = (float)B.f("opt_timestep", 1)[0];
  D.tolerance = (float)B.f("opt_tolerance", 1)[0];
  D.impratio = (float)B.f("opt_impratio", 1)[0];
  D.mpr_tolerance = (float)B.f("opt_mpr_tolerance", 1)[0];
  D.meaninertia = (float)B.f("stat_meaninertia", 1)[0];
  { auto g = B.f("opt_gravity", 3); for (int k = 0; k < 3; k++) D.gravity[k] = (float)g[k]; }
  D.iterations = B.i("opt_iterations", 1)[0];
  D.mpr_iterations = B.i("opt_mpr_iterations", 1)[0];
  auto parent = B.i("body_parentid", nb), rootid = B.i("body_rootid", nb), weld = B.i("body_weldid", nb);
  auto dofadr = B.i("body_dofadr", nb), dofnum = B.i("body_dofnum", nb), jntadr = B.i("body_jntadr", nb), jntnum = B.i("body_jntnum", nb);
  auto geomadr = B.i("body_geomadr", nb), geomnum = B.i("body_geomnum", nb);
  auto bpos = B.f("body_pos", nb * 3), bquat = B.f("body_quat", nb * 4), bipos = B.f("body_ipos", nb * 3), biquat = B.f("body_iquat", nb * 4);
  auto bmass = B.f("body_mass", nb), binertia = B.f("body_inertia", nb * 3), bsub = B.f("body_subtreemass", nb), binvw = B.f("body_invweight0", nb * 2);
  auto jtype = B.i("jnt_type", nj), jbody = B.i("jnt_bodyid", nj), jqadr = B.i("jnt_qposadr", nj), jdadr = B.i("jnt_dofadr", nj), jlim = B.i("jnt_limited", nj);
  auto jpos = B.f("jnt_pos", nj * 3), jaxis = B.f("jnt_axis", nj * 3), jrange = B.f("jnt_range", nj * 2), jmargin = B.f("jnt_margin", nj);
  auto jsolref = B.f("jnt_solref", nj * 2), jsolimp = B.f("jnt_solimp", nj * 5);
  auto dbody = B.i("dof_bodyid", nv), djnt = B.i("dof_jntid", nv), dparent = B.i("dof_parentid", nv);
  auto darm = B.f("dof_armature", nv), ddamp = B.f("dof_damping", nv), dfl = B.f("dof_frictionloss", nv), dinvw = B.f("dof_invweight0", nv);
  auto dsolref = B.f("dof_solref", nv * 2), dsolimp = B.f("dof_solimp", nv * 5);
  auto qpos0 = B.f("qpos0", D.nq);
  auto gtype = B.i("geom_type", ng), gbody = B.i("geom_bodyid", ng), gcontype = B.i("geom_contype", ng), gconaff = B.i("geom_conaffinity", ng);
  auto gcondim = B.i("geom_condim", ng), gvadr = B.i("geom_vertadr", ng), gvnum = B.i("geom_vertnum", ng);
  auto gpos = B.f("geom_pos", ng * 3), gquat = B.f("geom_quat", ng * 4), gsize = B.f("geom_size", ng * 3), grbound = B.f("geom_rbound", ng);
  auto gfri = B.f("geom_friction", ng * 3), gsolref = B.f("geom_solref", ng * 2), gsolimp = B.f("geom_solimp", ng * 5), gsolmix = B.f("geom_solmix", ng);
  auto gmargin = B.f("geom_margin", ng), ggap = B.f("geom_gap", ng);
  std::vector<double> mv;
  { int dt; size_t n; const void* p = hoic_blob_find(&B.v, "mesh_vert", &dt, &n); if (!p || dt != HOIC_BLOB_F64) { g_err = "blob field mesh_vert"; return false; } mv.resize(n); memcpy(mv.data(), p, n * 8); }
  auto actdof = B.i("actuator_dofid", D.nu);
  std::vector<int> pg1, pg2, pdim, psig;
  std::vector<double> pfri, psolref, psolimp, pmargin, pgap;
  { int dt; size_t n; const void* p = hoic_blob_find(&B.v, "pair_geom1", &dt, &n);
    if (p && n > 0) { pg1 = B.i("pair_geom1"); pg2 = B.i("pair_geom2"); pdim = B.i("pair_dim"); psig = B.i("pair_signature");
      pfri = B.f("pair_friction"); psolref = B.f("pair_solref"); psolimp = B.f("pair_solimp"); pmargin = B.f("pair_margin"); pgap = B.f("pair_gap"); } }
  auto hbi = B.i("env_hand_body_idx");
  auto ids = B.i("env_ids", 11);
  auto base_pose = B.f("env_base_pose", D.nu), ctrl_scale = B.f("env_ctrl_scale", D.nu), jkp = B.f("env_jkp", D.nu), jkd = B.f("env_jkd", D.nu), tlim = B.f("env_torque_lim", D.nu);
  auto sc = B.f("env_scalars", 8);
  if (!B.ok) return false;
  for (int u = 0; u < D.nu; u++) if (actdof[u] != u) { g_err = "actuators must drive dofs 0..nu-1 in order"; return false; }

  D.obj_body = ids[0]; D.palm_body = ids[1]; D.hand_geom0 = ids[2]; D.hand_geom1 = ids[3]; D.obj_geom0 = ids[4]; D.obj_geom1 = ids[5];
  D.sim_step = ids[6]; D.w_size = ids[7]; D.residual_force = ids[8]; D.surface_contact = ids[9]; D.explain_force = ids[10];
  D.last_geom = ng - 1;
  D.nhand_q = D.nq - 7; D.nhand_v = nv - 6; D.n_hand_body = (int)hbi.size();
  if (D.hand_geom1 - D.hand_geom0 + 1 > H_NHG || D.n_hand_body > HN) { g_err = "too many hand geoms/bodies"; return false; }
  D.hand_mass = (float)sc[0]; D.rf_scale = (float)sc[1]; D.rt_scale = (float)sc[2]; D.pos_thresh = (float)sc[3]; D.rot_thresh = (float)sc[4];
  D.jpos_thresh = (float)sc[5]; D.obj_pos_thresh = (float)sc[6]; D.obj_rot_thresh = (float)sc[7];
  D.obj_mass = (float)bmass[nb - 1];
  for (int k = 0; k < 3; k++) D.obj_inertia[k] = (float)binertia[(nb - 1) * 3 + k];
  for (int i = 0; i < D.n_hand_body; i++) D.hand_body_idx[i] = hbi[i];
  for (int i = 0; i < D.nu; i++) { D.base_pose[i] = (float)base_pose[i]; D.ctrl_scale[i] = (float)ctrl_scale[i]; D.jkp[i] = (float)jkp[i]; D.jkd[i] = (float)jkd[i]; D.torque_lim[i] = (float)tlim[i]; }
  // origin: the first three qpos must be the slides of the hand root (world-aligned translation)
  D.org_qadr = 0;
  // bodies
  D.nroot = 0;
  std::vector<int> rootslot_of_body(nb, -1);
  for (int b = 1; b < nb; b++) {
    if (rootid[b] == b) {
      bool moving = false;
      for (int c = b; c < nb; c++) if (rootid[c] == b && dofnum[c] > 0) moving = true;
      if (moving) {
        if (D.nroot >= 4) { g_err = "more than 4 moving kinematic trees"; return false; }
        D.root_body[D.nroot] = b;
        D.root_invmass[D.nroot] = bsub[b] > 1e-15 ? (float)(1.0 / bsub[b]) : 0.f;
        rootslot_of_body[b] = D.nroot++;
      }
    }
  }
  int maxlevel = 0;
  for (int b = 0; b < nb; b++) {
    D.body_parent[b] = parent[b];
    int lvl = 0;
    for (int c = b; c != 0; c = parent[c]) lvl++;
    D.body_level[b] = lvl;
    if (lvl > maxlevel) maxlevel = lvl;
    D.body_rootslot[b] = b ? rootslot_of_body[rootid[b]] : -1;
    D.body_dofadr[b] = dofadr[b]; D.body_dofnum[b] = dofnum[b]; D.body_jntadr[b] = jntadr[b]; D.body_jntnum[b] = jntnum[b];
    D.body_weld[b] = weld[b];
    D.body_firstchild[b] = -1; D.body_nextsib[b] = -1;
    for (int k = 0; k < 3; k++) { D.body_pos[b][k] = (float)bpos[3 * b + k]; D.body_ipos[b][k] = (float)bipos[3 * b + k]; D.body_inertia[b][k] = (float)binertia[3 * b + k]; }
    for (int k = 0; k < 4; k++) D.body_quat[b][k] = (float)bquat[4 * b + k];
    quat2mat_h(&biquat[4 * b], D.body_imat[b]);
    D.body_mass[b] = (float)bmass[b];
  }
  D.nlevel = maxlevel;
  for (int b = nb - 1; b >= 1; b--) { int p = parent[b]; D.body_nextsib[b] = D.body_firstchild[p]; D.body_firstchild[p] = b; }
  for (int b = 0; b < nb; b++) {
    unsigned mask = 0;
    int bb = b;
    while (bb && !dofnum[bb]) bb = parent[bb];
    if (bb) for (int d = dofadr[bb] + dofnum[bb] - 1; d >= 0; d = dparent[d]) mask |= 1u << d;
    D.body_chain[b] = mask;
  }
  for (int j = 0; j < nj; j++) {
    D.jnt_type[j] = jtype[j]; D.jnt_body[j] = jbody[j]; D.jnt_qposadr[j] = jqadr[j]; D.jnt_dofadr[j] = jdadr[j]; D.jnt_limited[j] = jlim[j];
    for (int k = 0; k < 3; k++) { D.jnt_pos[j][k] = (float)jpos[3 * j + k]; D.jnt_axis[j][k] = (float)jaxis[3 * j + k]; }
    D.jnt_range[j][0] = (float)jrange[2 * j]; D.jnt_range[j][1] = (float)jrange[2 * j + 1];
    D.jnt_margin[j] = (float)jmargin[j];
    for (int k = 0; k < 2; k++) D.jnt_solref[j][k] = (float)jsolref[2 * j + k];
    for (int k = 0; k < 5; k++) D.jnt_solimp[j][k] = (float)jsolimp[5 * j + k];
    D.jnt_invweight[j] = (float)dinvw[jdadr[j]];
  }
  for (int k = 0; k < D.nq; k++) D.qpos0[k] = (float)qpos0[k];
  for (int d = 0; d < nv; d++) {
    D.dof_body[d] = dbody[d]; D.dof_jnt[d] = djnt[d]; D.dof_parent[d] = dparent[d];
    D.dof_rootslot[d] = rootslot_of_body[rootid[dbody[d]]];
    {
      unsigned anc = 0;
      for (int e = dparent[d]; e >= 0; e = dparent[e]) anc |= 1u << e;
      D.dof_anc[d] = anc;
      const int j = djnt[d];
      if (jtype[j] == J_FREE) {
        const int k = d - jdadr[j];
        unsigned base = 0;
        for (int e = dparent[jdadr[j]]; e >= 0; e = dparent[e]) base |= 1u << e;
        D.dof_dotzero[d] = k < 3;
        D.dof_premask[d] = k < 3 ? 0u : (base | (7u << jdadr[j]));
      } else {
        D.dof_dotzero[d] = 0;
        D.dof_premask[d] = anc;
      }
    }
    D.dof_armature[d] = (float)darm[d]; D.dof_damping[d] = (float)ddamp[d]; D.dof_frictionloss[d] = (float)dfl[d];
    // friction-loss row constants: pos = margin = 0 -> impedance solimp[0]; K = 0
    const double* si = &dsolimp[5 * d];
    const double* sr = &dsolref[2 * d];
    double imp = impedance_h(si, 0, 0), dmax = si[1] < 1e-4 ? 1e-4 : (si[1] > 0.9999 ? 0.9999 : si[1]);
    double R = (1 - imp) * dinvw[d] / imp;
    if (R < 1e-15) R = 1e-15;
    double Bc;
    if (sr[0] > 0) { double tc = sr[0] > 2 * D.timestep ? sr[0] : 2 * (double)D.timestep; Bc = 2 / (dmax * tc); } else Bc = -sr[1] / dmax;
    D.fr_D[d] = (float)(1.0 / R); D.fr_B[d] = (float)Bc; D.fr_Rf[d] = (float)(R * dfl[d]);
  }
  // dof-tree tables for tree_solve (dev_math.cuh)
  {
    std::vector<int> depth(HN, -1);
    std::vector<std::vector<int>> kids(nv);
    int maxd = 0, hub = -1;
    for (int d = 0; d < nv; d++) {
      depth[d] = dparent[d] < 0 ? 0 : depth[dparent[d]] + 1;
      if (dparent[d] >= d) { g_err = "dof parents must precede their children"; return false; }
      if (dparent[d] >= 0) kids[dparent[d]].push_back(d);
      if (depth[d] > maxd) maxd = depth[d];
    }
    for (int d = 0; d < nv; d++)
      if (kids[d].size() > 1) {
        if (hub >= 0) { g_err = "unsupported dof tree: more than one branching dof"; return false; }
        hub = d;
      }
    if (maxd >= 10) { g_err = "unsupported dof tree: deeper than 10 dofs"; return false; }
    D.tree_maxdepth = maxd;
    D.hub_depth = hub >= 0 ? depth[hub] : -1;
    if (hub >= 0 && depth[hub] != T_HUBD) { g_err = "unsupported dof tree: the branching dof must have 5 ancestors"; return false; }
    D.branch_mask = 0;
    for (int w = 0; w < 4; w++) D.dof_depth_nib[w] = 0;
    for (int d = 0; d < nv; d++) D.dof_depth_nib[d / 8] |= (unsigned)depth[d] << (4 * (d % 8));
    for (int i = 0; i < HN; i++) {
      D.dof_depth[i] = i < nv ? depth[i] : -1;
      D.dof_is_trunk[i] = 0;
      D.dof_desc[i] = 0;
      for (int w = 0; w < 3; w++) { D.dof_anc_at[i][w] = 0xffffffffu; D.dof_desc_at[i][w] = 0xffffffffu; }
    }
    auto setbyte = [](unsigned* w, int i, int v) { w[i / 4] = (w[i / 4] & ~(0xffu << (8 * (i % 4)))) | ((unsigned)v << (8 * (i % 4))); };
    for (int d = hub; d >= 0; d = dparent[d]) D.dof_is_trunk[d] = 1;
    for (int k = 0; k < nv; k++) {
      bool below_hub = false;
      for (int a = dparent[k]; a >= 0; a = dparent[a]) {
        setbyte(D.dof_anc_at[k], depth[a], a);
        D.dof_desc[a] |= 1u << k;
        if (a == hub) below_hub = true;
      }
      if (below_hub) D.branch_mask |= 1u << k;
      // k is the unique descendant at depth[k] of every ancestor up to (excluding) the hub
      for (int a = dparent[k]; a >= 0 && !(below_hub && D.dof_is_trunk[a]); a = dparent[a]) setbyte(D.dof_desc_at[a], depth[k], k);
    }
    D.branch_lo = 0; D.branch_n = 0;
    if (D.branch_mask) {
      D.branch_lo = __builtin_ctz(D.branch_mask);
      D.branch_n = __builtin_popcount(D.branch_mask);
      if (D.branch_mask != (((D.branch_n < 32 ? (1u << D.branch_n) : 0u) - 1u) << D.branch_lo)) { g_err = "unsupported dof tree: dofs below the branching dof must be contiguous"; return false; }
    }
  }
  for (int g = 0; g < ng; g++) {
    D.geom_type[g] = gtype[g]; D.geom_body[g] = gbody[g]; D.geom_vertadr[g] = gvadr[g]; D.geom_vertnum[g] = gvnum[g];
    for (int k = 0; k < 3; k++) { D.geom_pos[g][k] = (float)gpos[3 * g + k]; D.geom_size[g][k] = (float)gsize[3 * g + k]; }
    quat2mat_h(&gquat[4 * g], D.geom_mat[g]);
    D.geom_rbound[g] = (float)grbound[g];
  }
  // mesh vertices: per mesh geom, reorder into clusters of 32 spatially close vertices (recursive median split along the
  // widest axis, split points on multiples of 32) and record each cluster's AABB; .w keeps the original index, which
  // breaks ties in the support function exactly like a serial scan in the original order.
  verts.assign(mv.size() / 3 + 1, make_float4(0.f, 0.f, 0.f, 0.f));
  cl_base = verts.size();
  for (int g = 0; g < ng; g++) {
    D.geom_cladr[g] = 0; D.geom_ncl[g] = 0;
    for (int k = 0; k < 3; k++) { D.geom_aabb_c[g][k] = 0.f; D.geom_aabb_h[g][k] = 0.f; }
    const int n = gvnum[g], adr = gvadr[g];
    if (gtype[g] != G_MESH || n <= 0 || adr < 0) continue;
    std::vector<int> idx(n);
    for (int i = 0; i < n; i++) idx[i] = i;
    auto coord = [&](int i, int ax) { return mv[3 * (size_t)(adr + i) + ax]; };
    std::vector<std::pair<int, int>> stack{{0, n}};
    while (!stack.empty()) {
      const auto [lo, hi] = stack.back();
      stack.pop_back();
      if (hi - lo <= 32) continue;
      double mn[3] = {1e30, 1e30, 1e30}, mx[3] = {-1e30, -1e30, -1e30};
      for (int i = lo; i < hi; i++)
        for (int ax = 0; ax < 3; ax++) { mn[ax] = std::min(mn[ax], coord(idx[i], ax)); mx[ax] = std::max(mx[ax], coord(idx[i], ax)); }
      int ax = 0;
      for (int k = 1; k < 3; k++) if (mx[k] - mn[k] > mx[ax] - mn[ax]) ax = k;
      const int ncl = (hi - lo + 31) / 32, mid = lo + 32 * ((ncl + 1) / 2);
      std::nth_element(idx.begin() + lo, idx.begin() + mid, idx.begin() + hi, [&](int a, int b) { return coord(a, ax) < coord(b, ax); });
      stack.push_back({lo, mid});
      stack.push_back({mid, hi});
    }
    const int ncl = (n + 31) / 32;
    D.geom_cladr[g] = (int)(verts.size() - cl_base);
    D.geom_ncl[g] = ncl;
    float glo[3] = {1e30f, 1e30f, 1e30f}, ghi[3] = {-1e30f, -1e30f, -1e30f};
    for (int c = 0; c < ncl; c++) {
      float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
      for (int i = 32 * c; i < std::min(n, 32 * c + 32); i++) {
        float4 v = make_float4((float)coord(idx[i], 0), (float)coord(idx[i], 1), (float)coord(idx[i], 2), __int_as_float_h(idx[i]));
        verts[adr + i] = v;
        const float xyz[3] = {v.x, v.y, v.z};
        for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], xyz[k]); hi[k] = std::max(hi[k], xyz[k]); }
      }
      verts.push_back(make_float4(lo[0], lo[1], lo[2], 0.f));
      verts.push_back(make_float4(hi[0], hi[1], hi[2], 0.f));
      for (int k = 0; k < 3; k++) { glo[k] = std::min(glo[k], lo[k]); ghi[k] = std::max(ghi[k], hi[k]); }
    }
    for (int k = 0; k < 3; k++) { D.geom_aabb_c[g][k] = 0.5f * (glo[k] + ghi[k]); D.geom_aabb_h[g][k] = 0.5f * (ghi[k] - glo[k]); }
    const float (*dop)[3] = h_dop;
    for (int k = 0; k < H_NDOP; k++) {
      double mn = 1e30, mx = -1e30;
      for (int i = 0; i < n; i++) {
        const double pr = coord(i, 0) * dop[k][0] + coord(i, 1) * dop[k][1] + coord(i, 2) * dop[k][2];
        mn = std::min(mn, pr); mx = std::max(mx, pr);
      }
      D.geom_dop[g][k][0] = (float)mn - 1e-6f; D.geom_dop[g][k][1] = (float)mx + 1e-6f;
    }
  }
  // collision candidates in signature order (explicit pairs merged; mj_collision driver)
  int ncand = 0;
  bool too_many = false;
  auto add_cand = [&](int g1, int g2, int pair) {
    if (gtype[g1] > gtype[g2]) { int t = g1; g1 = g2; g2 = t; }
    const int t1 = gtype[g1], t2 = gtype[g2];
    double margin, gap, solref[2], solimp[5], fr[3];
    int condim;
    if (pair < 0) {
      if (gbody[g1] == gbody[g2]) return;
      if (!((gcontype[g1] & gconaff[g2]) || (gcontype[g2] & gconaff[g1]))) return;
      margin = gmargin[g1] > gmargin[g2] ? gmargin[g1] : gmargin[g2];
      gap = ggap[g1] > ggap[g2] ? ggap[g1] : ggap[g2];
      double s1 = gsolmix[g1], s2 = gsolmix[g2], mix = (s1 + s2 > 0) ? s1 / (s1 + s2) : 0.5;
      condim = gcondim[g1] > gcondim[g2] ? gcondim[g1] : gcondim[g2];
      for (int k = 0; k < 3; k++) fr[k] = gfri[3 * g1 + k] > gfri[3 * g2 + k] ? gfri[3 * g1 + k] : gfri[3 * g2 + k];
      if (gsolref[2 * g1] > 0 && gsolref[2 * g2] > 0) for (int k = 0; k < 2; k++) solref[k] = mix * gsolref[2 * g1 + k] + (1 - mix) * gsolref[2 * g2 + k];
      else for (int k = 0; k < 2; k++) solref[k] = gsolref[2 * g1 + k] < gsolref[2 * g2 + k] ? gsolref[2 * g1 + k] : gsolref[2 * g2 + k];
      for (int k = 0; k < 5; k++) solimp[k] = mix * gsolimp[5 * g1 + k] + (1 - mix) * gsolimp[5 * g2 + k];
    } else {
      margin = pmargin[pair]; gap = pgap[pair]; condim = pdim[pair];
      fr[0] = pfri[5 * pair]; fr[1] = pfri[5 * pair + 2]; fr[2] = pfri[5 * pair + 3];
      for (int k = 0; k < 2; k++) solref[k] = psolref[2 * pair + k];
      for (int k = 0; k < 5; k++) solimp[k] = psolimp[5 * pair + k];
    }
    int kind;
    if (t1 == G_PLANE && t2 == G_CAPSULE) kind = K_PLANE_CAPSULE;
    else if (t1 == G_PLANE && t2 == G_BOX) kind = K_PLANE_BOX;
    else if (t1 == G_PLANE && t2 == G_MESH) kind = K_PLANE_MESH;
    else if (t1 == G_CAPSULE && t2 == G_CAPSULE) kind = K_CAP_CAP;
    else if (t1 == G_CAPSULE && t2 == G_BOX) kind = K_CAP_BOX;
    else if (t1 == G_BOX && t2 == G_BOX) kind = K_BOX_BOX;
    else if (t2 == G_MESH && t1 != G_PLANE) kind = K_CONVEX;
    else return;
    if (condim != 1 && condim != 3 && condim != 4) { too_many = true; g_err = "unsupported condim"; return; }
    if (ncand >= H_MAXCAND) { too_many = true; g_err = "too many collision candidates"; return; }
    DevCand& c = D.cand[ncand++];
    c.g1 = g1; c.g2 = g2; c.kind = kind; c.condim = condim;
    c.margin = (float)margin; c.includemargin = (float)(margin - gap);
    for (int k = 0; k < 3; k++) c.friction[k] = (float)fr[k];
    for (int k = 0; k < 2; k++) c.solref[k] = (float)solref[k];
    for (int k = 0; k < 5; k++) c.solimp[k] = (float)solimp[k];
    c.dA_tran = (float)(binvw[2 * gbody[g1]] + binvw[2 * gbody[g2]]);
    c.dA_rot = (float)(binvw[2 * gbody[g1] + 1] + binvw[2 * gbody[g2] + 1]);
    c.chain1 = D.body_chain[gbody[g1]]; c.chain2 = D.body_chain[gbody[g2]];
    if (__builtin_popcount(c.chain1 ^ c.chain2) > B_STRIDE - 1) { too_many = true; g_err = "a contact pair moves more than 16 dofs (compact Jacobian rows)"; return; }
    c.hand_slot = (g1 >= D.hand_geom0 && g1 <= D.hand_geom1 && g2 >= D.obj_geom0 && g2 <= D.obj_geom1) ? g1 - D.hand_geom0 : -1;
  };
  {
    size_t pairadr = 0;
    for (int b1 = 0; b1 < nb; b1++)
      for (int b2 = b1 + 1; b2 < nb; b2++) {
        const int sig = (b1 << 16) + b2;
        bool merged = false;
        while (pairadr < pg1.size() && psig[pairadr] <= sig) { if (psig[pairadr] == sig) merged = true; add_cand(pg1[pairadr], pg2[pairadr], (int)pairadr); pairadr++; }
        if (merged) continue;
        if (!geomnum[b1] || !geomnum[b2]) continue;
        if (weld[b1] == 0 && weld[b2] == 0) continue;
        if (weld[b1] == weld[b2]) continue;
        const int w1 = weld[b1], w2 = weld[b2], pw1 = weld[parent[w1]], pw2 = weld[parent[w2]];
        if ((w1 != 0 && w2 != 0) && (pw1 == w2 || pw2 == w1)) continue;
        for (int g1 = geomadr[b1]; g1 < geomadr[b1] + geomnum[b1]; g1++)
          for (int g2 = geomadr[b2]; g2 < geomadr[b2] + geomnum[b2]; g2++) add_cand(g1, g2, -1);
      }
    while (pairadr < pg1.size()) { add_cand(pg1[pairadr], pg2[pairadr], (int)pairadr); pairadr++; }
  }
  if (too_many) return false;
  D.ncand = ncand;
  // leading bytes k_step stages into shared memory: everything up to the used candidates; large meshes keep the L1 instead
  D.stage_bytes = mv.size() / 3 <= 1024 ? (int)((offsetof(DevModel, cand) + (size_t)ncand * sizeof(DevCand) + 7) & ~(size_t)7) : 0;
  return true;
}

}  // namespace

extern "C" hoic_model* hoic_model_create(const void* blob_host, size_t nbytes, int device) {
  hoic_model* M = new hoic_model();
  std::vector<float4> verts;
  size_t cl_base = 0;
  if (!build_dev_model(blob_host, nbytes, M->host, verts, cl_base)) { delete M; return nullptr; }
  M->device = device;
  cudaError_t e = cudaSetDevice(device);
  if (e == cudaSuccess) e = cudaMalloc(&M->mesh_vert, verts.size() * sizeof(float4));
  if (e == cudaSuccess) e = cudaMemcpy(M->mesh_vert, verts.data(), verts.size() * sizeof(float4), cudaMemcpyHostToDevice);
  M->host.mesh_vert = M->mesh_vert;
  M->host.mesh_cl = M->mesh_vert + cl_base;
  if (e == cudaSuccess) e = cudaMalloc(&M->dev, sizeof(DevModel));
  if (e == cudaSuccess) e = cudaMemcpy(M->dev, &M->host, sizeof(DevModel), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { set_err(std::string("hoic_model_create: ") + cudaGetErrorString(e)); delete M; return nullptr; }
  return M;
}

extern "C" void hoic_model_destroy(hoic_model* m) {
  if (!m) return;
  cudaSetDevice(m->device);
  cudaFree(m->dev);
  cudaFree(m->mesh_vert);
  delete m;
}

__global__ void k_get_work(BatchDev B, int* out);
extern "C" hoic_batch* hoic_batch_create(const hoic_model* const* models, int n_models, const int* env_model_id_host, int n_envs, int device) {
  if (!models || n_models < 1 || n_envs < 1) { set_err("hoic_batch_create: bad arguments"); return nullptr; }
  hoic_batch* b = new hoic_batch();
  memset(&b->b, 0, sizeof b->b);
  b->device = device; b->n_envs = n_envs;
  b->env_model_host.resize(n_envs);
  for (int i = 0; i < n_envs; i++) {
    int id = env_model_id_host ? env_model_id_host[i] : 0;
    if (id < 0 || id >= n_models || models[id]->device != device) { set_err("hoic_batch_create: bad model id / device"); delete b; return nullptr; }
    b->env_model_host[i] = models[id]->dev;
  }
  CUDA_OKP(cudaSetDevice(device));
  CUDA_OKP(cudaMalloc(&b->env_model_dev, n_envs * sizeof(DevModel*)));
  CUDA_OKP(cudaMemcpy(b->env_model_dev, b->env_model_host.data(), n_envs * sizeof(DevModel*), cudaMemcpyHostToDevice));
  {
    std::vector<int> idx(n_envs);
    for (int i = 0; i < n_envs; i++) idx[i] = env_model_id_host ? env_model_id_host[i] : 0;
    CUDA_OKP(cudaMalloc(&b->env_model_idx, n_envs * sizeof(int)));
    CUDA_OKP(cudaMemcpy(b->env_model_idx, idx.data(), n_envs * sizeof(int), cudaMemcpyHostToDevice));
    CUDA_OKP(cudaMalloc(&b->order, n_envs * sizeof(int)));
    b->b.env_model_idx = b->env_model_idx;
    b->b.order = b->order;
    b->n_models = n_models;
    b->balance = n_envs > 7;
  }
  CUDA_OKP(cudaMalloc(&b->b.state, (size_t)n_envs * ST_STRIDE * sizeof(float)));
  CUDA_OKP(cudaMemset(b->b.state, 0, (size_t)n_envs * ST_STRIDE * sizeof(float)));
  CUDA_OKP(cudaMalloc(&b->b.old_geom, (size_t)n_envs * (HN * 12 + 4) * sizeof(float)));
  CUDA_OKP(cudaMalloc(&b->b.csum, (size_t)n_envs * CSUM_STRIDE * sizeof(float)));
  CUDA_OKP(cudaMalloc(&b->b.csum_carry, (size_t)n_envs * CSUM_STRIDE * sizeof(float)));
  if (capi_env_int("HOIC_FULL_LIST", 1)) {
    CUDA_OKP(cudaMalloc(&b->b.full_list, (size_t)2 * HOIC_MAX_PARTS * n_envs * sizeof(int)));
    CUDA_OKP(cudaMalloc(&b->b.carry_list, (size_t)HOIC_MAX_PARTS * n_envs * sizeof(int)));
    CUDA_OKP(cudaMalloc(&b->b.late_stamp, (size_t)n_envs * sizeof(int)));
    CUDA_OKP(cudaMemset(b->b.late_stamp, 0, (size_t)n_envs * sizeof(int)));
    b->b.solve_budget = capi_env_int("HOIC_SOLVE_BUDGET", 0);
    b->ticks = std::min(std::max(capi_env_int("HOIC_TICKS", 0), 0), 256);
    b->b.tk_budget = std::max(capi_env_int("HOIC_TICK_BUDGET", 4), 1);
    if (b->ticks > 0) {
      CUDA_OKP(cudaMalloc(&b->b.tk_k, (size_t)n_envs * sizeof(int)));
      CUDA_OKP(cudaMalloc(&b->b.tk_list, (size_t)6 * n_envs * sizeof(int)));
      CUDA_OKP(cudaMalloc(&b->b.tk_count, (size_t)(b->ticks + 2) * 4 * sizeof(int)));
    }
    CUDA_OKP(cudaMalloc(&b->b.full_count, (size_t)HOIC_COUNT_INTS * sizeof(int)));
    CUDA_OKP(cudaMemset(b->b.full_count, 0, (size_t)HOIC_COUNT_INTS * sizeof(int)));
    CUDA_OKP(cudaMalloc(&b->b.reset_list, (size_t)n_envs * sizeof(int)));
    b->b.carry_count = b->b.full_count + HOIC_MAX_PARTS * FULL_COUNT_STRIDE;
    b->b.reset_count = b->b.full_count + 2 * HOIC_MAX_PARTS * FULL_COUNT_STRIDE;
  }
  CUDA_OKP(cudaMalloc(&b->b.counters, 16 * sizeof(unsigned long long)));
  CUDA_OKP(cudaMemset(b->b.counters, 0, 16 * sizeof(unsigned long long)));
  CUDA_OKP(cudaMalloc(&b->reward_w, 24 * sizeof(float)));
  CUDA_OKP(cudaMemset(b->reward_w, 0, 24 * sizeof(float)));
  b->b.n_envs = n_envs;
  b->b.env_model = b->env_model_dev;
  b->b.reward_w = b->reward_w;
  b->b.mode_train = 1;
  b->smem_bytes = (size_t)S_TOTAL * sizeof(float);
  CUDA_OKP(cudaMalloc(&b->b.scratch, (size_t)n_envs * G_TOTAL * sizeof(float)));
  CUDA_OKP(cudaMemset(b->b.scratch, 0, (size_t)n_envs * G_TOTAL * sizeof(float)));
  for (int i = 0; i < n_models; i++) b->nsub = std::max(b->nsub, models[i]->host.sim_step);
  b->b.model_smem = 0;
  {
    // parts only pay off while a part still fills the GPU about once; HOIC_STEP_PARTS overrides
    int want = n_envs >= 1024 ? 2 : 1;
    if (const char* e = getenv("HOIC_STEP_PARTS")) want = atoi(e);
    b->nparts = std::min(std::max(want, 1), HOIC_MAX_PARTS);
    // fused solve + k_p1 launches (k_a.cu) are an opt-in experiment (HOIC_FUSE=1): measured SLOWER than the separate kernels
    // (box, 4096 envs: 5.37 vs 4.32 ms per step; 16 384 envs: 18.0 vs 13.6 ms) -- warps of one SM then run the solver's and
    // k_p1's code at the same time and the instruction cache hit rate falls from 93 % / 91 % to 73 % (profiles/README.md)
    b->fuse = false;
    b->part_delay_ns = 1000 * capi_env_int("HOIC_PART_DELAY_US", 0);
    b->p1_wide = capi_env_int("HOIC_P1_WIDE", n_envs >= 8192 ? 1 : 0) != 0;
    if (const char* e = getenv("HOIC_FUSE")) b->fuse = atoi(e) != 0;
    b->fuse12 = capi_env_int("HOIC_FUSE12", 0) != 0 && !b->fuse;
    if (b->fuse || b->fuse12 || !b->b.full_list) b->b.solve_budget = 0;      // carry mode needs the separate kernels and the env lists
    if (b->fuse || b->fuse12 || b->b.solve_budget > 0 || !b->b.full_list) b->ticks = 0;
    CUDA_OKP(cudaEventCreateWithFlags(&b->ev_fork, cudaEventDisableTiming));
    CUDA_OKP(cudaEventCreateWithFlags(&b->ev_tail, cudaEventDisableTiming));
    for (int j = 0; j < b->nparts; j++) {
      // experiment knobs (tools/ab_r2b.sh): HOIC_PRIO_PART1 / HOIC_PRIO_AUX = CUDA stream priority (lower = more urgent)
      int lo = 0, hi = 0;
      cudaDeviceGetStreamPriorityRange(&lo, &hi);
      const int pp = j == 1 ? capi_env_int("HOIC_PRIO_PART1", 0) : 0, pa = capi_env_int("HOIC_PRIO_AUX", 0);
      CUDA_OKP(cudaStreamCreateWithPriority(&b->part_stream[j], cudaStreamNonBlocking, std::max(hi, std::min(lo, pp))));
      CUDA_OKP(cudaStreamCreateWithPriority(&b->aux_stream[j], cudaStreamNonBlocking, std::max(hi, std::min(lo, pa))));
      // the carry stream runs the few late envs' short launches, which sit on the critical path: most urgent priority by default
      const int pc = capi_env_int("HOIC_PRIO_CARRY", -100);
      CUDA_OKP(cudaStreamCreateWithPriority(&b->carry_stream[j], cudaStreamNonBlocking, std::max(hi, std::min(lo, pc))));
      CUDA_OKP(cudaEventCreateWithFlags(&b->ev_p3s[j], cudaEventDisableTiming));
      CUDA_OKP(cudaEventCreateWithFlags(&b->ev_carry[j], cudaEventDisableTiming));
      CUDA_OKP(cudaEventCreateWithFlags(&b->ev_join[j], cudaEventDisableTiming));
      CUDA_OKP(cudaEventCreateWithFlags(&b->ev_rows[j], cudaEventDisableTiming));
      CUDA_OKP(cudaEventCreateWithFlags(&b->ev_aux[j], cudaEventDisableTiming));
    }
  }
  {
    static bool configured[64] = {false};
    if (!configured[device & 63]) {
      CUDA_OKP(hoic_configure_p1()); CUDA_OKP(hoic_configure_p2()); CUDA_OKP(hoic_configure_p3()); CUDA_OKP(hoic_configure_p3s()); CUDA_OKP(hoic_configure_p3sb()); CUDA_OKP(hoic_configure_p3c()); CUDA_OKP(hoic_configure_p3fc()); CUDA_OKP(hoic_configure_tail());
      CUDA_OKP(hoic_configure_a()); CUDA_OKP(hoic_configure_as()); CUDA_OKP(hoic_configure_p1w()); CUDA_OKP(hoic_configure_p12());
      configured[device & 63] = true;
    }
  }
  CUDA_OKP(cudaFuncSetAttribute(k_reset, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b->smem_bytes));
  { cudaFuncAttributes fa; CUDA_OKP(cudaFuncGetAttributes(&fa, k_get_work)); }   // load now: launches may later be graph-captured
  CUDA_OKP(cudaFuncSetAttribute(k_forward_debug, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b->smem_bytes));
  CUDA_OKP(cudaFuncSetAttribute(k_substep_debug, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b->smem_bytes));
  return b;
}

extern "C" void hoic_batch_destroy(hoic_batch* b) {
  if (!b) return;
  cudaSetDevice(b->device);
  cudaDeviceSynchronize();
  if (b->ev_fork) cudaEventDestroy(b->ev_fork);
  if (b->ev_tail) cudaEventDestroy(b->ev_tail);
  for (int j = 0; j < HOIC_MAX_PARTS; j++) {
    if (b->ev_join[j]) cudaEventDestroy(b->ev_join[j]);
    if (b->ev_p3s[j]) cudaEventDestroy(b->ev_p3s[j]);
    if (b->ev_carry[j]) cudaEventDestroy(b->ev_carry[j]);
    if (b->carry_stream[j]) cudaStreamDestroy(b->carry_stream[j]);
    if (b->ev_rows[j]) cudaEventDestroy(b->ev_rows[j]);
    if (b->ev_aux[j]) cudaEventDestroy(b->ev_aux[j]);
    if (b->part_stream[j]) cudaStreamDestroy(b->part_stream[j]);
    if (b->aux_stream[j]) cudaStreamDestroy(b->aux_stream[j]);
  }
  cudaFree(b->env_model_dev); cudaFree(b->b.scratch); cudaFree(b->b.state); cudaFree(b->b.old_geom); cudaFree(b->b.counters); cudaFree(b->b.csum); cudaFree(b->b.csum_carry);
  cudaFree(b->b.full_list); cudaFree(b->b.full_count); cudaFree(b->b.reset_list); cudaFree(b->b.carry_list); cudaFree(b->b.late_stamp);
  cudaFree(b->b.tk_k); cudaFree(b->b.tk_list); cudaFree(b->b.tk_count);
  cudaFree(b->env_model_idx); cudaFree(b->samp_off); cudaFree(b->samp_ids); cudaFree(b->order);
  cudaFree(b->reward_w); cudaFree(b->frames); cudaFree(b->motion_offset); cudaFree(b->motion_len);
  delete b;
}

extern "C" int hoic_batch_size(const hoic_batch* b) { return b ? b->n_envs : -1; }

extern "C" int hoic_set_motions(hoic_batch* b, const float* frames_host, const int* motion_offset_host, const int* motion_len_host, int n_motions) {
  if (!b || !frames_host || n_motions < 1) return set_err("hoic_set_motions: bad arguments");
  size_t total = 0;
  b->min_motion_len = motion_len_host[0];
  for (int i = 0; i < n_motions; i++) {
    if (motion_len_host[i] < 2) return set_err("hoic_set_motions: motion shorter than 2 frames");
    b->min_motion_len = std::min(b->min_motion_len, motion_len_host[i]);
    size_t e = (size_t)motion_offset_host[i] + motion_len_host[i];
    if (e > total) total = e;
  }
  CUDA_OK(cudaSetDevice(b->device));
  CUDA_OK(cudaDeviceSynchronize());
  if (total != b->frames_cap || n_motions != b->motions_cap) {
    // (a table of unchanged shape is refreshed in place: the sliding-window tracking env rewrites its frames every step)
    cudaFree(b->frames); cudaFree(b->motion_offset); cudaFree(b->motion_len);
    b->frames = nullptr; b->motion_offset = nullptr; b->motion_len = nullptr; b->frames_cap = 0; b->motions_cap = 0;
    CUDA_OK(cudaMalloc(&b->frames, total * H_FRAME * sizeof(float)));
    CUDA_OK(cudaMalloc(&b->motion_offset, n_motions * sizeof(int)));
    CUDA_OK(cudaMalloc(&b->motion_len, n_motions * sizeof(int)));
    b->frames_cap = total; b->motions_cap = n_motions;
  }
  CUDA_OK(cudaMemcpy(b->frames, frames_host, total * H_FRAME * sizeof(float), cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(b->motion_offset, motion_offset_host, n_motions * sizeof(int), cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(b->motion_len, motion_len_host, n_motions * sizeof(int), cudaMemcpyHostToDevice));
  b->b.frames = b->frames; b->b.motion_offset = b->motion_offset; b->b.motion_len = b->motion_len; b->b.n_motions = n_motions;
  return 0;
}

extern "C" int hoic_set_reward_weights(hoic_batch* b, const float* w_host) {
  if (!b || !w_host) return set_err("hoic_set_reward_weights: bad arguments");
  CUDA_OK(cudaSetDevice(b->device));
  CUDA_OK(cudaMemcpy(b->reward_w, w_host, 24 * sizeof(float), cudaMemcpyHostToDevice));
  return 0;
}

extern "C" int hoic_set_record(hoic_batch* b, float* frames, float* surface) {
  if (!b) return set_err("hoic_set_record: null batch");
  if ((frames == nullptr) != (surface == nullptr)) return set_err("hoic_set_record: pass both buffers or neither");
  b->b.rec_frames = frames;
  b->b.rec_surface = surface;
  return 0;
}

extern "C" int hoic_set_mode(hoic_batch* b, int train) {
  if (!b) return set_err("hoic_set_mode: null batch");
  b->b.mode_train = train ? 1 : 0;
  return 0;
}

#define LAUNCH_CHECK() do { g_launches++; cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return set_err(std::string("kernel launch: ") + cudaGetErrorString(e_)); } while (0)

extern "C" int hoic_set_sampler(hoic_batch* b, const int* model_motion_offset_host, const int* motion_ids_host, int random_start,
                                int min_episode, unsigned long long seed) {
  if (!b || !model_motion_offset_host || !motion_ids_host) return set_err("hoic_set_sampler: bad arguments");
  if (!b->b.frames) return set_err("hoic_set_sampler: call hoic_set_motions first");
  const int total = model_motion_offset_host[b->n_models];
  if (model_motion_offset_host[0] != 0 || total < 1) return set_err("hoic_set_sampler: bad motion table");
  for (int i = 0; i < b->n_models; i++)
    if (model_motion_offset_host[i + 1] <= model_motion_offset_host[i]) return set_err("hoic_set_sampler: every model needs at least one motion");
  for (int i = 0; i < total; i++)
    if (motion_ids_host[i] < 0 || motion_ids_host[i] >= b->b.n_motions) return set_err("hoic_set_sampler: motion id out of range");
  CUDA_OK(cudaSetDevice(b->device));
  CUDA_OK(cudaDeviceSynchronize());
  cudaFree(b->samp_off); cudaFree(b->samp_ids);
  b->samp_off = b->samp_ids = nullptr;
  CUDA_OK(cudaMalloc(&b->samp_off, (b->n_models + 1) * sizeof(int)));
  CUDA_OK(cudaMalloc(&b->samp_ids, total * sizeof(int)));
  CUDA_OK(cudaMemcpy(b->samp_off, model_motion_offset_host, (b->n_models + 1) * sizeof(int), cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(b->samp_ids, motion_ids_host, total * sizeof(int), cudaMemcpyHostToDevice));
  b->b.samp_off = b->samp_off; b->b.samp_ids = b->samp_ids;
  b->b.samp_random_start = random_start ? 1 : 0;
  b->b.samp_min_episode = min_episode;
  b->b.samp_seed = seed;
  return 0;
}

extern "C" int hoic_reset(hoic_batch* b, const uint8_t* mask, const int* motion_id, const int* start_idx, float* obs_out, void* stream) {
  if (!b) return set_err("hoic_reset: null batch");
  if (!b->b.frames) return set_err("hoic_reset: call hoic_set_motions first");
  CUDA_OK(cudaSetDevice(b->device));
  k_reset<<<b->n_envs, 32, b->smem_bytes, (cudaStream_t)stream>>>(b->b, mask, motion_id, start_idx, obs_out, 0, 0);
  LAUNCH_CHECK();
  return 0;
}

extern "C" int hoic_reset_sampled(hoic_batch* b, const uint8_t* mask, float* obs_out, void* stream) {
  if (!b) return set_err("hoic_reset_sampled: null batch");
  if (!b->b.samp_off) return set_err("hoic_reset_sampled: call hoic_set_sampler first");
  CUDA_OK(cudaSetDevice(b->device));
  k_reset<<<b->n_envs, 32, b->smem_bytes, (cudaStream_t)stream>>>(b->b, mask, nullptr, nullptr, obs_out, 1, 0);
  LAUNCH_CHECK();
  return 0;
}

__global__ void k_get_episode(BatchDev B, int* motion_id, int* start_idx, int* cur_t) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B.n_envs) return;
  const int* ints = (const int*)(B.state + (size_t)i * ST_STRIDE + ST_INTS);
  if (cur_t) cur_t[i] = ints[0];
  if (motion_id) motion_id[i] = ints[1];
  if (start_idx) start_idx[i] = ints[2];
}
extern "C" int hoic_get_episode(hoic_batch* b, int* motion_id, int* start_idx, int* cur_t, void* stream) {
  if (!b) return set_err("hoic_get_episode: null batch");
  CUDA_OK(cudaSetDevice(b->device));
  k_get_episode<<<(b->n_envs + 255) / 256, 256, 0, (cudaStream_t)stream>>>(b->b, motion_id, start_idx, cur_t);
  LAUNCH_CHECK();
  return 0;
}

namespace {
// event pair around a launch of kernel class `cls` on `sm` (no-ops unless hoic_set_timing is on)
struct TimedLaunch {
  hoic_batch* b; cudaStream_t sm; bool on; size_t idx = 0;
  TimedLaunch(hoic_batch* b_, int cls, cudaStream_t sm_) : b(b_), sm(sm_), on(b_->timing) {
    if (!on) return;
    idx = b->tev_used;
    b->tev_used += 2;
    while (b->tev.size() < b->tev_used) { cudaEvent_t e; cudaEventCreate(&e); b->tev.push_back(e); }
    if (b->tev_class.size() < b->tev_used / 2) b->tev_class.resize(b->tev_used / 2);
    b->tev_class[idx / 2] = cls;
    cudaEventRecord(b->tev[idx], sm);
  }
  ~TimedLaunch() { if (on) cudaEventRecord(b->tev[idx + 1], sm); }
};
void timing_collect(hoic_batch* b) {
  if (!b->tev_used) return;
  cudaDeviceSynchronize();
  for (size_t i = 0; i + 1 < b->tev_used; i += 2) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, b->tev[i], b->tev[i + 1]) == cudaSuccess) b->tacc[b->tev_class[i / 2]] += ms;
  }
  b->tev_used = 0;
}
}  // namespace

extern "C" int hoic_set_timing(hoic_batch* b, int on) {
  if (!b) return set_err("hoic_set_timing: null batch");
  CUDA_OK(cudaSetDevice(b->device));
  timing_collect(b);
  b->timing = on != 0;
  for (int i = 0; i < HOIC_TIMING_CLASSES; i++) b->tacc[i] = 0;
  b->tsteps = 0;
  return 0;
}
extern "C" int hoic_get_timing(hoic_batch* b, double* ms_host, int* steps_host) {
  if (!b || !ms_host) return set_err("hoic_get_timing: bad arguments");
  CUDA_OK(cudaSetDevice(b->device));
  timing_collect(b);
  for (int i = 0; i < HOIC_TIMING_CLASSES; i++) ms_host[i] = b->tacc[i];
  if (steps_host) *steps_host = b->tsteps;
  return 0;
}

// tick mode: every env starts substep 0 in the k_p12 list of tick 0 (in the step's env order)
__global__ void k_tick_init(BatchDev B) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B.n_envs) return;
  const int env = B.perm ? B.perm[i] : i;
  B.tk_k[env] = 0;
  TK_LIST(B, 0, 0)[i] = env;
  if (i == 0) *TK_COUNT(B, 0, 0) = B.n_envs;
}
// tick mode: an env that is not through its substeps when the ticks are used up fails like a diverged simulation
__global__ void k_tick_finish(BatchDev B) {
  const int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= B.n_envs) return;
  int* misc = (int*)(B.scratch + (size_t)env * G_TOTAL + G_OFF(MISC));
  if (B.tk_k[env] < B.env_model[env]->sim_step && !(misc[MI_FLAGS] & 3)) misc[MI_FLAGS] |= 1;
}
__global__ void k_delay(int ns) {
  const long long t0 = clock64();
  while (clock64() - t0 < (long long)ns * 2) __nanosleep(200);     // ~2 cycles per ns at 1.9 GHz
}

extern "C" int hoic_step(hoic_batch* b, const float* actions, float* obs, float* reward, float* reward_info, uint8_t* done,
                         uint8_t* fail, uint8_t* end, float* percent, int auto_reset, void* stream) {
  if (!b || !actions) return set_err("hoic_step: bad arguments");
  if (!b->b.frames) return set_err("hoic_step: call hoic_set_motions first");
  if (auto_reset && (!done || !b->b.samp_off)) return set_err("hoic_step: auto_reset needs the done output and hoic_set_sampler");
  CUDA_OK(cudaSetDevice(b->device));
  cudaStream_t user = (cudaStream_t)stream;
  if (b->timing) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(user, &cs);
    if (cs != cudaStreamCaptureStatusNone) return set_err("hoic_step: per-kernel timing cannot run under stream capture");
    timing_collect(b);      // the previous call's events
  }
  StepOut o = {obs, reward, reward_info, done, fail, end, percent};
  const int np = b->nparts;
#define STEP_LAUNCH(cls, sm_, call) do { TimedLaunch tl_(b, cls, sm_); cudaError_t e_ = (call); g_launches++; if (e_ != cudaSuccess) return set_err(std::string("kernel launch (" #call "): ") + cudaGetErrorString(e_)); } while (0)
#define STEP_PARTS(cls, call) do { for (int j = 0; j < np; j++) { const StepPart sp = {j, np}; cudaStream_t sm = b->part_stream[j]; STEP_LAUNCH(cls, sm, call); } } while (0)
  {
    TimedLaunch whole(b, 7, user);
    if (b->user_perm) b->b.perm = b->user_perm;
    else if (b->balance) {
      // the order of THIS step was computed at the end of the previous one (below), beside the auto-reset; only a batch's
      // first step sorts at its head
      if (!b->order_valid) {
        TimedLaunch tl(b, 6, user);
        k_order<<<1, 1024, 0, user>>>(b->b, b->order);
        LAUNCH_CHECK();
      }
      b->b.perm = b->order;
    } else b->b.perm = nullptr;
    if (b->b.full_count) CUDA_OK(cudaMemsetAsync(b->b.full_count, 0, (size_t)HOIC_COUNT_INTS * sizeof(int), user));
    if (b->b.solve_budget > 0) CUDA_OK(cudaMemsetAsync(b->b.late_stamp, 0, (size_t)b->n_envs * sizeof(int), user));
    // fork: every part runs on its own stream; the caller's stream waits for all of them at the end
    CUDA_OK(cudaEventRecord(b->ev_fork, user));
    for (int j = 0; j < np; j++) CUDA_OK(cudaStreamWaitEvent(b->part_stream[j], b->ev_fork, 0));
    if (b->part_delay_ns > 0 && np > 1) { k_delay<<<1, 1, 0, b->part_stream[1]>>>(b->part_delay_ns); LAUNCH_CHECK(); }
    if (b->fuse12) STEP_PARTS(0, hoic_launch_p12(b->b, sp, actions, -1, sm));
    else {
      STEP_PARTS(0, b->p1_wide ? hoic_launch_p1w(b->b, sp, actions, -1, sm) : hoic_launch_p1(b->b, sp, actions, -1, sm));
      STEP_PARTS(1, hoic_launch_p2(b->b, sp, -1, sm));
    }
    const bool tick = b->ticks > 0 && b->b.tk_k;
    if (tick) {
      // Tick mode (opt-in experiment, HOIC_TICKS=T HOIC_TICK_BUDGET=M): an env-asynchronous schedule out of plain launches.
      // The parts' pass -1 launches above have run; from here ONE chain: every tick runs k_p12 (smooth dynamics + collision +
      // rows of an env's own next substep) for the envs that are due and a solver launch of at most M Newton iterations for the
      // envs that are in a solve, side by side; an env moves to the other list when it is through.  Bit-identical results;
      // MEASURED SLOWER (box, 4 096 envs: 5.9 ms with T = 64, M = 3 against 4.3 ms): a tick's k_p12 over the ~2 000 envs that are due
      // is a throughput-bound 110 us launch, not one env's 60 us chain (profiles/README.md).
      cudaStream_t M0 = b->part_stream[0], A = b->aux_stream[0], C = b->carry_stream[0];
      for (int j = 1; j < np; j++) { CUDA_OK(cudaEventRecord(b->ev_join[j], b->part_stream[j])); CUDA_OK(cudaStreamWaitEvent(M0, b->ev_join[j], 0)); }
      CUDA_OK(cudaMemsetAsync(b->b.tk_count, 0, (size_t)(b->ticks + 2) * 4 * sizeof(int), M0));
      k_tick_init<<<(b->n_envs + 255) / 256, 256, 0, M0>>>(b->b);
      LAUNCH_CHECK();
      const StepPart tk = {0, 1, 3};
      for (int t = 0; t < b->ticks; t++) {
        CUDA_OK(cudaEventRecord(b->ev_rows[0], M0));
        CUDA_OK(cudaStreamWaitEvent(A, b->ev_rows[0], 0));
        CUDA_OK(cudaStreamWaitEvent(C, b->ev_rows[0], 0));
        STEP_LAUNCH(0, M0, hoic_launch_p12(b->b, tk, actions, t, M0));
        STEP_LAUNCH(2, A, hoic_launch_p3c(b->b, tk, t, A));
        STEP_LAUNCH(3, C, hoic_launch_p3fc(b->b, tk, t, C));
        CUDA_OK(cudaEventRecord(b->ev_aux[0], A));
        CUDA_OK(cudaEventRecord(b->ev_carry[0], C));
        CUDA_OK(cudaStreamWaitEvent(M0, b->ev_aux[0], 0));
        CUDA_OK(cudaStreamWaitEvent(M0, b->ev_carry[0], 0));
      }
      k_tick_finish<<<(b->n_envs + 255) / 256, 256, 0, M0>>>(b->b);
      LAUNCH_CHECK();
      // the other parts' streams continue after the ticks (their k_tail launches below)
      CUDA_OK(cudaEventRecord(b->ev_rows[0], M0));
      for (int j = 1; j < np; j++) CUDA_OK(cudaStreamWaitEvent(b->part_stream[j], b->ev_rows[0], 0));
    }
    const bool carry = !tick && b->b.solve_budget > 0 && !b->fuse && !b->fuse12;
    for (int k = 0; carry && k < b->nsub; k++) {
      // Carry mode (opt-in experiment, HOIC_SOLVE_BUDGET=n; bit-identical results).  A launch of the solver lasts as long as its
      // slowest solve (7-8 Newton iterations in SOME env of the part, every substep), and k_p1 / k_p2 of the next substep wait for
      // it.  Here k_p3s stops after `solve_budget` iterations; the few solves that are still open (and the full tier's envs) are
      // finished on the carry stream while k_p1 / k_p2 of substep k + 1 already run for everybody else (group 1), followed by
      // k_p1 / k_p2 for those few (group 2).  MEASURED SLOWER (box, 4 096 envs, budget 3 / 4 / 6: 4.88 / 4.67 / 4.48 vs 4.29 ms per
      // step): the late group's chain -- the straggler's own solve, now split, + its own k_p1 + k_p2 + three more launches -- is as
      // long as the lock-step chain it replaces, because the same env is rarely the straggler twice in a row and every substep
      // still joins.  What the split solver and the list-mode kernels are for is a schedule without that join.
      auto p1 = [&](const StepPart& sp, int kk, cudaStream_t sm) {
        return b->p1_wide ? hoic_launch_p1w(b->b, sp, actions, kk, sm) : hoic_launch_p1(b->b, sp, actions, kk, sm);
      };
      for (int j = 0; j < np; j++) {
        const StepPart all = {j, np, 0}, early = {j, np, 1}, late = {j, np, 2};
        cudaStream_t S = b->part_stream[j], X = b->aux_stream[j], Y = b->carry_stream[j];
        if (k == 0) {
          STEP_LAUNCH(0, S, p1(all, 0, S));
          STEP_LAUNCH(1, S, hoic_launch_p2(b->b, all, 0, S));
        } else CUDA_OK(cudaStreamWaitEvent(S, b->ev_carry[j], 0));          // the late group's rows of substep k
        CUDA_OK(cudaEventRecord(b->ev_rows[j], S));
        CUDA_OK(cudaStreamWaitEvent(X, b->ev_rows[j], 0));
        STEP_LAUNCH(3, X, hoic_launch_p3(b->b, all, k, X));
        CUDA_OK(cudaEventRecord(b->ev_aux[j], X));
        STEP_LAUNCH(2, S, hoic_launch_p3sb(b->b, all, k, S));
        CUDA_OK(cudaEventRecord(b->ev_p3s[j], S));
        CUDA_OK(cudaStreamWaitEvent(Y, b->ev_p3s[j], 0));
        CUDA_OK(cudaStreamWaitEvent(Y, b->ev_aux[j], 0));
        STEP_LAUNCH(2, Y, hoic_launch_p3c(b->b, all, k, Y));
        if (k + 1 < b->nsub) {
          STEP_LAUNCH(0, Y, p1(late, k + 1, Y));
          STEP_LAUNCH(1, Y, hoic_launch_p2(b->b, late, k + 1, Y));
          STEP_LAUNCH(0, S, p1(early, k + 1, S));
          STEP_LAUNCH(1, S, hoic_launch_p2(b->b, early, k + 1, S));
        }
        CUDA_OK(cudaEventRecord(b->ev_carry[j], Y));
        if (k + 1 == b->nsub) CUDA_OK(cudaStreamWaitEvent(S, b->ev_carry[j], 0));
      }
    }
    for (int k = 0; !carry && !tick && k < b->nsub; k++) {
      // fused mode: k_p1 of substep k >= 1 ran inside the previous substep's solve launch (k_a.cu)
      if (b->fuse12) STEP_PARTS(0, hoic_launch_p12(b->b, sp, actions, k, sm));
      else {
        if (!(b->fuse && k > 0)) STEP_PARTS(0, b->p1_wide ? hoic_launch_p1w(b->b, sp, actions, k, sm) : hoic_launch_p1(b->b, sp, actions, k, sm));
        STEP_PARTS(1, hoic_launch_p2(b->b, sp, k, sm));
      }
      // the two capacity tiers of the solver take disjoint envs: side by side (the few long solves of the full tier would
      // otherwise sit on every substep's critical path)
      for (int j = 0; j < np; j++) {
        const StepPart sp = {j, np};
        CUDA_OK(cudaEventRecord(b->ev_rows[j], b->part_stream[j]));
        CUDA_OK(cudaStreamWaitEvent(b->aux_stream[j], b->ev_rows[j], 0));
        if (b->fuse) STEP_LAUNCH(3, b->aux_stream[j], hoic_launch_a(b->b, sp, actions, k, b->aux_stream[j]));
        else STEP_LAUNCH(3, b->aux_stream[j], hoic_launch_p3(b->b, sp, k, b->aux_stream[j]));
        CUDA_OK(cudaEventRecord(b->ev_aux[j], b->aux_stream[j]));
        if (b->fuse) STEP_LAUNCH(2, b->part_stream[j], hoic_launch_as(b->b, sp, actions, k, b->part_stream[j]));
        else STEP_LAUNCH(2, b->part_stream[j], hoic_launch_p3s(b->b, sp, k, b->part_stream[j]));
        CUDA_OK(cudaStreamWaitEvent(b->part_stream[j], b->ev_aux[j], 0));
      }
    }
    STEP_PARTS(4, hoic_launch_tail(b->b, sp, actions, o, sm));
    for (int j = 0; j < np; j++) { CUDA_OK(cudaEventRecord(b->ev_join[j], b->part_stream[j])); CUDA_OK(cudaStreamWaitEvent(user, b->ev_join[j], 0)); }
    const bool order_next = !b->user_perm && b->balance;
    if (order_next) {
      // env order of the NEXT step from this step's cost proxies (k_tail wrote them), on a side stream beside the auto-reset
      CUDA_OK(cudaEventRecord(b->ev_tail, user));
      CUDA_OK(cudaStreamWaitEvent(b->aux_stream[0], b->ev_tail, 0));
      {
        TimedLaunch tl(b, 6, b->aux_stream[0]);
        k_order<<<1, 1024, 0, b->aux_stream[0]>>>(b->b, b->order);
        LAUNCH_CHECK();
      }
      CUDA_OK(cudaEventRecord(b->ev_aux[0], b->aux_stream[0]));
      b->order_valid = true;
    }
    if (auto_reset) {
      // the sampler's episode loop (agent_handmimic.py:437-454): a finished env restarts on a freshly drawn (motion, start)
      // and its returned observation is the reset observation; only the envs k_tail listed as finished get a CTA
      TimedLaunch tl(b, 5, user);
      if (b->b.reset_list) k_reset<<<std::min(b->n_envs, 296), 32, b->smem_bytes, user>>>(b->b, nullptr, nullptr, nullptr, obs, 1, 1);
      else k_reset<<<b->n_envs, 32, b->smem_bytes, user>>>(b->b, done, nullptr, nullptr, obs, 1, 0);
      LAUNCH_CHECK();
    }
    if (order_next) CUDA_OK(cudaStreamWaitEvent(user, b->ev_aux[0], 0));
  }
#undef STEP_PARTS
#undef STEP_LAUNCH
  if (b->timing) b->tsteps++;
  return 0;
}

extern "C" int hoic_get_state(hoic_batch* b, float* qpos, float* qvel, void* stream) {
  if (!b) return set_err("hoic_get_state: null batch");
  CUDA_OK(cudaSetDevice(b->device));
  k_get_state<<<(b->n_envs * 33 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(b->b, qpos, qvel, nullptr);
  LAUNCH_CHECK();
  return 0;
}
extern "C" int hoic_set_state(hoic_batch* b, const float* qpos, const float* qvel, void* stream) {
  if (!b) return set_err("hoic_set_state: null batch");
  CUDA_OK(cudaSetDevice(b->device));
  k_set_state<<<(b->n_envs * 33 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(b->b, qpos, qvel, nullptr, nullptr);
  LAUNCH_CHECK();
  return 0;
}
extern "C" int hoic_set_cur_t(hoic_batch* b, const int* cur_t, void* stream) {
  if (!b || !cur_t) return set_err("hoic_set_cur_t: bad arguments");
  CUDA_OK(cudaSetDevice(b->device));
  k_set_state<<<(b->n_envs * 33 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(b->b, nullptr, nullptr, nullptr, cur_t);
  LAUNCH_CHECK();
  return 0;
}
extern "C" int hoic_set_episode(hoic_batch* b, const int* motion_id, const int* start_idx, const int* cur_t, void* stream) {
  if (!b) return set_err("hoic_set_episode: null batch");
  if (!b->b.frames) return set_err("hoic_set_episode: call hoic_set_motions first");
  CUDA_OK(cudaSetDevice(b->device));
  k_set_episode<<<(b->n_envs + 255) / 256, 256, 0, (cudaStream_t)stream>>>(b->b, motion_id, start_idx, cur_t);
  LAUNCH_CHECK();
  return 0;
}
extern "C" int hoic_push_frame(hoic_batch* b, const float* frame, int per_motion, void* stream) {
  if (!b || !frame) return set_err("hoic_push_frame: bad arguments");
  if (!b->b.frames) return set_err("hoic_push_frame: call hoic_set_motions first");
  if (b->min_motion_len < 4) return set_err("hoic_push_frame: every motion must hold a window of at least 4 frames [w0, w0, .., wK, wK]");
  CUDA_OK(cudaSetDevice(b->device));
  k_push_frame<<<b->b.n_motions, 32, 0, (cudaStream_t)stream>>>(b->frames, b->motion_offset, b->motion_len, b->b.n_motions, frame, per_motion);
  LAUNCH_CHECK();
  return 0;
}
extern "C" int hoic_get_warmstart(hoic_batch* b, float* warm, void* stream) {
  if (!b || !warm) return set_err("hoic_get_warmstart: bad arguments");
  CUDA_OK(cudaSetDevice(b->device));
  k_get_state<<<(b->n_envs * 33 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(b->b, nullptr, nullptr, warm);
  LAUNCH_CHECK();
  return 0;
}
extern "C" int hoic_set_warmstart(hoic_batch* b, const float* warm, void* stream) {
  if (!b || !warm) return set_err("hoic_set_warmstart: bad arguments");
  CUDA_OK(cudaSetDevice(b->device));
  k_set_state<<<(b->n_envs * 33 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(b->b, nullptr, nullptr, warm, nullptr);
  LAUNCH_CHECK();
  return 0;
}
extern "C" int hoic_get_contacts(hoic_batch* b, int* ncon, int* geom_pairs, float* dist_pos_frame, void* stream) {
  if (!b || !ncon) return set_err("hoic_get_contacts: bad arguments");
  CUDA_OK(cudaSetDevice(b->device));
  k_forward_debug<<<b->n_envs, 32, b->smem_bytes, (cudaStream_t)stream>>>(b->b, nullptr, nullptr, nullptr, nullptr, ncon, geom_pairs, dist_pos_frame);
  LAUNCH_CHECK();
  return 0;
}
extern "C" int hoic_get_derived(hoic_batch* b, float* xpos, float* xquat, float* qM, float* qfrc_bias, void* stream) {
  if (!b) return set_err("hoic_get_derived: null batch");
  CUDA_OK(cudaSetDevice(b->device));
  k_forward_debug<<<b->n_envs, 32, b->smem_bytes, (cudaStream_t)stream>>>(b->b, xpos, xquat, qM, qfrc_bias, nullptr, nullptr, nullptr);
  LAUNCH_CHECK();
  return 0;
}
extern "C" int hoic_substep_debug(hoic_batch* b, const float* ctrl, const float* qfrc_applied, float* qacc, float* qfrc_constraint,
                                  int* stats, void* stream) {
  if (!b || !ctrl || !qfrc_applied) return set_err("hoic_substep_debug: bad arguments");
  CUDA_OK(cudaSetDevice(b->device));
  k_substep_debug<<<b->n_envs, 32, b->smem_bytes, (cudaStream_t)stream>>>(b->b, ctrl, qfrc_applied, qacc, qfrc_constraint, stats);
  LAUNCH_CHECK();
  return 0;
}
extern "C" int hoic_get_rfc(hoic_batch* b, float* out, void* stream) {
  if (!b || !out) return set_err("hoic_get_rfc: bad arguments");
  CUDA_OK(cudaSetDevice(b->device));
  k_get_rfc<<<(b->n_envs * 8 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(b->b, out);
  LAUNCH_CHECK();
  return 0;
}
__global__ void k_get_work(BatchDev B, int* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B.n_envs) out[i] = ((const int*)(B.state + (size_t)i * ST_STRIDE + ST_INTS))[5];
}
extern "C" int hoic_get_work(hoic_batch* b, int* work, void* stream) {
  if (!b || !work) return set_err("hoic_get_work: bad arguments");
  CUDA_OK(cudaSetDevice(b->device));
  k_get_work<<<(b->n_envs + 255) / 256, 256, 0, (cudaStream_t)stream>>>(b->b, work);
  LAUNCH_CHECK();
  return 0;
}
extern "C" int hoic_set_env_order(hoic_batch* b, const int* perm) {
  if (!b) return set_err("hoic_set_env_order: null batch");
  b->user_perm = perm;
  return 0;
}
extern "C" int hoic_set_balance(hoic_batch* b, int on) {
  if (!b) return set_err("hoic_set_balance: null batch");
  b->balance = on ? 1 : 0;
  return 0;
}
extern "C" int hoic_get_counters(hoic_batch* b, hoic_counters* out_host) {
  if (!b || !out_host) return set_err("hoic_get_counters: bad arguments");
  CUDA_OK(cudaSetDevice(b->device));
  CUDA_OK(cudaDeviceSynchronize());
  unsigned long long h[16];
  CUDA_OK(cudaMemcpy(h, b->b.counters, sizeof h, cudaMemcpyDeviceToHost));
  out_host->env_steps = h[0]; out_host->substeps = h[1]; out_host->ncon = h[2]; out_host->nefc = h[3]; out_host->newton_iters = h[4];
  out_host->ls_evals = h[5]; out_host->mpr_calls = h[6]; out_host->mpr_iters = h[7]; out_host->qp_iters = h[8];
  out_host->fail_physics = h[9]; out_host->fail_overflow = h[10]; out_host->chol_calls = h[11];
  out_host->max_ncon = h[12]; out_host->max_nbasis = h[13]; out_host->max_nedge = h[14];
  return 0;
}

// ---- FP32 FMA peak micro-benchmark: 16 independent accumulators per thread, full-chip grid
__global__ void k_fma_peak(float* out, int iters, float a, float b) {
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; i++) x[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = fmaf(x[i], a, b);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; i++) s += x[i];
  if (s == 12345.678f) out[0] = s;
}
extern "C" int hoic_solve_wrench_qp(int device, int n_problems, int max_contacts, int surface_contact, const float* contacts,
                                    const int* n_contacts, const float* weights, const float* penalties, const float* target,
                                    const float* center, float mu, float w_t, float* rest, float* score, float* surface, int* iters,
                                    void* stream) {
  if (n_problems <= 0 || !contacts || !n_contacts || !target || !center) return set_err("hoic_solve_wrench_qp: bad arguments");
  const int ec = surface_contact ? 5 : 1;
  if (max_contacts < 1 || max_contacts > H_NHG) return set_err("hoic_solve_wrench_qp: max_contacts must be in [1, 19] (380 QP variables)");
  if (!(mu > 0.f) || !(w_t > 0.f)) return set_err("hoic_solve_wrench_qp: mu and w_t must be positive");
  CUDA_OK(cudaSetDevice(device));
  k_wrench_qp<<<n_problems, 32, (size_t)S_TOTAL * sizeof(float), (cudaStream_t)stream>>>(n_problems, max_contacts, ec, contacts, n_contacts,
                                                                                         weights, penalties, target, center, mu, w_t, rest,
                                                                                         score, surface, iters);
  LAUNCH_CHECK();
  return 0;
}

extern "C" int hoic_preprocess_seq(const hoic_model* m, const float* hand_pose_seq, const float* obj_pose_seq, int T, float motion_freq,
                                   int clip, float* frames_out, int* contact_mask, float* contact_pos, void* stream) {
  if (!m || !hand_pose_seq || !obj_pose_seq || !frames_out) return set_err("hoic_preprocess_seq: bad arguments");
  if (T < 2) return set_err("hoic_preprocess_seq: a sequence needs at least 2 frames (finite differences)");
  if ((contact_mask == nullptr) != (contact_pos == nullptr)) return set_err("hoic_preprocess_seq: pass both contact outputs or neither");
  CUDA_OK(cudaSetDevice(m->device));
  k_preprocess<<<T, 32, (size_t)S_TOTAL * sizeof(float), (cudaStream_t)stream>>>(m->dev, hand_pose_seq, obj_pose_seq, T, motion_freq, clip,
                                                                                 frames_out, contact_mask, contact_pos);
  LAUNCH_CHECK();
  return 0;
}

extern "C" int hoic_measure_fp32_peak(int device, double* tflops_host) {
  if (!tflops_host) return set_err("hoic_measure_fp32_peak: null output");
  CUDA_OK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_OK(cudaGetDeviceProperties(&prop, device));
  float* d;
  CUDA_OK(cudaMalloc(&d, 4));
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0);
    k_fma_peak<<<blocks, threads>>>(d, iters, 0.999f, 0.001f);
    g_launches++;
    cudaEventRecord(e1);
    CUDA_OK(cudaEventSynchronize(e1));
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double tf = 2.0 * 16.0 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(d);
  *tflops_host = best;
  return 0;
}

/* hoic.h -- C ABI of libhoic_b200.so: batched HandMimic environments on one B200 (sm_100a).
 *
 * This is the drop-in boundary for the reference's native layer on the rollout hot path.  The reference has no
 * plugin registry: its env (uhc/envs/ho_im4.py) drives MuJoCo 2.1 through mujoco_py objects.  Each entry point below
 * names the reference interface it replaces (paths relative to the reference root).  INTEGRATION.md shows the
 * ctypes binding a maintainer adds on the reference side.
 *
 * Conventions: every function returns 0 on success and <0 on error (`hoic_last_error()` gives the message);
 * all array pointers are DEVICE pointers owned by the caller unless the name ends in `_host`; launches are
 * asynchronous on the caller-supplied stream (a `cudaStream_t` passed as void*; NULL = legacy default stream);
 * one host thread per batch, no internal threads, no global state besides the handles and the error string.
 * Floating-point I/O is float32; the arithmetic type of the kernels is float32 (float64 only inside the
 * residual-force QP).
 */
#ifndef HOIC_H
#define HOIC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HOIC_OBS_DIM 617     /* get_full_obs_v5, ho_im4.py:280-356 */
#define HOIC_ACT_DIM 32      /* 26 PD targets + 6 residual force/torque, ho_im4.py:139-161 */
#define HOIC_FRAME_DIM 212   /* one expert frame, dataset_singledepth.py:239-288 (layout: hoic_b200/motion.py) */
#define HOIC_INFO_DIM 9      /* ho_mimic_reward_9 info vector, ho_reward.py:1043-1046 */
#define HOIC_NQ 33
#define HOIC_NV 32
#define HOIC_NHG 19          /* hand collision geoms (hand_geom_range, ho_im4.py:80-87) */
#define HOIC_MAX_CONTACTS 32 /* contact slots per env in hoic_get_contacts buffers = the kernels' contact capacity; they hold 32
                              * contacts, 80 contact basis rows and 128 pyramid edges per env (the MJCF's nconmax / njmax are
                              * 100 / 500): an env that needs more is failed like MuJoCo's full-buffer warning does, see
                              * hoic_counters.max_* (high-water marks) and fail_overflow */
/* test-mode recording (MotionData, ho_im4.py:27-41).  One frame per physics substep (compute_data_frame :599-609):
 *   [0,63) hand_kps = body_xpos of the 21 hand bodies   [63,89) hand_qpos   [89,95) compensate_ft (obj_vf, obj_vt)
 *   [95,102) obj_pose   [102,108) obj_vel   [108,114) obj_acc = qacc[-6:]   [114] number of hand-object contacts
 *   [115] 1 if the substep ran   then per contact 19 floats: pos[3] frame[9] mj_contactForce[6] hand geom id */
#define HOIC_REC_MAXCON 32
#define HOIC_REC_CON 116
#define HOIC_REC_STRIDE (HOIC_REC_CON + 19 * HOIC_REC_MAXCON)
/* ... and one record per env step from solve_rfc (ho_im4.py:1072-1081):
 *   [0] number of expanded contact points  [1,7) rest_obj_ft  [7] 1 if the QP ran (there was a contact)
 *   then per point 15 floats: surface_contact_frames pos[3] frame[9], surface_contact_force[3] */
#define HOIC_SURF_MAXPT 95
#define HOIC_SURF_PT 8
#define HOIC_SURF_STRIDE (HOIC_SURF_PT + 15 * HOIC_SURF_MAXPT + 3)

typedef struct hoic_model hoic_model;
typedef struct hoic_batch hoic_batch;

typedef struct hoic_counters {
  unsigned long long env_steps, substeps, ncon, nefc, newton_iters, ls_evals, mpr_calls, mpr_iters, qp_iters;
  unsigned long long fail_physics, fail_overflow, chol_calls;
  unsigned long long max_ncon, max_nbasis, max_nedge;   /* high-water marks against the kernels' per-env capacities */
} hoic_counters;

const char* hoic_last_error(void);
int hoic_version(void);

/* replaces mujoco_py.load_model_from_path (uhc/khrylib/rl/envs/common/mujoco_env.py:15-34): the MJCF is compiled on
 * the host by hoic_b200.mjcf into a flat blob (include/hoic_blob.h); this uploads it. `blob_host` is host memory. */
hoic_model* hoic_model_create(const void* blob_host, size_t nbytes, int device);
void hoic_model_destroy(hoic_model* m);

/* replaces one mujoco_py.MjSim per forked sampler process (uhc/agents/agent_handmimic.py:421-428, 503-535):
 * n_envs environments on `device`, env i using models[env_model_id_host[i]] (NULL = all model 0). */
hoic_batch* hoic_batch_create(const hoic_model* const* models, int n_models, const int* env_model_id_host, int n_envs,
                              int device);
void hoic_batch_destroy(hoic_batch* b);
int hoic_batch_size(const hoic_batch* b);

/* replaces env.set_expert(load_seq(...)) (ho_im4.py:135-137, agent_handmimic.py:450-452): all motions, concatenated
 * frames_host[n_frames_total][212] (float32, host), motion m = frames [offset[m], offset[m]+len[m]). */
int hoic_set_motions(hoic_batch* b, const float* frames_host, const int* motion_offset_host, const int* motion_len_host,
                     int n_motions);

/* replaces DatasetSingleDepth.preprocess_seq for one raw sequence (uhc/data_loaders/dataset_singledepth.py:78-142): clips the
 * hand dofs to the joint ranges when clip != 0 (:100; the raw sequence of :109 is processed with clip = 0), finite-difference velocities (compute_vel_from_seq :152-185), FK of the 21 hand
 * bodies (compute_body_pos_quat_from_seq :222-237) and, when the contact outputs are given, the per-frame hand-object
 * contact pre-pass (compute_contact_info :187-220; contact_info[g1] = contact.pos, indexed here by hand geom slot
 * g1 - hand_geom0).  hand_pose_seq[T][26], obj_pose_seq[T][7] (xyz + wxyz) -> frames_out[T][HOIC_FRAME_DIM],
 * contact_mask[T][HOIC_NHG], contact_pos[T][HOIC_NHG][3]; all device pointers; one warp per frame. */
int hoic_preprocess_seq(const hoic_model* m, const float* hand_pose_seq, const float* obj_pose_seq, int T, float motion_freq,
                        int clip, float* frames_out, int* contact_mask, float* contact_pos, void* stream);

/* replaces PhysMetrics.solve_force (scripts/metrics.py:156-196) and exposes the QP of solve_rfc (uhc/envs/ho_im4.py:1003-1070) on
 * caller-supplied contacts, batched, one warp per problem:
 *   min_x |F - J_f x|^2 + w_t |T - J_t x|^2 + c.x + 0.5e-7 |x|^2,  x >= 0,
 * x = 4 friction-pyramid edge forces (coefficient mu) at each contact, expanded to 5 points (+-0.0025 along the tangents)
 * when surface_contact != 0; torques about center[b].  contacts[B][max_contacts][12] = pos[3] frame[9] (rows: normal,
 * tangent, bitangent), n_contacts[B] <= max_contacts <= 19, weights[B][max_contacts] (NULL = 1: the `ts` factor of
 * ho_im4.py:1012-1013), penalties[B][max_contacts * points * 4] (NULL = 0: p_n + p_t of :1028-1042), target[B][6] = F, T.
 * Outputs (any may be NULL): rest[B][6] = F - J_f x, T - J_t x; score[B] = |rest_f| + sqrt(w_t) |rest_t|;
 * surface[B][HOIC_SURF_STRIDE] (layout above: points, their frames and forces); iters[B] = Newton iterations.
 * The reference's settings: solve_rfc mu 0.75, w_t 1e4; solve_force mu 1, w_t 1, no weights, no penalties. */
int hoic_solve_wrench_qp(int device, int n_problems, int max_contacts, int surface_contact, const float* contacts,
                         const int* n_contacts, const float* weights, const float* penalties, const float* target,
                         const float* center, float mu, float w_t, float* rest, float* score, float* surface, int* iters,
                         void* stream);

/* replaces cfg.reward_weights after cfg.update_adaptive_params (uhc/utils/config_utils/handmimic_config.py:157-195):
 * w_host[24] = w_p w_wp w_v w_j w_op w_or w_ov w_orfc k_p k_wp k_v k_j k_op k_or k_ov k_orfc, rest reserved. */
int hoic_set_reward_weights(hoic_batch* b, const float* w_host);

/* replaces env.set_mode (ho_im4.py:132): train=1 applies the tracking-error termination of ho_im4.py:655-656 */
int hoic_set_mode(hoic_batch* b, int train);

/* replaces the MotionData lists filled in mode == 'test' (ho_im4.py:547-548 compute_data_frame, :1072-1081): every
 * following hoic_step writes frames[N][sim_step][HOIC_REC_STRIDE] and surface[N][HOIC_SURF_STRIDE] (device buffers kept
 * by reference; NULL, NULL switches recording off).  Contact forces are mj_contactForce of the pyramidal rows. */
int hoic_set_record(hoic_batch* b, float* frames, float* surface);

/* replaces env.reset() (mujoco_env.py:95-105 + ho_im4.py:690-716): for every env with mask[i]!=0 (mask NULL = all)
 * take expert = motion[motion_id[i]][start_idx[i]:], state <- its frame 0, mj_resetData + forward, cur_t = 0,
 * obs_out[i] <- observation.  motion_id/start_idx NULL = keep the env's current assignment. */
int hoic_reset(hoic_batch* b, const uint8_t* mask, const int* motion_id, const int* start_idx, float* obs_out,
               void* stream);

/* replaces the sampler's draw of the next episode (uhc/agents/agent_handmimic.py:437-452: seq_idx = randint(0, seq_num - 1),
 * start_idx = randint(0, len - 200), env.set_expert(load_seq(...))) with an on-device draw: env i picks a uniform motion
 * among those of its model, motion_ids_host[model_motion_offset_host[m] .. model_motion_offset_host[m + 1]) for model m
 * (n_models + 1 offsets), and start_idx ~ U{0 .. len - min_episode - 1} (0 when random_start == 0).  The draws are a
 * counter-based hash of (seed, env, episode number of the env): reproducible and independent of the batch order. */
int hoic_set_sampler(hoic_batch* b, const int* model_motion_offset_host, const int* motion_ids_host, int random_start,
                     int min_episode, unsigned long long seed);
/* hoic_reset with (motion, start_idx) drawn by the sampler for every env with mask[i] != 0 (mask NULL = all) */
int hoic_reset_sampled(hoic_batch* b, const uint8_t* mask, float* obs_out, void* stream);
/* env.cur_t / the current expert assignment of every env (any pointer may be NULL) */
int hoic_get_episode(hoic_batch* b, int* motion_id /*[N]*/, int* start_idx /*[N]*/, int* cur_t /*[N]*/, void* stream);

/* replaces env.step(a) + custom_reward (ho_im4.py:611-662, ho_reward.py:943-1046, agent_handmimic.py:467-473):
 * 15 x [stable PD, gravity compensation + residual wrench, contact recording, mj_step], residual-force QP,
 * termination, reward 9, observation.  Any output pointer may be NULL.
 * auto_reset != 0 also runs the sampler's episode loop (agent_handmimic.py:437-454) for the envs that finished: they are
 * reset to a (motion, start_idx) drawn by hoic_set_sampler and obs[i] is the reset observation (reward / done / fail /
 * end / percent still describe the finished transition).  Needs `done`.
 * A simulation that blows up (non-finite state, failed factorisation: MuJoCo's mjWARN_BADQACC -> MujocoException ->
 * fail, ho_im4.py:635-637) is reported as fail with reward 0 and zero info / observation, and the env keeps the state it
 * had before the step.
 * The call launches a pipeline of kernels (per physics substep: smooth dynamics, collision + constraint rows, Newton
 * solve) on internal streams forked from / joined to `stream`; it is capturable into a CUDA graph. */
int hoic_step(hoic_batch* b, const float* actions /*[N][32]*/, float* obs /*[N][617]*/, float* reward /*[N]*/,
              float* reward_info /*[N][9]*/, uint8_t* done /*[N]*/, uint8_t* fail /*[N]*/, uint8_t* end /*[N]*/,
              float* percent /*[N]*/, int auto_reset, void* stream);

/* replaces sim.get_state() / data.qpos, data.qvel reads (mujoco_env.py:109, ho_im4.py:794-804) */
int hoic_get_state(hoic_batch* b, float* qpos /*[N][33]*/, float* qvel /*[N][32]*/, void* stream);
/* replaces env.set_state(qpos, qvel) = sim.set_state + sim.forward (mujoco_env.py:107-114); keeps cur_t */
int hoic_set_state(hoic_batch* b, const float* qpos, const float* qvel, void* stream);
/* sets env.cur_t (parity tests start from arbitrary expert frames) */
int hoic_set_cur_t(hoic_batch* b, const int* cur_t, void* stream);
/* env.cur_t and the expert assignment (motion, start_idx) of every env without touching its physics state; any pointer NULL */
int hoic_set_episode(hoic_batch* b, const int* motion_id, const int* start_idx, const int* cur_t, void* stream);
/* replaces HandObjMimicTest.insert_new_frame (uhc/envs/ho_im_test.py:34-38; driven by InferenceServer/RLTest.py:254-303) on the
 * device: every motion of the batch is a sliding expert window stored as [w0, w0, w1, ..., wK, wK]; the call drops w0 and
 * appends the new frame -> [w1, w1, w2, ..., wK, f, f].  frame: [n_motions][212] (per_motion != 0) or one [212] frame for all,
 * device memory.  With that layout `hoic_reset(start_idx = 1)` sees [w0 .. wK, wK] and a step from (start_idx 0, cur_t 0) sees
 * [w0, w0 .. wK]: exactly the window-relative accessors of ho_im_test.py:40-75 (hoic_b200/track.py). */
int hoic_push_frame(hoic_batch* b, const float* frame, int per_motion, void* stream);
/* qacc_warmstart in/out (data.qacc_warmstart); parity entry */
int hoic_get_warmstart(hoic_batch* b, float* warm /*[N][32]*/, void* stream);
int hoic_set_warmstart(hoic_batch* b, const float* warm, void* stream);

/* replaces reading data.contact[:ncon] (ho_im4.py:883-889) after the LAST forward: ncon[N], geom pairs [N][32][2],
 * and [N][32][13] = dist, pos[3], frame[9] */
int hoic_get_contacts(hoic_batch* b, int* ncon, int* geom_pairs, float* dist_pos_frame, void* stream);

/* one mj_step (sim.step(), ho_im4.py:545) from the current state with explicit ctrl[N][26] and qfrc_applied[N][32];
 * optional outputs qacc[N][32], qfrc_constraint[N][32], stats[N][4] = nefc, ncon, newton iterations, fail */
int hoic_substep_debug(hoic_batch* b, const float* ctrl, const float* qfrc_applied, float* qacc, float* qfrc_constraint,
                       int* stats, void* stream);
/* derived quantities of the last forward (data.body_xpos / body_xquat / qM / qfrc_bias), any pointer may be NULL:
 * xpos[N][25][3], xquat[N][25][4], qM[N][32][32], qfrc_bias[N][32] */
int hoic_get_derived(hoic_batch* b, float* xpos, float* xquat, float* qM, float* qfrc_bias, void* stream);
/* per-step scalars of the last hoic_step: [N][8] = rfc_score, rest_force[3], rest_torque[3], n contact geoms */
int hoic_get_rfc(hoic_batch* b, float* out, void* stream);

/* load balancing: every kernel of the step pipeline schedules one env per CTA in `order`; by default (batches of 8+
 * envs) hoic_step sorts the envs by the cost proxy of their last step, most expensive first (longest-processing-time
 * order), in a small kernel of its own.  hoic_set_balance(0) switches that off (identity order); hoic_set_env_order
 * installs a caller-owned device permutation instead (NULL = back to the library's choice); hoic_get_work exports the
 * cost proxy work[N] (solver iterations, contacts, MPR iterations, QP size).  Results do not depend on the order. */
int hoic_get_work(hoic_batch* b, int* work /*[N]*/, void* stream);
int hoic_set_env_order(hoic_batch* b, const int* perm /*[N] device, kept by reference*/);
int hoic_set_balance(hoic_batch* b, int on);

/* replaces running_state(state) = ZFilter.__call__ / RunningStat.push of the sampler (uhc/khrylib/utils/zfilter.py:14-24, 59-68;
 * agent_handmimic.py:455-457, 469-470) for the whole batch obs[n][dim]: with update != 0 the batch is merged into the running
 * statistics n / mean[dim] / M2[dim] (double, device; Chan's parallel Welford merge = what n successive push calls give) and,
 * if given, into a second since-last-sync triple dn / dmean / dM2 (multi-rank pooling); then
 * x_out = clip((obs - mean) / (std + 1e-8), -clip, clip) with std = sqrt(M2 / (n - 1)) (|mean| while n <= 1).
 * scratch: double[64 dim + 1], zero-initialised once by the caller; shift_scale: float[2 dim] work buffer.  Results are bitwise
 * reproducible (fixed summation order). */
int hoic_zfilter(const float* obs, int n, int dim, double* st_n, double* mean, double* m2, double* dn, double* dmean, double* dm2,
                 double* scratch, float* shift_scale, float clip, int update, float* x_out, void* stream);
/* hoic_zfilter with a second output for the policy MLP's first GEMM on the tensor cores: x16_out[n][ld16] = bfloat16(x_out)
 * (round to nearest even), columns dim..ld16-1 zero (ld16 >= dim).  x16_out NULL = hoic_zfilter. */
int hoic_zfilter16(const float* obs, int n, int dim, double* st_n, double* mean, double* m2, double* dn, double* dmean, double* dm2,
                   double* scratch, float* shift_scale, float clip, int update, float* x_out, void* x16_out, int ld16, void* stream);
/* replaces PolicyGaussian.select_action + the sampler's mean_action switch (uhc/khrylib/rl/core/policy_gaussian.py:35-37,
 * agent_handmimic.py:463-466, 484) for the batch: explore[i] = u[i] < *noise_rate ? 1 : 0 (u uniform in [0,1): mean_action =
 * binomial(1, 1 - noise_rate)), action[i][:] = mean[i][:] + exp(log_std[:]) * noise[i][:] * explore[i].  mean: [n][act_dim]
 * bfloat16 (mean_is_bf16 != 0, the MLP's tensor-core output) or float32; noise: standard-normal draws of the caller's RNG. */
int hoic_sample_action(const void* mean, int mean_is_bf16, const float* log_std, const float* noise, const float* u,
                       const float* noise_rate, int n, int act_dim, float* action, float* explore, void* stream);
/* replaces LoggerRL.step / end_episode (uhc/khrylib/rl/core/logger_rl.py:28-44) for one batched step: sums[13] += steps,
 * finished episodes, c_reward, reward_info[9], reward of the episodes that finished; minmax[4] = c_reward min / max, episode
 * reward min / max; episode_reward[n] (may be NULL) = per-env running sums, cleared where done. */
int hoic_logger_step(const float* reward, const float* reward_info, const uint8_t* done, int n, double* sums, double* minmax,
                     double* episode_reward, void* stream);
const char* hoic_glue_last_error(void);

/* per-kernel timing of hoic_step for bench.py's roofline (off by default; not under stream capture): with it on, every
 * launch of the step pipeline is bracketed by CUDA events on the stream it runs on.  hoic_get_timing returns the summed
 * durations since hoic_set_timing(b, 1) in ms_host[8] = k_p1, k_p2, k_p3s (solver, small tier), k_p3 (full tier), k_tail,
 * k_reset, k_order, the whole call on the caller's stream (fork to join), and the number of hoic_step calls.  The batch
 * runs as interleaved parts on parallel streams, so the per-kernel sums count concurrent time more than once. */
int hoic_set_timing(hoic_batch* b, int on);
int hoic_get_timing(hoic_batch* b, double* ms_host /*[8]*/, int* steps_host);

/* integer work counters accumulated by the kernels since creation (synchronises the device) */
int hoic_get_counters(hoic_batch* b, hoic_counters* out_host);
/* FP32 FMA micro-benchmark on `device` (roofline denominator for the non-tensor pipe): 2*FMAs/s in TFLOP/s */
int hoic_measure_fp32_peak(int device, double* tflops_host);
/* number of kernel launches issued by this library since load (for bench.py's gpu_launches) */
unsigned long long hoic_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* HOIC_H */

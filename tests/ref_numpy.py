"""Literal numpy transcription of the reference's pure-Python observation / reward / termination code, used to pin the
oracle (and through it the kernels).  Sources: uhc/envs/ho_im4.py:280-356 (get_full_obs_v5), :664-688 (calc_ho_diff),
uhc/envs/ho_reward.py:943-1046 (ho_mimic_reward_9), uhc/utils/math_utils.py:103-131, 210-236,
uhc/utils/transformation.py:1344-1360, 1478-1515, transforms3d.quaternions.quat2mat (restated)."""
import math

import numpy as np


def quaternion_matrix(q):
    q = np.array(q, dtype=np.float64, copy=True)
    n = np.dot(q, q)
    if n < np.finfo(float).eps * 4.0:
        return np.identity(4)
    q *= math.sqrt(2.0 / n)
    q = np.outer(q, q)
    return np.array([
        [1.0 - q[2, 2] - q[3, 3], q[1, 2] - q[3, 0], q[1, 3] + q[2, 0], 0.0],
        [q[1, 2] + q[3, 0], 1.0 - q[1, 1] - q[3, 3], q[2, 3] - q[1, 0], 0.0],
        [q[1, 3] - q[2, 0], q[2, 3] + q[1, 0], 1.0 - q[1, 1] - q[2, 2], 0.0],
        [0.0, 0.0, 0.0, 1.0]])


def quat2mat(q):  # transforms3d.quaternions.quat2mat
    w, x, y, z = q
    Nq = w * w + x * x + y * y + z * z
    if Nq < np.finfo(np.float64).eps:
        return np.eye(3)
    s = 2.0 / Nq
    X, Y, Z = x * s, y * s, z * s
    wX, wY, wZ = w * X, w * Y, w * Z
    xX, xY, xZ = x * X, x * Y, x * Z
    yY, yZ, zZ = y * Y, y * Z, z * Z
    return np.array([[1.0 - (yY + zZ), xY - wZ, xZ + wY], [xY + wZ, 1.0 - (xX + zZ), yZ - wX], [xZ - wY, yZ + wX, 1.0 - (xX + yY)]])


def quaternion_multiply(quaternion1, quaternion0):
    w0, x0, y0, z0 = quaternion0
    w1, x1, y1, z1 = quaternion1
    return np.array([-x1 * x0 - y1 * y0 - z1 * z0 + w1 * w0, x1 * w0 + y1 * z0 - z1 * y0 + w1 * x0,
                     -x1 * z0 + y1 * w0 + z1 * x0 + w1 * y0, x1 * y0 - y1 * x0 + z1 * w0 + w1 * z0], dtype=np.float64)


def quaternion_inverse(quaternion):
    q = np.array(quaternion, dtype=np.float64, copy=True)
    q[1:] *= -1.0
    return q / np.dot(q, q)


def transform_vec(v, q):
    rot = quaternion_matrix(q)[:3, :3]
    return rot.T.dot(v[:, None]).ravel()


def transform_vec_batch(v_b, q):
    rot = quaternion_matrix(q)[:3, :3]
    return rot.T.dot(v_b[:, :, None]).squeeze()


def multi_quat_diff(nq1, nq0):
    nq_diff = np.zeros_like(nq0)
    for i in range(nq1.shape[0] // 4):
        ind = slice(4 * i, 4 * i + 4)
        nq_diff[ind] = quaternion_multiply(nq1[ind], quaternion_inverse(nq0[ind]))
    return nq_diff


def multi_quat_norm_v2(nq):
    _diff = []
    for i in range(nq.shape[0] // 4):
        q = nq[4 * i:4 * (i + 1)]
        d = np.array([abs(q[0]) - 1.0, q[1], q[2], q[3]])
        _diff.append(np.linalg.norm(d))
    return np.array(_diff)


class EnvView:
    """What the reference code reads from `env`: data.qpos/qvel, body_xpos/xquat (of the LAST forward), expert dict."""

    def __init__(self, qpos, qvel, body_xpos, body_xquat, expert, cur_t, hand_body_idx, palm_body, w_size=5, rfc_score=0.0):
        self.qpos, self.qvel, self.body_xpos, self.body_xquat = qpos, qvel, body_xpos, body_xquat
        self.expert, self.cur_t, self.start_ind = expert, cur_t, 0
        self.hand_body_idx, self.palm_body, self.w_size, self.rfc_score = hand_body_idx, palm_body, w_size, rfc_score
        self.expert_len = expert["hand_dof_seq"].shape[0]
        self.hand_qpos_dim, self.hand_qvel_dim, self.qpos_dim, self.qvel_dim = qpos.size - 7, qvel.size - 6, qpos.size, qvel.size

    def get_expert_attr(self, attr, ind):
        ind = min(ind, self.expert_len - 1)
        return self.expert[attr][ind].copy()

    def get_hand_root_pose(self):
        return np.concatenate((self.body_xpos[self.palm_body], self.body_xquat[self.palm_body]))

    def get_wbody_pos(self):
        return self.body_xpos[self.hand_body_idx].copy().ravel()

    def get_wbody_quat(self):
        return self.body_xquat[self.hand_body_idx].copy().ravel()


def get_full_obs_v5(env):
    w_size = env.w_size
    qpos, qvel = env.qpos.copy(), env.qvel.copy()
    cur_root_pose = env.get_hand_root_pose()
    obs = []
    hand_rot_mat = quat2mat(cur_root_pose[3:])
    obs.append(hand_rot_mat[:, 0:2].flatten())
    obs.append(qpos[6:env.hand_qpos_dim])
    for idx in range(0, w_size):
        target_hand_qpos = env.get_expert_attr("hand_dof_seq", env.cur_t + 1 + idx)
        obs.append(target_hand_qpos[6:] - qpos[6:env.hand_qpos_dim])
    obs.append(qvel[6:env.hand_qvel_dim])
    obs.append(transform_vec(qvel[:3], cur_root_pose[3:]))
    obs.append(transform_vec(qvel[3:6], cur_root_pose[3:]))
    for idx in range(0, w_size):
        ind = env.cur_t + 1 + idx
        target_root_pose = np.concatenate([env.get_expert_attr("body_pos_seq", ind)[0], env.get_expert_attr("body_quat_seq", ind)[0]])
        rel_root_pos = transform_vec(target_root_pose[:3] - cur_root_pose[:3], cur_root_pose[3:])
        rel_root_rot = quat2mat(quaternion_multiply(quaternion_inverse(cur_root_pose[3:]), target_root_pose[3:]))
        obs.append(rel_root_pos)
        obs.append(rel_root_rot[:, :2].flatten())
    curr_jpos = env.get_wbody_pos().reshape(-1, 3)
    r_jpos = transform_vec_batch(curr_jpos[1:] - qpos[None, :3], cur_root_pose[3:])
    obs.append(r_jpos.ravel())
    for idx in range(0, w_size):
        target_jpos = env.get_expert_attr("body_pos_seq", env.cur_t + 1 + idx).reshape(-1, 3)
        diff_jpos = transform_vec_batch(target_jpos[1:] - curr_jpos[1:], cur_root_pose[3:])
        obs.append(diff_jpos.ravel())
    obj_root_pose = qpos[env.hand_qpos_dim:]
    obs.append(transform_vec(obj_root_pose[:3] - cur_root_pose[:3], cur_root_pose[3:]))
    obs.append(quat2mat(quaternion_multiply(quaternion_inverse(cur_root_pose[3:]), obj_root_pose[3:]))[:, 0:2].flatten())
    obs.append(transform_vec(qvel[env.hand_qvel_dim:env.hand_qvel_dim + 3], cur_root_pose[3:]))
    obs.append(transform_vec(qvel[env.hand_qvel_dim + 3:env.hand_qvel_dim + 6], cur_root_pose[3:]))
    for idx in range(0, w_size):
        target_obj_pose = env.get_expert_attr("obj_pose_seq", env.cur_t + 1 + idx)
        rel_obj_pos = transform_vec(target_obj_pose[:3] - obj_root_pose[:3], cur_root_pose[3:])
        rel_obj_rot = quaternion_multiply(target_obj_pose[3:], quaternion_inverse(obj_root_pose[3:]))
        rel_obj_rot = quaternion_multiply(quaternion_inverse(cur_root_pose[3:]), rel_obj_rot)
        obs.append(rel_obj_pos)
        obs.append(quat2mat(rel_obj_rot)[:, 0:2].flatten())
    return np.concatenate(obs)


def calc_ho_diff(env):
    ind = env.cur_t
    expert_root_pose = np.concatenate([env.get_expert_attr("body_pos_seq", ind)[0], env.get_expert_attr("body_quat_seq", ind)[0]])
    expert_jpos = env.get_expert_attr("body_pos_seq", ind).reshape(-1, 3)
    cur_root_pose = env.get_hand_root_pose()
    cur_jpos = env.get_wbody_pos().reshape(-1, 3)
    pos_diff = np.linalg.norm(cur_root_pose[:3] - expert_root_pose[:3])
    jpos_diff = np.linalg.norm(cur_jpos - expert_jpos, axis=1).mean()
    quat_diff = quaternion_multiply(expert_root_pose[3:], quaternion_inverse(cur_root_pose[3:]))
    rot_diff = 2.0 * np.arcsin(np.clip(np.linalg.norm(quat_diff[1:]), 0, 1))
    epose_obj = env.get_expert_attr("obj_pose_seq", ind)
    obj_dist = np.linalg.norm(env.qpos[env.hand_qpos_dim:][:3] - epose_obj[:3])
    obj_quat_diff = quaternion_multiply(epose_obj[3:], quaternion_inverse(epose_obj[3:]))
    obj_rot_diff = 2.0 * np.arcsin(np.clip(np.linalg.norm(obj_quat_diff[1:]), 0, 1))
    return np.array([pos_diff, rot_diff, jpos_diff, obj_dist, obj_rot_diff])


def ho_mimic_reward_9(env, ws):
    g = ws.get
    w_p, w_wp, w_v, w_j, w_op, w_or, w_ov, w_orfc = (g("w_p", 0.4), g("w_wp", 0.4), g("w_v", 0.005), g("w_j", 100), g("w_op", 0.45),
                                                     g("w_or", 0.45), g("w_ov", 0.1), g("w_orfc", 0.2))
    k_p, k_wp, k_v, k_j, k_op, k_or, k_ov, k_orfc = (g("k_p", 0.4), g("k_wp", 0.4), g("k_v", 0.005), g("k_j", 100), g("k_op", 100.0),
                                                     g("k_or", 5.0), g("k_ov", 0.05), g("k_orfc", 1))
    ind = env.cur_t + env.start_ind
    cur_qpos, cur_qvel = env.qpos[:env.hand_qpos_dim], env.qvel[:env.hand_qvel_dim]
    cur_wbquat = env.get_wbody_quat().reshape(-1, 4)
    cur_wbpos = env.get_wbody_pos().reshape(-1, 3)
    e_qpos, e_qvel = env.get_expert_attr("hand_dof_seq", ind), env.get_expert_attr("hand_dof_vel_seq", ind)
    e_wbquat = env.get_expert_attr("body_quat_seq", ind).reshape(-1, 4)
    e_wbpos = env.get_expert_attr("body_pos_seq", ind).reshape(-1, 3)
    pose_reward = math.exp(-k_p * np.abs(cur_qpos[6:] - e_qpos[6:]).mean())
    wpose_diff = multi_quat_norm_v2(multi_quat_diff(cur_wbquat.flatten(), e_wbquat.flatten()))
    wpose_reward = math.exp(-k_wp * np.abs(wpose_diff).mean())
    vel_reward = math.exp(-k_v * np.abs(cur_qvel - e_qvel).mean())
    jpos_reward = math.exp(-k_j * np.linalg.norm(cur_wbpos - e_wbpos, axis=1).mean())
    epose_obj = env.get_expert_attr("obj_pose_seq", ind)
    obj_pose = env.qpos[env.hand_qpos_dim:]
    obj_pos_reward = math.exp(-k_op * np.linalg.norm(obj_pose[:3] - epose_obj[:3]))
    obj_rot_diff = multi_quat_norm_v2(multi_quat_diff(obj_pose[3:], epose_obj[3:]))
    obj_rot_reward = math.exp(-k_or * abs(obj_rot_diff[0]))
    evel_obj = np.concatenate([env.get_expert_attr("obj_vel_seq", ind), env.get_expert_attr("obj_angle_vel_seq", ind)])
    obj_vel_reward = math.exp(-k_ov * abs(env.qvel[env.hand_qvel_dim:] - evel_obj).mean())
    obj_rfc_reward = math.exp(-k_orfc * env.rfc_score)
    hand_reward = (w_p * pose_reward + w_wp * wpose_reward + w_j * jpos_reward + w_v * vel_reward) / (w_p + w_wp + w_j + w_v)
    obj_reward = (w_op * obj_pos_reward + w_or * obj_rot_reward + w_ov * obj_vel_reward + w_orfc * obj_rfc_reward) / (w_op + w_or + w_ov + w_orfc)
    return hand_reward * obj_reward, np.array([pose_reward, wpose_reward, jpos_reward, vel_reward, obj_pos_reward, obj_rot_reward,
                                               obj_vel_reward, obj_rfc_reward, 1.0])


def solve_rfc_numpy(F, T, frames, ts_counts, geom_xpos_g1, gvel, gangvel, obj_p, obj_v, obj_w, frame_skip=15, surface_contact=True):
    """ho_im4.py:983-1083 with scipy's bounded least squares instead of daqp.  frames: [k][12] averaged contact frames."""
    from scipy.optimize import minimize
    dx, mu, w_t = 0.0025, 0.75, 1e4
    e_cps, rel_vn_arr, rel_vt_arr = [], [], []
    ec = 5 if surface_contact else 1
    for i in range(len(frames)):
        pos, frame = frames[i][:3], frames[i][3:12].reshape(3, 3)
        delta_pos = np.array([np.zeros(3), frame[1], -frame[1], frame[2], -frame[2]]) * dx
        for j in range(ec):
            e_cps.append(np.concatenate([pos + delta_pos[j], frames[i][3:12]]))
            cr_h, cr_o = pos + delta_pos[j] - geom_xpos_g1[i], pos + delta_pos[j] - obj_p
            rel_v = (obj_v + np.cross(obj_w, cr_o)) - (gvel[i] + np.cross(gangvel[i], cr_h))
            rel_vn = np.dot(frame[0], rel_v) * frame[0]
            rel_vn_arr.append(rel_vn)
            rel_vt_arr.append(rel_v - rel_vn)
    n_c = len(e_cps)
    A_arr, rA_arr, p_n_arr, p_t_arr = [], [], [], []
    for i in range(n_c):
        pos, frame = e_cps[i][:3], e_cps[i][3:12].reshape(3, 3)
        ts = ts_counts[i // 5] / frame_skip
        xs = [(frame[0] + mu * frame[1]), (frame[0] - mu * frame[1]), (frame[0] + mu * frame[2]), (frame[0] - mu * frame[2])]
        A = np.stack([x / np.sqrt(1 + mu ** 2) for x in xs]).T
        A_arr.append(ts * A)
        r = pos - obj_p
        r_x = np.array([[0, -r[2], r[1]], [r[2], 0, -r[0]], [-r[1], r[0], 0]])
        rA_arr.append(ts * np.matmul(r_x, A))
        p_n_arr.append(np.linalg.norm(rel_vn_arr[i]) * np.ones(4) if np.dot(rel_vn_arr[i], frame[0]) <= 0 else np.zeros(4))
        dirs = np.array([np.dot(-rel_vt_arr[i], frame[1]), np.dot(-rel_vt_arr[i], -frame[1]), np.dot(-rel_vt_arr[i], frame[2]), np.dot(-rel_vt_arr[i], -frame[2])])
        p_t = np.linalg.norm(rel_vt_arr[i]) * np.ones(4)
        p_t[np.argmax(dirs)] = 0
        p_t_arr.append(p_t)
    J_f, J_t = np.concatenate(A_arr, axis=1), np.concatenate(rA_arr, axis=1)
    p_n, p_t = np.concatenate(p_n_arr), np.concatenate(p_t_arr)
    Q = 2 * (J_f.T @ J_f + w_t * J_t.T @ J_t) + 1e-7 * np.eye(4 * n_c)
    p = -2 * J_f.T @ F - 2 * w_t * J_t.T @ T + p_n + p_t
    res = minimize(lambda x: 0.5 * x @ Q @ x + p @ x, np.zeros(4 * n_c), jac=lambda x: Q @ x + p, bounds=[(0, None)] * (4 * n_c),
                   method="L-BFGS-B", options=dict(maxiter=20000, maxfun=200000, ftol=1e-16, gtol=1e-12))
    x = res.x
    rest_f, rest_t = F - J_f @ x, T - J_t @ x
    return np.linalg.norm(rest_f) + math.sqrt(w_t) * np.linalg.norm(rest_t), 0.5 * x @ Q @ x + p @ x, Q, p


def compute_torque_numpy(qpos, qvel, ref_qpos, ctrl, M, C, jkp, jkd, ctrl_scale, torque_lim, dt):
    """Literal transcription of HandObjMimic4.compute_torque + compute_desired_accel + the clip of do_simulation
    (uhc/envs/ho_im4.py:393-486, :521) for action_type 'position', pd_type 'rel', grot_type 'euler', no meta-PD:
    qpos / qvel / ref_qpos are the 26 hand entries, M and C the hand block of mj_fullM(qM) and qfrc_bias."""
    from scipy.linalg import cho_factor, cho_solve
    qpos, qvel = qpos.copy(), qvel.copy()
    target_pos = np.zeros_like(qpos)
    ctrl_pos, ctrl_rot, ctrl_loco = ctrl[:3], ctrl[3:6], ctrl[6:26]
    target_pos[0:3] = ref_qpos[0:3] + 0.1 * ctrl_pos
    target_pos[6:] = ref_qpos[6:] + ctrl_scale[6:] * ctrl_loco
    target_pos[3:6] = ref_qpos[3:6] + 0.3 * ctrl_rot
    k_p, k_d = np.zeros(qvel.shape[0]), np.zeros(qvel.shape[0])
    k_p[:], k_d[:] = jkp, jkd
    qpos_err = qpos[:] + qvel[:] * dt - target_pos
    qvel_err = qvel
    rot_qpos_err = qpos_err[3:].copy()
    while np.any(rot_qpos_err > np.pi):
        rot_qpos_err[rot_qpos_err > np.pi] -= 2 * np.pi
    while np.any(rot_qpos_err < -np.pi):
        rot_qpos_err[rot_qpos_err < -np.pi] += 2 * np.pi
    qpos_err[3:] = rot_qpos_err[:]
    K_p, K_d = np.diag(k_p), np.diag(k_d)
    q_accel = cho_solve(cho_factor(M + K_d * dt, overwrite_a=True, check_finite=False),
                        -C[:, None] - K_p.dot(qpos_err[:, None]) - K_d.dot(qvel_err[:, None]), overwrite_b=True, check_finite=False).squeeze()
    qvel_err += q_accel * dt
    torque = -jkp * qpos_err - jkd * qvel_err
    return np.clip(torque, -torque_lim, torque_lim)

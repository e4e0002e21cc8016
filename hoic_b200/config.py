"""YAML -> Config attributes, mirroring `uhc/utils/config_utils/handmimic_config.py:12-195` and
`base_config.py:10-61` for the keys the rollout path reads (same attribute names, same defaults, same
`update_adaptive_params` schedule).  Directory creation, wandb and checkpoint paths are out of scope.
"""
from __future__ import annotations

import glob
import json
import os

import numpy as np
import yaml

REWARD_KEYS = ("w_p", "w_wp", "w_v", "w_j", "w_op", "w_or", "w_ov", "w_orfc",
               "k_p", "k_wp", "k_v", "k_j", "k_op", "k_or", "k_ov", "k_orfc")
# defaults of ho_mimic_reward_9 (ho_reward.py:947-967) for keys missing from `reward_weights`
REWARD_DEFAULTS = dict(w_p=0.4, w_wp=0.4, w_v=0.005, w_j=100, w_op=0.45, w_or=0.45, w_ov=0.1, w_orfc=0.2,
                       k_p=0.4, k_wp=0.4, k_v=0.005, k_j=100, k_op=100.0, k_or=5.0, k_ov=0.05, k_orfc=1)


class Config:
    def __init__(self, cfg_id=None, base_dir=".", cfg_dict=None, **kwargs):
        if cfg_dict is None:
            # base_config.py:17-23: glob config/**/<id>.yml, must be unique
            files = glob.glob(os.path.join(base_dir, "config", "**", f"{cfg_id}.yml"), recursive=True)
            if len(files) != 1:
                raise FileNotFoundError(f"{len(files)} config files match {cfg_id} under {base_dir}/config")
            with open(files[0]) as f:
                cfg_dict = yaml.safe_load(f)
        self.id = cfg_id
        self.base_dir = base_dir
        self.cfg_dict = c = cfg_dict
        g = c.get
        # training
        self.gamma, self.tau = g("gamma", 0.95), g("tau", 0.95)
        self.policy_htype, self.policy_hsize = g("policy_htype", "relu"), g("policy_hsize", [300, 200])
        self.policy_lr, self.policy_weightdecay = g("policy_lr", 5e-5), g("policy_weightdecay", 0.0)
        self.value_htype, self.value_hsize = g("value_htype", "relu"), g("value_hsize", [300, 200])
        self.value_lr, self.value_weightdecay = g("value_lr", 3e-4), g("value_weightdecay", 0.0)
        self.adv_clip = g("adv_clip", np.inf)
        self.clip_epsilon = g("clip_epsilon", 0.2)
        self.log_std, self.fix_std = g("log_std", -2.3), g("fix_std", False)
        self.num_optim_epoch = g("num_optim_epoch", 10)
        self.min_batch_size = g("min_batch_size", 50000)
        self.mini_batch_size = g("mini_batch_size", self.min_batch_size)
        self.num_epoch = g("num_epoch", 20000)
        self.seed = g("seed", 1)
        self.save_n_epochs = g("save_n_epochs", 100)
        self.reward_type = g("reward_type", 0)
        self.reward_weights = dict(g("reward_weights", {}) or {})
        self.end_reward = g("end_reward", False)
        self.data_specs = g("data_specs", {})
        self.mujoco_model = g("mujoco_model", None)
        # adaptive schedule (handmimic_config.py:48-74)
        self.adp_iter_cp = np.array(g("adp_iter_cp", [0]))
        pad = lambda a: np.pad(np.array(a, dtype=float), (0, self.adp_iter_cp.size - len(a)), "edge")
        self.adp_noise_rate_cp = pad(g("adp_noise_rate_cp", [1.0]))
        self.adp_log_std_cp = pad(g("adp_log_std_cp", [self.log_std]))
        self.adp_policy_lr_cp = pad(g("adp_policy_lr_cp", [self.policy_lr]))
        self.adp_noise_rate = self.adp_log_std = self.adp_policy_lr = None
        # env
        self.future_w_size = g("future_w_size", 5)
        self.pos_diff_thresh = g("pos_diff_thresh", 0.1)
        self.rot_diff_thresh = g("rot_diff_thresh", 1.0)
        self.jpos_diff_thresh = g("jpos_diff_thresh", 0.1)
        self.obj_pos_diff_thresh = g("obj_pos_diff_thresh", 0.1)
        self.obj_rot_diff_thresh = g("obj_rot_diff_thresh", 1.0)
        self.random_start = g("random_start", False)
        self.sim_step = g("sim_step", 15)
        self.obs_type = g("obs_type", 0)
        self.noise_future_pose = g("noise_future_pose", False)
        self.action_type = g("action_type", "position")
        self.render = False
        # residual force
        self.residual_force = g("residual_force", False)
        self.surface_contact = g("surface_contact", True)
        self.explain_force = g("explain_force", True)
        self.residual_force_scale = g("residual_force_scale", 2.0)
        self.residual_torque_scale = g("residual_torque_scale", 0.1)
        self.meta_pd, self.meta_pd_joint = g("meta_pd", False), g("meta_pd_joint", False)
        self.pd_type, self.grot_type = g("pd_type", "base"), g("grot_type", "euler")
        if "joint_params" in c:
            cols = list(zip(*c["joint_params"]))
            self.joint_names = list(cols[0])
            self.jkp = np.array(cols[1], dtype=float) * g("jkp_multiplier", 1.0)
            self.jkd = np.array(cols[2], dtype=float) * g("jkd_multiplier", g("jkp_multiplier", 1.0))
            self.torque_lim = np.array(cols[3], dtype=float) * g("torque_limit_multiplier", 1.0)
        for k, v in kwargs.items():
            setattr(self, k, v)
        unsupported = []
        if self.obs_type == 3 or self.noise_future_pose:
            unsupported.append("obs_type 3 / noise_future_pose")
        if self.meta_pd or self.meta_pd_joint or self.action_type != "position":
            unsupported.append("meta_pd / torque actions")
        if self.pd_type != "rel" or self.grot_type != "euler":
            unsupported.append("pd_type != rel or grot_type != euler")
        if self.reward_type != 9:
            unsupported.append("reward_type != 9")
        self.unsupported = unsupported  # release YAMLs: empty (checked by HandMimicVecEnv)

    def get(self, key, default=None):
        return getattr(self, key, default)

    def update_adaptive_params(self, i_iter):
        """handmimic_config.py:157-195."""
        cp = self.adp_iter_cp
        ind = np.where(i_iter >= cp)[0][-1]
        nind = ind + int(ind < len(cp) - 1)
        t = (i_iter - cp[ind]) / (cp[nind] - cp[ind]) if nind > ind else 0.0
        self.adp_noise_rate = self.adp_noise_rate_cp[ind] * (1 - t) + self.adp_noise_rate_cp[nind] * t
        self.adp_log_std = self.adp_log_std_cp[ind] * (1 - t) + self.adp_log_std_cp[nind] * t
        self.adp_policy_lr = self.adp_policy_lr_cp[ind] * (1 - t) + self.adp_policy_lr_cp[nind] * t
        ws = self.reward_weights
        it = min(i_iter, ws.get("end_grad_iter", 3000))
        if ws.get("w_rfc_init") is not None and ws.get("k_rfc_init") is not None:
            ws["w_orfc"] = ws["w_rfc_init"] * np.exp(ws.get("rfc_grad", 0) * it)
            ws["k_orfc"] = ws["k_rfc_init"] * np.exp(ws.get("rfc_grad", 0) * it)
        if all(ws.get(k) is not None for k in ("k_p_init", "k_j_init", "k_wp_init")):
            ws["k_p"] = ws["k_p_init"] * np.exp(ws.get("hand_grad", 0) * it)
            ws["k_j"] = ws["k_j_init"] * np.exp(ws.get("hand_grad", 0) * it)
            ws["k_wp"] = ws["k_wp_init"] * np.exp(ws.get("hand_grad", 0) * it)

    def reward_vector(self) -> np.ndarray:
        """24 floats for hoic_set_reward_weights: REWARD_KEYS order, rest reserved (0)."""
        w = np.zeros(24)
        for i, k in enumerate(REWARD_KEYS):
            w[i] = self.reward_weights.get(k, REWARD_DEFAULTS[k])
        return w

    def to_json(self) -> str:
        return json.dumps(self.cfg_dict)

    @classmethod
    def from_json(cls, s, cfg_id=None):
        return cls(cfg_id=cfg_id, cfg_dict=json.loads(s))

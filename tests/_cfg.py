"""Default configuration values the hot path reads (the knobs of the reference's
configs/default.yaml that reach model/, with the same nesting), written out as a literal so the
tests do not need the reference tree."""
import copy

_DEFAULT = {
    "model": {"hidden_dim": 256, "occ_activation": "softplus", "pos_enc_levels": 10, "dir_enc_levels": 4,
              "network_type": "official"},
    "rendering": {"type": "nope_nerf", "n_max_network_queries": 64000, "white_background": False, "radius": 4.0,
                  "num_points": 128, "depth_range": [0.01, 10], "dist_alpha": False, "use_ray_dir": True,
                  "normalise_ray": True, "normal_loss": False, "sample_option": "uniform", "outside_steps": 0},
    "depth": {"type": "None"},
    "pose": {"learn_pose": True, "learn_R": True, "learn_t": True, "init_pose": False, "learn_focal": False},
    "distortion": {"learn_distortion": True, "fix_scaleN": True, "learn_scale": True, "learn_shift": True},
    "training": {"type": "nope_nerf", "n_training_points": 1024, "learning_rate": 0.001, "pose_lr": 0.0005,
                 "distortion_lr": 0.0005, "rgb_weight": [1.0, 1.0], "depth_weight": [0.04, 0.0],
                 "weight_dist_2nd_loss": [0.0, 0.0], "weight_dist_1st_loss": [0.0, 0.0], "pc_weight": [1.0, 0.0],
                 "rgb_s_weight": [1.0, 0.0], "depth_consistency_weight": [0.0, 0.0], "rgb_loss_type": "l1",
                 "depth_loss_type": "l1", "with_auto_mask": False, "vis_geo": True, "with_ssim": False,
                 "detach_gt_depth": False, "match_method": "dense", "pc_ratio": 4, "shift_first": False,
                 "detach_ref_img": True, "scale_pcs": True, "detach_rgbs_scale": False, "vis_reprojection_every": 5000,
                 "nearest_limit": 0.01, "annealing_epochs": 2000, "scheduling_start": 10000},
    "eval_pose": {"n_points": 1024, "type": "nope_nerf"},
}


def default_cfg():
    return copy.deepcopy(_DEFAULT)

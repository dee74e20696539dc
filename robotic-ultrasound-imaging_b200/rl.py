"""Training / evaluation driver with the interface of the reference's ``src/rl.py`` (reads the same ``rl_config.yaml``).

  python -m rui_b200.rl --config rl_config.yaml [--num-envs 4096] [--n-steps 32] [--total-timesteps 4e7]
  torchrun --nproc-per-node 8 -m rui_b200.rl --config rl_config.yaml --num-envs 65536   # envs sharded over the ranks

``sb_config.num_cpu`` (64 worker processes in the reference, rl.py:130) becomes the number of batched envs unless
``--num-envs`` is given.  ``n_steps`` must shrink with thousands of envs (2048 x 65536 exceeds total_timesteps).
"""
from __future__ import annotations

import argparse
import os

import torch
import torch.distributed as dist
import yaml

from .dist import shard_range
from .env import BatchedUltrasound
from .model import cylinder_torso_params
from .ppo import PPO

_ENV_KEYS = ("controller_configs", "control_freq", "horizon", "early_termination", "torso_solref_randomization",
             "initial_probe_pos_randomization", "deterministic_trajectory")


def env_from_config(cfg: dict, num_envs: int, seed: int, device, env_id_offset: int = 0) -> BatchedUltrasound:
    rs = dict(cfg["robosuite"])
    assert rs.pop("env_id", "Ultrasound") == "Ultrasound" and rs.get("robots", "Panda") in ("Panda", "UR5e")
    if rs.get("use_camera_obs") or rs.get("has_renderer") or rs.get("has_offscreen_renderer"):
        raise NotImplementedError("rendering is out of scope of the hot path")
    extra = {} if rs.get("use_box_torso", True) else {"scene_params": cylinder_torso_params()}
    if rs.get("robots", "Panda") == "UR5e":
        from .model import ur5e_params
        extra = {"scene_params": ur5e_params(extra.get("scene_params"))}
    return BatchedUltrasound(num_envs, device=device, seed=seed, env_id_offset=env_id_offset, **{k: rs[k] for k in _ENV_KEYS if k in rs}, **extra)


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="rl_config.yaml")
    ap.add_argument("--num-envs", type=int, default=0, help="total envs over all ranks (default: sb_config.num_cpu)")
    ap.add_argument("--n-steps", type=int, default=32)
    ap.add_argument("--batch-size", type=int, default=0)
    ap.add_argument("--total-timesteps", type=float, default=0)
    args = ap.parse_args(argv)
    with open(args.config) as f:
        cfg = yaml.safe_load(f)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank, local = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    total_envs = args.num_envs or int(cfg["sb_config"]["num_cpu"])
    off, n = shard_range(total_envs, rank, world)
    fh = cfg["file_handling"]
    env = env_from_config(cfg, n, int(cfg["seed"]), torch.device(f"cuda:{local}"), off)
    model = PPO(env, n_steps=args.n_steps, batch_size=args.batch_size or None, net_arch=cfg["sb_policy"].get("net_arch"), seed=int(cfg["seed"]),
                verbose=1)
    save_path = os.path.join(fh["save_model_folder"], fh["save_model_filename"])
    if cfg.get("training", True):
        if fh.get("continue_training_model_filename"):
            model.load(os.path.join(fh["continue_training_model_folder"], fh["continue_training_model_filename"]))
        interval = float(cfg["sb_config"].get("check_pt_interval", 0))
        state = {"next": interval}

        def checkpoint(m):  # CheckpointCallback(save_freq=check_pt_interval, save_path='./checkpoints/')
            if interval and m.num_timesteps >= state["next"]:
                m.save(os.path.join("checkpoints", f"{fh['save_model_filename']}_{m.num_timesteps}_steps"))
                state["next"] += interval

        model.learn(int(args.total_timesteps or float(cfg["sb_config"]["total_timesteps"])), callback=checkpoint)
        model.save(save_path)
    else:
        model.load(os.path.join(fh["load_model_folder"], fh["load_model_filename"]))
        model.norm.training = False
        obs, eprew = env.reset(), torch.zeros(n, device=env.device)
        for _ in range(int(cfg["robosuite"]["horizon"])):
            obs, r, d, _ = env.step(model.predict(obs), auto_reset=True)
            eprew += r
        if rank == 0:
            print(f"mean return over {n} envs: {float(eprew.mean()):.2f}")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""GPU tests of the caller side: the reference's shipped `tracking` policy drives the new env (closed-loop behavioural
probe) and the PPO driver trains on device-resident rollouts."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import CC_TRACK, ROOT

pytestmark = pytest.mark.gpu


def test_reference_policy_closed_loop_statistics():
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    from closed_loop_probe import run
    out = run(envs=1024, steps=1200)
    # same order as the reference's 40 M-sample statistics (not equality: probe geometry / arm constants are assumptions)
    assert 0.6 * out["art_step_reward"] <= out["step_reward"] <= 12.0
    assert out["ep_len_mean"] >= 0.5 * out["art_ep_len_mean"]
    assert abs(out["reset_in_contact_fraction"] - out["art_reset_in_contact_fraction"]) < 0.15
    np.testing.assert_allclose(out["reset_pos_err_mean"], out["art_reset_pos_err_mean"], atol=1.5e-3)
    np.testing.assert_allclose(out["reset_pos_err_std"], out["art_reset_pos_err_std"], rtol=0.35)
    m, am = np.array(out["obs_mean"]), np.array(out["art_obs_mean"])
    assert m[2] > 1.0 and 0.2 < m[2] / am[2] < 5.0           # the policy keeps the probe pressed: mean Fz same order
    assert abs(m[12] - am[12]) < 3e-3 and abs(m[13] - am[13]) < 3e-3 and abs(m[14] - am[14]) < 8e-3  # tracking error (m)
    assert m[15] < -0.5                                      # quaternion dot ~ -1 (A-QUAT-1)
    assert out["art_reset_pos_err_mean"][2] > 0.004          # the systematic +z reset offset exists in the artifacts


def test_ppo_trains_on_device():
    from rui_b200.env import BatchedUltrasound
    from rui_b200.ppo import PPO
    env = BatchedUltrasound(1024, device=0, controller_configs=CC_TRACK, control_freq=500, horizon=1000, early_termination=True,
                            torso_solref_randomization=True, initial_probe_pos_randomization=True, seed=3)
    model = PPO(env, n_steps=32, seed=3)
    model.learn(1024 * 32 * 2)
    r0 = model.last_stats["step_reward_mean"]
    model.learn(1024 * 32 * 14)
    r1 = model.last_stats["step_reward_mean"]
    assert np.isfinite(r0) and np.isfinite(r1) and r1 > r0 + 0.1, (r0, r1)
    assert model.num_timesteps == 1024 * 32 * 14
    obs = env.reset()
    a = model.predict(obs)
    assert a.shape == (1024, 6) and bool((a >= 0).all()) and bool((a <= 1).all())
    env.close()


def test_graphed_update_matches_eager_update():
    """The CUDA-graph replay of a PPO minibatch step (ppo.PPO._train_graphed) does the same arithmetic as the eager loop."""
    import torch

    from rui_b200.env import BatchedUltrasound
    from rui_b200.ppo import PPO
    env = BatchedUltrasound(256, device=0, controller_configs=CC_TRACK, control_freq=500, horizon=1000, early_termination=True,
                            torso_solref_randomization=True, initial_probe_pos_randomization=True, seed=5)
    g = PPO(env, n_steps=16, batch_size=512, seed=7, cuda_graph=True)
    e = PPO(env, n_steps=16, batch_size=512, seed=7, cuda_graph=False)
    g._setup()
    batch = g.collect_rollouts()
    for model in (g, e):
        torch.manual_seed(123)  # the same minibatch permutations
        model.train(batch)
        torch.manual_seed(123)
        model.train(batch)      # second call: pure replay on the graphed side
    for (n1, p1), (n2, p2) in zip(g.policy.named_parameters(), e.policy.named_parameters()):
        assert n1 == n2 and torch.allclose(p1, p2, rtol=1e-4, atol=1e-6), (n1, float((p1 - p2).abs().max()))
    assert g._n_updates == e._n_updates == 20
    assert abs(g.last_stats["value_loss"] - e.last_stats["value_loss"]) <= 1e-4 * max(1.0, abs(e.last_stats["value_loss"]))
    env.close()

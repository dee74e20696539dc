"""CPU tests of the PPO driver (caller side of the hot path): SB3-compatible policy layout, GAE, normaliser,
checkpoint round trip, learning on a toy batched env, and the 2-rank gloo gradient / moment all-reduce."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT
from rui_b200.ppo import PPO, MlpPolicy, RunningMeanStd, compute_gae, load_sb3_zip


class ToyEnv:
    """Batched env with the BatchedUltrasound interface: reward = 1 - |a - target(obs)|, 8-step episodes."""

    def __init__(self, n, obs_dim=19, act_dim=6, seed=0):
        self.num_envs, self.action_dim, self.device = n, act_dim, torch.device("cpu")
        self.g = torch.Generator().manual_seed(seed)
        self.obs = torch.zeros(n, obs_dim)
        self.t = torch.zeros(n)
        self.action_spec = (np.zeros(act_dim), np.ones(act_dim))

    def _draw(self):
        self.obs = torch.rand(self.num_envs, self.obs.shape[1], generator=self.g) * 4 - 2

    def reset(self, mask=None):
        self._draw(); self.t.zero_()
        return self.obs

    def step(self, a, auto_reset=True):
        target = torch.sigmoid(self.obs[:, : self.action_dim])
        r = 1 - (a - target).abs().mean(1)
        self.t += 1
        d = (self.t >= 8).to(torch.uint8)
        self.t[d.bool()] = 0
        self._draw()
        return self.obs, r, d, self.obs


def test_policy_layout_matches_sb3_and_art(art):
    pol = MlpPolicy(19, 6)
    keys = set(pol.state_dict().keys())
    assert keys == {"log_std", "mlp_extractor.policy_net.0.weight", "mlp_extractor.policy_net.0.bias", "mlp_extractor.policy_net.2.weight",
                    "mlp_extractor.policy_net.2.bias", "mlp_extractor.value_net.0.weight", "mlp_extractor.value_net.0.bias",
                    "mlp_extractor.value_net.2.weight", "mlp_extractor.value_net.2.bias", "action_net.weight", "action_net.bias",
                    "value_net.weight", "value_net.bias"}  # SURVEY App. D [ART]
    assert sum(p.numel() for p in pol.parameters()) == 76941  # [ART] parameter count of the shipped policies
    assert pol.state_dict()["mlp_extractor.policy_net.0.weight"].shape == (256, 19)
    assert (art["tracking"]["n_steps"], art["tracking"]["n_epochs"], art["tracking"]["gamma"], art["tracking"]["gae_lambda"]) == (2048, 10, 0.99, 0.95)


def test_gae_matches_naive_loop():
    T, N = 7, 5
    g = torch.Generator().manual_seed(0)
    r, v, lv = torch.rand(T, N, generator=g), torch.rand(T, N, generator=g), torch.rand(N, generator=g)
    d = (torch.rand(T, N, generator=g) < 0.3).float()
    adv, ret = compute_gae(r, v, d, lv, 0.99, 0.95)
    for n in range(N):
        last = 0.0
        for t in reversed(range(T)):
            nv = lv[n] if t == T - 1 else v[t + 1, n]
            nt = 1 - d[t, n]
            delta = r[t, n] + 0.99 * nv * nt - v[t, n]
            last = delta + 0.99 * 0.95 * nt * last
            assert abs(float(adv[t, n]) - float(last)) < 1e-5
    assert torch.allclose(ret, adv + v)


def test_running_mean_std_matches_numpy():
    rms = RunningMeanStd((3,))
    g = torch.Generator().manual_seed(1)
    chunks = [torch.randn(50, 3, generator=g) * 3 + 1 for _ in range(4)]
    for c in chunks:
        rms.update(c)
    allx = torch.cat(chunks).double().numpy()
    np.testing.assert_allclose(rms.mean.numpy(), allx.mean(0), atol=1e-3)
    np.testing.assert_allclose(rms.var.numpy(), allx.var(0), rtol=1e-3)


def test_manual_backward_matches_autograd():
    """The graphed update differentiates the minibatch loss by hand (PPO._manual_grads: explicit GEMMs into the flat gradient buffer).
    Same loss terms and the same gradient as autograd on ``_minibatch_loss``, entropy term included, to fp32 round-off."""
    torch.manual_seed(0)
    ppo = PPO(ToyEnv(4), n_steps=4, cuda_graph=False)
    ppo.ent_coef = 0.01
    B = 1024  # (split over 16 slabs of 64 samples in the weight-gradient products)
    obs, act = torch.randn(B, 19), torch.rand(B, 6)
    with torch.no_grad():
        for p in ppo.policy.parameters():
            p.add_(0.05 * torch.randn_like(p))
        mean, _ = ppo.policy(obs)
        old_lp = ppo.policy.log_prob(mean, ppo.policy.log_std, act) + 0.3 * torch.randn(B)  # ratios on both sides of the clip range
    adv, ret = torch.randn(B), torch.randn(B)
    loss, pl, vl, kl = ppo._minibatch_loss(obs, act, old_lp, adv, ret)
    ppo.policy.zero_grad()
    loss.backward()
    ref = {n: p.grad.clone() for n, p in ppo.policy.named_parameters()}
    ratio = torch.exp(ppo.policy.evaluate(obs, act)[1] - old_lp)
    assert float((ratio < 0.8).float().mean()) > 0.1 and float((ratio > 1.2).float().mean()) > 0.1
    for p in ppo.policy.parameters():
        p.grad = torch.full_like(p, 7.0)  # every entry must be overwritten
    pl2, vl2, kl2 = ppo._manual_grads(obs, act, old_lp, adv, ret)
    assert abs(float(pl) - float(pl2)) < 1e-6 and abs(float(vl) - float(vl2)) < 1e-6 and abs(float(kl) - float(kl2)) < 1e-6
    for n, p in ppo.policy.named_parameters():
        assert float((p.grad - ref[n]).abs().max()) <= 1e-6 * max(1.0, float(ref[n].abs().max())), n


def test_ppo_learns_and_checkpoints(tmp_path):
    env = ToyEnv(64)
    model = PPO(env, n_steps=16, batch_size=256, seed=0)
    model._setup()
    model.train(model.collect_rollouts())
    r0 = model.last_stats["step_reward_mean"]
    model.learn(64 * 16 * 13)
    r1 = model.last_stats["step_reward_mean"]
    assert r1 > r0 + 0.02, (r0, r1)
    assert abs(model.last_stats["ep_len_mean"] - 8) < 1e-9
    path = str(tmp_path / "model")
    model.save(path)
    sd, data, opt = load_sb3_zip(path + ".zip")
    assert data["num_timesteps"] == model.num_timesteps and opt is not None and set(sd) == set(model.policy.state_dict())
    import zipfile
    assert set(zipfile.ZipFile(path + ".zip").namelist()) == {"data", "policy.pth", "policy.optimizer.pth", "pytorch_variables.pth", "_stable_baselines3_version"}
    m2 = PPO(ToyEnv(64), n_steps=16, batch_size=256, seed=5).load(path)
    for a, b in zip(model.policy.parameters(), m2.policy.parameters()):
        assert torch.equal(a, b)
    np.testing.assert_allclose(m2.norm.obs_rms.mean.numpy(), model.norm.obs_rms.mean.numpy())
    obs = env.reset()
    assert torch.equal(model.predict(obs), m2.predict(obs))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from rui_b200.ppo import PPO
    model = PPO(ToyEnv(32, seed=10 + rank), n_steps=8, batch_size=128, seed=0)
    p0 = torch.cat([p.detach().reshape(-1) for p in model.policy.parameters()]).clone()
    model.learn(2 * 32 * 8 * 2)
    p1 = torch.cat([p.detach().reshape(-1) for p in model.policy.parameters()])
    q.put((rank, p0.numpy(), p1.numpy(), model.norm.obs_rms.mean.numpy(), float(model.norm.obs_rms.count), model.num_timesteps))
    dist.barrier(); dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_training_keeps_replicas_identical():
    ctx = mp.get_context("spawn")
    q, port = ctx.Queue(), _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60); assert p.exitcode == 0
    a, b = res
    np.testing.assert_array_equal(a[1], b[1])  # same initial weights (broadcast)
    assert np.abs(a[2] - a[1]).max() > 1e-4     # training moved them
    np.testing.assert_allclose(a[2], b[2], atol=1e-6)  # gradient all-reduce keeps the replicas in lock-step
    np.testing.assert_allclose(a[3], b[3], atol=1e-12)  # normaliser moments merged over ranks
    assert a[4] == b[4] and a[5] == b[5] == 2 * 32 * 8 * 2

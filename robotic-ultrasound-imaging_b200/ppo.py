"""PPO driver for the batched Ultrasound env — the caller side of the hot path (rl.py, BASELINE config 5).

Mirrors what ``rl.py:130-167`` does with stable-baselines3 (not installed here), with the
hyper-parameters decoded from the shipped models (SURVEY §6 / App. C.6 [ART]):
``PPO("MlpPolicy", net_arch=[dict(pi=[256,128], vf=[256,128])])``, tanh MLPs, state-independent ``log_std``,
lr 3e-4 (Adam eps 1e-5), gamma 0.99, GAE lambda 0.95, clip 0.2, ent 0, vf 0.5, max-grad-norm 0.5, 10 epochs,
``VecNormalize(clip_obs=10, clip_reward=10, gamma=0.99, epsilon=1e-8)``.

Everything stays on the device: the env is stepped through the device-pointer C ABI, the rollout buffer, the
normaliser and the policy are CUDA tensors.  Multi-GPU: one process per GPU, each rank owns a contiguous slice of
the global env ids; NCCL is used ONLY for (a) the flat gradient all-reduce per optimiser step, (b) the merge of the
normaliser moments and episode statistics once per rollout, (c) the initial parameter broadcast.

Documented deviations from SB3 1.1.0a5: with thousands of envs ``n_steps`` must shrink (2048 x 65536 > total
timesteps) and the minibatch grows with the batch (``batch_size`` is a parameter; SB3's 64 is kept as the default
only for num_envs <= 64); the running normalisers are updated once per rollout from the merged batch moments
instead of once per env step (identical statistics, frozen within a rollout).
"""
from __future__ import annotations

import io
import json
import math
import os
import pickle
import time
import zipfile
from typing import Any, Dict, Optional

import numpy as np
import torch
import torch.distributed as dist
import torch.nn as nn

from .dist import allreduce_episode_stats, allreduce_moments

SB3_VERSION = "1.1.0a5"


class RunningMeanStd:
    """SB3 ``RunningMeanStd`` (parallel-variance update) on torch tensors."""

    def __init__(self, shape=(), device="cpu", epsilon: float = 1e-4):
        self.mean = torch.zeros(shape, dtype=torch.float64, device=device)
        self.var = torch.ones(shape, dtype=torch.float64, device=device)
        self.count = torch.tensor(epsilon, dtype=torch.float64, device=device)

    def update_from_moments(self, batch_mean, batch_var, batch_count):
        delta = batch_mean - self.mean
        tot = self.count + batch_count
        new_mean = self.mean + delta * batch_count / tot
        m2 = self.var * self.count + batch_var * batch_count + delta * delta * self.count * batch_count / tot
        # in place: the graphed rollout step reads these tensors by address
        self.mean.copy_(new_mean); self.var.copy_(m2 / tot); self.count.copy_(tot)

    def update(self, x: torch.Tensor, sync: bool = False):
        x = x.reshape(-1, *self.mean.shape).to(torch.float64)
        n = torch.tensor(float(x.shape[0]), dtype=torch.float64, device=x.device)
        mean = x.mean(0)
        m2 = ((x - mean) ** 2).sum(0)
        if sync:
            n, mean, m2 = allreduce_moments(n, mean, m2)
        self.update_from_moments(mean, m2 / n, n)

    def state(self):
        return {"mean": self.mean.cpu().numpy(), "var": self.var.cpu().numpy(), "count": float(self.count)}

    def load(self, st):
        dev = self.mean.device
        self.mean.copy_(torch.as_tensor(np.asarray(st["mean"]), dtype=torch.float64, device=dev).reshape(self.mean.shape))
        self.var.copy_(torch.as_tensor(np.asarray(st["var"]), dtype=torch.float64, device=dev).reshape(self.var.shape))
        self.count.copy_(torch.tensor(float(st["count"]), dtype=torch.float64, device=dev))


class VecNormalizeState:
    """Observation / reward normalisation with SB3 ``VecNormalize`` semantics."""

    def __init__(self, num_envs, obs_dim, device, gamma=0.99, clip_obs=10.0, clip_reward=10.0, epsilon=1e-8):
        self.obs_rms = RunningMeanStd((obs_dim,), device)
        self.ret_rms = RunningMeanStd((), device)
        self.returns = torch.zeros(num_envs, dtype=torch.float64, device=device)
        self.gamma, self.clip_obs, self.clip_reward, self.epsilon = gamma, clip_obs, clip_reward, epsilon
        self.training, self.norm_reward = True, True

    def normalize_obs(self, obs):
        o = (obs.to(torch.float64) - self.obs_rms.mean) / torch.sqrt(self.obs_rms.var + self.epsilon)
        return torch.clamp(o, -self.clip_obs, self.clip_obs).to(torch.float32)

    def normalize_reward(self, rew):
        if not self.norm_reward:
            return rew
        r = rew.to(torch.float64) / torch.sqrt(self.ret_rms.var + self.epsilon)
        return torch.clamp(r, -self.clip_reward, self.clip_reward).to(torch.float32)

    def state(self):
        return {"obs_rms": self.obs_rms.state(), "ret_rms": self.ret_rms.state(), "gamma": self.gamma, "clip_obs": self.clip_obs,
                "clip_reward": self.clip_reward, "epsilon": self.epsilon}

    def load(self, st):
        self.obs_rms.load(st["obs_rms"])
        self.ret_rms.load(st["ret_rms"])
        self.gamma, self.clip_obs, self.clip_reward = st.get("gamma", 0.99), st.get("clip_obs", 10.0), st.get("clip_reward", 10.0)


class MlpPolicy(nn.Module):
    """SB3 ``ActorCriticPolicy`` with ``net_arch=[dict(pi=[256,128], vf=[256,128])]``; state_dict keys match SB3's."""

    def __init__(self, obs_dim=19, act_dim=6, pi=(256, 128), vf=(256, 128), log_std_init=0.0):
        super().__init__()

        def mlp(sizes):
            layers, d = [], obs_dim
            for h in sizes:
                layers += [nn.Linear(d, h), nn.Tanh()]
                d = h
            return nn.Sequential(*layers)

        self.mlp_extractor = nn.Module()
        self.mlp_extractor.policy_net = mlp(pi)
        self.mlp_extractor.value_net = mlp(vf)
        self.action_net = nn.Linear(pi[-1], act_dim)
        self.value_net = nn.Linear(vf[-1], 1)
        self.log_std = nn.Parameter(torch.ones(act_dim) * log_std_init)
        for mod, gain in ((self.mlp_extractor.policy_net, math.sqrt(2)), (self.mlp_extractor.value_net, math.sqrt(2)),
                          (self.action_net, 0.01), (self.value_net, 1.0)):
            for m in mod.modules() if isinstance(mod, nn.Sequential) else [mod]:
                if isinstance(m, nn.Linear):
                    nn.init.orthogonal_(m.weight, gain=gain)
                    nn.init.zeros_(m.bias)

    def forward(self, obs):
        mean = self.action_net(self.mlp_extractor.policy_net(obs))
        value = self.value_net(self.mlp_extractor.value_net(obs)).squeeze(-1)
        return mean, value

    @staticmethod
    def log_prob(mean, log_std, actions):
        var = torch.exp(2 * log_std)
        return (-((actions - mean) ** 2) / (2 * var) - log_std - 0.5 * math.log(2 * math.pi)).sum(-1)

    def act(self, obs, deterministic=False):
        mean, value = self(obs)
        if deterministic:
            return mean, value, torch.zeros_like(value)
        a = mean + torch.randn_like(mean) * torch.exp(self.log_std)
        return a, value, self.log_prob(mean, self.log_std, a)

    def evaluate(self, obs, actions):
        mean, value = self(obs)
        ent = (0.5 + 0.5 * math.log(2 * math.pi) + self.log_std).sum().expand(obs.shape[0])
        return value, self.log_prob(mean, self.log_std, actions), ent


def compute_gae(rewards, values, dones, last_values, gamma, lam):
    """SB3 ``RolloutBuffer.compute_returns_and_advantage``.  rewards/values/dones: [T, N]; dones[t] = episode ended AT step t."""
    T = rewards.shape[0]
    adv = torch.zeros_like(rewards)
    last = torch.zeros_like(last_values)
    for t in reversed(range(T)):
        next_v = last_values if t == T - 1 else values[t + 1]
        nonterminal = 1.0 - dones[t]
        delta = rewards[t] + gamma * next_v * nonterminal - values[t]
        last = delta + gamma * lam * nonterminal * last
        adv[t] = last
    return adv, adv + values


class PPO:
    def __init__(self, env, n_steps=32, batch_size=None, n_epochs=10, learning_rate=3e-4, gamma=0.99, gae_lambda=0.95, clip_range=0.2,
                 ent_coef=0.0, vf_coef=0.5, max_grad_norm=0.5, net_arch=None, seed=0, normalize=True, verbose=0, cuda_graph=None):
        self.env = env  # BatchedUltrasound
        self.device = env.device
        self.N, self.obs_dim, self.act_dim = env.num_envs, env.obs.shape[1], env.action_dim
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank() if self.world > 1 else 0
        self.n_steps, self.n_epochs = n_steps, n_epochs
        self.batch_size = batch_size or (64 if self.N * n_steps <= 64 * 2048 and self.N <= 64 else max(64, self.N * n_steps // 32))
        self.lr, self.gamma, self.lam, self.clip, self.ent_coef, self.vf_coef, self.max_grad_norm = (
            learning_rate, gamma, gae_lambda, clip_range, ent_coef, vf_coef, max_grad_norm)
        arch = (net_arch or [dict(pi=[256, 128], vf=[256, 128])])[0]
        torch.manual_seed(seed)  # same initial weights on every rank; sampling streams are then offset by rank
        self.policy = MlpPolicy(self.obs_dim, self.act_dim, tuple(arch["pi"]), tuple(arch["vf"])).to(self.device)
        if self.world > 1:
            for p in self.policy.parameters():
                dist.broadcast(p.data, 0)
        torch.manual_seed(seed + 1000 * (self.rank + 1))
        # On a CUDA device one minibatch step is replayed as two CUDA graphs (launch bound otherwise: ~100 small kernels per step)
        self.cuda_graph = (self.device.type == "cuda") if cuda_graph is None else bool(cuda_graph)
        # fused multi-tensor Adam on CUDA: one kernel for the 12 parameter tensors instead of ~10 foreach kernels per step
        self.opt = torch.optim.Adam(self.policy.parameters(), lr=learning_rate, eps=1e-5, capturable=self.cuda_graph,
                                    fused=(self.device.type == "cuda"))
        self._graphs = None
        # graphed update: hand-written backward pass (USIM_PPO_AUTOGRAD=1: autograd, for A/B measurements)
        self.manual_backward = os.environ.get("USIM_PPO_AUTOGRAD", "0") != "1"
        self._splitk, self._ones_row, self._one = int(os.environ.get("USIM_PPO_SPLITK", "16")), None, None
        self._rgraphs, self.graph_rollout = None, os.environ.get("USIM_PPO_EAGER_ROLLOUT", "0") != "1"
        if os.environ.get("USIM_PPO_TF32", "0") == "1":  # developer knob: TF32 tensor-core GEMMs in the update (default: true fp32, as SB3)
            torch.backends.cuda.matmul.allow_tf32 = True
        self.norm = VecNormalizeState(self.N, self.obs_dim, self.device, gamma=gamma) if normalize else None
        lo, hi = env.action_spec
        self.act_lo = torch.as_tensor(lo, dtype=torch.float32, device=self.device)
        self.act_hi = torch.as_tensor(hi, dtype=torch.float32, device=self.device)
        self.num_timesteps, self._n_updates, self.verbose = 0, 0, verbose
        self.ep_ret = torch.zeros(self.N, dtype=torch.float64, device=self.device)
        self.ep_len = torch.zeros(self.N, dtype=torch.float64, device=self.device)
        self.last_stats: Dict[str, float] = {}
        self._last_obs = None
        # opt-in device timing of the NCCL gradient all-reduce (bench.py --workload ppo): pairs of CUDA events, drained by allreduce_ms()
        self.profile_allreduce = False
        self._ar_events = []

    # ------------------------------------------------------------------ rollouts
    def _setup(self):
        self._last_obs = self.env.reset().clone()
        if self.norm is not None:
            self.norm.obs_rms.update(self._last_obs, sync=True)

    def _capture_rollout(self):
        """One rollout step as two CUDA graphs around the env step (which launches through the C ABI): A = normalise + policy +
        clamp, B = reward / episode bookkeeping.  The eager loop issues ~60 small torch ops per step and is bound by the HOST's launch
        rate (45 ms per 32 steps on a quiet box, 82 ms on a busy one); replayed, a step costs the host two graph launches."""
        N, dev = self.N, self.device
        st = dict(obs=self._last_obs.clone(), ep_r=torch.zeros((), dtype=torch.float64, device=dev),
                  ep_l=torch.zeros((), dtype=torch.float64, device=dev), ep_n=torch.zeros((), dtype=torch.float64, device=dev))

        def part_a():
            nobs = self.norm.normalize_obs(st["obs"]) if self.norm else st["obs"] * 1.0
            a, v, lp = self.policy.act(nobs)
            return nobs, a, v, lp, torch.max(torch.min(a, self.act_hi), self.act_lo)

        def part_b(o, r, d):
            d32 = d.to(torch.float32)
            self.ep_ret.add_(r.to(torch.float64))
            self.ep_len.add_(1)
            trace = None
            if self.norm is not None:
                self.norm.returns.mul_(self.gamma).add_(r.to(torch.float64))
                trace = self.norm.returns.clone()
                self.norm.returns.mul_(1 - d32.to(torch.float64))
            dm = d32.to(torch.float64)
            st["ep_r"].add_((self.ep_ret * dm).sum()); st["ep_l"].add_((self.ep_len * dm).sum()); st["ep_n"].add_(dm.sum())
            self.ep_ret.mul_(1 - dm)
            self.ep_len.mul_(1 - dm)
            st["obs"].copy_(o)
            return r * 1.0, d32, trace

        o, r, d = self.env.obs, self.env.rew, self.env.done
        saved = [x.clone() for x in (self.ep_ret, self.ep_len, st["obs"])] + ([self.norm.returns.clone()] if self.norm else [])
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(3):
                part_a(); part_b(o, r, d)
        torch.cuda.current_stream(dev).wait_stream(side)
        ga, gb = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.no_grad():
            with torch.cuda.graph(ga):
                out_a = part_a()
            with torch.cuda.graph(gb, pool=ga.pool()):
                out_b = part_b(o, r, d)
        self.ep_ret.copy_(saved[0]); self.ep_len.copy_(saved[1]); st["obs"].copy_(saved[2])
        if self.norm:
            self.norm.returns.copy_(saved[3])
        for k in ("ep_r", "ep_l", "ep_n"):
            st[k].zero_()
        self._rgraphs = dict(a=ga, b=gb, out_a=out_a, out_b=out_b, st=st)

    def _collect_rollouts_graphed(self):
        T, N, dev = self.n_steps, self.N, self.device
        if self._rgraphs is None:
            self._capture_rollout()
        g = self._rgraphs
        st = g["st"]
        obs_b = torch.empty(T, N, self.obs_dim, device=dev)
        raw_b = torch.empty(T, N, self.obs_dim, device=dev)
        act_b = torch.empty(T, N, self.act_dim, device=dev)
        rew_b, val_b, logp_b, done_b = (torch.empty(T, N, device=dev) for _ in range(4))
        ret_trace = torch.empty(T, N, dtype=torch.float64, device=dev)
        st["obs"].copy_(self._last_obs)
        for k in ("ep_r", "ep_l", "ep_n"):
            st[k].zero_()
        nobs, a, v, lp, ac = g["out_a"]
        r1, d32, trace = g["out_b"]
        with torch.no_grad():
            for t in range(T):
                g["a"].replay()
                raw_b[t].copy_(st["obs"]); obs_b[t].copy_(nobs); act_b[t].copy_(a); val_b[t].copy_(v); logp_b[t].copy_(lp)
                self.env.step(ac, auto_reset=True)
                g["b"].replay()
                rew_b[t].copy_(r1); done_b[t].copy_(d32)
                if trace is not None:
                    ret_trace[t].copy_(trace)
            self._last_obs = st["obs"].clone()
        return self._finish_rollout(obs_b, raw_b, act_b, rew_b, val_b, logp_b, done_b, ret_trace, st["ep_r"].clone(), st["ep_l"].clone(), st["ep_n"].clone())

    @torch.no_grad()
    def _finish_rollout(self, obs_b, raw_b, act_b, rew_b, val_b, logp_b, done_b, ret_trace, ep_r, ep_l, ep_n):
        """Bootstrap value, one merge of the rollout's moments (NCCL all-reduce when world > 1), reward normalisation, GAE, statistics."""
        T, N = rew_b.shape
        obs = self._last_obs
        last_v = self.policy(self.norm.normalize_obs(obs) if self.norm else obs)[1]
        if self.norm is not None:
            self.norm.ret_rms.update(ret_trace, sync=True)
            rew_n = self.norm.normalize_reward(rew_b)
            self.norm.obs_rms.update(raw_b, sync=True)
        else:
            rew_n = rew_b
        adv, ret = compute_gae(rew_n, val_b, done_b, last_v, self.gamma, self.lam)
        ep_r, ep_l, ep_n = allreduce_episode_stats(ep_r, ep_l, ep_n)
        self.num_timesteps += T * N * self.world
        self.last_stats.update(ep_rew_mean=float(ep_r / ep_n) if ep_n > 0 else float("nan"),
                               ep_len_mean=float(ep_l / ep_n) if ep_n > 0 else float("nan"), episodes=float(ep_n),
                               step_reward_mean=float(rew_b.mean()))
        flat = lambda x: x.reshape(T * N, *x.shape[2:])
        return flat(obs_b), flat(act_b), flat(val_b), flat(logp_b), flat(adv), flat(ret)

    def collect_rollouts(self):
        if self.cuda_graph and self.graph_rollout:
            return self._collect_rollouts_graphed()
        T, N, dev = self.n_steps, self.N, self.device
        obs_b = torch.empty(T, N, self.obs_dim, device=dev)
        raw_b = torch.empty(T, N, self.obs_dim, device=dev)
        act_b = torch.empty(T, N, self.act_dim, device=dev)
        rew_b, val_b, logp_b, done_b = (torch.empty(T, N, device=dev) for _ in range(4))
        ret_trace = torch.empty(T, N, dtype=torch.float64, device=dev)
        ep_r, ep_l, ep_n = (torch.zeros((), dtype=torch.float64, device=dev) for _ in range(3))
        obs = self._last_obs
        with torch.no_grad():
            for t in range(T):
                nobs = self.norm.normalize_obs(obs) if self.norm else obs
                a, v, lp = self.policy.act(nobs)
                raw_b[t], obs_b[t], act_b[t], val_b[t], logp_b[t] = obs, nobs, a, v, lp
                o, r, d, _ = self.env.step(torch.max(torch.min(a, self.act_hi), self.act_lo), auto_reset=True)
                d = d.to(torch.float32)
                rew_b[t], done_b[t] = r, d
                self.ep_ret += r.to(torch.float64)
                self.ep_len += 1
                if self.norm is not None:
                    self.norm.returns = self.norm.returns * self.gamma + r.to(torch.float64)
                    ret_trace[t] = self.norm.returns
                    self.norm.returns = self.norm.returns * (1 - d.to(torch.float64))
                dm = d.to(torch.float64)
                ep_r += (self.ep_ret * dm).sum(); ep_l += (self.ep_len * dm).sum(); ep_n += dm.sum()
                self.ep_ret *= 1 - dm
                self.ep_len *= 1 - dm
                obs = o.clone()
            self._last_obs = obs
        return self._finish_rollout(obs_b, raw_b, act_b, rew_b, val_b, logp_b, done_b, ret_trace, ep_r, ep_l, ep_n)

    # ------------------------------------------------------------------ update
    def _allreduce_grads(self):
        if self.world == 1:
            return
        grads = [p.grad for p in self.policy.parameters() if p.grad is not None]
        flat = torch.cat([g.reshape(-1) for g in grads])  # one flat bucket (~308 KB fp32), latency bound
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat /= self.world
        off = 0
        for g in grads:
            n = g.numel()
            g.copy_(flat[off:off + n].view_as(g))
            off += n

    def _timed_allreduce(self, flat):
        if self.profile_allreduce:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
            e1.record()
            self._ar_events.append((e0, e1))
        else:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)

    def allreduce_ms(self):
        """(total device ms, count) of the gradient all-reduces timed since the last call (profile_allreduce)."""
        torch.cuda.synchronize(self.device)
        ms, n = sum(a.elapsed_time(b) for a, b in self._ar_events), len(self._ar_events)
        self._ar_events = []
        return ms, n

    def _minibatch_loss(self, obs, act, old_lp, adv, ret):
        a = (adv - adv.mean()) / (adv.std() + 1e-8)
        v, lp, ent = self.policy.evaluate(obs, act)
        ratio = torch.exp(lp - old_lp)
        pl = -torch.min(a * ratio, a * torch.clamp(ratio, 1 - self.clip, 1 + self.clip)).mean()
        vl = ((ret - v) ** 2).mean()
        loss = pl + self.ent_coef * (-ent.mean()) + self.vf_coef * vl
        return loss, pl, vl, (old_lp - lp).mean()

    @torch.no_grad()
    def _manual_grads(self, obs, act, old_lp, adv, ret):
        """The minibatch loss of ``_minibatch_loss`` and its gradient WITHOUT an autograd graph: forward through the two tanh MLPs,
        the PPO head differentiated by hand, backward as explicit GEMMs written straight into the ``.grad`` views of the flat
        gradient buffer.  A third of the kernels autograd launches for the same arithmetic (the update is launch bound: ~77 k
        parameters, 8192-sample minibatches).  Gradients agree with autograd to fp32 round-off (tests/test_ppo_cpu.py).
        Returns (policy loss, value loss, approx kl)."""
        pol, B = self.policy, obs.shape[0]

        def forward(net, head):
            l1, l2 = net[0], net[2]
            h1 = torch.addmm(l1.bias, obs, l1.weight.t()).tanh_()
            h2 = torch.addmm(l2.bias, h1, l2.weight.t()).tanh_()
            return h1, h2, torch.addmm(head.bias, h2, head.weight.t())

        # Weight gradients are [out, in] = g^T h with the BATCH as the contraction: a tall-skinny product whose output (at most
        # 256 x 128) is a handful of tiles -- cuBLAS runs it on ~8 CTAs of 148 SMs (58 us for the 256 x 128 layer).  Split the batch
        # into S slabs (one batched GEMM: S times the CTAs) and add the partial products.  Bias gradients ride on a GEMM with a row
        # of ones instead of a column-reduction kernel.
        S = self._splitk if B % self._splitk == 0 and B >= 64 * self._splitk else 1
        if self._ones_row is None or self._ones_row.shape[1] != B:
            self._ones_row = torch.ones(1, B, device=obs.device)

        def wgrad(g, h, lin):
            if S > 1:
                torch.sum(torch.bmm(g.view(S, B // S, -1).transpose(1, 2), h.view(S, B // S, -1)), 0, out=lin.weight.grad)
            else:
                torch.mm(g.t(), h, out=lin.weight.grad)
            torch.mm(self._ones_row, g, out=lin.bias.grad.view(1, -1))

        def backward(net, head, h1, h2, g_out):
            l1, l2 = net[0], net[2]
            wgrad(g_out, h2, head)
            g = torch.mm(g_out, head.weight).mul_(torch.addcmul(self._one, h2, h2, value=-1.0))  # through tanh: 1 - h^2
            wgrad(g, h1, l2)
            g = torch.mm(g, l2.weight).mul_(torch.addcmul(self._one, h1, h1, value=-1.0))
            wgrad(g, obs, l1)

        if self._one is None:
            self._one = torch.ones((), device=obs.device)
        p1, p2, mean = forward(pol.mlp_extractor.policy_net, pol.action_net)
        v1, v2, value = forward(pol.mlp_extractor.value_net, pol.value_net)
        value = value.squeeze(-1)
        # head (SB3 PPO.train): normalised advantage, clipped surrogate, value loss, entropy of the state-independent Gaussian
        a = (adv - adv.mean()) / (adv.std() + 1e-8)
        inv_var = torch.exp(-2.0 * pol.log_std)
        d = act - mean
        dv = d * inv_var
        lp = -0.5 * (d * dv).sum(-1) - pol.log_std.sum() - 0.5 * math.log(2 * math.pi) * act.shape[1]
        ratio = torch.exp(lp - old_lp)
        s1 = a * ratio
        s2 = a * torch.clamp(ratio, 1 - self.clip, 1 + self.clip)
        pl = -torch.minimum(s1, s2).mean()
        verr = value - ret
        vl = (verr * verr).mean()
        kl = (old_lp - lp).mean()
        # d(-min(s1, s2)) / d lp: the unclipped branch carries the gradient wherever it is the smaller one or the clip is inactive
        dlp = torch.where(s1 <= s2, s1, torch.zeros_like(s1)).mul_(-1.0 / B)
        g_mean = dv * dlp.unsqueeze(-1)
        torch.sum((d * dv - 1.0) * dlp.unsqueeze(-1), 0, out=pol.log_std.grad)
        if self.ent_coef != 0.0:
            pol.log_std.grad.sub_(self.ent_coef)
        backward(pol.mlp_extractor.policy_net, pol.action_net, p1, p2, g_mean)
        backward(pol.mlp_extractor.value_net, pol.value_net, v1, v2, (verr * (2.0 * self.vf_coef / B)).unsqueeze(-1))
        return pl, vl, kl

    def _pack(self, packed, batch):
        od, ad = self.obs_dim, self.act_dim
        packed[:, :od].copy_(batch[0]); packed[:, od:od + ad].copy_(batch[1])
        packed[:, od + ad].copy_(batch[3]); packed[:, od + ad + 1].copy_(batch[4]); packed[:, od + ad + 2].copy_(batch[5])

    def _capture(self, batch):
        """Capture one minibatch step as two CUDA graphs sharing a memory pool: A = gather + forward + loss + backward + flat
        gradient bucket, B = (bucket / world) -> grads, clip, Adam.  Between them the bucket is all-reduced eagerly (NCCL)."""
        n, B, dev = batch[0].shape[0], self.batch_size, self.device
        # the rollout as ONE packed array [n][obs | act | old_lp | adv | ret]: a minibatch is one row gather, its fields are column views
        od, ad = self.obs_dim, self.act_dim
        packed = torch.empty(n, od + ad + 3, device=dev)
        self._pack(packed, batch)
        idx = torch.arange(B, device=dev)

        def fields():
            mb = packed[idx]
            return mb[:, :od], mb[:, od:od + ad], mb[:, od + ad], mb[:, od + ad + 1], mb[:, od + ad + 2]

        params = [p for p in self.policy.parameters()]
        # the warm-up steps below are real optimiser steps on real data: put parameters and Adam state back afterwards
        saved_p = [p.detach().clone() for p in params]
        saved_s = {p: {k: v.clone() for k, v in st.items() if torch.is_tensor(v)} for p, st in self.opt.state.items()}
        stats = [torch.zeros((), device=dev) for _ in range(3)]

        # every p.grad is a VIEW into one flat buffer: backward accumulates into it in place, the NCCL all-reduce works on the
        # buffer itself, clipping is one norm + one scale of the buffer -- no concatenation, no copies back
        flat_grad = torch.zeros(sum(p.numel() for p in params), device=dev)
        off = 0
        for p in params:
            k = p.numel()
            p.grad = flat_grad[off:off + k].view_as(p)
            off += k

        def part_a():
            if self.manual_backward:  # every gradient entry is overwritten: no zeroing, no autograd graph
                pl, vl, kl = self._manual_grads(*fields())
            else:
                flat_grad.zero_()
                loss, pl, vl, kl = self._minibatch_loss(*fields())
                loss.backward()
            for t, v in zip(stats, (pl, vl, kl)):
                t.copy_(v.detach())
            return flat_grad

        def part_b(flat):
            if self.world > 1:
                flat.div_(self.world)
            flat.mul_(torch.clamp(self.max_grad_norm / (flat.norm() + 1e-6), max=1.0))  # clip_grad_norm_ on the flat buffer
            self.opt.step()

        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(3):
                part_b(part_a())
        torch.cuda.current_stream(dev).wait_stream(side)
        ga, gb = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(ga):
            flat = part_a()
        with torch.cuda.graph(gb, pool=ga.pool()):
            part_b(flat)
        with torch.no_grad():
            for p, q in zip(params, saved_p):
                p.copy_(q)
            for p, st in self.opt.state.items():
                for k, v in st.items():
                    if torch.is_tensor(v):
                        v.copy_(saved_s[p][k]) if p in saved_s and k in saved_s[p] else v.zero_()
        self._graphs = dict(a=ga, b=gb, flat=flat, idx=idx, packed=packed, stats=stats, n=n)

    def _train_graphed(self, batch):
        if self._graphs is None or self._graphs["n"] != batch[0].shape[0]:
            self._capture(batch)
        g = self._graphs
        self._pack(g["packed"], batch)
        n = g["n"]
        for _ in range(self.n_epochs):
            perm = torch.randperm(n, device=self.device)
            for s in range(0, n - self.batch_size + 1, self.batch_size):
                g["idx"].copy_(perm[s:s + self.batch_size])
                g["a"].replay()
                if self.world > 1:
                    self._timed_allreduce(g["flat"])
                g["b"].replay()
            self._n_updates += 1
        pl, vl, kl = (float(t) for t in g["stats"])
        self.last_stats.update(policy_loss=pl, value_loss=vl, approx_kl=kl, n_updates=self._n_updates)

    def train(self, batch):
        if self.cuda_graph:
            return self._train_graphed(batch)
        obs, act, old_v, old_lp, adv, ret = batch
        n = obs.shape[0]
        pl = vl = kl = torch.zeros((), device=self.device)
        for _ in range(self.n_epochs):
            perm = torch.randperm(n, device=self.device)
            for s in range(0, n - self.batch_size + 1, self.batch_size):
                idx = perm[s:s + self.batch_size]
                a = adv[idx]
                a = (a - a.mean()) / (a.std() + 1e-8)
                v, lp, ent = self.policy.evaluate(obs[idx], act[idx])
                ratio = torch.exp(lp - old_lp[idx])
                pl = -torch.min(a * ratio, a * torch.clamp(ratio, 1 - self.clip, 1 + self.clip)).mean()
                vl = ((ret[idx] - v) ** 2).mean()
                loss = pl + self.ent_coef * (-ent.mean()) + self.vf_coef * vl
                self.opt.zero_grad(set_to_none=True)
                loss.backward()
                self._allreduce_grads()
                nn.utils.clip_grad_norm_(self.policy.parameters(), self.max_grad_norm)
                self.opt.step()
                with torch.no_grad():
                    kl = (old_lp[idx] - lp).mean()
            self._n_updates += 1
        self.last_stats.update(policy_loss=float(pl.detach()), value_loss=float(vl.detach()), approx_kl=float(kl), n_updates=self._n_updates)

    def learn(self, total_timesteps: int, log_interval: int = 1, callback=None):
        if self._last_obs is None:
            self._setup()
        it, t0 = 0, time.time()
        while self.num_timesteps < total_timesteps:
            self.train(self.collect_rollouts())
            it += 1
            if self.verbose and self.rank == 0 and it % log_interval == 0:
                fps = self.num_timesteps / max(time.time() - t0, 1e-9)
                print(f"[ppo] it {it} steps {self.num_timesteps} fps {fps:.0f} " + " ".join(f"{k} {v:.4g}" for k, v in self.last_stats.items()), flush=True)
            if callback is not None:
                callback(self)
        return self

    @torch.no_grad()
    def predict(self, obs, deterministic=True):
        nobs = self.norm.normalize_obs(obs) if self.norm else obs
        a = self.policy.act(nobs, deterministic=deterministic)[0]
        return torch.max(torch.min(a, self.act_hi), self.act_lo)

    # ------------------------------------------------------------------ checkpoints (SB3 zip layout, SURVEY §5 [ART])
    def save(self, path: str):
        if self.rank != 0:
            return
        path = path if path.endswith(".zip") else path + ".zip"
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        data = {"num_timesteps": self.num_timesteps, "_n_updates": self._n_updates, "n_envs": self.N * self.world, "n_steps": self.n_steps,
                "batch_size": self.batch_size, "n_epochs": self.n_epochs, "gamma": self.gamma, "gae_lambda": self.lam, "ent_coef": self.ent_coef,
                "vf_coef": self.vf_coef, "max_grad_norm": self.max_grad_norm, "learning_rate": self.lr, "clip_range": self.clip,
                "policy_class": "MlpPolicy", "net_arch": [dict(pi=[256, 128], vf=[256, 128])]}
        with zipfile.ZipFile(path, "w") as z:
            z.writestr("data", json.dumps(data))
            for name, obj in (("policy.pth", self.policy.state_dict()), ("policy.optimizer.pth", self.opt.state_dict()), ("pytorch_variables.pth", {})):
                buf = io.BytesIO()
                torch.save(obj, buf)
                z.writestr(name, buf.getvalue())
            z.writestr("_stable_baselines3_version", SB3_VERSION)
        if self.norm is not None:
            with open(os.path.join(os.path.dirname(os.path.abspath(path)), "vec_normalize_" + os.path.basename(path)[:-4] + ".pkl"), "wb") as f:
                pickle.dump(self.norm.state(), f)

    def load(self, path: str, load_optimizer: bool = True):
        path = path if path.endswith(".zip") else path + ".zip"
        sd, data, opt = load_sb3_zip(path)
        self.policy.load_state_dict({k: v.to(self.device) for k, v in sd.items()})
        if load_optimizer and opt is not None:
            try:
                self.opt.load_state_dict(opt)
            except Exception:
                pass
        self.num_timesteps, self._n_updates = int(data.get("num_timesteps", 0)), int(data.get("_n_updates", 0))
        vn = os.path.join(os.path.dirname(os.path.abspath(path)), "vec_normalize_" + os.path.basename(path)[:-4] + ".pkl")
        if self.norm is not None and os.path.exists(vn):
            self.norm.load(load_vecnormalize(vn))
        return self


def load_sb3_zip(path: str):
    """(policy state_dict, data dict, optimizer state or None) from an SB3-layout zip (the reference's trained_rl_models/*.zip)."""
    with zipfile.ZipFile(path) as z:
        data = json.loads(z.read("data"))
        sd = torch.load(io.BytesIO(z.read("policy.pth")), map_location="cpu", weights_only=False)
        opt = None
        if "policy.optimizer.pth" in z.namelist():
            try:
                opt = torch.load(io.BytesIO(z.read("policy.optimizer.pth")), map_location="cpu", weights_only=False)
            except Exception:
                opt = None
    return sd, data, opt


class _StubUnpickler(pickle.Unpickler):
    """Reads SB3's VecNormalize pickle without SB3/gym installed (SURVEY App. D recipe)."""

    def find_class(self, module, name):
        if module.startswith("numpy") or module in ("collections", "builtins", "_codecs"):
            return super().find_class(module, name)
        return type(name, (), {"__setstate__": lambda self, st: self.__dict__.update(st if isinstance(st, dict) else {})})


def load_vecnormalize(path: str) -> Dict[str, Any]:
    """Normaliser statistics from our own pickle (plain dict) or from an SB3 ``VecNormalize.save`` pickle."""
    with open(path, "rb") as f:
        obj = _StubUnpickler(f).load()
    if isinstance(obj, dict):
        return obj
    rms = lambda r: {"mean": np.asarray(r.mean), "var": np.asarray(r.var), "count": float(r.count)}
    return {"obs_rms": rms(obj.obs_rms), "ret_rms": rms(obj.ret_rms), "gamma": float(obj.gamma), "clip_obs": float(obj.clip_obs),
            "clip_reward": float(getattr(obj, "clip_reward", 10.0)), "epsilon": float(getattr(obj, "epsilon", 1e-8))}

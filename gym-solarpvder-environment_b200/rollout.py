"""Policy-in-the-loop rollout (BASELINE.json config 5): the loop shape of the reference's tf-agents
DQN demo (examples/gym_PVDER_environment_tf_agents_DQN_demo.ipynb: QNetwork fc(100,), one collect
step per iteration, replay buffer) with every piece on the same GPU and stream:

    obs[N,11] f32 --Q-net (11-100-5)--> argmax / eps-greedy --> actions[N] i32 --step kernel--> obs'

Nothing crosses PCIe or NVLink per step; the whole iteration (policy forward, action selection,
env step, replay write) can be captured once in a CUDA graph and replayed.

policy="fused" (default for the demo's 11-h-5 ReLU network in fp32): one hand-written kernel per
iteration next to the step kernel (``pvder_qnet_collect``: weights in shared memory, one thread per env,
Philox epsilon-greedy, and the replay-ring writes of the transition just completed and of the one being
started); policy="torch": the same network and ring through torch ops (library GEMMs, five copy
kernels), kept as the cross-check and for arbitrary modules.
"""
from __future__ import annotations

import time


def make_qnet(hidden=100, device="cuda", dtype=None):
    """QNetwork(fc_layer_params=(100,)) of the reference demo (ipynb cell 'q_net')."""
    import torch

    net = torch.nn.Sequential(torch.nn.Linear(11, hidden), torch.nn.ReLU(), torch.nn.Linear(hidden, 5))
    return net.to(device=device, dtype=dtype or torch.float32)


class DQNRollout:
    """Collect driver over a PVDERVecEnv (env should be created with auto_reset=True)."""

    def __init__(self, venv, qnet=None, epsilon=0.1, replay_steps=8, use_cuda_graph=True, seed=0, policy="auto"):
        import torch

        self.torch = torch
        self.venv = venv
        self.dev = venv.device
        self.qnet = qnet if qnet is not None else make_qnet(device=self.dev)
        self.epsilon = float(epsilon)
        n = venv.num_envs
        self.actions = torch.zeros(n, dtype=torch.int32, device=self.dev)
        # replay ring in SoA layout: [T, N, ...]
        self.T = int(replay_steps)
        self.rb_obs = torch.zeros((self.T, n, 11), dtype=torch.float32, device=self.dev)
        self.rb_next = torch.zeros((self.T, n, 11), dtype=torch.float32, device=self.dev)
        self.rb_act = torch.zeros((self.T, n), dtype=torch.int32, device=self.dev)
        self.rb_rew = torch.zeros((self.T, n), dtype=torch.float32, device=self.dev)
        self.rb_done = torch.zeros((self.T, n), dtype=torch.bool, device=self.dev)
        self.slot = torch.zeros((), dtype=torch.int64, device=self.dev)
        self.use_graph = bool(use_cuda_graph)
        self.graph = None
        self.steps_done = 0
        self.seed = int(seed)
        self._primed = False
        fusable = self._fusable(self.qnet)
        if policy == "auto":
            policy = "fused" if fusable else "torch"
        if policy not in ("fused", "torch"):
            raise ValueError("policy must be 'auto', 'fused' or 'torch'")
        if policy == "fused" and not fusable:
            raise ValueError("policy='fused' needs Sequential(Linear(11, h), ReLU, Linear(h, 5)) in float32 with h <= 256")
        self.policy = policy

    def _fusable(self, net):
        t = self.torch
        try:
            l1, act, l2 = net[0], net[1], net[2]
            return (len(net) == 3 and isinstance(l1, t.nn.Linear) and isinstance(act, t.nn.ReLU)
                    and isinstance(l2, t.nn.Linear) and l1.in_features == 11 and l2.out_features == 5
                    and l1.out_features == l2.in_features <= 256 and l1.weight.dtype == t.float32
                    and l1.bias is not None and l2.bias is not None)
        except Exception:
            return False

    def _select_actions_fused(self, obs):
        import ctypes as C
        from . import _cabi

        l1, l2 = self.qnet[0], self.qnet[2]
        p = lambda x: C.c_void_p(x.data_ptr())
        _cabi.check(_cabi.load().pvder_qnet_policy(
            p(obs), p(l1.weight), p(l1.bias), p(l2.weight), p(l2.bias), l1.out_features, self.epsilon,
            self.seed & 0xFFFFFFFFFFFFFFFF, 0, p(self.slot), p(self.actions), None, self.venv.num_envs,
            self.venv.env_offset, C.c_void_p(self.torch.cuda.current_stream(self.dev).cuda_stream)))

    def _select_actions(self, obs):
        t = self.torch
        if self.policy == "fused":
            return self._select_actions_fused(obs)
        with t.no_grad():
            q = self.qnet(obs.to(next(self.qnet.parameters()).dtype))
            greedy = q.argmax(dim=1).to(t.int32)
            if self.epsilon > 0.0:
                u = t.rand(obs.shape[0], device=self.dev)
                rnd = t.randint(0, 5, (obs.shape[0],), device=self.dev, dtype=t.int32)
                greedy = t.where(u < self.epsilon, rnd, greedy)
            self.actions.copy_(greedy)

    def _collect_fused(self, finish):
        import ctypes as C
        from . import _cabi

        venv = self.venv
        l1, l2 = self.qnet[0], self.qnet[2]
        p = lambda x: C.c_void_p(x.data_ptr()) if x is not None else None
        rew_f64 = None if venv.cfg.DISCRETE_REWARD else venv.reward
        rew_i32 = venv.reward_i if venv.cfg.DISCRETE_REWARD else None
        _cabi.check(_cabi.load().pvder_qnet_collect(
            p(venv.obs), p(l1.weight), p(l1.bias), p(l2.weight), p(l2.bias), l1.out_features, self.epsilon,
            self.seed & 0xFFFFFFFFFFFFFFFF, p(self.slot), p(self.actions), p(self.rb_obs), p(self.rb_next), p(self.rb_act),
            p(self.rb_rew), p(self.rb_done), self.T, p(rew_f64), p(rew_i32), p(venv.done), 1 if finish else 0,
            venv.num_envs, venv.env_offset, C.c_void_p(self.torch.cuda.current_stream(self.dev).cuda_stream)))

    def _iteration(self):
        """One collect step for every env: policy -> step kernel -> replay write."""
        t = self.torch
        venv = self.venv
        if self.policy == "fused":
            # the action for the current obs is already in self.actions (prologue / previous iteration):
            # env step, count the completed transition, then ONE kernel closes its ring slot, evaluates the
            # policy on the new obs and opens the next slot
            if not self._primed:
                self._collect_fused(finish=False)
                self._primed = True
            venv.step(self.actions)
            self.slot += 1
            self._collect_fused(finish=True)
            return
        slot = self.slot % self.T
        self.rb_obs.index_copy_(0, slot.view(1), venv.obs.unsqueeze(0))
        self._select_actions(venv.obs)
        obs, rew, done, _ = venv.step(self.actions)
        self.rb_act.index_copy_(0, slot.view(1), self.actions.unsqueeze(0))
        self.rb_rew.index_copy_(0, slot.view(1), rew.to(t.float32).unsqueeze(0))
        self.rb_done.index_copy_(0, slot.view(1), done.unsqueeze(0))
        self.rb_next.index_copy_(0, slot.view(1), obs.unsqueeze(0))
        self.slot += 1

    def collect(self, n_steps, warmup=3):
        """Run n_steps collect iterations; returns timing (CUDA events on the launching stream)."""
        t = self.torch
        with t.cuda.device(self.dev):
            stream = t.cuda.current_stream(self.dev)
            if self.use_graph and self.graph is None:
                side = t.cuda.Stream(device=self.dev)
                side.wait_stream(stream)
                with t.cuda.stream(side):
                    for _ in range(max(1, warmup)):
                        self._iteration()
                stream.wait_stream(side)
                t.cuda.synchronize(self.dev)
                self.graph = t.cuda.CUDAGraph()
                with t.cuda.graph(self.graph):
                    self._iteration()
            else:
                for _ in range(warmup):
                    self._iteration()
            e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
            t.cuda.synchronize(self.dev)
            w0 = time.perf_counter()
            e0.record()
            for _ in range(n_steps):
                if self.graph is not None:
                    self.graph.replay()
                else:
                    self._iteration()
            e1.record()
            t.cuda.synchronize(self.dev)
            wall = time.perf_counter() - w0
        ms = e0.elapsed_time(e1)
        self.steps_done += n_steps
        n = self.venv.num_envs
        return {"env_steps_per_s": n * n_steps / (ms * 1e-3), "ms_per_iteration": ms / n_steps,
                "wall_s": wall, "n_envs": n, "steps": n_steps, "cuda_graph": self.graph is not None, "policy": self.policy,
                "obs_bytes_per_step": n * 44, "action_bytes_per_step": n * 4}

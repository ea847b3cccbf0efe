"""Action selection, mirrors src/policy.py:7-87 (EpsilonGreedy) of the reference."""
import numpy as np
import torch

from . import _lib
from .model import DQN


class EpsilonGreedy:
    """policy.py:7-87.  Compat calls (numpy obs of one env) consume the global legacy
    `np.random` stream exactly like the reference (randint(n_act,size=A) THEN rand(A),
    policy.py:46-47); batched calls (CUDA tensors) draw from a device Philox stream and never
    touch the host."""

    def __init__(self, env, model, action_space, args, seed=0) -> None:
        self._env = env
        self._enable_action_mask = hasattr(self._env, "enable_action_mask") and self._env.enable_action_mask
        self._model = model
        self._action_space = action_space
        self._args = args
        self._epsilon = args.epsilon
        self._step = 0
        self._epsilon_tmp = None
        self._seed = seed
        self._dev_step = None  # _lib.DeviceCounter in CUDA-graph mode (rollout.Rollout)

    def _decay(self):
        if (self._epsilon > 0 and self._step > self._args.step_before_train
                and self._step % self._args.epsilon_update_freq == 0):
            self._epsilon *= self._args.epsilon_decay
            if self._epsilon < 0.01:
                self._epsilon = 0.01

    def __call__(self, obs, adj):
        self._step += 1
        batched = isinstance(obs, tuple) or torch.is_tensor(obs)
        device = next(self._model.parameters()).device
        with torch.no_grad():
            if batched:
                obs_a, obs_g = obs if isinstance(obs, tuple) else (obs, None)
                mask = None
                if self._enable_action_mask:
                    mask = self._env._out.get("action_mask_out")
                if isinstance(self._model, DQN):
                    step, step_dev = self._step, None
                    if self._dev_step is not None:
                        step, step_dev = self._dev_step.offset(self._step), self._dev_step.ptr()
                    _, actions = self._model.act(obs_a, obs_g, action_mask=mask, epsilon=self._epsilon,
                                                 seed=self._seed, step=step, want_q=False, step_dev=step_dev)
                else:
                    x = obs_a if obs_g is None else torch.cat((obs_a, obs_g), -1)
                    q = self._model(x, adj.float())
                    if mask is not None:
                        q = q.masked_fill(mask.bool(), float("-inf"))
                    g = torch.Generator(device=q.device).manual_seed(self._seed * 1000003 + self._step)
                    ra = torch.randint(self._action_space, q.shape[:-1], device=q.device, generator=g)
                    filt = torch.rand(q.shape[:-1], device=q.device, generator=g, dtype=torch.float64) < self._epsilon
                    actions = torch.where(filt, ra, q.argmax(-1)).int()
                self._decay()
                return actions
            # ---- compat: one env, numpy in / numpy out (policy.py:20-64) ----
            A = obs.shape[0]
            mask = self._env.action_mask if self._enable_action_mask else None
            if isinstance(self._model, DQN):
                o = torch.as_tensor(np.ascontiguousarray(obs, dtype=np.float32)).to(device)
                random_actions = np.random.randint(self._action_space, size=A)
                random_u = np.random.rand(A)
                ra = torch.as_tensor(random_actions.astype(np.int32)).to(device)
                ru = torch.as_tensor(random_u).to(device)
                m = None if mask is None else torch.as_tensor(np.ascontiguousarray(mask, dtype=np.uint8)).to(device)
                _, act = self._model.act(o, None, action_mask=m, epsilon=self._epsilon, rand_action=ra, rand_u=ru,
                                         want_q=False)
                actions = act.cpu().numpy().astype(np.int64)
            else:
                o = torch.tensor(obs, dtype=torch.float32).unsqueeze(0).to(device)
                a = torch.tensor(adj, dtype=torch.float32).unsqueeze(0).to(device)
                q_values = self._model(o, a).cpu().squeeze(0).detach().numpy()
                if mask is not None:
                    q_values[mask.nonzero()] = float("-inf")
                random_actions = np.random.randint(self._action_space, size=A)
                random_filter = np.random.rand(A) < self._epsilon
                actions = np.argmax(q_values, axis=-1) * ~random_filter + random_filter * random_actions
        self._decay()
        return actions

    def eval(self):
        self._eps_tmp = self._epsilon
        self._epsilon = 0

    def reset(self, agents_to_reset):
        if hasattr(self._model, "state") and self._model.state is not None:
            self._model.state = self._model.state * ~torch.tensor(
                agents_to_reset, dtype=bool, device=self._model.state.device).unsqueeze(-1)

    def train(self):
        # reference quirk kept: `_epsilon_tmp` is never set by eval(), so
        # train() does not restore epsilon.
        if self._epsilon_tmp is not None:
            self._epsilon_tmp = None
            self._epsilon = self._epsilon_tmp

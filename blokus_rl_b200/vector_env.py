"""Boundary B2: a gymnasium-style vector env over the batched engine (duck-typed; gymnasium is not imported).

What the reference's PPO loop calls (``blokus_rl/ppo/trainer.py:36-38, 68, 111, 146-173, 380-386``; spaces at
``ppo/agent.py:112-114, 205-206`` and ``ppo/memory.py:18-25``):

    obs, info = envs.reset()
    envs.get_attr("ai_possible_indexes")           # list[list[int]]: the agent's legal action ids per env
    obs, reward, terminated, truncated, info = envs.step(actions)        # np int actions
    info["final_info"][i]["episode"]["r" | "l"]    # gymnasium-0.29 autoreset episode statistics
    envs.single_observation_space.shape, envs.single_action_space.n / .shape, envs.close()

The reference env (``blokus_gym:blokus-simple-v0``, absent) is single-agent: the agent is player 0 and random
bots play the other colours inside ``step`` (docs/README.md:47-51).  Reward: 0 until the game
ends, then +1 win / 0 draw / -1 loss for the agent.  Observation: the ``[N, N]`` board contents (0 empty,
1..P colour), which is what ``CnnAgent`` unsqueezes to one channel (``ppo/agent.py:96-99, 112-113``).

All envs live on the GPU; opponents are stepped by the engine's on-device Philox sampler, so one ``step`` is
at most P kernel launches regardless of ``num_envs``.  GPU-native callers can skip the NumPy surface entirely:
``step_device`` takes/returns CUDA tensors and ``action_mask`` is the bool ``[num_envs, A]`` tensor that
``FilterLegalMoves`` (``ppo/agent.py:33-42``) would otherwise rebuild from index lists.
"""
from __future__ import annotations

import numpy as np
import torch


class _Box:
    def __init__(self, shape, dtype=np.float32, low=0, high=4):
        self.shape, self.dtype, self.low, self.high = tuple(shape), dtype, low, high


class _Discrete:
    def __init__(self, n):
        self.n, self.shape, self.dtype = int(n), (), np.int64


class BlokusVectorEnv:
    metadata = {"render_modes": ["rgb_array"]}

    def __init__(self, num_envs: int, board_size: int = 7, num_players: int = 2, seed: int = 0, engine=None,
                 device=None, agent_player: int = 0):
        if engine is None:
            from .engine import BlokusEngine          # GPU engine; no CPU fallback
            engine = BlokusEngine(board_size, num_players, device=device)
        self.eng = engine
        self.num_envs, self.agent = num_envs, agent_player
        self.N, self.P, self.A = engine.board_size, engine.num_players, engine.num_actions
        self.single_observation_space = _Box((self.N, self.N))
        self.single_action_space = _Discrete(self.A)
        self.observation_space = _Box((num_envs, self.N, self.N))
        self.action_space = _Box((num_envs,), np.int64, 0, self.A - 1)
        self.seed = seed
        self._meta = self.P * self.N + self.P
        self.states = None
        self.action_mask = None            # bool [num_envs, A] on the device: the agent's legal actions
        self._ep_len = torch.zeros(num_envs, dtype=torch.int64, device=engine.device)
        self._ep_ret = torch.zeros(num_envs, dtype=torch.float32, device=engine.device)
        self._epoch = 0

    # ---- helpers ----------------------------------------------------------------------------------------
    def _mover(self):
        return (self.states[:, self._meta] & 15).long()

    def _done(self):
        return ((self.states[:, self._meta] >> 4) & 1).bool()

    def _play_opponents(self):
        """Random bots move (one rollout launch, no host sync) until it is the agent's turn or the game is over."""
        self._epoch += 1
        self.eng.rollout(self.states, 1, seed=self.seed + self._epoch, stop_player=self.agent, out_states=self.states)

    def _refresh_mask(self):
        self.action_mask = self.eng.step(self.states, None, mask="bytes", want_count=False, want_terminal=False,
                                         want_scores=False).mask

    def _obs(self):
        return self.eng.board_contents(self.states)

    # ---- gym surface ----------------------------------------------------------------------------------------
    def reset(self, seed=None, options=None):
        if seed is not None:
            self.seed = seed
        self.states = self.eng.new_states(self.num_envs)
        self._ep_len.zero_()
        self._ep_ret.zero_()
        if self.agent != 0:
            self._play_opponents()
        self._refresh_mask()
        return self._obs().cpu().numpy().astype(np.float32), {}

    def step_device(self, actions: torch.Tensor, check: bool = True):
        """Device-side step: int32 CUDA actions in; (obs uint8 [E,N,N], reward f32 [E], terminated bool [E],
        (episode returns, episode lengths, final observations) for the envs that finished) out, all on the
        device.  With ``check=False`` nothing synchronises with the host (illegal actions then leave their env
        unchanged, as ``blk_step`` defines)."""
        eng = self.eng
        out = eng.step(self.states, actions.to(torch.int32).contiguous(), mask=None, want_count=False,
                       want_terminal=False, want_scores=False)
        if check and bool((out.flags & 2).any()):
            raise ValueError("illegal action passed to BlokusVectorEnv.step")
        self._play_opponents()
        flags, term, _ = eng.game_ended(self.states)
        done = (flags & 1).bool()
        mine = term[:, self.agent]
        reward = torch.where(done, torch.where(mine == 3, 1.0, torch.where(mine == 1, 0.0, -1.0)), 0.0).float()
        self._ep_len += 1
        self._ep_ret += reward
        fin_ret, fin_len = self._ep_ret.clone(), self._ep_len.clone()
        final_obs = self._obs()
        # gymnasium-0.29 autoreset: finished envs restart at once and return the NEW episode's first observation
        fresh = eng.new_states(self.num_envs)
        self.states = torch.where(done[:, None], fresh, self.states).contiguous()
        self._ep_len = torch.where(done, torch.zeros_like(self._ep_len), self._ep_len)
        self._ep_ret = torch.where(done, torch.zeros_like(self._ep_ret), self._ep_ret)
        if self.agent != 0:
            self._play_opponents()
        self._refresh_mask()
        return self._obs(), reward, done, (fin_ret, fin_len, final_obs)

    def step(self, actions):
        acts = torch.as_tensor(np.asarray(actions), dtype=torch.int32, device=self.eng.device)
        obs, reward, done, (fin_ret, fin_len, final_obs) = self.step_device(acts)
        done_h = done.cpu().numpy()
        info = {}
        if done_h.any():
            r, l = fin_ret.cpu().numpy(), fin_len.cpu().numpy()
            fo = final_obs.cpu().numpy().astype(np.float32)
            info["final_info"] = np.array([{"episode": {"r": float(r[i]), "l": int(l[i])}} if done_h[i] else None
                                           for i in range(self.num_envs)], dtype=object)
            info["final_observation"] = np.array([fo[i] if done_h[i] else None for i in range(self.num_envs)],
                                                 dtype=object)
            info["_final_info"] = done_h.copy()
        return (obs.cpu().numpy().astype(np.float32), reward.cpu().numpy(), done_h,
                np.zeros(self.num_envs, dtype=bool), info)

    def get_attr(self, name: str):
        if name == "ai_possible_indexes":                # ppo/trainer.py:385
            nz = torch.nonzero(self.action_mask).cpu().numpy()
            splits = np.searchsorted(nz[:, 0], np.arange(self.num_envs + 1))
            return [nz[splits[i]: splits[i + 1], 1].tolist() for i in range(self.num_envs)]
        return [getattr(self, name)] * self.num_envs

    def close(self):
        self.states = None

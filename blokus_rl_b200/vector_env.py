"""Boundary B2: a gymnasium-style vector env over the batched engine (duck-typed; gymnasium is not imported).

What the reference's PPO loop calls (``blokus_rl/ppo/trainer.py:36-38, 68, 111, 146-173, 380-386``; spaces at
``ppo/agent.py:112-114, 205-206`` and ``ppo/memory.py:18-25``):

    obs, info = envs.reset()
    envs.get_attr("ai_possible_indexes")           # per env: the agent's legal action ids
    obs, reward, terminated, truncated, info = envs.step(actions)        # np int actions
    info["final_info"][i]["episode"]["r" | "l"]    # gymnasium-0.29 autoreset episode statistics
    envs.single_observation_space.shape, envs.single_action_space.n / .shape, envs.close()

The reference env (``blokus_gym:blokus-simple-v0``, absent) is single-agent: the agent is player 0 and random
bots play the other colours inside ``step`` (docs/README.md:47-51).  Reward: 0 until the game
ends, then +1 win / 0 draw / -1 loss for the agent.  Observation: the ``[N, N]`` board contents (0 empty,
1..P colour), which is what ``CnnAgent`` unsqueezes to one channel (``ppo/agent.py:96-99, 112-113``).

All envs live on the GPU; opponents are stepped by the engine's on-device Philox sampler, so one ``step`` is
a fixed handful of kernel launches regardless of ``num_envs``.  The NumPy surface moves only what the caller
reads: per step ONE pinned staging buffer comes back (board cells, rewards, done flags, and the agent's legal ids
in the sparse ``BLK_MASK_INDICES`` form -- a few hundred bytes per env instead of the 30 KB dense mask).
GPU-native callers skip the NumPy surface entirely: ``step_device`` takes/returns CUDA tensors and ``action_mask``
is the bool ``[num_envs, A]`` tensor that ``FilterLegalMoves`` (``ppo/agent.py:33-42``) would otherwise rebuild
from index lists (it is produced on first use after a step).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import BLK_FLAG_ILLEGAL, BLK_FLAG_TRUNCATED, BLK_MASK_INDICES, BLK_MASK_NONE


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
        dev = engine.device
        self.states = None
        self._next = torch.empty((num_envs, engine.state_words), dtype=torch.int32, device=dev)   # step target (commit on success)
        self._fresh = engine.new_states(num_envs)
        self._mask = None                  # dense bool mask of the current states, built on demand
        self._ep_len = torch.zeros(num_envs, dtype=torch.int64, device=dev)
        self._ep_ret = torch.zeros(num_envs, dtype=torch.float32, device=dev)
        self._epoch = 0
        # sparse legal ids of the agent: row width grows when a position has more legal moves than fit
        self._idx_stride = 128 if self.N <= 7 else 512
        # `step` on the real engine goes to the C ABI directly (argument blocks filled once, one staging copy back): at the
        # reference's default of 4 envs a step is launch-latency-bound, and ~40 torch calls cost 4x what the kernels do
        from .engine import BlokusEngine
        self._direct = isinstance(engine, BlokusEngine)
        self._alloc_idx()
        self._ids_host = None              # (counts, ids) of the current states on the host, filled on demand
        self._h_reward = self._host(num_envs, torch.float32)
        self._h_done = self._host(num_envs, torch.bool)
        self._h_fret, self._h_flen = self._host(num_envs, torch.float32), self._host(num_envs, torch.int64)
        self._np_ep_len = np.zeros(num_envs, np.int64)          # episode statistics of the NumPy surface (`step`)
        self._np_ep_ret = np.zeros(num_envs, np.float32)
        if self._direct:
            self._h_acts = self._host(num_envs, torch.int32)
            self._d_acts = torch.empty(num_envs, dtype=torch.int32, device=dev)
            self._cudart = _cudart()
            self._args_a = _lib.BlkStepArgs(num_envs, None, None, self._d_acts.data_ptr(), None, BLK_MASK_NONE, 0, None, None, None,
                                            None, None, 0, 0, 0, None, None, None, None)
            self._args_i = _lib.BlkStepArgs(num_envs, None, None, None, None, BLK_MASK_INDICES, 0, None, None, None,
                                            None, None, 0, 0, 0, None, None, None, None)
            self._rargs = _lib.BlkRolloutArgs(num_envs, None, 1, 0, 0, None, None, None, None, 88, None, int(self.agent), None, 0)

    def _alloc_idx(self):
        """One device buffer and one (pinned) host buffer hold everything a step returns, so it comes back in ONE copy:
        flags of the agent's move | winners of finished games | flags of the id launch | legal counts | final board cells |
        board cells | legal ids (``_idx_stride`` per env; grows when a position has more legal moves)."""
        E, dev, NN = self.num_envs, self.eng.device, self.N * self.N
        a16 = lambda x: (x + 15) & ~15
        o_fa, o_win, o_if = 0, a16(E), 2 * a16(E)
        o_cnt = 3 * a16(E)
        o_fobs = a16(o_cnt + 4 * E)
        o_obs = a16(o_fobs + E * NN)
        o_idx = a16(o_obs + E * NN)
        self._pack_bytes = o_idx + 2 * E * self._idx_stride
        self._dpack = torch.zeros(self._pack_bytes, dtype=torch.uint8, device=dev)
        self._hpack = self._host(self._pack_bytes, torch.uint8)

        def carve(pack):
            return (pack[o_fa: o_fa + E], pack[o_win: o_win + E], pack[o_if: o_if + E],
                    pack[o_cnt: o_cnt + 4 * E].view(torch.int32), pack[o_fobs: o_fobs + E * NN].view(E, self.N, self.N),
                    pack[o_obs: o_obs + E * NN].view(E, self.N, self.N),
                    pack[o_idx:].view(torch.int16).view(E, self._idx_stride))
        self._d_flags_a, self._d_win, self._idx_flags, self._cnt, self._d_fobs, self._d_obs, self._idx = carve(self._dpack)
        self._h_flags_a, self._h_win, self._h_idx_flags, self._h_cnt, self._h_fobs, self._h_board, self._h_idx = carve(self._hpack)

    def _host(self, shape, dtype):
        # pinned staging memory (the CPU stand-in engine of the host-logic tests has none)
        return torch.empty(shape, dtype=dtype, pin_memory=self.eng.device.type == "cuda")

    # ---- helpers ----------------------------------------------------------------------------------------
    def _sync(self):
        if self.eng.device.type == "cuda":
            torch.cuda.current_stream(self.eng.device).synchronize()

    def _play_opponents(self):
        """Random bots move (one rollout launch, no host sync) until it is the agent's turn or the game is over."""
        self._epoch += 1
        self.eng.rollout(self.states, 1, seed=self.seed + self._epoch, stop_player=self.agent, out_states=self.states)

    @property
    def action_mask(self) -> torch.Tensor:
        """bool [num_envs, A] on the device: the agent's legal actions in the current states."""
        if self._mask is None:
            self._mask = self.eng.step(self.states, None, mask="bytes", want_count=False, want_terminal=False,
                                       want_scores=False).mask
        return self._mask

    def _launch_ids(self):
        """Legal ids of the current states (sparse form) -> pinned host buffers; the caller synchronises."""
        self.eng.step(self.states, None, mask=self._idx, want_terminal=False, want_scores=False,
                      buffers=_Bufs(self._cnt, self._idx_flags))
        self._h_idx.copy_(self._idx, non_blocking=True)
        self._h_cnt.copy_(self._cnt, non_blocking=True)
        self._h_idx_flags.copy_(self._idx_flags, non_blocking=True)

    def _collect_ids(self):
        """After a synchronise: (counts, ids) on the host; widens the id rows and retries when a row was truncated."""
        while (self._h_idx_flags.numpy() & BLK_FLAG_TRUNCATED).any():
            self._idx_stride *= 2
            self._alloc_idx()
            self._launch_ids()
            self._sync()
        self._ids_host = (self._h_cnt.numpy(), self._h_idx.numpy().view(np.uint16))

    def _obs(self):
        return self.eng.board_contents(self.states)

    # ---- gym surface ----------------------------------------------------------------------------------------
    def reset(self, seed=None, options=None):
        if seed is not None:
            self.seed = seed
        self.states = self.eng.new_states(self.num_envs)
        self._ep_len.zero_()
        self._ep_ret.zero_()
        self._np_ep_len[:] = 0
        self._np_ep_ret[:] = 0
        if self.agent != 0:
            self._play_opponents()
        self._mask = self._ids_host = None
        return self._obs().cpu().numpy().astype(np.float32), {}

    def step_device(self, actions: torch.Tensor, check: bool = True):
        """Device-side step: int32 CUDA actions in; (obs uint8 [E,N,N], reward f32 [E], terminated bool [E],
        (episode returns, episode lengths, final observations) for the envs that finished) out, all on the
        device.  With ``check=True`` the env only advances when every action is legal (an illegal action raises
        and leaves ALL envs as they were); with ``check=False`` nothing synchronises with the host and an illegal
        action leaves just its own env unchanged, as ``blk_step`` defines."""
        eng = self.eng
        out = eng.step(self.states, actions.to(torch.int32).contiguous(), out_states=self._next, mask=None,
                       want_count=False, want_terminal=False, want_scores=False)
        if check and bool((out.flags & BLK_FLAG_ILLEGAL).any()):
            raise ValueError("illegal action passed to BlokusVectorEnv.step")
        self.states, self._next = self._next, self.states              # commit
        self._mask = self._ids_host = None
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
        self.states = torch.where(done[:, None], self._fresh, self.states)
        self._ep_len = torch.where(done, torch.zeros_like(self._ep_len), self._ep_len)
        self._ep_ret = torch.where(done, torch.zeros_like(self._ep_ret), self._ep_ret)
        if self.agent != 0:
            self._play_opponents()
        return self._obs(), reward, done, (fin_ret, fin_len, final_obs)

    def _step_direct(self, actions):
        """`step` on the CUDA engine: the same launches as `step_device` + `_launch_ids`, issued through the C ABI with
        prefilled argument blocks into the scratch state buffer, one copy back, one synchronize; committed only when every
        action was legal.  Rewards and episode statistics are kept on the host (they are E numbers)."""
        eng, E, lib = self.eng, self.num_envs, self.eng._lib
        self._h_acts.numpy()[:] = np.asarray(actions)
        S, T = self.states, self._next
        tp = T.data_ptr()
        with torch.cuda.device(eng.device):
            st = eng._stream()
            self._d_acts.copy_(self._h_acts, non_blocking=True)
            a, i, r = self._args_a, self._args_i, self._rargs
            a.state_in, a.state_out, a.flags = S.data_ptr(), tp, self._d_flags_a.data_ptr()
            _lib.check(lib.blk_step(eng._h, C.byref(a), st))
            # random bots move until it is the agent's turn or the game is over; `winners` != 0 marks the finished games
            r.roots = r.state_out = tp
            r.winners = self._d_win.data_ptr()
            r.seed = (self.seed + self._epoch + 1) & 0xFFFFFFFFFFFFFFFF
            _lib.check(lib.blk_rollout(eng._h, C.byref(r), st))
            _lib.check(lib.blk_board_contents(eng._h, tp, self._d_fobs.data_ptr(), E, st))
            # gymnasium-0.29 autoreset: finished envs restart at once
            torch.where(self._d_win.bool()[:, None], self._fresh, T, out=T)
            epochs = 1
            if self.agent != 0:
                r.winners = None
                r.seed = (self.seed + self._epoch + 2) & 0xFFFFFFFFFFFFFFFF
                _lib.check(lib.blk_rollout(eng._h, C.byref(r), st))
                epochs = 2
            i.state_in = i.state_out = tp
            i.mask, i.mask_stride = self._idx.data_ptr(), self._idx_stride
            i.legal_count, i.flags = self._cnt.data_ptr(), self._idx_flags.data_ptr()
            _lib.check(lib.blk_step(eng._h, C.byref(i), st))
            _lib.check(lib.blk_board_contents(eng._h, tp, self._d_obs.data_ptr(), E, st))
            rt = self._cudart
            if rt is not None:
                if rt.cudaMemcpyAsync(self._hpack.data_ptr(), self._dpack.data_ptr(), self._pack_bytes, 2, st) or \
                        rt.cudaStreamSynchronize(st):
                    raise _lib.EngineError("device-to-host copy of a vector-env step failed")
            else:
                self._hpack.copy_(self._dpack, non_blocking=True)
                torch.cuda.current_stream().synchronize()
        if (self._h_flags_a.numpy() & BLK_FLAG_ILLEGAL).any():
            self._ids_host = None                          # the staging buffer now holds the abandoned step's ids
            raise ValueError("illegal action passed to BlokusVectorEnv.step")          # nothing was committed
        self.states, self._next = T, S
        self._epoch += epochs
        self._mask = None
        win = self._h_win.numpy()
        done = win != 0
        mine = (win >> self.agent) & 1
        reward = np.where(done, np.where(mine != 0, np.where((win & (win - 1)) == 0, 1.0, 0.0), -1.0), 0.0).astype(np.float32)
        self._np_ep_len += 1
        self._np_ep_ret += reward
        obs = self._h_board.numpy().astype(np.float32)
        info = {}
        idx = np.flatnonzero(done)
        if len(idx):
            # (7x7 games last ~5 agent moves: at 4,096 envs ~800 finish per step, so this is built without per-env NumPy calls)
            fo = self._h_fobs.numpy()[idx].astype(np.float32)
            final_info = np.full(E, None, dtype=object)
            final_obs_h = np.full(E, None, dtype=object)
            where = idx.tolist()
            for j, r, l, o in zip(where, self._np_ep_ret[idx].tolist(), self._np_ep_len[idx].tolist(), fo):
                final_info[j] = {"episode": {"r": r, "l": l}}
                final_obs_h[j] = o
            info = {"final_info": final_info, "final_observation": final_obs_h, "_final_info": done.copy()}
            self._np_ep_len[idx] = 0
            self._np_ep_ret[idx] = 0
        if (self._h_idx_flags.numpy() & BLK_FLAG_TRUNCATED).any():
            self._collect_ids()                                    # widens the id rows (new staging buffers) and asks again
        else:
            self._ids_host = (self._h_cnt.numpy(), self._h_idx.numpy().view(np.uint16))
        return obs, reward, done, np.zeros(E, dtype=bool), info

    def step(self, actions):
        if self._direct:
            return self._step_direct(actions)
        acts = torch.as_tensor(np.asarray(actions), dtype=torch.int32).to(self.eng.device, non_blocking=True)
        obs, reward, done, (fin_ret, fin_len, final_obs) = self.step_device(acts)
        # one round trip: board cells, rewards, done flags, episode statistics and the agent's next legal ids travel together
        for dst, src in ((self._h_board, obs), (self._h_reward, reward), (self._h_done, done), (self._h_fret, fin_ret),
                         (self._h_flen, fin_len), (self._h_fobs, final_obs)):
            dst.copy_(src, non_blocking=True)
        self._launch_ids()
        self._sync()
        board_h = self._h_board.numpy().astype(np.float32)
        fo = self._h_fobs.numpy().copy()
        self._collect_ids()                                # (may replace the staging buffers: board cells were read first)
        done_h = self._h_done.numpy().copy()
        info = {}
        idx = np.flatnonzero(done_h)
        if len(idx):                                       # gymnasium-0.29 autoreset bookkeeping, only for the envs that finished
            r, l = self._h_fret.numpy(), self._h_flen.numpy()
            final_info = np.full(self.num_envs, None, dtype=object)
            final_obs_h = np.full(self.num_envs, None, dtype=object)
            for i in idx:
                final_info[i] = {"episode": {"r": float(r[i]), "l": int(l[i])}}
                final_obs_h[i] = fo[i].astype(np.float32)
            info = {"final_info": final_info, "final_observation": final_obs_h, "_final_info": done_h.copy()}
        return (board_h, self._h_reward.numpy().copy(), done_h, np.zeros(self.num_envs, dtype=bool), info)

    def legal_ids_padded(self):
        """The agent's legal action ids as they come off the device: ``(ids uint16 [num_envs, width], counts int32
        [num_envs])``, row i valid up to ``counts[i]``, ascending.  Views of the pinned staging buffers (valid until the next
        ``step`` / ``reset``): the form a vectorised consumer wants."""
        if self._ids_host is None:
            self._launch_ids()
            self._sync()
            self._collect_ids()
        cnt, ids = self._ids_host
        return ids, cnt

    def legal_id_arrays(self):
        """The agent's legal action ids per env as int64 arrays (views of one flat array): what
        ``mask[i, possible_move] = 1`` indexes with (ppo/agent.py:36-37), without building Python lists."""
        ids, cnt = self.legal_ids_padded()
        keep = np.arange(ids.shape[1], dtype=np.int32)[None, :] < cnt[:, None]
        return ids[keep].astype(np.int64), cnt

    def get_attr(self, name: str):
        if name == "ai_possible_indexes":                # ppo/trainer.py:385 -> list[list[int]]
            flat, cnt = self.legal_id_arrays()
            flat, ends = flat.tolist(), np.cumsum(cnt).tolist()
            return [flat[a:b] for a, b in zip([0] + ends[:-1], ends)]
        return [getattr(self, name)] * self.num_envs

    def close(self):
        self.states = None


def _cudart():
    from .backend import _cudart as find
    return find()


class _Bufs:
    """The two per-step outputs the sparse-mask launch needs, in the shape ``BlokusEngine.step(buffers=...)`` reads."""
    mask_raw = terminal = scores = next_action = obs = None

    def __init__(self, legal_count, flags):
        self.legal_count, self.flags = legal_count, flags

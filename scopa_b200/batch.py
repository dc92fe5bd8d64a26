"""Batched Miniscopa on the GPU: the performance API underneath the drop-in classes.

One row = one game.  States are torch.int32 tensors of shape [n, 4] on a CUDA device (the 16-byte
packed layout of csrc/ms_state.cuh); every method enqueues one kernel of libscopa_b200.so on torch's
current stream through the C ABI (include/scopa_b200.h).

Replaces, for n games at once: MiniScopaEnv.reset/step (reference src/envs/mini_scopa_game.py:131-167),
MiniScopaState.legal_actions / information_state_string (src/envs/openspiel_mini_scopa.py:22-47, :86-95).
"""
import torch

from . import _lib


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _check_dev(t, dtype, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.is_contiguous() and t.dtype == dtype):
        raise ValueError(f"{name}: expected a contiguous CUDA tensor of dtype {dtype}")


class BatchedMiniScopa:
    """n concurrent games resident in HBM."""

    def __init__(self, device="cuda"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.MsError("scopa_b200 runs on CUDA devices only (no CPU fallback)")
        self.lib = _lib.load()
        self.states = None       # [n, 4] uint32
        self.hand_order = None   # [n] uint32

    @property
    def n(self):
        return 0 if self.states is None else self.states.shape[0]

    # ----------------------------------------------------------------------------------- reset
    def reset(self, seeds):
        """MiniScopaEnv.reset(seed) per row; seeds: int64 tensor/sequence (seed 0 means 42)."""
        seeds = torch.as_tensor(seeds, dtype=torch.int64).to(self.device, non_blocking=True).contiguous()
        n = seeds.numel()
        with torch.cuda.device(self.device):
            self.states = torch.empty((n, 4), dtype=torch.int32, device=self.device)
            self.hand_order = torch.empty((n,), dtype=torch.int32, device=self.device)
            _lib.check(self.lib.ms_deal_from_seeds(seeds.data_ptr(), n, self.states.data_ptr(),
                                                   self.hand_order.data_ptr(), _lib.stream_ptr()))
        return self

    def set_states(self, states, hand_order):
        _check_dev(states, torch.int32, "states")
        _check_dev(hand_order, torch.int32, "hand_order")
        self.states, self.hand_order = states, hand_order
        return self

    # ----------------------------------------------------------------------------------- step
    def step(self, actions, rewards=None, done=None):
        """MiniScopaEnv.step per row, in place.  actions: uint8 [n].  Returns (rewards [n,2] f32, done [n] u8)."""
        _check_dev(actions, torch.uint8, "actions")
        n = self.n
        with torch.cuda.device(self.device):
            if rewards is None:
                rewards = torch.empty((n, 2), dtype=torch.float32, device=self.device)
            if done is None:
                done = torch.empty((n,), dtype=torch.uint8, device=self.device)
            _lib.check(self.lib.ms_step(self.states.data_ptr(), actions.data_ptr(), rewards.data_ptr(), done.data_ptr(),
                                        n, _lib.stream_ptr()))
        return rewards, done

    def legal_actions(self, player=-1, want_capture=False):
        """-> mask [n] int16 (bit pattern of a uint16, bit = action id), ordered [n,4] uint8 (hand order, 0xFF pad), count [n] uint8
        (and, if want_capture, capture [n,4] uint8: table-position mask each legal action would take)."""
        n = self.n
        with torch.cuda.device(self.device):
            mask = torch.empty((n,), dtype=torch.int16, device=self.device)
            ordered = torch.empty((n, 4), dtype=torch.uint8, device=self.device)
            count = torch.empty((n,), dtype=torch.uint8, device=self.device)
            cap = torch.empty((n, 4), dtype=torch.uint8, device=self.device) if want_capture else None
            _lib.check(self.lib.ms_legal_actions(self.states.data_ptr(), self.hand_order.data_ptr(), player,
                                                 mask.data_ptr(), ordered.data_ptr(), count.data_ptr(), _ptr(cap), n,
                                                 _lib.stream_ptr()))
        return (mask, ordered, count, cap) if want_capture else (mask, ordered, count)

    def capture(self, cards):
        """MiniScopaGame.card_in_table per row -> table-position masks [n] uint8."""
        _check_dev(cards, torch.uint8, "cards")
        with torch.cuda.device(self.device):
            out = torch.empty((self.n,), dtype=torch.uint8, device=self.device)
            _lib.check(self.lib.ms_capture(self.states.data_ptr(), cards.data_ptr(), out.data_ptr(), self.n,
                                           _lib.stream_ptr()))
        return out

    def infoset_keys(self, player=-1):
        with torch.cuda.device(self.device):
            keys = torch.empty((self.n,), dtype=torch.int64, device=self.device)
            _lib.check(self.lib.ms_infoset_keys(self.states.data_ptr(), player, keys.data_ptr(), self.n,
                                                _lib.stream_ptr()))
        return keys

    # ----------------------------------------------------------------------------------- rollouts
    def rollout_random(self, philox_seed=0, game_offset=0, want_final=False, actions=None, rewards=None):
        """Play every game to the end with a uniform-random legal policy (one fused kernel).
        -> actions [n,8] uint8, rewards [n,2] f32 (, final states [n,4] uint32)."""
        n = self.n
        with torch.cuda.device(self.device):
            if actions is None:
                actions = torch.empty((n, 8), dtype=torch.uint8, device=self.device)
            if rewards is None:
                rewards = torch.empty((n, 2), dtype=torch.float32, device=self.device)
            final = torch.empty((n, 4), dtype=torch.int32, device=self.device) if want_final else None
            _lib.check(self.lib.ms_rollout_random(self.states.data_ptr(), self.hand_order.data_ptr(), n, philox_seed,
                                                  game_offset, actions.data_ptr(), rewards.data_ptr(), _ptr(final),
                                                  _lib.stream_ptr()))
        return (actions, rewards, final) if want_final else (actions, rewards)


def rollout_random_host(seeds, philox_seed=0, game_offset=0, actions=None, rewards=None):
    """End-to-end call with HOST buffers: seeds (numpy int64 / pinned torch tensor) in, actions [n,8] uint8
    and rewards [n,2] float32 out.  Reset (deal) + 8 steps per game run on the device; H2D/D2H inside."""
    import numpy as np
    lib = _lib.load()
    if isinstance(seeds, torch.Tensor):
        assert seeds.dtype == torch.int64 and not seeds.is_cuda and seeds.is_contiguous()
        n, sp = seeds.numel(), seeds.data_ptr()
    else:
        seeds = np.ascontiguousarray(seeds, dtype=np.int64)
        n, sp = seeds.size, seeds.ctypes.data
    if actions is None:
        actions = torch.empty((n, 8), dtype=torch.uint8).pin_memory()
    if rewards is None:
        rewards = torch.empty((n, 2), dtype=torch.float32).pin_memory()
    _lib.check(lib.ms_rollout_random_host(sp, n, philox_seed, game_offset, actions.data_ptr(), rewards.data_ptr()))
    return actions, rewards

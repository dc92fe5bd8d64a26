"""Batched 2v2 team Miniscopa on the GPU (csrc/ms_team.cu) and the codec of its 32-byte packed state.

One row = one game; states are torch.int32 [n, 8] (bit patterns of the uint32 words), hand_order torch.int64 [n]
(the shuffled deck: nibble 4p+i = i-th card dealt to player p).  Replaces, n games at a time,
TeamMiniScopaEnv.reset/step of /root/reference/src/envs/team_mini_scopa_game.py:151-205.
"""
import torch

from . import _lib, codec


def unpack_team_state(words):
    w = [int(x) & 0xFFFFFFFF for x in words]
    meta = w[3]
    lct = ((meta >> 12) & 3) - 1
    return {
        "hand_mask": [w[0] & 0xFFFF, w[0] >> 16, w[1] & 0xFFFF, w[1] >> 16],
        "table": codec.nibbles(w[2], meta & 0xF),
        "cap_mask": [w[4] & 0xFFFF, w[4] >> 16, w[5] & 0xFFFF, w[5] >> 16],
        "scopas": [(w[6] >> (4 * p)) & 0xF for p in range(4)],
        "step_count": (meta >> 4) & 0x1F, "cur": (meta >> 9) & 3, "terminal": bool((meta >> 11) & 1),
        "last_capture_team": None if lct < 0 else lct, "max_steps": (meta >> 14) & 0x1F,
    }


def pack_team_state(hand_mask, table, cap_mask, scopas, step_count, cur, terminal, last_capture_team, max_steps=16):
    meta = (len(table) & 0xF) | ((step_count & 0x1F) << 4) | ((cur & 3) << 9) | ((1 if terminal else 0) << 11) \
        | (((-1 if last_capture_team is None else last_capture_team) + 1) << 12) | ((max_steps & 0x1F) << 14)
    return (hand_mask[0] | (hand_mask[1] << 16), hand_mask[2] | (hand_mask[3] << 16), codec.pack_nibbles(table), meta,
            cap_mask[0] | (cap_mask[1] << 16), cap_mask[2] | (cap_mask[3] << 16),
            sum((scopas[p] & 0xF) << (4 * p) for p in range(4)), 0)


def team_hand_in_order(hand_mask, hand_order, player):
    out = []
    for c in codec.nibbles(int(hand_order) >> (16 * player), 4):
        if (hand_mask >> c) & 1 and c not in out:
            out.append(c)
    return out


class BatchedTeamMiniScopa:
    def __init__(self, device="cuda"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.MsError("scopa_b200 runs on CUDA devices only (no CPU fallback)")
        self.lib = _lib.load()
        self.states = None
        self.hand_order = None

    @property
    def n(self):
        return 0 if self.states is None else self.states.shape[0]

    def reset(self, seeds):
        seeds = torch.as_tensor(seeds, dtype=torch.int64).to(self.device).contiguous()
        n = seeds.numel()
        with torch.cuda.device(self.device):
            self.states = torch.empty((n, 8), dtype=torch.int32, device=self.device)
            self.hand_order = torch.empty((n,), dtype=torch.int64, device=self.device)
            _lib.check(self.lib.ms_team_deal_from_seeds(seeds.data_ptr(), n, self.states.data_ptr(),
                                                        self.hand_order.data_ptr(), _lib.stream_ptr()))
        return self

    def step(self, actions):
        """-> rewards [n, 4] f32 (team rewards t0, t0, t1, t1; zero while running), done [n] u8."""
        n = self.n
        with torch.cuda.device(self.device):
            rewards = torch.empty((n, 4), dtype=torch.float32, device=self.device)
            done = torch.empty((n,), dtype=torch.uint8, device=self.device)
            _lib.check(self.lib.ms_team_step(self.states.data_ptr(), actions.data_ptr(), rewards.data_ptr(), done.data_ptr(),
                                             n, _lib.stream_ptr()))
        return rewards, done

    def rollout_random(self, philox_seed=0, game_offset=0):
        """16 uniform-random legal plies per game -> actions [n, 16] u8, rewards [n, 4] f32, final states [n, 8]."""
        n = self.n
        with torch.cuda.device(self.device):
            actions = torch.empty((n, 16), dtype=torch.uint8, device=self.device)
            rewards = torch.empty((n, 4), dtype=torch.float32, device=self.device)
            final = torch.empty((n, 8), dtype=torch.int32, device=self.device)
            _lib.check(self.lib.ms_team_rollout_random(self.states.data_ptr(), self.hand_order.data_ptr(), n, philox_seed,
                                                       game_offset, actions.data_ptr(), rewards.data_ptr(), final.data_ptr(),
                                                       _lib.stream_ptr()))
        return actions, rewards, final

"""open_spiel.python.policy.Policy stand-in (the reference only uses its constructor)."""
try:  # pragma: no cover - open_spiel is not part of this image
    from open_spiel.python.policy import Policy
except ImportError:
    class Policy:
        def __init__(self, game, player_ids):
            self.game = game
            self.player_ids = player_ids


def root_of(game):
    """(packed root words, hand_order) of game.new_initial_state() for the device solver."""
    state = game.new_initial_state()
    return state.env.packed()


_MAX_STEPS_FIELD = 0x1F << 19      # codec.pack_state: bits 19-23 of the meta word hold the env's max_steps


def root_id(words, order):
    """Identity of a packed state for "is this the solver's root?" checks.  The step limit is not part of it: a
    `clone()` of a fresh root carries max_steps = 16 where the original has 8 (the reference's clone quirk,
    openspiel_mini_scopa.py:108), and both are the same root to _cfr_recursive / _external_sampling_cfr."""
    w = [int(x) & 0xFFFFFFFF for x in words]
    w[3] &= ~_MAX_STEPS_FIELD
    return tuple(w), int(order)

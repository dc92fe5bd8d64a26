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

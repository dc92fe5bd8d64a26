class Policy:
    def __init__(self, game, player_ids):
        self.game = game
        self.player_ids = player_ids

class AECEnv:
    def __init__(self):
        pass

    def _was_dead_step(self, action):
        return None

"""evaluate_agent for both tabular trainers (the reference defines the same function twice:
/root/reference/src/algorithms/vanilla_cfr.py:157-216 and src/algorithms/mc_cfr.py:146-206).

Protocol: `num_episodes` games from the root; the trained policy sits in seat 0 while
`episode < num_episodes / 2` and in seat 1 afterwards; returns (average reward of the trained policy, its
running average per episode, scopa statistics with running averages).  Two execution paths with the same result
shape: all episodes on the GPU in two launches when both policies can export a per-infoset table
(`_device_table`), otherwise an episode-by-episode loop over the state API.
"""
import numpy as np


def _pack_result(rewards, trained_scopas, opponent_scopas):
    """Per-episode arrays (already from the trained policy's point of view) -> the reference's return triple."""
    n = len(rewards)
    k = np.arange(1, n + 1)
    cum_t, cum_o = np.cumsum(trained_scopas), np.cumsum(opponent_scopas)
    history = {"trained": (cum_t / k).tolist(), "opponent": (cum_o / k).tolist(), "diff": ((cum_t - cum_o) / k).tolist()}
    avg_t = float(cum_t[-1]) / n if n else 0.0
    avg_o = float(cum_o[-1]) / n if n else 0.0
    stats = {"trained_avg": avg_t, "opponent_avg": avg_o, "difference": avg_t - avg_o, "history": history,
             "data_collected": n > 0}
    return (float(np.sum(rewards)) / n if n else 0.0), (np.cumsum(rewards) / k).tolist(), stats


def _first_half(num_episodes):
    return int(np.ceil(num_episodes / 2))           # number of episodes with episode < num_episodes / 2


def _on_device(solver, trained, opponent, num_episodes, seed):
    t_tab, o_tab = trained._device_table(solver), opponent._device_table(solver)
    n0 = _first_half(num_episodes)
    r_a, s_a = solver.evaluate(t_tab, o_tab, n0, philox_seed=seed, first_game=0)
    r_b, s_b = solver.evaluate(o_tab, t_tab, num_episodes - n0, philox_seed=seed, first_game=n0)
    s_a, s_b = s_a.cpu().numpy().astype(np.int64), s_b.cpu().numpy().astype(np.int64)
    rewards = np.concatenate([r_a.cpu().numpy().astype(np.float64), -r_b.cpu().numpy().astype(np.float64)])
    return _pack_result(rewards, np.concatenate([s_a[:, 0], s_b[:, 1]]), np.concatenate([s_a[:, 1], s_b[:, 0]]))


def _play_episode(game, seat_policies):
    state = game.new_initial_state()
    while not state.is_terminal():
        dist = seat_policies[state.current_player()].action_probabilities(state)
        acts = list(dist)
        state.apply_action(np.random.choice(acts, p=[dist[a] for a in acts]))
    return state


def _on_host(game, trained, opponent, num_episodes):
    n0 = _first_half(num_episodes)
    rewards, ts, os_ = np.zeros(num_episodes), np.zeros(num_episodes), np.zeros(num_episodes)
    for ep in range(num_episodes):
        seat = 0 if ep < n0 else 1
        final = _play_episode(game, [trained, opponent] if seat == 0 else [opponent, trained])
        rewards[ep] = final.rewards()[seat]
        players = final.env.game.players
        ts[ep], os_[ep] = players[seat].scopas, players[1 - seat].scopas
    return _pack_result(rewards, ts, os_)


def evaluate_agent(game, trained_policy, opponent_policy, num_episodes=10000, seed=None):
    """Drop-in for the reference's evaluate_agent.  `seed` selects the Philox stream of the GPU path (default:
    drawn from np.random, so np.random.seed() still makes a run repeatable)."""
    if game.num_players() != 2:
        raise ValueError("evaluate_agent only supports 2-player games")
    solver = getattr(trained_policy, "_solver", None) or getattr(opponent_policy, "_solver", None)
    exportable = hasattr(trained_policy, "_device_table") and hasattr(opponent_policy, "_device_table")
    if solver is not None and exportable and num_episodes > 0:
        if seed is None:
            seed = int(np.random.randint(0, 2 ** 31 - 1))
        return _on_device(solver, trained_policy, opponent_policy, num_episodes, seed)
    return _on_host(game, trained_policy, opponent_policy, num_episodes)

"""Drop-in for the reference's src/algorithms/deep_cfr/deep_cfr.py ("SDCFR").

Same classes and surface (/root/reference/src/algorithms/deep_cfr/deep_cfr.py:24-504):
AdvantageNetwork(.net, .buffer, get_advantages, add_experience, train), StrategyBuffer, RandomPolicy,
DeepCFR(.advantage_nets, .strategy_buffers, .training_history, _state_to_features,
_get_legal_actions_mask, _external_sampling_cfr, evaluate_vs_random, train, get_policy).

What changes underneath: the traversal no longer calls a batch-1 MLP (and crosses PCIe) per tree node.
`_external_sampling_cfr` runs `traversals_per_iteration` traversals level by level on the GPU with the
frontier of each level as one batched inference (csrc/ms_sdcfr.cu; `precision="fp32"` = CUDA-core path in
the reference's precision, `"bf16"` = tcgen05 tensor-core path), and the samples land in a device-resident
replay buffer.  The optimiser (Adam, masked MSE, clip-norm 1.0, as the reference) is PyTorch's by default;
`optimizer="fused"` runs all epochs of a train() call in one launch of sd_train_kernel (csrc/ms_sd_train.cuh):
same arithmetic in fp32, minibatches drawn without replacement from torch's CUDA generator; `"fused-cluster"` is the
8-CTA cluster form of that kernel (csrc/ms_sd_train_cluster.cuh).
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
import torch.optim as optim

from ... import codec, sdcfr
from .._policy_base import root_id, root_of
from .nets import FlexibleNet, positive_regret_policy

HIDDEN = [128, 64]


class DeviceReplayBuffer:
    """deque(maxlen=100000) of (features, advantages, mask) kept as three CUDA ring tensors."""

    def __init__(self, maxlen=100000, device="cuda"):
        self.maxlen, self.device = maxlen, device
        self.feat = torch.zeros((maxlen, 34), dtype=torch.float32, device=device)
        self.target = torch.zeros((maxlen, 16), dtype=torch.float32, device=device)
        self.mask = torch.zeros((maxlen, 16), dtype=torch.float32, device=device)
        self.count = 0          # total ever appended

    def __len__(self):
        return min(self.count, self.maxlen)

    def clear(self):
        self.count = 0

    def add_batch(self, feat, target, mask):
        n = feat.shape[0]
        if n > self.maxlen:
            feat, target, mask, n = feat[-self.maxlen:], target[-self.maxlen:], mask[-self.maxlen:], self.maxlen
        start = self.count % self.maxlen                    # ring position of the first new row
        head = min(n, self.maxlen - start)                  # rows that fit before the ring wraps
        for dst, src in ((self.feat, feat), (self.target, target), (self.mask, mask)):
            dst[start:start + head] = src[:head]            # plain slice copies: no index tensors, no scatter kernels
            if head < n:
                dst[:n - head] = src[head:]
        self.count += n

    def append(self, item):
        f, a, m = (torch.as_tensor(np.asarray(x), dtype=torch.float32, device=self.device).reshape(1, -1) for x in item)
        self.add_batch(f, a, m)

    def sample(self, batch_size):
        idx = torch.randperm(len(self), device=self.device)[:batch_size]     # random.sample: without replacement
        return self.feat[idx], self.target[idx], self.mask[idx]

    def __iter__(self):
        n = len(self)
        f, a, m = self.feat[:n].cpu().numpy(), self.target[:n].cpu().numpy(), self.mask[:n].cpu().numpy()
        return iter([(f[i], a[i], m[i]) for i in range(n)])


class AdvantageNetwork:
    """Manages the advantage network for one player."""

    def __init__(self, input_dim, num_actions, device="cuda", lr=5e-4, precision="fp32", optimizer="torch",
                 sampler_seed=None):
        if optimizer not in ("torch", "fused", "fused-cluster"):
            raise ValueError(f"optimizer must be 'torch', 'fused' or 'fused-cluster', not {optimizer!r}")
        self.device = device
        self.num_actions = num_actions
        self.precision = sdcfr.TENSOR_CORE if precision == "bf16" else sdcfr.FP32
        self.net = FlexibleNet(mode="mlp", input_shape=(input_dim,), output_dim=num_actions, mlp_hidden=HIDDEN,
                               mlp_act="relu", mlp_norm="none", mlp_dropout=0.0).to(device)
        for layer in self.net.modules():
            if isinstance(layer, nn.Linear):
                nn.init.xavier_uniform_(layer.weight)
                nn.init.constant_(layer.bias, 0.1)
        self.optimizer = optim.Adam(self.net.parameters(), lr=lr)
        self.criterion = nn.MSELoss()
        self.buffer = DeviceReplayBuffer(100000, device)
        self._fused = None
        # fused optimisers only: None = minibatch rows from torch's CUDA generator (rand + top-k), an int = rows from the
        # Philox stream of ms_sdcfr_sample_rows (one launch, repeatable, independent of torch's generator state)
        self.sampler_seed = sampler_seed
        self._sample_entry = None       # tests only (emulated ms_sdcfr_sample_rows)
        if optimizer != "torch":
            # the module's parameters become views of one flat blob that the kernel updates in place
            self._fused = sdcfr.FusedAdam(sdcfr.flatten_parameters_(self.net), lr=lr,
                                          kernel="cluster" if optimizer == "fused-cluster" else "cta")

    def blob(self):
        return self._fused.blob if self._fused is not None else sdcfr.flatten_net(self.net)

    def get_advantages(self, state_features, legal_actions_mask):
        """Get advantages for a batch of states (our batched inference kernel)."""
        f = torch.as_tensor(np.asarray(state_features), dtype=torch.float32, device=self.device)
        m = torch.as_tensor(np.asarray(legal_actions_mask), dtype=torch.float32, device=self.device)
        if f.dim() == 1:
            f, m = f.unsqueeze(0), m.unsqueeze(0)
        adv, _ = sdcfr.mlp_forward(self.blob(), f, m, self.precision)
        return adv.cpu().numpy()

    def add_experience(self, state_features, advantages, legal_actions_mask):
        advantages = np.asarray(advantages, dtype=np.float32)
        if np.max(np.abs(advantages)) > 0:
            advantages = advantages / (np.max(np.abs(advantages)) + 1e-8)
        self.buffer.append((state_features, advantages, legal_actions_mask))

    def train(self, batch_size=128, epochs=1):
        if len(self.buffer) < batch_size:
            batch_size = min(len(self.buffer), 32)
            if batch_size == 0:
                return 0.0
        if self._fused is not None:
            return self._train_fused(batch_size, epochs)
        total_loss = 0.0
        for _ in range(epochs):
            states, target_adv, masks = self.buffer.sample(batch_size)
            self.optimizer.zero_grad()
            pred_adv = self.net(states)
            loss = self.criterion(pred_adv * masks, target_adv * masks)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(self.net.parameters(), max_norm=1.0)
            self.optimizer.step()
            total_loss += loss.item()
        return total_loss / epochs

    def _train_fused(self, batch_size, epochs):
        """All epochs in one launch; one device->host read (the mean loss) per call instead of one per epoch."""
        b = self.buffer
        loss = self._fused.step(b.feat, b.target, b.mask, self._sample_rows(batch_size, epochs))
        return float(loss.mean().item())

    def _sample_rows(self, batch_size, epochs):
        """[epochs, batch_size] int32: per epoch `batch_size` distinct buffer rows (random.sample) = the top-k positions
        of iid uniforms."""
        if self.sampler_seed is not None:
            return self._fused.sample_rows(batch_size, epochs, len(self.buffer), self.sampler_seed, _entry=self._sample_entry)
        u = torch.rand((epochs, len(self.buffer)), device=self.device)
        return u.topk(batch_size, dim=1).indices.to(torch.int32).contiguous()


class NetSnapshot:
    """A stored strategy net as ONE cloned weight blob (a single device copy) instead of a freshly constructed
    FlexibleNet + load_state_dict (about 0.8 ms of host time per snapshot, more than a whole fused train() call).
    Calling it forwards a batch like the module would; anything else a caller might ask of the stored net
    (`parameters()`, `state_dict()`, `.backbone`, ...) materialises the FlexibleNet once, with its parameters as views
    of the blob."""

    _SHAPES = ((128, 34), (128,), (64, 128), (64,), (16, 64), (16,))

    def __init__(self, blob, input_dim=34):
        assert blob.numel() == sdcfr.NET_FLOATS and input_dim == 34
        self._blob = blob.detach().clone()
        self._input_dim = input_dim
        self._module = None

    def _views(self):
        out, off = [], 0
        for shp in self._SHAPES:
            n = int(np.prod(shp))
            out.append(self._blob[off:off + n].view(shp))
            off += n
        return out

    def __call__(self, x):
        w1, b1, w2, b2, w3, b3 = self._views()
        return F.linear(F.relu(F.linear(F.relu(F.linear(x, w1, b1)), w2, b2)), w3, b3)

    def module(self):
        if self._module is None:
            net = FlexibleNet(mode="mlp", input_shape=(self._input_dim,), output_dim=16, mlp_hidden=HIDDEN, mlp_act="relu",
                              mlp_norm="none").to(self._blob.device)
            params = [net.backbone[0].fc.weight, net.backbone[0].fc.bias, net.backbone[1].fc.weight,
                      net.backbone[1].fc.bias, net.head.weight, net.head.bias]
            for p, v in zip(params, self._views()):
                p.data = v
            self._module = net
        return self._module

    def parameters(self):
        return self.module().parameters()

    def state_dict(self, *args, **kwargs):
        return self.module().state_dict(*args, **kwargs)

    def __getattr__(self, name):                 # reached only when normal lookup fails
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.module(), name)


class StrategyBuffer:
    """Stores past strategies for final policy computation.

    Same lists as the reference (`strategies`, `weights`, `max_size`; deep_cfr.py:118-160).  When the stored nets live on
    a CUDA device the average policy of ALL of them is two kernel launches (`ms_sdcfr_average_policy`: one CTA per net)
    instead of one batch-1 forward per net; the nets are snapshots, so their flattened weights are cached per module."""

    _entry = None           # tests only: emulated entry point (see sdcfr.average_policy)

    def __init__(self, max_size=100):
        self.strategies = []
        self.weights = []
        self.max_size = max_size
        self._blob_of = {}          # id(module) -> (module, flat fp32 blob)
        self._stack_key, self._stack = None, None

    def add_strategy(self, strategy_net, iteration):
        if len(self.strategies) >= self.max_size:
            self.strategies.pop(0)
            self.weights.pop(0)
        self.strategies.append(strategy_net)
        self.weights.append(iteration + 1)

    def _stacked(self):
        """([K, 13776] blobs, [K] fp32 weights / total) of the current lists; rebuilt only when the lists changed."""
        key = (tuple(id(s) for s in self.strategies), tuple(self.weights))
        if key != self._stack_key:
            live = set(key[0])
            self._blob_of = {i: v for i, v in self._blob_of.items() if i in live}
            for s_ in self.strategies:
                if id(s_) not in self._blob_of:
                    blob = getattr(s_, "_blob", None)            # NetSnapshot: already flat
                    self._blob_of[id(s_)] = (s_, blob if blob is not None else sdcfr.flatten_net(s_))
            nets = torch.stack([self._blob_of[i][1] for i in key[0]]).contiguous()
            total_weight = sum(self.weights)
            # the reference multiplies a float32 array by the Python float weight / total_weight
            w = torch.tensor([np.float32(w_ / total_weight) for w_ in self.weights], dtype=torch.float32, device=nets.device)
            self._stack_key, self._stack = key, (nets, w)
        return self._stack

    def average_policy_batch(self, feat, mask):
        """feat [n, 34], mask [n, 16] tensors on the nets' device -> [n, 16] average policy (needs >= 1 strategy)."""
        nets, w = self._stacked()
        return sdcfr.average_policy(nets, w, feat.to(torch.float32).contiguous(), mask.to(torch.float32).contiguous(),
                                    _entry=self._entry)

    def get_average_policy(self, state_features, legal_actions_mask):
        if not self.strategies:
            mask = np.asarray(legal_actions_mask).astype(np.float32)
            return mask / mask.sum()
        first = self.strategies[0]
        dev = first._blob.device if isinstance(first, NetSnapshot) else next(first.parameters()).device
        x = torch.as_tensor(np.asarray(state_features), dtype=torch.float32, device=dev).unsqueeze(0)
        m = torch.as_tensor(np.asarray(legal_actions_mask), dtype=torch.float32, device=dev).unsqueeze(0)
        if dev.type == "cuda" or self._entry is not None:
            return self.average_policy_batch(x, m)[0].cpu().numpy().astype(np.float32)
        # strategy nets that a caller keeps on the CPU: the reference's own loop (torch on the CPU)
        policy = torch.zeros_like(m)
        total_weight = sum(self.weights)
        with torch.no_grad():
            for strategy, weight in zip(self.strategies, self.weights):
                policy += positive_regret_policy(strategy(x), m) * (weight / total_weight)
        return policy[0].cpu().numpy().astype(np.float32)


class RandomPolicy:
    """Simple random policy for evaluation."""

    def action_probabilities(self, state, player_id=None):
        if state.is_terminal():
            return {}
        if player_id is None:
            player_id = state.current_player()
        legal_actions = state.legal_actions(player_id)
        prob = 1.0 / len(legal_actions)
        return {action: prob for action in legal_actions}


class DeepCFR:
    """Main Deep CFR algorithm."""

    def __init__(self, game, num_players=2, device="cuda", precision="fp32", traversals_per_iteration=1, seed=0,
                 verbose=False, optimizer="torch", device_eval=False, sampler_seed=None):
        self.game = game
        self.num_players = num_players
        self.device = device
        self.precision = precision
        self.traversals_per_iteration = int(traversals_per_iteration)
        self.seed = int(seed)
        self.verbose = verbose
        self.device_eval = bool(device_eval)
        self._trav_count = 0
        self.input_dim = self._estimate_input_dim()
        if verbose:
            print(f"Estimated input dimension: {self.input_dim}")
        self.advantage_nets = [AdvantageNetwork(self.input_dim, 16, device, precision=precision, optimizer=optimizer,
                                                sampler_seed=None if sampler_seed is None else int(sampler_seed) * 2 + p)
                               for p in range(num_players)]
        self.strategy_buffers = [StrategyBuffer() for _ in range(num_players)]
        self.training_history = {
            "losses": [[] for _ in range(num_players)],
            "values": [[] for _ in range(num_players)],
            "buffer_sizes": [[] for _ in range(num_players)],
            "eval_rewards": [],
            "eval_scopas": [],
        }
        words, order = root_of(game)
        self._root = root_id(words, order)
        self._root_words = (tuple(int(w) for w in words), int(order))
        self._traverser = sdcfr.Traverser(words, order, device=device)
        self._root_state = None

    def _estimate_input_dim(self):
        test_state = self.game.new_initial_state()
        return len(self._state_to_features(test_state, 0))

    def _state_to_features(self, state, player):
        """hand one-hot[16] | table one-hot[16] | [player == current_player, 0.0]  (reference :213-275)."""
        features = np.zeros(34, dtype=np.float32)
        if state.is_terminal() or player < 0:
            return features                                    # "TERMINAL" has no H[ / T[ parts -> zeros
        g = state.env.game
        for c in g.players[player].hand:
            features[codec.card_id(c.rank, c.suit)] = 1.0
        for c in g.table:
            features[16 + codec.card_id(c.rank, c.suit)] = 1.0
        features[32] = float(player == state.current_player())
        return features

    def _get_legal_actions_mask(self, state, player):
        legal_actions = state.legal_actions(player)
        mask = np.zeros(16, dtype=np.float32)
        mask[legal_actions] = 1.0
        return mask

    def _external_sampling_cfr(self, state, player, depth=0, prob=1.0):
        """External sampling CFR from the root: runs `traversals_per_iteration` traversals on the GPU, adds
        their samples to the traverser's replay buffer, returns the (mean) root value."""
        if state.is_terminal():
            return float(state.rewards()[player])
        words, order = state.env.packed()
        if root_id(words, order) != self._root:
            raise NotImplementedError("_external_sampling_cfr on a non-root state")
        n = self.traversals_per_iteration
        blobs = [a.blob() for a in self.advantage_nets]
        prec = sdcfr.TENSOR_CORE if self.precision == "bf16" else sdcfr.FP32
        feat, target, mask, value = self._traverser.run(player, blobs, n, philox_seed=self.seed,
                                                        first_trav=self._trav_count, precision=prec)
        self._trav_count += n
        self.advantage_nets[player].buffer.add_batch(feat, target, mask)
        return float(value.mean().item())

    def _evaluate_vs_random_device(self, num_episodes):
        """All episodes in two launches of the policy-evaluation kernel (`ms_eval_policies`, the path evaluate_agent uses
        for the tabular trainers): seat 0 for the first half of the episodes, seat 1 afterwards, like the loop below.
        The average policy of every infoset comes from `average_policy_table()`; actions are drawn from a Philox stream
        whose seed is taken from np.random (so np.random.seed() still makes a run repeatable)."""
        tab = self.average_policy_table()
        sv = self._solver
        uni = sv.uniform_policy()
        n0 = int(np.ceil(num_episodes / 2))                       # episodes with episode < num_episodes / 2
        seed = int(np.random.randint(0, 2 ** 31 - 1))
        r_a, s_a = sv.evaluate(tab, uni, n0, philox_seed=seed, first_game=0)
        r_b, s_b = sv.evaluate(uni, tab, num_episodes - n0, philox_seed=seed, first_game=n0)
        total_reward = float(r_a.double().sum().item()) - float(r_b.double().sum().item())   # zero-sum: r1 = -r0
        s_a, s_b = s_a.long(), s_b.long()
        trained = int(s_a[:, 0].sum().item()) + int(s_b[:, 1].sum().item())
        opponent = int(s_a[:, 1].sum().item()) + int(s_b[:, 0].sum().item())
        avg_reward = total_reward / num_episodes
        scopas = [trained / num_episodes, opponent / num_episodes]
        self.training_history["eval_rewards"].append(avg_reward)
        self.training_history["eval_scopas"].append(scopas)
        return avg_reward, scopas

    def evaluate_vs_random(self, num_episodes=100, on_device=None):
        if (self.device_eval if on_device is None else on_device) and num_episodes > 0:
            return self._evaluate_vs_random_device(num_episodes)
        total_reward = 0.0
        total_trained_scopas = 0
        total_random_scopas = 0
        random_policy = RandomPolicy()
        for episode in range(num_episodes):
            trained_seat, random_seat = (0, 1) if episode < num_episodes / 2 else (1, 0)
            state = self.game.new_initial_state()
            while not state.is_terminal():
                current_player = state.current_player()
                if current_player == trained_seat:
                    policy_probs = self.get_policy(state, current_player)
                    legal_actions = state.legal_actions(current_player)
                    action_probs = np.array([policy_probs[a] for a in legal_actions])
                    if np.any(np.isnan(action_probs)) or np.sum(action_probs) <= 0:
                        action_probs = np.ones(len(legal_actions)) / len(legal_actions)
                    else:
                        action_probs = action_probs / np.sum(action_probs)
                    action = np.random.choice(legal_actions, p=action_probs)
                else:
                    action_probs = random_policy.action_probabilities(state, current_player)
                    actions, probs = zip(*action_probs.items())
                    action = np.random.choice(actions, p=probs)
                state.apply_action(action)
            total_reward += state.rewards()[trained_seat]
            game = state.env.game
            total_trained_scopas += game.players[trained_seat].scopas
            total_random_scopas += game.players[random_seat].scopas
        avg_reward = total_reward / num_episodes
        scopas = [total_trained_scopas / num_episodes, total_random_scopas / num_episodes]
        self.training_history["eval_rewards"].append(avg_reward)
        self.training_history["eval_scopas"].append(scopas)
        return avg_reward, scopas

    def train(self, iterations=100, advantage_epochs=10, eval_freq=5, eval_episodes=50):
        for iteration in range(iterations):
            for player in range(self.num_players):
                # the traversal only reads the root (the reference builds a fresh one per traversal, :443; here that
                # is a deal on the device plus a read-back, so one root object serves every iteration)
                if self._root_state is None:
                    self._root_state = self.game.new_initial_state()
                value = self._external_sampling_cfr(self._root_state, player)
                loss = self.advantage_nets[player].train(epochs=advantage_epochs)
                self.training_history["losses"][player].append(loss)
                self.training_history["values"][player].append(value)
                self.training_history["buffer_sizes"][player].append(len(self.advantage_nets[player].buffer))
            if iteration > 0:
                for player in range(self.num_players):
                    adv = self.advantage_nets[player]
                    if adv._fused is not None:               # fused configuration: one device copy per snapshot
                        strategy_net = NetSnapshot(adv.blob(), self.input_dim)
                    else:
                        strategy_net = FlexibleNet(mode="mlp", input_shape=(self.input_dim,), output_dim=16,
                                                   mlp_hidden=HIDDEN, mlp_act="relu", mlp_norm="none").to(self.device)
                        strategy_net.load_state_dict(adv.net.state_dict())
                    self.strategy_buffers[player].add_strategy(strategy_net, iteration)
            if iteration % eval_freq == 0 and eval_episodes > 0:
                self.evaluate_vs_random(num_episodes=eval_episodes)

    # -- whole-game views of the average policy (device-batched; not in the reference) --------------------
    def _policy_inputs(self):
        """Features / masks / legal-action lists of every infoset of the deal as device tensors (static: built once)."""
        if getattr(self, "_pol_in", None) is not None:
            return self._pol_in
        from ...solver import Solver
        if getattr(self, "_solver", None) is None:
            self._solver = Solver(self._root_words[0], self._root_words[1], device=self.device)
        sv = self._solver
        st = sv.static_table()
        S = sv.n_slots
        feat = np.zeros((S, 34), dtype=np.float32)
        mask = np.zeros((S, 16), dtype=np.float32)
        for s in range(S):
            f = codec.key_fields(st["keys"][s])
            for c in range(16):
                feat[s, c] = (f["hand_mask"] >> c) & 1
            for c in f["table"]:
                feat[s, 16 + c] = 1.0
            feat[s, 32] = 1.0
            mask[s] = feat[s, :16]
        x = torch.from_numpy(feat).to(self.device)
        m = torch.from_numpy(mask).to(self.device)
        legal = torch.from_numpy(np.where(st["legal"] == 255, 0, st["legal"]).astype(np.int64)).to(self.device)
        nl = torch.from_numpy(st["nlegal"].astype(np.int64)).to(self.device)
        player = torch.from_numpy(st["player"]).to(self.device)
        self._pol_in = (x, m, legal, nl, player)
        return self._pol_in

    def average_policy_table(self):
        """[S, 4] float64 CUDA tensor: get_policy() of every infoset of the deal at once, restricted to the legal
        actions in hand order and normalised the way evaluate_vs_random does (uniform if the sum is <= 0)."""
        x, m, legal, nl, player = self._policy_inputs()
        S = x.shape[0]
        pol16 = torch.zeros((S, 16), dtype=torch.float32, device=self.device)
        with torch.no_grad():
            for p in range(self.num_players):
                buf = self.strategy_buffers[p]
                rows = player == p
                if not buf.strategies:
                    pol16[rows] = m[rows] / m[rows].sum(1, keepdim=True)
                    continue
                pol16[rows] = buf.average_policy_batch(x[rows], m[rows])
        tab = torch.gather(pol16.double(), 1, legal)
        valid = torch.arange(4, device=self.device).unsqueeze(0) < nl.unsqueeze(1)
        tab = tab * valid
        ssum = tab.sum(1, keepdim=True)
        uni = valid.double() / nl.unsqueeze(1).double()
        return torch.where(ssum > 0, tab / ssum.clamp_min(1e-300), uni)

    def exploitability(self):
        """Best-response exploitability of the average policy (device sweep, restated open_spiel BR)."""
        tab = self.average_policy_table()
        self._solver.import_table(strategy=tab.cpu().numpy())
        return self._solver.exploitability(0)

    def get_policy(self, state, player):
        """Get average policy for a state."""
        return self.strategy_buffers[player].get_average_policy(self._state_to_features(state, player),
                                                                self._get_legal_actions_mask(state, player))

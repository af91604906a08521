"""Learner data path: what learner/data_loader.cpp:134-200 assembles for every training sample of a board game —
the rotated feature planes of a sampled position, its policy target and its value — with the feature reconstruction
(BaseEnvLoader::getFeatures, environment/base/base_env.h:235-241: "replays the game again to get features") done for the whole
batch in one device call (Engine.replay_features -> mz_replay_features) instead of one environment replay per sample on a host thread.

Reference interfaces mirrored (paths relative to /root/reference/minizero):
  environment/base/base_env.h:149-205   BaseEnvLoader::loadFromString       -> parse_record
  environment/base/base_env.h:243-269   BaseEnvLoader::getPolicy            -> Record.policy
  environment/go/go.h:137, othello.h:79 getValue = the game's return        -> Record.value
  utils/rotation.h:51-93                getPositionByRotating               -> rotate_action
  learner/data_loader.cpp:134-157       setAlphaZeroTrainingData            -> alphazero_batch
  learner/data_loader.cpp:159-200       setMuZeroTrainingData               -> muzero_batch (board games: go.cpp:725-736, othello.cpp:264-275 action planes)
"""
import re

import numpy as np


def rotate_action(action, rotation, board_size):
    """utils::getPositionByRotating (rotation.h:51-93); the pass (= board_size^2) does not rotate. Doubled integer coordinates
    instead of the reference's float centre: the same cell for every board size."""
    n = board_size
    if action == n * n:
        return action
    x, y = 2 * (action % n) - (n - 1), 2 * (action // n) - (n - 1)
    rx, ry = [(x, y), (y, -x), (-x, -y), (-y, x), (x, -y), (-y, -x), (-x, y), (y, x)][rotation]
    return ((ry + n - 1) // 2) * n + (rx + n - 1) // 2


class Record:
    """One game record as the zero server hands it to the learner: tags, actions with their players and per-move info tags."""

    def __init__(self, tags, actions, players, infos):
        self.tags, self.actions, self.players, self.infos = tags, actions, players, infos

    def __len__(self):
        return len(self.actions)

    def data_range(self):
        """BaseEnvLoader::getDataRange (base_env.h:271-279)"""
        dlen = self.tags.get("DLEN", "")
        if not dlen:
            return 0, max(0, len(self.actions) - 1)
        return int(dlen.split("-")[0]), int(dlen.split("-")[1])

    def policy(self, pos, rotation, action_size, board_size, rotates=True):
        """BaseEnvLoader::getPolicy: normalised visit counts of the P tag at the rotated action ids; a move without P tag is one-hot;
        positions past the end are uniform (absorbing states). rotates=False: the game's getRotateAction is the identity (Hex)."""
        out = np.zeros(action_size, np.float32)
        rot = (lambda a: rotate_action(a, rotation, board_size)) if rotates else (lambda a: a)
        if pos < len(self.actions):
            dist = self.infos[pos].get("P", "")
            if not dist:
                out[rot(self.actions[pos])] = 1.0
            else:
                total = np.float32(0.0)
                for tok in dist.split(","):
                    a, c = tok.split(":")
                    out[rot(int(a))] = np.float32(float(c))
                    total = np.float32(total + np.float32(float(c)))
                out /= total
        else:
            out[:] = np.float32(1.0) / np.float32(action_size)
        return out

    def value(self, pos):
        """GoEnvLoader / OthelloEnvLoader / ... ::getValue: the game's return (RE tag)"""
        return np.float32(float(self.tags["RE"]))

    def reward(self, pos):
        """BaseEnvLoader::getReward (base_env.h:279): the move's R tag, 0 past the end of the game"""
        return np.float32(float(self.infos[pos]["R"])) if pos < len(self.actions) else np.float32(0.0)

    def action_plane(self, pos, rotation, board_size, rand_int, rotates=True):
        """GoEnvLoader / OthelloEnvLoader::getActionFeatures (go.cpp:725-736, othello.cpp:264-275): one-hot cell of the rotated move (a pass is the empty
        plane); past the end of the game a random action id drawn with the learner's generator (rand_int() = Random::randInt()), set only when it is
        smaller than the game's length — the reference's own condition, kept as it is"""
        n2 = board_size * board_size
        out = np.zeros(n2, np.float32)
        if pos < len(self.actions):
            a = self.actions[pos]
            if a != n2:
                out[rotate_action(a, rotation, board_size) if rotates else a] = 1.0
        else:
            a = rand_int() % (n2 + 1)
            if a < len(self.actions) and a < n2:  # (the reference writes index n2 of an n2-element vector when the draw is n2 and the game is longer: skipped here)
                out[a] = 1.0
        return out


_TAG = re.compile(r"([A-Z]+)\[((?:\\.|[^\]\\])*)\]")


def parse_record(text):
    """BaseEnvLoader::loadFromString on a record `(;GM[..]RE[..]...;B[id]P[..]V[..]R[..];W[id]...)`, or on a whole `SelfPlay ... #` line."""
    if text.startswith("SelfPlay "):
        text = text.split(" ", 5)[5].rsplit(" #", 1)[0]
    assert text.startswith("(;") and text.endswith(")"), "not a game record"
    nodes = text[2:-1].split(";")
    unescape = lambda v: re.sub(r"\\(.)", r"\1", v)
    tags = {k: unescape(v) for k, v in _TAG.findall(nodes[0])}
    actions, players, infos = [], [], []
    for node in nodes[1:]:
        kv = _TAG.findall(node)
        if not kv:
            continue
        players.append(1 if kv[0][0] == "B" else 2)
        actions.append(int(kv[0][1]))
        infos.append({k: unescape(v) for k, v in kv[1:]})
    return Record(tags, actions, players, infos)


def alphazero_batch(engine, records, picks):
    """DataLoaderThread::setAlphaZeroTrainingData for a list of picks (record index, position, rotation): returns
    features [n][C*H*W] (rebuilt on the device), policy [n][A], value [n]. len(picks) <= the engine's number of games."""
    from .engine import GAME_HEX
    rotates = (engine.game != GAME_HEX)  # HexEnvLoader::getRotateAction is the identity (hex.h:65)
    n = len(picks)
    max_len = max(1, max(len(records[r]) for r, _, _ in picks))
    actions = np.full((n, max_len), -1, np.int32)
    for j, (r, _, _) in enumerate(picks):
        actions[j, :len(records[r])] = records[r].actions
    pos = np.array([p for _, p, _ in picks], np.int32)
    rot = np.array([q for _, _, q in picks], np.uint8)
    feats = engine.replay_features(actions, pos, rot)
    policy = np.stack([records[r].policy(p, q, engine.A, engine.board_size, rotates) for r, p, q in picks])
    value = np.array([records[r].value(p) for r, p, _ in picks], np.float32)
    return feats, policy, value


def muzero_batch(engine, records, picks, unrolling_step, rand_int):
    """DataLoaderThread::setMuZeroTrainingData (data_loader.cpp:159-200) for board games: the root planes of every pick rebuilt on the device, then per
    unroll step the action plane, policy target, value (the game's return) and reward. Returns features [n][C*H*W], action_features [n][K][N*N],
    policy [n][K + 1][A], value [n][K + 1], reward [n][K]."""
    from .engine import GAME_HEX
    rotates = (engine.game != GAME_HEX)
    n, K, N = len(picks), int(unrolling_step), engine.board_size
    feats, _, _ = alphazero_batch(engine, records, picks)
    act = np.zeros((n, K, N * N), np.float32)
    policy = np.zeros((n, K + 1, engine.A), np.float32)
    value = np.zeros((n, K + 1), np.float32)
    reward = np.zeros((n, K), np.float32)
    for j, (r, p, q) in enumerate(picks):
        rec = records[r]
        for step in range(K + 1):  # the reference's order of calls per step: action plane, policy, value, reward (the random draws follow it)
            if step < K:
                act[j, step] = rec.action_plane(p + step, q, N, rand_int, rotates)
            policy[j, step] = rec.policy(p + step, q, engine.A, N, rotates)
            value[j, step] = rec.value(p + step)
            if step < K:
                reward[j, step] = rec.reward(p + step)
    return feats, act, policy, value, reward

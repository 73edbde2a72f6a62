"""Self-play training-sample export: the reference's TrainDataExporter (engine/src/rl/
traindataexporter.cpp:32-300) with the same on-disk layout -- a zarr v2 directory store with the datasets

    x              int16   [N, C, 8, 8]     un-normalised input planes of every searched position
    y_value        int16   [N]              game result from the mover's point of view (+1 / 0 / -1)
    y_policy       float32 [N, NB_LABELS]   MCTS posterior scattered to the classic label indices (mirrored for black)
    y_best_move_q  float32 [N]              Q of the selected move
    plys_to_end    int16   [N]              plies until the end of the game
    phase_vector   int16   [N]              game phase id (single-phase build: 0)
    start_indices  int32   [N]              first sample of every game

chunked by `chunk_size` samples (default 128, rl_config.py) and written uncompressed (`"compressor": null`), so any zarr
reader opens it; no zarr package is needed to write it."""
import json
import os

import numpy as np

from .labels import classic_index, uci_labels

WHITE_WIN, BLACK_WIN, DRAWN = 0, 1, 2


class _ZArray:
    def __init__(self, root, name, shape, chunks, dtype):
        self.dir = os.path.join(root, name)
        os.makedirs(self.dir, exist_ok=True)
        self.shape, self.chunks, self.dtype = tuple(shape), tuple(chunks), np.dtype(dtype)
        meta = {"zarr_format": 2, "shape": list(self.shape), "chunks": list(self.chunks), "dtype": self.dtype.str,
                "compressor": None, "fill_value": 0, "order": "C", "filters": None}
        with open(os.path.join(self.dir, ".zarray"), "w") as f:
            json.dump(meta, f)

    def _chunk_path(self, i):
        return os.path.join(self.dir, ".".join([str(i)] + ["0"] * (len(self.shape) - 1)))

    def write(self, start, data):
        """Rows [start, start+len(data)) along the first axis (chunks are whole along the other axes)."""
        data = np.ascontiguousarray(data, self.dtype)
        c = self.chunks[0]
        pos = 0
        while pos < len(data):
            row = start + pos
            ci, off = divmod(row, c)
            n = min(c - off, len(data) - pos)
            path = self._chunk_path(ci)
            if os.path.exists(path):
                chunk = np.fromfile(path, self.dtype).reshape((c,) + self.shape[1:])
            else:
                chunk = np.zeros((c,) + self.shape[1:], self.dtype)
            chunk[off:off + n] = data[pos:pos + n]
            chunk.tofile(path)
            pos += n


class TrainDataExporter:
    def __init__(self, path, mode, channels, number_chunks=200, chunk_size=128):
        self.path, self.mode = path, mode
        self.n_labels = len(uci_labels(mode))
        self.number_samples = number_chunks * chunk_size
        os.makedirs(path, exist_ok=True)
        with open(os.path.join(path, ".zgroup"), "w") as f:
            json.dump({"zarr_format": 2}, f)
        N, cs = self.number_samples, chunk_size
        self.d_x = _ZArray(path, "x", (N, channels, 8, 8), (cs, channels, 8, 8), "<i2")
        self.d_value = _ZArray(path, "y_value", (N,), (cs,), "<i2")
        self.d_policy = _ZArray(path, "y_policy", (N, self.n_labels), (cs, self.n_labels), "<f4")
        self.d_q = _ZArray(path, "y_best_move_q", (N,), (cs,), "<f4")
        self.d_plys = _ZArray(path, "plys_to_end", (N,), (cs,), "<i2")
        self.d_phase = _ZArray(path, "phase_vector", (N,), (cs,), "<i2")
        self.d_start = _ZArray(path, "start_indices", (N,), (cs,), "<i4")
        self.start_idx = 0
        self.game_idx = 0
        self.d_start.write(0, np.array([0], np.int32))  # save_start_idx() of create_new_dataset_file

    def is_file_full(self):
        return self.start_idx >= self.number_samples

    def new_game(self):
        return dict(x=[], value=[], policy=[], q=[], phase=[])

    def save_sample(self, game, planes, legal_uci, policy, best_move_q, side_to_move, phase=0):
        """planes: un-normalised float planes [C,8,8] of the position; legal_uci / policy: EvalInfo.legalMoves and
        policyProbSmall; side_to_move: 0 white, 1 black (its moves are mirrored, traindataexporter.cpp:195-218)."""
        game["x"].append(np.asarray(planes).astype(np.int16))
        row = np.zeros(self.n_labels, np.float32)
        for u, p in zip(legal_uci, policy):
            row[classic_index(self.mode, u, side_to_move == 1)] = p
        game["policy"].append(row)
        game["q"].append(np.float32(best_move_q))
        game["value"].append(np.int16(-(side_to_move * 2 - 1)))  # +1 for white to move, -1 for black
        game["phase"].append(np.int16(phase))

    def export_game_samples(self, game, result):
        n = len(game["x"])
        if n == 0 or self.is_file_full():
            return 0
        n = min(n, self.number_samples - self.start_idx)
        value = np.array(game["value"][:n], np.int16)
        if result == BLACK_WIN:
            value = -value
        elif result == DRAWN:
            value = value * 0
        plys = (len(game["x"]) - np.arange(n)).astype(np.int16)  # (idx - curSampleIdx) * -1
        s = self.start_idx
        self.d_x.write(s, np.stack(game["x"][:n]))
        self.d_value.write(s, value)
        self.d_q.write(s, np.array(game["q"][:n], np.float32))
        self.d_policy.write(s, np.stack(game["policy"][:n]))
        self.d_plys.write(s, plys)
        self.d_phase.write(s, np.array(game["phase"][:n], np.int16))
        self.start_idx += n
        self.game_idx += 1
        self.d_start.write(self.game_idx, np.array([self.start_idx], np.int32))
        return n


def read_dataset(path, name):
    """Minimal reader of the uncompressed store written above (tests / inspection)."""
    d = os.path.join(path, name)
    meta = json.load(open(os.path.join(d, ".zarray")))
    shape, chunks, dtype = tuple(meta["shape"]), tuple(meta["chunks"]), np.dtype(meta["dtype"])
    out = np.zeros(shape, dtype)
    for ci in range((shape[0] + chunks[0] - 1) // chunks[0]):
        p = os.path.join(d, ".".join([str(ci)] + ["0"] * (len(shape) - 1)))
        if os.path.exists(p):
            c = np.fromfile(p, dtype).reshape(chunks)
            lo = ci * chunks[0]
            out[lo:lo + chunks[0]] = c[:shape[0] - lo]
    return out

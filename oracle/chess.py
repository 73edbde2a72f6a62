"""ctypes wrapper around the C chess oracle (oracle/chess.c) -- TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

VARIANTS = {"chess": 0, "crazyhouse": 1, "kingofthehill": 2, "koth": 2, "3check": 3, "threecheck": 3, "anti": 4,
            "atomic": 5, "horde": 6, "racingkings": 7}
T_LOSS, T_DRAW, T_WIN, T_CUSTOM, T_NONE = 0, 1, 2, 3, 4


def build():
    subprocess.run(["make", "-s", "-C", _HERE], check=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "liboracle.so")
        if not os.path.exists(path):
            build()
        L = ctypes.CDLL(path)
        L.opos_sizeof.restype = ctypes.c_size_t
        L.opos_perft.restype = ctypes.c_uint64
        L.opos_compute_key.restype = ctypes.c_uint64
        L.opos_checkers_bb.restype = ctypes.c_uint64
        L.opos_pieces_bb.restype = ctypes.c_uint64
        L.opos_uci_to_move.restype = ctypes.c_uint32
        L.opos_start_fen.restype = ctypes.c_char_p
        L.opos_set.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int]
        L.opos_copy.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.opos_fen.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
        L.opos_legal_moves.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.opos_do_move.argtypes = [ctypes.c_void_p, ctypes.c_uint32]
        L.opos_move_to_uci.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_char_p]
        L.opos_uci_to_move.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
        L.opos_in_check.argtypes = [ctypes.c_void_p]
        L.opos_checkers_bb.argtypes = [ctypes.c_void_p]
        L.opos_gives_check.argtypes = [ctypes.c_void_p, ctypes.c_uint32]
        L.opos_is_terminal.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.opos_number_repetitions.argtypes = [ctypes.c_void_p]
        L.opos_compute_key.argtypes = [ctypes.c_void_p]
        L.opos_perft.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.opos_pieces_bb.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        L.opos_start_fen.argtypes = [ctypes.c_int]
        _LIB = L
    return _LIB


class Position:
    """One game state (BoardState of the reference: engine/src/environments/chess_related/boardstate.h)."""

    def __init__(self, fen=None, variant="chess", is960=False):
        L = lib()
        self.variant = VARIANTS[variant] if isinstance(variant, str) else int(variant)
        self.is960 = bool(is960)
        self._buf = ctypes.create_string_buffer(L.opos_sizeof())
        if fen is None:
            fen = L.opos_start_fen(self.variant).decode()
        if L.opos_set(self._buf, fen.encode(), self.variant, int(self.is960)) != 0:
            raise ValueError(f"bad FEN: {fen}")

    @property
    def ptr(self):
        return ctypes.addressof(self._buf)

    def clone(self):
        q = Position.__new__(Position)
        q.variant, q.is960 = self.variant, self.is960
        q._buf = ctypes.create_string_buffer(len(self._buf))
        lib().opos_copy(q._buf, self._buf)
        return q

    def fen(self):
        b = ctypes.create_string_buffer(256)
        lib().opos_fen(self._buf, b)
        return b.value.decode()

    def legal_moves(self):
        arr = (ctypes.c_uint32 * 512)()
        n = lib().opos_legal_moves(self._buf, arr)
        return list(arr[:n])

    def uci(self, move):
        b = ctypes.create_string_buffer(8)
        lib().opos_move_to_uci(self._buf, move, b)
        return b.value.decode()

    def legal_uci(self):
        return [self.uci(m) for m in self.legal_moves()]

    def move_from_uci(self, s):
        m = lib().opos_uci_to_move(self._buf, s.encode())
        if m == 0:
            raise ValueError(f"illegal move {s} in {self.fen()}")
        return m

    def do_move(self, move):
        lib().opos_do_move(self._buf, move)

    def push_uci(self, *moves):
        for s in moves:
            self.do_move(self.move_from_uci(s))
        return self

    def in_check(self):
        return bool(lib().opos_in_check(self._buf))

    def gives_check(self, move):
        return bool(lib().opos_gives_check(self._buf, move))

    def terminal(self, n_legal=None):
        if n_legal is None:
            n_legal = len(self.legal_moves())
        return lib().opos_is_terminal(self._buf, n_legal)

    def number_repetitions(self):
        return lib().opos_number_repetitions(self._buf)

    def key(self):
        return lib().opos_compute_key(self._buf)

    def perft(self, depth):
        return lib().opos_perft(self._buf, depth)

    def pieces_bb(self, color, pt=0):
        return lib().opos_pieces_bb(self._buf, color, pt)

    def side_to_move(self):
        return ctypes.c_int.from_buffer(self._buf, 64 + 64 + 2 * 7 * 4).value

    def check_result(self):
        """State::check_result (engine/src/state.h): +1 white win, -1 black win, 0 draw, None if not over."""
        t = self.terminal()
        if t == T_NONE:
            return None
        if t == T_DRAW:
            return 0
        stm_white = self.side_to_move() == 0
        if t == T_WIN:
            return 1 if stm_white else -1
        return -1 if stm_white else 1


MODES = {"crazyhouse": 0, "chess": 1, "lichess": 2}


def planes(pos, mode, version, normalize):
    """board_to_planes (inputrepresentation.cpp:628-680) -> float32 array [C, 8, 8]."""
    import numpy as np
    L = lib()
    L.oplanes_channels.argtypes = [ctypes.c_int, ctypes.c_int]
    L.oplanes_encode.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    m = MODES[mode] if isinstance(mode, str) else mode
    c = L.oplanes_channels(m, version)
    if c < 0:
        raise ValueError("unsupported mode/version")
    out = np.zeros((c, 8, 8), np.float32)
    got = L.oplanes_encode(pos._buf, m, version, int(bool(normalize)), out.ctypes.data_as(ctypes.c_void_p))
    if got != c:
        raise RuntimeError(f"plane encoder channel mismatch {got} != {c}")
    return out


def plane_stats(p):
    """get_stats_from_input_planes (engine/tests/tests.cpp:67-80)."""
    import numpy as np
    v = p.reshape(-1).astype(np.float64)
    mx, arg = 0.0, 0
    for i, x in enumerate(v):
        if x > mx:
            mx, arg = x, i
    return dict(sum=float(v.sum()), max=float(mx), argmax=int(arg), key=float((np.arange(v.size) * v).sum()))

"""Builds and loads the 1-lane host emulation of the device rules/search code (test scaffolding only)."""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libhostemu.so")
        srcs = [os.path.join(_HERE, "hostemu.cpp")] + [os.path.join(_ROOT, "crazyara_b200", "csrc", f) for f in
                                                       os.listdir(os.path.join(_ROOT, "crazyara_b200", "csrc"))
                                                       if f.endswith((".cuh", ".h"))]
        if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
            subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-ffp-contract=off",
                            "-I" + os.path.join(_ROOT, "crazyara_b200", "csrc"), "-I" + os.path.join(_ROOT, "include"),
                            "-x", "c++", os.path.join(_HERE, "hostemu.cpp"), "-o", so], check=True)
        L = ctypes.CDLL(so)
        L.he_new.restype = ctypes.c_void_p
        L.he_new.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int]
        L.he_clone.restype = ctypes.c_void_p
        for n in ("he_free", "he_clone", "he_in_check", "he_repetition", "he_terminal"):
            getattr(L, n).argtypes = [ctypes.c_void_p]
        L.he_legal_moves.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.he_move_uci.argtypes = [ctypes.c_void_p, ctypes.c_uint16, ctypes.c_char_p]
        L.he_uci_move.restype = ctypes.c_uint16
        L.he_uci_move.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
        L.he_do_move.argtypes = [ctypes.c_void_p, ctypes.c_uint16]
        L.he_fen.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
        L.he_key.restype = ctypes.c_ulonglong
        L.he_key.argtypes = [ctypes.c_void_p]
        L.he_key_scratch.restype = ctypes.c_ulonglong
        L.he_key_scratch.argtypes = [ctypes.c_void_p]
        L.he_policy_index.argtypes = [ctypes.c_void_p, ctypes.c_uint16]
        L.he_planes.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        L.he_board.restype = ctypes.c_void_p
        L.he_board.argtypes = [ctypes.c_void_p]
        _LIB = L
    return _LIB


class HeState:
    def __init__(self, fen, variant, is960=False, _h=None):
        self.L = lib()
        self.h = _h if _h is not None else self.L.he_new(fen.encode(), variant, int(is960))
        if not self.h:
            raise ValueError("bad fen")

    def __del__(self):
        if getattr(self, "h", None):
            self.L.he_free(self.h)
            self.h = None

    def clone(self):
        return HeState(None, 0, _h=self.L.he_clone(self.h))

    def legal_moves(self):
        arr = (ctypes.c_uint16 * 512)()
        n = self.L.he_legal_moves(self.h, arr)
        return list(arr[:n])

    def uci(self, m):
        b = ctypes.create_string_buffer(8)
        self.L.he_move_uci(self.h, m, b)
        return b.value.decode()

    def move_from_uci(self, s):
        m = self.L.he_uci_move(self.h, s.encode())
        if m == 0:
            raise ValueError("illegal " + s)
        return m

    def do_move(self, m):
        self.L.he_do_move(self.h, m)

    def fen(self):
        b = ctypes.create_string_buffer(256)
        self.L.he_fen(self.h, b)
        return b.value.decode()

    def key(self):
        return self.L.he_key(self.h)

    def key_scratch(self):
        return self.L.he_key_scratch(self.h)

    def terminal(self):
        return self.L.he_terminal(self.h)

    def repetition(self):
        return self.L.he_repetition(self.h)

    def policy_index(self, m):
        return self.L.he_policy_index(self.h, m)

    def planes(self, mode, version, normalize):
        import numpy as np
        out = np.full((80, 8, 8), np.nan, np.float32)
        c = self.L.he_planes(self.h, mode, version, int(normalize), out.ctypes.data)
        return out[:c]

    def board_bytes(self):
        return ctypes.string_at(self.L.he_board(self.h), 128)

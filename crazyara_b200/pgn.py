"""PGN export of self-play games: the reference's GamePGN record and its stream layout (engine/src/rl/gamepgn.{h,cpp}),
filled the way rl/selfplay.cpp does (set_game_pgn_header :90-115, play_move_and_update :38-55, set_game_result :327-330).

Moves are SAN as `State::action_to_san` spells them (ara_state_move_to_san); a move that ends the game with a win gets
'#' (replacing a trailing '+'), as play_move_and_update does."""
import datetime
import os

UCI_VARIANT_NAMES = {0: "chess", 1: "crazyhouse", 2: "kingofthehill", 3: "3check"}


def result_string(result):
    """selfplay.cpp:327-330: the PGN result token of a game outcome (crazyara_b200.export.WHITE_WIN / BLACK_WIN / DRAWN)."""
    from .export import BLACK_WIN, WHITE_WIN
    return "1-0" if result == WHITE_WIN else ("0-1" if result == BLACK_WIN else "1/2-1/2")


class GamePGN:
    def __init__(self, uci_variant="chess", is960=False, white="?", black="?", date=None):
        """set_game_pgn_header: `uci_variant` is the UCI_Variant option value (or the variant id of the C-ABI)."""
        if not isinstance(uci_variant, str):
            uci_variant = UCI_VARIANT_NAMES[int(uci_variant)]
        self.variant = "standard" if (not is960 and uci_variant == "chess") else uci_variant + ("960" if is960 else "")
        self.event, self.site, self.round = "SelfPlay", "Darmstadt, GER", "?"
        self.date = date if date is not None else datetime.datetime.now().strftime("%Y.%m.%d %X")
        self.white, self.black = white, black
        self.time_control = "?"
        self.is960 = is960
        self.fen = "?"
        self.new_game()

    def new_game(self):  # GamePGN::new_game
        self.game_moves = []
        self.result = "?"

    def play_move(self, state, action):
        """play_move_and_update: SAN of `action` in `state`, the move is made, a winning move ends in '#'.
        Returns the TerminalType of the position after the move."""
        from .engine import TERMINAL_LOSS, TERMINAL_WIN
        san = state.action_to_san(action)
        if isinstance(action, str):
            state.do_uci(action)
        else:
            state.do_action(action)
        term = state.is_terminal()
        if term in (TERMINAL_LOSS, TERMINAL_WIN):
            san = san[:-1] + "#" if san.endswith("+") else san + "#"
        self.game_moves.append(san)
        return term

    def __str__(self):  # operator<<(ostream&, const GamePGN&)
        out = [f'[Variant "{self.variant}"]', f'[Event "{self.event}"]', f'[Date "{self.date}"]', f'[Site "{self.site}"]',
               f'[Round "{self.round}"]', f'[FEN "{self.fen}"]', f'[White "{self.white}"]', f'[Black "{self.black}"]',
               f'[Result "{self.result}"]', f'[PlyCount "{len(self.game_moves)}"]',
               f'[TimeControl "{self.time_control}"]', ""]
        body = ""
        for ply, mv in enumerate(self.game_moves):
            if ply % 2 == 0:
                body += f"{ply // 2 + 1}. "
            body += mv + " "
            if (ply + 1) % 8 == 0:
                body += "\n"
        body += self.result + "\n\n"
        return "\n".join(out) + "\n" + body

    def write(self, path):
        """write_game_to_pgn (selfplay.cpp:315-324): append the game, followed by an empty line."""
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        with open(path, "a") as f:
            f.write(str(self) + "\n")

"""Records the search oracle's results for the parity cases (fake backend) as tests/golden/search_oracle.json.
The reference has no test that runs a search ("the oracle's own dumps become the golden files"): this file
does not pin the oracle to the reference, it pins it to ITSELF -- a change of oracle/mcts.c that moves a visit count shows
up as a diff of this file, in review, instead of silently moving the target the device code is compared with.
    python tests/golden/gen_search_golden.py"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def run_case(case):
    from oracle import search as osr
    from oracle.chess import Position
    variant, vid, mode, fen, is960, premoves, batch, sims, extra = case
    from tests.test_search_hostemu import case_settings
    st = case_settings(mode, batch, sims, extra)
    pos = Position(fen, variant, is960)
    pos.push_uci(*premoves)
    S = osr.Search(st)
    r = S.run(pos, osr.fake_net(S.n_labels), with_keys=True)
    f32 = lambda a: [int(x) for x in np.asarray(a, np.float32).view(np.uint32)]  # noqa: E731  (exact bit patterns)
    return dict(moves=r["moves"], visits=[int(v) for v in r["visits"]], q_bits=f32(r["q"]), prior_bits=f32(r["prior"]),
                root_value_bits=f32([r["root_value"]])[0], best_move=r.get("best_move"), nodes=int(r["nodes"]),
                visit_sum=int(r["visit_sum"]), free_visits=int(r["free_visits"]), tree_nodes=int(r["tree_nodes"]),
                iterations=int(r["iterations"]), sum_depth=int(r["sum_depth"]), node_type=int(r["node_type"]))


def main():
    from tests.test_search_hostemu import CASES
    out = [dict(case=list(c[:8]) + [c[8]], result=run_case(c)) for c in CASES]
    json.dump(out, open(os.path.join(HERE, "search_oracle.json"), "w"), separators=(",", ":"))
    print(len(out), "cases")


if __name__ == "__main__":
    main()

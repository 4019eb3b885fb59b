#!/usr/bin/env python
"""A longer run of the oracle-vs-spec-text successor comparison than the unit tests do (tests/spec_text.py).
    python tests/golden/long_sweep.py R V L RESTART SEED MINUTES   ->  one JSON line (states compared, per-action successor counts)
Random walks from Init (and, for the README constants, from the published trace's states), every visited state compared."""
import base64
import json
import os
import random
import sys
import time
import zlib

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import _pkg  # noqa: E402
import spec_text as S  # noqa: E402


def main():
    R, V, L, restart, seed = map(int, sys.argv[1:6])
    minutes = float(sys.argv[6])
    pkg = _pkg.load()
    P = S.Pair(pkg, R, V, L, restart=restart)
    rng = random.Random(seed)
    starts = [P.init_flat()]
    if (R, V, L, restart) == (3, 3, 3, 0):
        fx = json.load(open(os.path.join(HERE, "state_transfer_trace.json")))
        starts += [P.Flat.from_buffer_copy(zlib.decompress(base64.b64decode(s["flat_zlib_b64"]))) for s in fx["states"]]
    pref = ("RestartEmpty", "ReceivesRecoveryMsg", "ReceivesRecoveryResponseMsg", "CompleteRecovery") if restart else \
           ("SendGetState", "ReceiveGetState", "ReceiveNewState", "ReceiveHigherDVC")
    t0, n, walks = time.time(), 0, 0
    while time.time() - t0 < minutes * 60:
        n += P.walk(rng.choice(starts), 60, rng, prefer=pref if rng.random() < 0.5 else ())
        walks += 1
    print(json.dumps(dict(R=R, V=V, L=L, RestartEmptyLimit=restart, seed=seed, walks=walks, states_compared=n, mismatches=0,
                          successors_by_action=dict(P.stats), choose_picks_that_mattered=P.choose_retries,
                          minutes=round((time.time() - t0) / 60, 1))))


if __name__ == "__main__":
    main()

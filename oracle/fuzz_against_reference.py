"""CPU oracle -- extra pinning runs (TEST INFRASTRUCTURE, build container only: needs /root/reference).

Runs the reference's OWN level5 envs (through oracle/make_golden_level5.py) on seeds that are NOT among the committed
recordings and replays them through oracle/level5_oracle.py with the checks of tests/test_oracle_golden_level5.py.
Nothing is written under tests/golden; a mismatch is an oracle bug (or an unrecorded reference path) to look at.

    python -m oracle.fuzz_against_reference stage03 exp02_vFinal 2001 4 800 0.9
    python -m oracle.fuzz_against_reference stage02 3001 2 600 0.9 150
    python -m oracle.fuzz_against_reference stage01 4001 1 500 0.8
    python -m oracle.fuzz_against_reference dumb 711 7 900 5
    python -m oracle.fuzz_against_reference eval 811 2 1200
    python -m oracle.fuzz_against_reference fusion 611 6 900 0.9
    python -m oracle.fuzz_against_reference c1 511 3 1000 0.5
    python -m oracle.fuzz_against_reference fusion 612 2 500 0.9 40      # chase probability, ram from step 40 on
"""
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import make_golden_level5 as m          # noqa: E402
from oracle.level5_oracle import Level5Oracle      # noqa: E402


def main(argv):
    kind = argv[0]
    if kind != "stage03":
        seed, env_index, steps = int(argv[1]), int(argv[2]), int(argv[3])
    extra = argv[4] if len(argv) > 4 and kind != "stage03" else None
    import tests.test_oracle_golden_level5 as T
    from tests.util import load_recording
    t0 = time.time()
    if kind == "stage03":              # stage03 <preset> <seed> <env> <steps> [chase_prob] [kamikaze_after]
        from oracle import make_golden as m3
        import tests.test_oracle_golden as T3
        preset, seed, env_index, steps = argv[1], int(argv[2]), int(argv[3]), int(argv[4])
        chase = float(argv[5]) if len(argv) > 5 else 0.9
        kami = int(argv[6]) if len(argv) > 6 else None
        rec = m3.run_reference(preset, seed, env_index, steps, seed + 1, 0.02, chase, kami)
        t1 = time.time()
        with tempfile.TemporaryDirectory() as d:
            path = os.path.join(d, "stage03_fuzz.npz")
            rec["lidar"] = rec["lidar"].astype(np.float32)
            np.savez_compressed(path, preset=np.array(preset), **rec)
            T3.test_oracle_matches_reference_recording(path)
        print(f"OK stage03 {preset} seed {seed} env {env_index} steps {steps} chase {chase} ram {kami}: episodes "
              f"{int(np.sum(rec['done']))}, reference {t1 - t0:.0f} s, replay {time.time() - t1:.0f} s", flush=True)
        return
    if kind in ("stage02", "stage01"):   # stage02 <seed> <env> <steps> [chase_prob] [ram_after];  stage01 <seed> <env> <steps> [chase_prob]
        chase = float(argv[4]) if len(argv) > 4 else 0.9
        if kind == "stage02":
            from oracle import make_golden_stage02 as mk
            import tests.test_oracle_golden_stage02 as Tk
            rec = mk.run_reference(seed, env_index, steps, seed + 1, 0.02, chase, int(argv[5]) if len(argv) > 5 else None)
            check = Tk.test_stage02_oracle_matches_reference_recording
        else:
            from oracle import make_golden_stage01 as mk
            import tests.test_oracle_golden_stage01 as Tk
            rec = mk.run_reference(seed, env_index, steps, seed + 1, 0.02, chase)
            check = Tk.test_stage01_oracle_matches_reference_recording
        t1 = time.time()
        with tempfile.TemporaryDirectory() as d:
            path = os.path.join(d, kind + "_fuzz.npz")
            np.savez_compressed(path, **rec)
            check(path)
        print(f"OK {kind} seed {seed} env {env_index} steps {steps} chase {chase}: episodes {int(np.sum(rec['done']))}, "
              f"reference {t1 - t0:.0f} s, replay {time.time() - t1:.0f} s", flush=True)
        return
    if kind == "dumb":
        rec = m.run_reference_dumb(seed, env_index, steps, 0.02, int(extra) if extra else None)
    elif kind == "eval":
        rec = m.run_reference_eval2bt(seed, env_index, steps, 0.02, int(extra) if extra else None)
    else:
        kamikaze_after = int(argv[5]) if len(argv) > 5 else None      # the scripted pilot rams from that step on (agent deaths)
        rec = m.run_reference(seed, env_index, steps, seed + 1, 0.02, float(extra) if extra else 0.9, kamikaze_after,
                              "fusion" if kind == "fusion" else "c1")
    t1 = time.time()
    with tempfile.TemporaryDirectory() as d:
        stem = {"dumb": "l5dumb_fuzz", "eval": "l5eval2bt_fuzz", "fusion": "l5fusion_fuzz", "c1": "level5_fuzz"}[kind]
        path = os.path.join(d, stem + ".npz")
        np.savez_compressed(path, **rec)
        make = lambda cfg, s, e: Level5Oracle(cfg, 1, seed=s, env_offset=e)   # noqa: E731
        if kind == "dumb":
            T.replay_dumb(load_recording(path), make)
        elif kind == "eval":
            T.replay_eval2bt(load_recording(path), make)
        else:
            T.test_level5_oracle_matches_reference_recording(path)
    print(f"OK {kind} seed {seed} env {env_index} steps {steps} extra {extra}: episodes {int(np.sum(rec['done']))}, "
          f"reference {t1 - t0:.0f} s, replay {time.time() - t1:.0f} s", flush=True)


if __name__ == "__main__":
    main(sys.argv[1:])

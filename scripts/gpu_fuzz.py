#!/usr/bin/env python
"""The random-parameter parity loop of tests/test_gpu_parity.py::test_random_parameters run by the clock:
CUDA path (C ABI) against the oracle, fresh seed, until the time is up.  Prints one JSON line; on a mismatch the
assertion names the round, so that `seed`/`round` reproduce it.
Usage (under gpurun): python scripts/gpu_fuzz.py <seed> <seconds> > gpurun_out/fuzz_<tag>.json"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import util  # noqa: E402
import test_gpu_parity as tg  # noqa: E402
import oracle  # noqa: E402  (checker)
import defuse_b200  # noqa: E402


def main():
    seed, seconds = int(sys.argv[1]), float(sys.argv[2])
    if not os.path.exists(oracle.PORT_LIB):
        oracle.build()
    ctx = defuse_b200.default_context(0)
    rng = np.random.default_rng(seed)
    t_end = time.time() + seconds
    rounds = tasks = 0
    while time.time() < t_end:
        try:
            tasks += util.check_random_parameter_round(rng, rounds, oracle, ctx, tg._check_split, tg._check_simple)
        except AssertionError as e:
            print(json.dumps({"seed": seed, "failed_round": rounds, "error": str(e)[:2000]}))
            return 1
        rounds += 1
    print(json.dumps({"seed": seed, "seconds": seconds, "rounds": rounds, "tasks_compared": tasks, "mismatches": 0,
                      "device": ctx.device_info()["name"]}))
    return 0


if __name__ == "__main__":
    sys.exit(main())

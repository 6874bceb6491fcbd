"""Golden draws of the reference's subset sampler (celeba19/train.py:87-142: enumerate_combinations + sample_combinations),
produced by the UNMODIFIED reference in the build container:

    python tests/golden/make_sampler_golden.py      # needs /root/reference

For each (seed, size) the numpy GLOBAL generator is seeded (the reference draws from it) and the returned [size, 19] bool
rows are recorded.  tests/test_sampler_golden_cpu.py replays them through the product's O(1) un-ranking sampler and the
oracle's restatement: bit-identical rows are required (same subsets AND same consumption of the RNG stream).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import load_ref  # noqa: E402  (same import shims: xrange, np.int, stub `datasets`)
import types  # noqa: E402

CASES = [(0, 1), (1, 1), (2, 1), (3, 2), (4, 3), (5, 5), (7, 1), (11, 5), (13, 8), (17, 1), (19, 4), (23, 16), (1234, 1),
         (1234, 7), (99991, 32)]


def main():
    cds = types.ModuleType("datasets"); cds.N_ATTRS = 18; cds.CelebAttributes = object
    m = load_ref("celeba19", "model", "ref_celeba19_model_s", {"datasets": cds})
    t = load_ref("celeba19", "train", "ref_celeba19_train_s", {"datasets": cds, "model": m})
    pool = t.enumerate_combinations(19)
    out = {"cases": np.array(CASES, np.int64), "pool_shape": np.array(pool.shape, np.int64),
           "pool_rowsum_hist": np.bincount(pool.sum(1), minlength=20).astype(np.int64)}
    for seed, size in CASES:
        np.random.seed(seed)
        rows = np.asarray(t.sample_combinations(pool, size=size)).astype(bool)
        out[f"draw_{seed}_{size}"] = rows
        out[f"next_{seed}_{size}"] = np.array(np.random.randint(0, 2 ** 31 - 1))   # where the RNG stream stands afterwards
    np.savez_compressed(os.path.join(HERE, "sampler_golden.npz"), **out)
    print("wrote sampler_golden.npz:", {k: v.shape for k, v in out.items() if k.startswith("draw_")})


if __name__ == "__main__":
    main()

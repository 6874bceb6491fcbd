"""The product's subset sampler (trainer_celeba19.sample_combinations: O(1) un-ranking, no 524,267 x 19 pool) and the
oracle's restatement against draws of the UNMODIFIED reference sampler (celeba19/train.py:111-142), recorded by
tests/golden/make_sampler_golden.py: bit-identical rows and the same position in the numpy RNG stream afterwards.
Host-side logic: no GPU needed (importing the sampler does not load the CUDA library)."""
import os

import numpy as np
import pytest

from oracle import celeba19_oracle as O19

G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(os.path.join(G, "sampler_golden.npz")))


def test_pool_the_reference_enumerates(gold):
    # all subsets of sizes 2..18 of 19 modalities (celeba19/train.py:87-108)
    assert tuple(gold["pool_shape"]) == (524267, 19)
    from math import comb
    assert [int(v) for v in gold["pool_rowsum_hist"]] == [0, 0] + [comb(19, k) for k in range(2, 19)] + [0]


@pytest.mark.parametrize("which", ["product", "oracle"])
def test_draws_match_the_reference_bit_for_bit(gold, which):
    if which == "product":
        from multimodal_vae_public_b200.trainer_celeba19 import sample_combinations as draw
    else:
        draw = O19.sample_combinations_fast
    for seed, size in gold["cases"]:
        rs = np.random.RandomState(int(seed))
        rows = np.asarray(draw(19, int(size), rs)).astype(bool)
        ref = gold[f"draw_{seed}_{size}"]
        assert rows.shape == ref.shape and np.array_equal(rows, ref), (which, seed, size)
        assert int(rs.randint(0, 2 ** 31 - 1)) == int(gold[f"next_{seed}_{size}"]), (which, seed, size, "RNG stream position")

"""The C restatement against the committed fixtures produced by the unmodified reference (tests/golden/*.npz)."""
import numpy as np
import pytest

import fixtures


@pytest.mark.parametrize("name", fixtures.NAMES)
def test_oracle_reproduces_reference_fixture(oracle, name):
    fx = fixtures.Fixture(name, oracle)
    res = oracle.dsk(fx.seqs, fx.k, fx.m, fx.repart, fx.nb_partitions, abundance_min=fx.abundance_min, nb_passes=fx.nb_passes)
    assert res["stats"][:4].tolist() == fx.z["stats"].tolist()
    for key in range(fx.nb_partitions * fx.nb_passes):
        for got, want in zip(res["parts"][key], fx.part(key)):
            assert (got == want).all()
        for got, want in zip(res["solid"][key], fx.solid(key)):
            assert (got == want).all()
    assert (res["histogram"] == fx.z["histogram"]).all()
    assert list(oracle.histogram_cutoff(res["histogram"])) == fx.z["cutoff"].tolist()
    # Bloom of the solid k-mers: sizing rule + bytes for the three kinds
    nb_solid = int(fx.z["stats"][3])
    assert list(oracle.bloom_params(fx.k, nb_solid)) == fx.z["bloom_size"].tolist()
    lo = np.concatenate([fx.solid(key)[0] for key in range(fx.nb_partitions * fx.nb_passes)])
    hi = np.concatenate([fx.solid(key)[1] for key in range(fx.nb_partitions * fx.nb_passes)])
    for kind in ("basic", "cache", "neighbor"):
        b, bitsize = oracle.bloom(kind, int(fx.z["bloom_size"][0]), int(fx.z["bloom_size"][1]), fx.k, fx.words, lo, hi if fx.words == 2 else None)
        assert bitsize == int(fx.z["bloom_%s_bitsize" % kind][0])
        assert (b == fx.z["bloom_" + kind]).all()

"""csrc/hostrng.cu against numpy itself: ``bore_b200.hostrng.uniform`` must return the numbers
``RandomState.uniform(low, high, size=(n, dim))`` returns (bore/mixins.py:49 is that call) and leave the generator
where numpy leaves it.  CPU only: the function is host code of the library."""
import numpy as np
import pytest

from bore_b200 import hostrng


def _pair(seed, burn):
    a, b = np.random.RandomState(seed), np.random.RandomState(seed)
    if burn:
        a.random_sample(burn); b.random_sample(burn)
    return a, b


@pytest.mark.parametrize("seed,burn,n,dim", [
    (0, 0, 65536, 50),       # BASELINE.json configs[2]
    (1, 7, 1024, 6),         # below FAST_MIN: numpy's own call
    (2, 311, 4096, 8),       # the state sits in the middle of a block: 311 doubles = 622 words
    (3, 1, 3000, 7),         # odd row length
    (4, 312, 20000, 2),      # pos == 624 exactly on entry
    (5, 100, 16384, 1),
])
def test_same_numbers_and_same_state_as_numpy(seed, burn, n, dim):
    a, b = _pair(seed, burn)
    rs = np.random.RandomState(100 + seed)
    low = rs.uniform(-3.0, 1.0, size=dim)
    high = low + rs.uniform(0.1, 5.0, size=dim)
    want = a.uniform(low=low, high=high, size=(n, dim))
    got = hostrng.uniform(b, low, high, n, dim)
    assert got.dtype == np.float64 and got.shape == (n, dim)
    assert np.array_equal(got, want)
    # the generators continue identically (uniform, integers, and the cached gaussian survives)
    assert np.array_equal(a.uniform(size=33), b.uniform(size=33))
    assert np.array_equal(a.randint(0, 1 << 30, size=17), b.randint(0, 1 << 30, size=17))
    assert np.array_equal(a.normal(size=5), b.normal(size=5))


def test_scalar_bounds_cached_gaussian_and_out_buffer():
    a, b = _pair(11, 0)
    a.normal(size=3); b.normal(size=3)  # odd count: one gaussian stays cached in the state
    want = a.uniform(low=0.0, high=1.0, size=(20000, 3))
    buf = np.empty((20000, 3))
    got = hostrng.uniform(b, 0.0, 1.0, 20000, 3, out=buf)
    assert got is buf and np.array_equal(got, want)
    assert np.array_equal(a.normal(size=4), b.normal(size=4))


def test_other_generators_keep_numpys_call():
    class Mine(np.random.RandomState):
        pass
    a, b = Mine(5), Mine(5)
    want = a.uniform(low=np.zeros(4), high=np.ones(4), size=(8192, 4))
    assert np.array_equal(hostrng.uniform(b, np.zeros(4), np.ones(4), 8192, 4), want)
    assert np.array_equal(a.uniform(size=3), b.uniform(size=3))

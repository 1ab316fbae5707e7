"""SumTree host logic (C heap) against the reference's own known-answer tests
(reagent/test/replay_memory/sum_tree_test.py:30-151) and the MT19937 stream against
Python's `random`.  No GPU needed."""
import random

import numpy as np
import pytest

from reagent_b200.replay_memory import sum_tree


@pytest.fixture
def tree():
    return sum_tree.SumTree(capacity=100)


def test_negative_capacity():
    with pytest.raises(ValueError):
        sum_tree.SumTree(capacity=-1)


def test_set_negative_value(tree):
    with pytest.raises(ValueError):
        tree.set(node_index=0, value=-1)


def test_small_capacity_constructor():
    assert len(sum_tree.SumTree(capacity=1).nodes) == 1
    assert len(sum_tree.SumTree(capacity=2).nodes) == 2


def test_set_value_small_capacity():
    t = sum_tree.SumTree(capacity=1)
    t.set(0, 1.5)
    assert t.get(0) == 1.5


def test_set_value(tree):  # sum_tree_test.py:53-62
    tree.set(node_index=0, value=1.0)
    assert tree.get(0) == 1.0
    for level in tree.nodes:
        assert level[0] == 1.0
        for i in range(1, len(level)):
            assert level[i] == 0.0


def test_capacity_greater_than_requested(tree):
    assert len(tree.nodes[-1]) >= 100


def test_sample_from_empty_tree(tree):
    with pytest.raises(Exception):
        tree.sample()


def test_sample_with_invalid_query_value(tree):
    tree.set(node_index=5, value=1.0)
    with pytest.raises(ValueError):
        tree.sample(query_value=-0.1)
    with pytest.raises(ValueError):
        tree.sample(query_value=1.1)


def test_sample_singleton(tree):
    tree.set(node_index=5, value=1.0)
    assert tree.sample() == 5


def test_sample_pair_with_uneven_probabilities(tree):  # :84-90
    tree.set(node_index=2, value=1.0)
    tree.set(node_index=3, value=3.0)
    for _ in range(200):
        random.seed(1)
        assert tree.sample() == 2


def test_sample_pair_with_query_value(tree):  # :92-97
    tree.set(node_index=2, value=1.0)
    tree.set(node_index=3, value=3.0)
    for _ in range(200):
        assert tree.sample(query_value=0.1) == 2


def test_sampling_with_seed_does_not_affect_future_calls(tree):  # :99-130
    seed = 1
    random.seed(seed)
    r = random.random()
    max_value, delta = 100, 0.01
    total_value = max_value / (1 - r - delta)
    min_value = r * total_value + delta
    tree.set(node_index=2, value=min_value)
    tree.set(node_index=3, value=max_value)
    for _ in range(200):
        random.seed(seed)
        assert tree.sample() == 2
    counts = {2: 0, 3: 0}
    for _ in range(2000):
        counts[tree.sample()] += 1
    assert counts[2] < counts[3]


def test_stratified_sampling_from_empty_tree(tree):
    with pytest.raises(Exception):
        tree.stratified_sample(5)


def test_stratified_sampling(tree):  # :136-143
    k = 32
    for i in range(k):
        tree.set(node_index=i, value=1)
    samples = tree.stratified_sample(k)
    assert samples == list(range(k))


def test_max_recorded_probability(tree):  # :145-151
    k = 32
    tree.set(node_index=0, value=0)
    assert tree.max_recorded_priority == 1
    for i in range(1, k):
        tree.set(node_index=i, value=i)
        assert tree.max_recorded_priority == i


def test_mt_stream_matches_python_random():
    """C MT19937 == CPython's `random`, including state hand-back and block refills."""
    random.seed(1234)
    expect = [random.random() for _ in range(1500)]
    tail = [random.uniform(2.0, 5.0) for _ in range(10)]
    random.seed(1234)
    got = sum_tree.MTStream.draw(700)
    got2 = sum_tree.MTStream.draw(800)
    assert list(got) + list(got2) == expect
    lo, hi = np.full(10, 2.0), np.full(10, 5.0)
    assert list(sum_tree.MTStream.draw(10, lo, hi)) == tail
    # the interpreter's stream continues where C stopped
    random.seed(1234)
    for _ in range(1510):
        random.random()
    a = random.random()
    random.seed(1234)
    sum_tree.MTStream.draw(1510)
    assert random.random() == a


def test_set_batch_matches_sequential_python_loop():
    """History-dependent fp64 delta propagation (sum_tree.py:181-187), batched in C."""
    rng = np.random.RandomState(0)
    cap = 1000
    t = sum_tree.SumTree(cap)
    depth = t.depth
    ref_nodes = [np.zeros(1 << l) for l in range(depth + 1)]
    idx = rng.randint(0, cap, size=5000)
    val = rng.uniform(0, 10, size=5000)
    for i, v in zip(idx, val):
        delta = v - ref_nodes[-1][i]
        n = i
        for lvl in reversed(ref_nodes):
            lvl[n] += delta
            n //= 2
    t.set_batch(idx, val)
    for l in range(depth + 1):
        assert np.array_equal(t.nodes[l], ref_nodes[l])  # bit-exact

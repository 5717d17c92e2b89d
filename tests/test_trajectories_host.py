"""Host logic of the batched trajectories (quantumflow_b200/trajectories.py): the vectorised branch draw is the
same function of numpy's global stream as sequential np.random.choice calls (quantumflow/channels.py:70-77)."""
import numpy as np

from quantumflow_b200.trajectories import choice_indices


def test_vectorised_draw_equals_sequential_numpy_choice():
    rng = np.random.RandomState(5)
    for nb, nk in ((1, 2), (8, 2), (16, 4), (64, 3)):
        probs = rng.random_sample((nb, nk)) + 1e-3
        probs[rng.random_sample((nb, nk)) < 0.2] = 0.0          # zero-probability branches
        probs[:, 0] += 1e-6
        probs /= probs.sum(axis=1, keepdims=True)
        np.random.seed(99)
        want = [np.random.choice(nk, p=p) for p in probs]
        after_seq = np.random.random_sample()
        np.random.seed(99)
        got = choice_indices(probs, np.random.random_sample(nb))
        after_vec = np.random.random_sample()
        assert list(got) == want
        assert after_seq == after_vec                            # the stream is left at the same position


def test_unnormalised_weights_are_normalised_like_kraus_run():
    # Kraus.run divides the branch probabilities by their sum before the draw
    probs = np.array([[0.2, 0.2], [3.0, 1.0]])
    got = choice_indices(probs, np.array([0.49, 0.76]))
    assert list(got) == [0, 1]

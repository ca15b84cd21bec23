"""T5X beam_search(num_decodes=1) vs the greedy loop (SURVEY D7, mt3/models.py:127): where they agree and the two ways
they differ, on crafted logits.  CPU only; the algorithm is restated in oracle/beam_search.py (parity unpinned: T5X is
not installable here)."""
import numpy as np

from oracle import beam_search as BS

V = 6            # ids: 0 pad, 1 EOS, 2.. regular
L = 8


def _table_fn(probs):
    """Step-indexed next-token distributions (independent of the prefix)."""
    t = np.log(np.asarray(probs, np.float64))
    return lambda prefixes, step: np.broadcast_to(t[min(step, len(t) - 1)], prefixes.shape[:2] + (V,))


def _p(**kw):
    """A distribution over V ids: named masses, the remainder spread over the unnamed non-EOS ids."""
    p = np.zeros(V)
    ids = {"pad": 0, "eos": 1, "a": 2, "b": 3, "c": 4, "d": 5}
    for k, v in kw.items():
        p[ids[k]] = v
    rest = [i for i in range(V) if p[i] == 0]
    p[rest] = (1.0 - p.sum()) / len(rest)
    return p


def test_agree_when_eos_is_decisive():
    probs = [_p(a=0.97, eos=0.001), _p(b=0.97, eos=0.001), _p(eos=0.97, c=0.01), _p(c=0.5, eos=0.3)]
    fn = _table_fn(probs)
    g = BS.greedy(fn, 1, L)
    b, _ = BS.beam_search(fn, 1, L)
    np.testing.assert_array_equal(g[0, :4], [2, 3, 1, 0])
    np.testing.assert_array_equal(b, g)


def test_runner_up_eos_wins_on_normalised_score():
    """EOS ranks SECOND at step 1; greedy walks on to a later EOS, beam-1 keeps the short hypothesis because
    (ln .9 + ln .35) / bp(2) = -1.053 beats (ln .9 + ln .6 + ln .5 + ln .9) / bp(4) = -1.109."""
    probs = [_p(a=0.9, eos=0.01), _p(a=0.6, eos=0.35), _p(b=0.5, eos=0.02), _p(eos=0.9, c=0.05), _p(eos=0.9, c=0.05)]
    fn = _table_fn(probs)
    g = BS.greedy(fn, 1, L)
    b, score = BS.beam_search(fn, 1, L)
    np.testing.assert_array_equal(g[0, :5], [2, 2, 3, 1, 0])
    np.testing.assert_array_equal(b[0, :3], [2, 1, 0])
    assert abs(score[0] - (np.log(0.9) + np.log(0.35)) / BS.brevity_penalty(0.6, 2)) < 1e-12


def test_first_ranked_eos_loses_to_a_later_finish():
    """EOS ranks FIRST at step 1 (greedy stops there); the runner-up keeps the beam alive and the longer hypothesis
    wins: (ln .9 + ln .45 + 3 ln .99) / bp(5) = -0.687 beats (ln .9 + ln .5) / bp(2) = -0.728."""
    probs = [_p(a=0.9, eos=0.01), _p(eos=0.5, a=0.45), _p(b=0.99, eos=0.002), _p(c=0.99, eos=0.002), _p(eos=0.99, c=0.002),
             _p(eos=0.99, c=0.002)]
    fn = _table_fn(probs)
    g = BS.greedy(fn, 1, L)
    b, score = BS.beam_search(fn, 1, L)
    np.testing.assert_array_equal(g[0, :3], [2, 1, 0])
    np.testing.assert_array_equal(b[0, :6], [2, 2, 3, 4, 1, 0])
    want = (np.log(0.9) + np.log(0.45) + 2 * np.log(0.99) + np.log(0.99)) / BS.brevity_penalty(0.6, 5)
    assert abs(score[0] - want) < 1e-12


def test_no_eos_returns_the_live_prefix_and_batch_elements_are_independent():
    never = [_p(a=0.7, eos=0.0001)] * 3
    decisive = [_p(a=0.97, eos=0.001), _p(eos=0.97, c=0.01), _p(c=0.5, eos=0.3)]
    t = np.log(np.stack([np.stack(never), np.stack(decisive)]))          # [batch 2, step 3, V]
    fn = lambda prefixes, step: t[:, None, min(step, 2), :]
    b, _ = BS.beam_search(fn, 2, L)
    np.testing.assert_array_equal(b[0], [2] * L)                          # ran to max_decode_len, nothing finished
    np.testing.assert_array_equal(b[1, :3], [2, 1, 0])
    b4, _ = BS.beam_search(fn, 2, L, num_decodes=4)                       # a wider beam finds the same decisive answer
    np.testing.assert_array_equal(b4[1, :3], [2, 1, 0])

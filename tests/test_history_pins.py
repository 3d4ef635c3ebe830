"""Pins of the keras-rl fork's fit() bookkeeping (its source is not in the reference) against the
14 training histories the reference ships: tests/golden/history_pins.npz holds, per agent, the episode lifetimes recovered
from the logged rolling averages and every column the fork derived from them (tests/golden/make_golden_pins.py).

Replayed through the product's host logic -- deepq_decoding_b200.episodes.EpisodeBook (what DQNAgent.fit calls per finished
episode) and agents.LinearAnnealedPolicy -- they must reproduce: episode_lifetimes_rolling_avg, best_rolling_avg, best_episode,
time_since_best, has_succeeded, stopped_improving (and that fit stops exactly at the logged last episode), and the
epsilon of the first training step.  No GPU, no CUDA library: pure host code.
"""
import importlib.util
import json
import os
import sys
import types

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
PINS = np.load(os.path.join(HERE, "golden", "history_pins.npz"))
NAMES = [str(n) for n in PINS["names"]]


def _load(modname):
    """episodes.py / the policy classes without importing torch-heavy siblings: episodes.py is plain numpy."""
    spec = importlib.util.spec_from_file_location(modname, os.path.join(os.path.dirname(HERE), "deepq_decoding_b200", modname + ".py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


episodes = _load("episodes")


def test_fixture_covers_the_14_shipped_agents():
    assert len(NAMES) == 14 and sum(n.startswith("d5_dp_") for n in NAMES) == 6 and sum(n.startswith("d5_x_") for n in NAMES) == 8


@pytest.mark.parametrize("name", NAMES)
def test_episode_bookkeeping_reproduces_shipped_history(name):
    g = lambda k: PINS["%s/%s" % (name, k)]
    L, success, patience, min_steps, _, _, _, max_steps = g("hyper")
    life, nb_steps = g("lifetime"), g("nb_steps")
    book = episodes.EpisodeBook(int(L), success, patience, min_steps)
    n = len(life)
    rolling, best, best_ep, since, succ, stopped = (np.zeros(n), np.zeros(n), np.zeros(n, np.int64), np.zeros(n, np.int64),
                                                    np.zeros(n, bool), np.zeros(n, bool))
    stop_at = None
    for k in range(n):
        e = book.finish_episode(int(life[k]), int(nb_steps[k]))
        assert e["episode"] == k
        rolling[k], best[k], best_ep[k], since[k] = e["episode_lifetimes_rolling_avg"], e["best_rolling_avg"], e["best_episode"], e["time_since_best"]
        succ[k], stopped[k] = e["has_succeeded"], e["stopped_improving"]
        if book.stop and stop_at is None:
            stop_at = k
    assert np.allclose(rolling, g("rolling"), rtol=1e-12, atol=1e-9)
    assert np.allclose(best, g("best_rolling"), rtol=1e-12, atol=1e-9)
    assert np.array_equal(best_ep, g("best_episode")) and np.array_equal(since, g("time_since_best"))
    assert np.array_equal(succ, g("has_succeeded")) and np.array_equal(stopped, g("stopped_improving"))
    # fit() ends after the first episode that sets a flag: either that is the last logged episode, or the run hit nb_steps
    if g("stopped_improving")[-1] or g("has_succeeded")[-1]:
        assert stop_at == n - 1
    else:
        assert stop_at is None and nb_steps[-1] <= max_steps


@pytest.mark.parametrize("name", NAMES)
def test_batched_bookkeeping_equals_the_per_episode_form_on_shipped_history(name):
    """`EpisodeBook.finish_many` (what the vectorised `fit` calls once per drain) over the shipped lifetimes in ragged batches: the same
    entries and the same final state, to the bit, as `finish_episode` one episode at a time."""
    g = lambda k: PINS["%s/%s" % (name, k)]
    L, success, patience, min_steps = g("hyper")[:4]
    life, nb_steps = np.asarray(g("lifetime")), np.asarray(g("nb_steps"))
    one, many = episodes.EpisodeBook(int(L), success, patience, min_steps), episodes.EpisodeBook(int(L), success, patience, min_steps)
    want = [one.finish_episode(int(a), int(b)) for a, b in zip(life, nb_steps)]
    rng, k, got = np.random.default_rng(len(life)), 0, {}
    while k < len(life):
        m = int(rng.integers(0, 700))
        for key, v in many.finish_many(life[k:k + m], nb_steps[k:k + m]).items():
            got.setdefault(key, []).extend(np.asarray(v).tolist())
        k += m
    for key in want[0]:
        assert got[key] == [w[key] for w in want], key
    assert (one.stop, one.episode, one.best_avg, one.best_episode, one.win_n, one.win_sum) == \
           (many.stop, many.episode, many.best_avg, many.best_episode, many.win_n, many.win_sum) and np.array_equal(one.win, many.win)
    assert np.allclose(got["episode_lifetimes_rolling_avg"], g("rolling"), rtol=1e-12, atol=1e-9)


@pytest.mark.parametrize("name", NAMES)
def test_epsilon_schedule_at_first_training_steps(name):
    """loss / mean_q / mean_eps are NaN for every episode that ends by step nb_steps_warmup; the first logged mean_eps is the mean
    of eps(s) over the 0-based step indices s = warmup+1 .. nb_steps-1 of the episode that crosses the warm-up: metrics (and
    training) start at `step > nb_steps_warmup`, and eps follows LinearAnnealedPolicy exactly (SPTS:110-114, :119-127)."""
    from deepq_decoding_b200.agents import LinearAnnealedPolicy, EpsGreedyQPolicy
    g = lambda k: PINS["%s/%s" % (name, k)]
    _, _, _, expl, max_eps, final_eps, warmup, _ = g("hyper")
    first, nb_at_first, nb_before, eps_logged = g("first_eps")
    pol = LinearAnnealedPolicy(EpsGreedyQPolicy(), "eps", max_eps, final_eps, 0.0, expl)
    assert nb_before <= warmup + 1 and nb_at_first > warmup + 1, "the first episode with metrics is the one that crosses the warm-up"
    steps = np.arange(int(warmup) + 1, int(nb_at_first))
    mean_eps = float(np.mean([pol.value(int(s)) for s in steps]))
    assert abs(mean_eps - eps_logged) < 1e-12, (mean_eps, eps_logged)
    assert pol.value(0) == max_eps and pol.value(10 ** 9) == final_eps and pol.value(5, training=False) == 0.0


@pytest.mark.reference
def test_fixture_matches_the_reference_files():
    """The committed fixture is what make_golden_pins.py derives from /root/reference today."""
    import glob
    files = sorted(glob.glob("/root/reference/trained_models/*/*/training_history.json"))
    assert len(files) == 14
    for f in files:
        model, rate = os.path.dirname(f).split("/")[-2:]
        name = "%s_%s" % (model, rate)
        h = json.load(open(f))
        assert np.array_equal(PINS[name + "/rolling"], np.array(h["episode_lifetimes_rolling_avg"]))
        assert np.array_equal(PINS[name + "/best_episode"], np.array(h["best_episode"]))
        assert np.array_equal(PINS[name + "/nb_steps"], np.array(h["nb_steps"]))
        assert set(h) == {"loss", "mean_q", "mean_eps", "episode_reward", "nb_episode_steps", "nb_steps", "episode_lifetimes_rolling_avg",
                          "best_rolling_avg", "best_episode", "time_since_best", "has_succeeded", "stopped_improving", "episode", "duration"}

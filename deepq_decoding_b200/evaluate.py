"""Evaluation sweep of a trained agent over physical error rates (logical lifetime / LER vs p).

Mirrors the tail of the reference's training driver (cluster_scripts/d5_dp/0.001/Single_Point_Training_Script.py:187-222):
for p = 0.001, 0.002, ... set env.p_phys = env.p_meas = p, run dqn.test, record the final cumulative mean lifetime under
the key str(p)[:5]; stop once the lifetime falls below the single-faulty-qubit threshold 1/p (or after `num_to_test`
points).  The returned dict has the layout of the reference's all_results.p; `detailed` holds the per-point cumulative
means (detailed_results/results_<p>.p).  With a process group the per-rank estimates are merged (parallel.reduce_lifetimes).
"""
from . import parallel


def test_sweep(dqn, env, nb_test_episodes, num_to_test=20, step=0.001, trained_at=None, stop_below_threshold=True, verbose=0):
    all_results, detailed, stats = {}, {}, {}
    for count in range(num_to_test):
        err_rate = (count + 1) * step
        env.p_phys = err_rate
        env.p_meas = err_rate
        key = str(err_rate)[:5]
        hist = dqn.test(env, nb_episodes=nb_test_episodes, visualize=False, verbose=verbose, interval=10, single_cycle=False).history
        mean, se, n = parallel.reduce_lifetimes(hist["episode_lifetime"])
        all_results[key] = mean
        detailed[key] = hist["episode_lifetimes_rolling_avg"]
        stats[key] = {"mean_lifetime": mean, "standard_error": se, "episodes": n, "logical_error_rate": 1.0 / mean if mean > 0 else float("inf"),
                      "threshold": 1.0 / err_rate}
        if stop_below_threshold and mean < 1.0 / err_rate:
            break
    return all_results, detailed, stats

"""Per-episode bookkeeping of the keras-rl fork's `fit` loop (host logic, no device code).

The fork's source is not in the reference repo; the rules below are the ones the 14 shipped
`trained_models/**/training_history.json` files obey exactly (replayed by
tests/test_history_pins.py):

  * `episode_lifetimes_rolling_avg` = mean of the last `episode_averaging_length` episode lifetimes
    (of all of them while fewer have finished);
  * `best_rolling_avg` / `best_episode` are updated from episode 0 on, on a strict improvement;
  * `time_since_best` = episode - best_episode;
  * `has_succeeded`     = rolling > success_threshold;
  * `stopped_improving` = (episode - best_episode > stopping_patience) and nb_steps >= min_nb_steps;
    `fit` ends after the first episode for which either flag is set.

Call sites in the reference: cluster_scripts/d5_dp/0.001/Single_Point_Training_Script.py:138-152
(`episode_averaging_length`, `success_threshold`, `stopping_patience`, `min_nb_steps`).
"""
import numpy as np


class EpisodeBook:
    def __init__(self, episode_averaging_length=1000, success_threshold=1e5, stopping_patience=1e9, min_nb_steps=0):
        self.L = max(1, int(episode_averaging_length))
        self.success_threshold, self.stopping_patience, self.min_nb_steps = success_threshold, stopping_patience, min_nb_steps
        self.win = np.zeros(self.L)          # ring of the last L lifetimes + running sum: O(1) rolling mean
        self.win_sum, self.win_n = 0.0, 0
        self.best_avg, self.best_episode, self.episode = -np.inf, 0, 0
        self.stop = False

    @property
    def rolling(self):
        return self.win_sum / self.win_n if self.win_n else float("nan")

    def finish_episode(self, lifetime, nb_steps):
        """Record one finished episode; returns the fork's per-episode history entries."""
        slot = self.episode % self.L
        self.win_sum += float(lifetime) - self.win[slot]
        self.win[slot] = lifetime
        self.win_n = min(self.win_n + 1, self.L)
        if self.win_n == self.L and slot == self.L - 1:      # once per pass over the ring: re-sum, so rounding does not accumulate
            self.win_sum = float(self.win.sum())
        rolling = self.win_sum / self.win_n
        if rolling > self.best_avg:
            self.best_avg, self.best_episode = rolling, self.episode
        succeeded = rolling > self.success_threshold
        stopped = (self.episode - self.best_episode > self.stopping_patience) and nb_steps >= self.min_nb_steps
        out = dict(episode_lifetimes_rolling_avg=rolling, best_rolling_avg=self.best_avg, best_episode=self.best_episode,
                   time_since_best=self.episode - self.best_episode, has_succeeded=bool(succeeded),
                   stopped_improving=bool(stopped), episode=self.episode)
        self.episode += 1
        self.stop = self.stop or succeeded or stopped
        return out

    def finish_many(self, lifetimes, nb_steps):
        """`finish_episode` for a batch of episodes in order, vectorised (a vectorised `fit` finishes thousands of episodes per drain;
        one Python call per episode was a third of its wall time).  Same entries, as arrays: lifetimes are integers, so every window
        sum is exact in float64 whatever the order of the additions, and the two forms agree to the bit
        (tests/test_history_pins.py replays the shipped histories through both)."""
        l = np.asarray(lifetimes, dtype=np.float64).reshape(-1)
        s = np.asarray(nb_steps).reshape(-1)
        m = l.size
        if m == 0:
            return {k: np.zeros(0) for k in ("episode_lifetimes_rolling_avg", "best_rolling_avg", "best_episode", "time_since_best",
                                             "has_succeeded", "stopped_improving", "episode")}
        L, e0 = self.L, self.episode
        p = min(e0, L)                                            # earlier lifetimes still inside some new window, oldest first
        prev = self.win[(np.arange(e0 - p, e0)) % L] if p else np.zeros(0)
        x = np.concatenate([prev, l])
        c = np.concatenate([[0.0], np.cumsum(x)])                 # c[t] = sum of x[:t]
        e = e0 + np.arange(m)
        n = np.minimum(e + 1, L)
        hi = p + np.arange(m) + 1
        rolling = (c[hi] - c[hi - n]) / n
        run = np.maximum.accumulate(np.concatenate([[self.best_avg], rolling]))
        improved = rolling > run[:-1]                             # strict improvement over everything before it
        best_avg = run[1:]
        best_ep = np.maximum.accumulate(np.where(improved, e, -1))
        best_ep = np.where(best_ep < 0, self.best_episode, best_ep)
        since = e - best_ep
        succeeded = rolling > self.success_threshold
        stopped = (since > self.stopping_patience) & (s >= self.min_nb_steps)
        # state after the batch
        keep = min(L, p + m)
        tail = x[-keep:]
        self.win[(np.arange(e0 + m - keep, e0 + m)) % L] = tail
        self.win_n = min(e0 + m, L)
        self.win_sum = float(tail[-self.win_n:].sum())
        self.best_avg, self.best_episode = float(best_avg[-1]), int(best_ep[-1])
        self.episode = e0 + m
        self.stop = bool(self.stop or succeeded.any() or stopped.any())
        return dict(episode_lifetimes_rolling_avg=rolling, best_rolling_avg=best_avg, best_episode=best_ep, time_since_best=since,
                    has_succeeded=succeeded, stopped_improving=stopped, episode=e)

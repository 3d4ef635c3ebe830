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

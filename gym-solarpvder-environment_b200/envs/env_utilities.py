"""Per-episode statistics of the single-env facade (reference gym_PVDER/envs/env_utilities.py:
counters 12-52; the matplotlib views 54-101 are host cosmetics and only the pprint paths exist)."""
import pprint


class Utilities:
    pp = pprint.PrettyPrinter(indent=4)

    def initialize_stats(self):
        self._action_stats = {i: 0 for i in range(self.action_space.n)}
        self._reward_stats = {"step": [], "reward": []}
        self._time_stats = {"step": [], "step_time": []}

    def update_action_stats(self, action):
        self._action_stats[action] = self._action_stats[action] + 1

    def update_reward_stats(self):
        self._reward_stats["step"].append(self._steps)
        self._reward_stats["reward"].append(self._reward)

    def update_time_stats(self):
        self._time_stats["step"].append(self._steps)
        self._time_stats["step_time"].append(self._step_time)

    @property
    def steps(self):
        return self._steps

    def show_step_time(self):
        print("Time for step:{:.3f}".format(self._step_time))

    def show_action_stats(self, SHOW_PLOT=False):
        self.pp.pprint(self._action_stats)

    def show_reward_stats(self, SHOW_PLOT=False):
        self.pp.pprint(self._reward_stats)

    def show_time_stats(self, SHOW_PLOT=False):
        self.pp.pprint(self._time_stats)

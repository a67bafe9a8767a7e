"""Stop-watch used by Kriging.optimize for its `optiTimer: ... s` line.

Interface of the reference's myutils/time_track.py:18-38 (``tic`` / ``toc(print_it)`` / ``print_time`` /
``get_time``; constructing the object starts the clock), implemented on ``time.perf_counter``.
"""
import time


class TimeTrack:

    def __init__(self, name='defaultTimer'):
        self._name = name
        self._t0 = time.perf_counter()
        self._lap = None  # seconds measured by the last toc()

    def tic(self):
        """(re)start the clock"""
        self._t0 = time.perf_counter()
        self._lap = None

    def toc(self, print_it=False):
        """seconds since construction / the last tic(); optionally prints them like the reference does"""
        self._lap = time.perf_counter() - self._t0
        if print_it:
            self.print_time()
        return self._lap

    def print_time(self):
        lap = self._lap if self._lap is not None else self.get_time()
        print('{}: {} s'.format(self._name, lap))

    def get_time(self):
        """running time without stopping"""
        return time.perf_counter() - self._t0

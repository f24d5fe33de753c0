"""Named, nestable wall-clock timers with the reference's interface
(pyro/util/profile_pyro.py:15-135: TimerCollection.timer(name) -> Timer.begin()/end(), report()).
Device work is asynchronous, so ``end()`` optionally synchronises the current CUDA stream
(``TimerCollection(sync=True)``) to attribute time to the right timer."""
import time


class TimerCollection:
    def __init__(self, sync=False):
        self.timers = []
        self.sync = sync

    def timer(self, name):
        for t in self.timers:
            if t.name == name:
                return t
        stack_count = sum(1 for t in self.timers if t.is_running)
        t = Timer(name, stack_count=stack_count, sync=self.sync)
        self.timers.append(t)
        return t

    def report(self):
        spacing = "   "
        for t in self.timers:
            print(t.stack_count * spacing + t.name + ": ", t.elapsed_time)


class Timer:
    def __init__(self, name, stack_count=0, sync=False):
        self.name = name
        self.stack_count = stack_count
        self.is_running = False
        self.start_time = 0.0
        self.elapsed_time = 0.0
        self.sync = sync

    def _sync(self):
        if self.sync:
            import torch
            if torch.cuda.is_available():
                torch.cuda.current_stream().synchronize()

    def begin(self):
        self._sync()
        self.start_time = time.time()
        self.is_running = True

    def end(self):
        self._sync()
        self.elapsed_time += time.time() - self.start_time
        self.is_running = False

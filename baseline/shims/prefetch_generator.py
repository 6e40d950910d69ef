"""Import shim (harness, NOT product code) for the third-party `prefetch_generator` package the reference imports
(/root/reference/lib/runtime/__init__.py:8) and that cannot be installed offline: a background thread that pulls items of
a generator into a bounded queue."""
import queue
import threading


class BackgroundGenerator(threading.Thread):
    def __init__(self, generator, max_prefetch=1):
        super().__init__(daemon=True)
        self.queue = queue.Queue(max_prefetch)
        self.generator = generator
        self.start()

    def run(self):
        for item in self.generator:
            self.queue.put(item)
        self.queue.put(None)

    def __iter__(self):
        return self

    def __next__(self):
        item = self.queue.get()
        if item is None:
            raise StopIteration
        return item

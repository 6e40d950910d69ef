class Server:  # single-node in-memory stand-in: enough for `import lib`
    def __init__(self, *args, **kwargs):
        self.storage = {}

    async def listen(self, port):
        return None

    async def bootstrap(self, peers):
        return None

    async def get(self, key):
        return self.storage.get(key)

    async def set(self, key, value):
        self.storage[key] = value
        return True

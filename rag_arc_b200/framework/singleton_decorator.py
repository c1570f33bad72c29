"""Process-wide singleton decorator (same contract as /root/reference
framework/singleton_decorator.py:1-6: first call constructs, later calls return that object)."""
import functools


def singleton(cls):
    cache = {}

    @functools.wraps(cls, updated=())
    def get(*args, **kwargs):
        try:
            return cache[cls]
        except KeyError:
            cache[cls] = cls(*args, **kwargs)
            return cache[cls]

    get._instances = cache      # test hook: lets a test drop the instance
    return get

import importlib
import tempfile
import time


class Stopwatch(object):
    def __init__(self):
        self._t0 = time.time()
        self.elapsed = 0.0

    def since(self):
        return time.time() - self._t0

    def __enter__(self):
        self._t0 = time.time()
        return self

    def __exit__(self, *a):
        self.elapsed = time.time() - self._t0
        return False


def tolist(x):
    if isinstance(x, (list, tuple)):
        return list(x)
    try:
        import numpy as np
        if isinstance(x, np.ndarray):
            return x.tolist() if x.ndim > 0 else [x.item()]
    except ImportError:
        pass
    return [x]


def try_import(pkg, name=None):
    try:
        importlib.import_module(pkg)
    except ImportError:
        raise ImportError('optional package "%s" not installed' % (name or pkg))


def groupbyasdict(items, f):
    d = {}
    for it in items:
        d.setdefault(f(it), []).append(it)
    return d


def flatlist(ll):
    return [x for l in ll for x in l]


def tempdir():
    return tempfile.gettempdir()


def save(obj, path):
    import pickle
    with open(path, 'wb') as f:
        pickle.dump(obj, f)
    return path


def load(path):
    import pickle
    with open(path, 'rb') as f:
        return pickle.load(f)

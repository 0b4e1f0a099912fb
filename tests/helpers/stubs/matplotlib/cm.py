from . import _Nothing


def __getattr__(name):
    return _Nothing()

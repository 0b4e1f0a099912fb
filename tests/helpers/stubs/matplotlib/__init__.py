"""Stand-in for matplotlib, which this image does not ship: the reference's tests_py scripts import it at module level
for plotting helpers (make_plot / makePlot) that the tests themselves never call.  Any attribute access works and
returns a do-nothing callable, so an accidental call would not crash but also cannot affect a result."""


class _Nothing:
    def __getattr__(self, name):
        return _Nothing()

    def __call__(self, *a, **k):
        return _Nothing()


def __getattr__(name):
    return _Nothing()

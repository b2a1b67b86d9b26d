import contextlib


@contextlib.contextmanager
def TraceAnnotation(name):
    yield

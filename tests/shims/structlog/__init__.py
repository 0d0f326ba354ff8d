"""Minimal structlog stand-in (tests only): key-value events forwarded to the stdlib logger."""
import logging

from . import dev, processors, stdlib, typing  # noqa: F401


class _Logger:
    def __init__(self, name):
        self._l = logging.getLogger(name or "semcode")

    def _emit(self, level, event, kw):
        self._l.log(level, "%s%s", event, "".join(f" {k}={v}" for k, v in kw.items()))

    def debug(self, event, **kw):
        self._emit(logging.DEBUG, event, kw)

    def info(self, event, **kw):
        self._emit(logging.INFO, event, kw)

    def warning(self, event, **kw):
        self._emit(logging.WARNING, event, kw)

    def error(self, event, **kw):
        self._emit(logging.ERROR, event, kw)

    def exception(self, event, **kw):
        self._emit(logging.ERROR, event, kw)

    def bind(self, **kw):
        return self


def get_logger(name=None):
    return _Logger(name)


def configure(**kwargs):
    return None


def make_filtering_bound_logger(min_level):
    return _Logger

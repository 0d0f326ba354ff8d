class BoundLogger:
    pass


class ProcessorFormatter:
    wrap_for_formatter = staticmethod(lambda logger, name, event_dict: event_dict)

    def __init__(self, processor=None, foreign_pre_chain=None):
        import logging

        self._f = logging.Formatter()

    def format(self, record):
        return self._f.format(record)


class LoggerFactory:
    def __call__(self, *args):
        import logging

        return logging.getLogger(*args)


class PositionalArgumentsFormatter:
    def __call__(self, logger, name, event_dict):
        return event_dict


def add_log_level(logger, name, event_dict):
    return event_dict


def add_logger_name(logger, name, event_dict):
    return event_dict

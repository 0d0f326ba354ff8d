class StackInfoRenderer:
    def __call__(self, logger, name, event_dict):
        return event_dict


class TimeStamper:
    def __init__(self, fmt=None):
        self.fmt = fmt

    def __call__(self, logger, name, event_dict):
        return event_dict


def format_exc_info(logger, name, event_dict):
    return event_dict

class ConsoleRenderer:
    def __init__(self, colors=False):
        self.colors = colors

    def __call__(self, logger, name, event_dict):
        return str(event_dict)

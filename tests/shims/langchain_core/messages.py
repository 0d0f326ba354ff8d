class _Message:
    def __init__(self, content=""):
        self.content = content


class HumanMessage(_Message):
    pass


class SystemMessage(_Message):
    pass

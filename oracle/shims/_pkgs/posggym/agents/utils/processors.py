class Processor:  # name only
    pass


class FlattenProcessor(Processor):
    def __init__(self, *a, **k):
        pass

class progress_bar:  # noqa: N801  (name fixed by the package being stood in for)
    def __init__(self, gen, total=None, display=True, **_):
        self._gen, self.total, self.display, self.comment = gen, total, display, ""

    def __iter__(self):
        return iter(self._gen)

    def update(self, *_):
        pass

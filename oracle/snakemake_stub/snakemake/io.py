class Namedlist(list):
    """list with attribute / string-key access to named items (snakemake.io.Namedlist)."""
    def __init__(self):
        super().__init__()
        self._names = {}

    def _add_name(self, name):
        self._names[name] = len(self) - 1
        setattr(self, name, self[-1])

    def __getitem__(self, key):
        if isinstance(key, str):
            return list.__getitem__(self, self._names[key])
        return list.__getitem__(self, key)

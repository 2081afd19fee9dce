"""How the pipeline hands a rule's variables to an external script: a JSON file named by the
environment variable SNAKEMAKE_PARAMS holding `input`, `output`, `wildcards`, `params` as lists of
[name-or-null, value] pairs and `threads` as an integer (tailseeker/powersnake.py:55-107).  This
module reads that file for the drop-in scripts of this directory without importing snakemake."""
from __future__ import annotations

import json
import os

ENVVAR = "SNAKEMAKE_PARAMS"


class Named(list):
    """A list whose named items are also attributes / string keys (snakemake.io.Namedlist as the
    reference's scripts use it: `input.signals`, `params['conf']`, `for f in input`)."""

    def __init__(self, pairs=()):
        super().__init__()
        self._names = {}
        for k, v in pairs:
            self.append(v)
            if k is not None:
                self._names.setdefault(k, []).append(len(self) - 1)

    def _get(self, name):
        idx = self._names[name]
        return self[idx[0]] if len(idx) == 1 else [self[i] for i in idx]

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        try:
            return self._get(name)
        except KeyError:
            raise AttributeError(name) from None

    def __getitem__(self, key):
        if isinstance(key, str):
            return self._get(key)
        return super().__getitem__(key)


def load():
    """{input, output, wildcards, params: Named; threads: int}.  The variable is removed from the
    environment so that child processes do not see it (powersnake.py:71)."""
    if ENVVAR not in os.environ:
        raise SystemExit("this script is run by the pipeline: %s is not set" % ENVVAR)
    with open(os.environ.pop(ENVVAR)) as f:
        options = json.load(f)
    out = {}
    for name, value in options.items():
        out[name] = value if isinstance(value, int) else Named(value)
    return out

"""stand-in: utils/parse.py:3 imports PatchCollection at module level (only its plotting helpers use it)"""


class PatchCollection:
    def __init__(self, *a, **k):
        raise RuntimeError("matplotlib shim: plotting is not available")

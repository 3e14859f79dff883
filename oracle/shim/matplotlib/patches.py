"""stand-in: utils/parse.py:2 imports Polygon at module level (only its plotting helpers use it)"""


class Polygon:
    def __init__(self, *a, **k):
        raise RuntimeError("matplotlib shim: plotting is not available")

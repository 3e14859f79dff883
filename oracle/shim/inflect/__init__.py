def engine():
    raise RuntimeError("inflect shim")

class FmIndexError(RuntimeError):
    pass


class FmIndex:  # placeholder, replaced below
    pass

class transformation(object):
    """Base class for transformations: transform(Input) -> (mu, cov|None)  (src/transformation/base.py:4-10)."""

    def __init__(self, params=None):
        self.params = params

    def transform(self, X_prev):
        raise NotImplementedError

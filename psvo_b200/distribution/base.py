class distribution(object):
    """Base class for distributions (src/distribution/base.py:4-22)."""

    def __init__(self, params=None):
        self.params = params

    def sample(self, Input):
        raise NotImplementedError

    def sample_and_log_prob(self, Input):
        raise NotImplementedError

    def log_prob(self, Input, output):
        raise NotImplementedError

    def mean(self, Input, output=None):
        raise NotImplementedError

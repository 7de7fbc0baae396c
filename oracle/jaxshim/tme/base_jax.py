def mean_and_cov(*_a, **_k):
    raise NotImplementedError('tme is not available (un-vendored third-party dependency, out of scope)')

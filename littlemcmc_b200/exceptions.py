"""Exceptions (mirror of reference littlemcmc/exceptions.py:22)."""


class SamplingError(RuntimeError):
    pass

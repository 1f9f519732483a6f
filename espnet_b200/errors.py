"""Exceptions mirrored from the reference."""


class TooShortUttError(Exception):
    """espnet2/legacy/nets/pytorch_backend/transformer/subsampling.py:14-28."""

    def __init__(self, message, actual_size, limit):
        super().__init__(message)
        self.actual_size = actual_size
        self.limit = limit

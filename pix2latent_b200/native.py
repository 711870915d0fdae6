"""Opaque-handle part of the C-ABI (include/p2l.h): declared here, used by model/ and loss."""


def declare(L):
    pass

"""Stub for the optional `termcolor` dependency of the reference's display module
(pygradflow/display.py:5).  Used only by tests/golden/make_golden.py in the build container."""


def colored(text, *args, **kwargs):
    return text

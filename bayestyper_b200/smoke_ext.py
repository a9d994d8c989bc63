"""Smoke checks for the later pipeline stages (filled in as they land)."""


def run(lib) -> None:
    pass

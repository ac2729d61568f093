"""Import-only stand-in for SUMO's `sumolib` (absent)."""


def checkBinary(name):
    return name

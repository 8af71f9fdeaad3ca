"""Stand-in module: the reference imports it at module scope but the solve path never calls it."""

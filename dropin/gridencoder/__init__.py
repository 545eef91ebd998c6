from .grid import GridEncoder, VarGridEncoder, grid_encode  # noqa: F401

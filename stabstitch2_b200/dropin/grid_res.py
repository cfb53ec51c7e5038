from stabstitch2_b200.grid_res import GRID_H, GRID_W  # noqa: F401

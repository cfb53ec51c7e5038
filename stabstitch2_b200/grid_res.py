"""grid_res.py:3-4 of the reference: the mesh is (6+1)x(8+1) control points."""
GRID_H = 6
GRID_W = 8

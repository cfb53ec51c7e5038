from stabstitch2_b200.spatial_network import H2Mesh, get_rigid_mesh, get_norm_mesh, build_SpatialNet, SpatialNet  # noqa: F401

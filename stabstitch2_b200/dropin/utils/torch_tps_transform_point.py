from stabstitch2_b200.utils.torch_tps_transform_point import transformer  # noqa: F401

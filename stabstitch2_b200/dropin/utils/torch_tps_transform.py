from stabstitch2_b200.utils.torch_tps_transform import transformer  # noqa: F401

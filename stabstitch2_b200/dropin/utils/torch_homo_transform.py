from stabstitch2_b200.utils.torch_homo_transform import transformer  # noqa: F401

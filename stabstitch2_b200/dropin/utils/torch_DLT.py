from stabstitch2_b200.utils.torch_DLT import tensor_DLT  # noqa: F401

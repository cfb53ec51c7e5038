from stabstitch2_b200.smooth_network import build_SmoothNet, SmoothNet, MotionPrediction  # noqa: F401

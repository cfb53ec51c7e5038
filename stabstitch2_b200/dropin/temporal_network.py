from stabstitch2_b200.temporal_network import build_TemporalNet, TemporalNet  # noqa: F401

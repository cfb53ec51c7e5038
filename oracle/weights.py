"""Seeded synthetic checkpoints with the reference's state-dict key names and shapes.

TEST INFRASTRUCTURE (see stabstitch_oracle.py header).  The trained weights are not in the
reference repo (Google-Drive links only, Full_model_inference/README.md:2-4) and there is no
network, so both the oracle and the CUDA path load the SAME seeded dicts made here.  Key
lists follow SURVEY.md Appendix B and are verified against the live reference modules'
`load_state_dict(strict=True)` in tests/golden/make_golden.py:132-134 (run in the build container).
"""
import math

import torch


def _conv(g, cout, cin, *k, gain=2.0):
    fan_in = cin
    for v in k:
        fan_in *= v
    return torch.randn(cout, cin, *k, generator=g) * math.sqrt(gain / fan_in)


def _bn(g, sd, pfx, c, gamma_scale=1.0):
    # non-trivial eval-mode statistics so that BN folding is really exercised
    sd[pfx + ".weight"] = (0.6 + 0.4 * torch.rand(c, generator=g)) * gamma_scale
    sd[pfx + ".bias"] = 0.05 * torch.randn(c, generator=g)
    sd[pfx + ".running_mean"] = 0.05 * torch.randn(c, generator=g)
    sd[pfx + ".running_var"] = 0.7 + 0.6 * torch.rand(c, generator=g)
    sd[pfx + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)


def _block(g, sd, pfx, cin, cout, down):
    sd[pfx + ".conv1.weight"] = _conv(g, cout, cin, 3, 3)
    _bn(g, sd, pfx + ".bn1", cout)
    sd[pfx + ".conv2.weight"] = _conv(g, cout, cout, 3, 3)
    _bn(g, sd, pfx + ".bn2", cout, gamma_scale=0.5)
    if down:
        sd[pfx + ".downsample.0.weight"] = _conv(g, cout, cin, 1, 1, gain=1.0)
        _bn(g, sd, pfx + ".downsample.1", cout, gamma_scale=0.7)


def _resnet(g, sd):
    p1, p2 = "feature_extractor_stage1", "feature_extractor_stage2"
    sd[p1 + ".0.weight"] = _conv(g, 64, 3, 7, 7)
    _bn(g, sd, p1 + ".1", 64)
    _block(g, sd, p1 + ".4.0", 64, 64, False)
    _block(g, sd, p1 + ".4.1", 64, 64, False)
    _block(g, sd, p1 + ".5.0", 64, 128, True)
    _block(g, sd, p1 + ".5.1", 128, 128, False)
    _block(g, sd, p2 + ".0.0", 128, 256, True)
    _block(g, sd, p2 + ".0.1", 256, 256, False)


def _linear(g, sd, pfx, cout, cin, wscale=1.0):
    sd[pfx + ".weight"] = torch.randn(cout, cin, generator=g) * math.sqrt(1.0 / cin) * wscale
    sd[pfx + ".bias"] = 0.01 * torch.randn(cout, generator=g)


def _regress2(g, sd, p1, p2, cin, out_scale):
    chans = [(cin, 64), (64, 64), (64, 128), (128, 128), (128, 128), (128, 128), (128, 256), (256, 256)]
    for idx, (ci, co) in zip((0, 2, 5, 7, 10, 12, 15, 17), chans):
        sd["%s.%d.weight" % (p1, idx)] = _conv(g, co, ci, 3, 3)
    _linear(g, sd, p2 + ".0", 1024, 1536)
    _linear(g, sd, p2 + ".2", 512, 1024)
    _linear(g, sd, p2 + ".4", 126, 512, wscale=out_scale)


def spatial_state_dict(seed=0, bias_x=-170.0, mesh_scale=1.0):
    """SpatialNet (spatial_network.py:144-271).  `bias_x` is SURVEY.md 8d's realism bias:
    x-offset of the four corners in 480-px units, so view 2 lands ~35% to the right."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    chans = [(2, 64), (64, 64), (64, 128), (128, 128), (128, 128), (128, 128)]
    for idx, (ci, co) in zip((0, 2, 5, 7, 10, 12), chans):
        sd["regressNet1_part1.%d.weight" % idx] = _conv(g, co, ci, 3, 3)
    _linear(g, sd, "regressNet1_part2.0", 512, 768)
    _linear(g, sd, "regressNet1_part2.2", 128, 512)
    _linear(g, sd, "regressNet1_part2.4", 8, 128)
    sd["regressNet1_part2.4.bias"][0::2] = bias_x
    _regress2(g, sd, "regressNet2_part1_ref", "regressNet2_part2_ref", 121, mesh_scale)
    _regress2(g, sd, "regressNet2_part1_tgt", "regressNet2_part2_tgt", 121, mesh_scale)
    _resnet(g, sd)
    return sd


def temporal_state_dict(seed=1, mesh_scale=1.0):
    """TemporalNet (temporal_network.py:62-116); feature_extractor_stage2 is present in the
    checkpoint but unused by forward."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    _regress2(g, sd, "regressNet2_part1", "regressNet2_part2", 49, mesh_scale)
    _resnet(g, sd)
    return sd


def smooth_state_dict(seed=2):
    """SmoothNet (smooth_network.py:106-136); embedding2 is present but unused."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    p = "MotionPre."
    # inputs are un-normalised pixel coordinates (up to ~500): keep the embeddings tame
    sd[p + "embedding1.0.weight"] = torch.randn(32, 2, generator=g) * 0.004
    sd[p + "embedding1.0.bias"] = 0.1 * torch.randn(32, generator=g)
    sd[p + "embedding2.0.weight"] = torch.randn(8, 1, generator=g)
    sd[p + "embedding2.0.bias"] = 0.1 * torch.randn(8, generator=g)
    sd[p + "embedding3.0.weight"] = torch.randn(32, 2, generator=g) * 0.2
    sd[p + "embedding3.0.bias"] = 0.1 * torch.randn(32, generator=g)
    for i in (0, 2, 4):
        sd[p + "MotionConv3D.%d.weight" % i] = _conv(g, 128, 128, 5, 3, 3)
        sd[p + "MotionConv3D.%d.bias" % i] = 0.01 * torch.randn(128, generator=g)
    sd[p + "decoding.0.weight"] = torch.randn(4, 128, generator=g) * math.sqrt(1.0 / 128)
    sd[p + "decoding.0.bias"] = 0.01 * torch.randn(4, generator=g)
    return sd

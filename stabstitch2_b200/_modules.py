"""Parameter containers with the reference's state-dict key names.

The nn.Module objects below hold parameters only (so .cuda(), .eval(), .state_dict(),
.load_state_dict(strict=True) behave like the reference's); no torch op runs in forward -
the weights are handed to libss2 (ss2_load_tensor / ss2_finalize_weights) and re-synced
whenever any parameter tensor changes."""
import torch
import torch.nn as nn


class _BasicBlock(nn.Module):
    """torchvision BasicBlock naming: conv1,bn1,conv2,bn2,downsample.{0,1}."""

    def __init__(self, cin, cout, stride):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(cout)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(cout)
        self.downsample = None
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False), nn.BatchNorm2d(cout))


def _layer(cin, cout, stride):
    return nn.Sequential(_BasicBlock(cin, cout, stride), _BasicBlock(cout, cout, 1))


def resnet18_feature_extractors():
    """spatial_network.py:123-139 get_res18_FeatureMap: (stage1, stage2) with the same
    Sequential indices: 0 conv1, 1 bn1, 2 relu, 3 maxpool, 4 layer1, 5 layer2 | 0 layer3."""
    stage1 = nn.Sequential(nn.Conv2d(3, 64, 7, 2, 3, bias=False), nn.BatchNorm2d(64), nn.ReLU(inplace=True),
                           nn.MaxPool2d(3, 2, 1), _layer(64, 64, 1), _layer(64, 128, 2))
    stage2 = nn.Sequential(_layer(128, 256, 2))
    return stage1, stage2


def regress_convs(cin, widths):
    """Conv3x3(no bias)+ReLU pairs with MaxPool2d(2,2) after every second conv
    (spatial_network.py:147-168,181-209): Sequential indices 0,2,5,7,10,12[,15,17]."""
    layers = []
    c = cin
    for i, w in enumerate(widths):
        layers += [nn.Conv2d(c, w, 3, padding=1, bias=False), nn.ReLU(inplace=True)]
        if i % 2 == 1:
            layers.append(nn.MaxPool2d(2, 2))
        c = w
    return nn.Sequential(*layers)


def regress_fc(fin, h1, h2, fout):
    return nn.Sequential(nn.Linear(fin, h1), nn.ReLU(inplace=True), nn.Linear(h1, h2), nn.ReLU(inplace=True),
                         nn.Linear(h2, fout))


class NativeNet(nn.Module):
    """Base: keeps libss2's packed copy of the weights in sync with the module's tensors."""
    NET_ID = -1

    def __init__(self):
        super().__init__()
        self._dirty_gen = 0

    def _signature(self):
        return tuple((k, t.data_ptr(), t._version) for k, t in self.state_dict(keep_vars=True).items())

    def mark_dirty(self):
        """Force a re-upload on the next forward.  Needed after in-place edits through `.data`
        (they do not bump the tensors' version counters, so _signature cannot see them)."""
        self._dirty_gen += 1

    def load_state_dict(self, *args, **kwargs):
        self.mark_dirty()
        return super().load_state_dict(*args, **kwargs)

    def _apply(self, fn, *args, **kwargs):
        self.mark_dirty()
        return super()._apply(fn, *args, **kwargs)

    def sync_weights(self, ctx):
        # libss2 keeps ONE packed weight set per NET_ID per context: the signature of what is loaded there lives on
        # the context and names its owner, so a second instance of the same class always triggers a reload
        sig = (id(self), self._dirty_gen) + self._signature()
        if ctx.synced.get(self.NET_ID) != sig:
            ctx.load_state_dict(self.NET_ID, self.state_dict())
            ctx.synced[self.NET_ID] = sig

    def init_reference_style(self):
        # spatial_network.py:261-266: kaiming-normal convs, BN weight 1 / bias 0
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight)
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()
        self.mark_dirty()

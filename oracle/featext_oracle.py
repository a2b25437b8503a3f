"""CPU restatement of the reference's feature extractor -- TEST INFRASTRUCTURE (a checker, never the product).

FeatExt.forward (code/utils/my_utils.py:705-708) = init_conv (:696-700) -> UNet.forward (:667-690) with
UNet(16, enc=2, dec=1, initial_scale=2, bottom=[], filters=[32, 64, 128], head=[]) (:701) -> final_conv_1..3 (:702-704),
BasicBlock (:531-576), eval-mode BatchNorm.  Written as plain functional calls on a state_dict with the reference's key
names; pinned against the live reference and tests/golden/featext_seed5.npz by tests/test_oracle.py."""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.nn.functional as F

EPS = 1e-5


def _bn(x, sd, p):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"], False, 0.0, EPS)


def _basic_block(x, sd, p, stride):
    """my_utils.py:558-576."""
    out = F.relu(_bn(F.conv2d(x, sd[p + ".conv1.weight"], None, stride, 1), sd, p + ".bn1"))
    out = _bn(F.conv2d(out, sd[p + ".conv2.weight"], None, 1, 1), sd, p + ".bn2")
    res = x
    if p + ".downsample.0.weight" in sd:
        res = _bn(F.conv2d(x, sd[p + ".downsample.0.weight"], None, stride, 0), sd, p + ".downsample.1")
    return F.relu(out + res)


def featext_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """x [n,3,H,W] -> (1/8, 1/4, 1/2 resolution maps, 32 channels each), NCHW like the reference."""
    out = F.relu(_bn(F.conv2d(x, sd["init_conv.0.weight"], None, 2, 2), sd, "init_conv.1"))
    enc = []
    for i, name in enumerate(("2d2_0", "2d4_1", "2d8_2")):
        p = "unet.enc_blocks." + name
        out = _basic_block(out, sd, p + ".0", 1 if i == 0 else 2)
        out = _basic_block(out, sd, p + ".1", 1)
        enc.append(out)
    dec = [out]
    for i, name in enumerate(("2d16_3", "2d8_4")):
        p = "unet.dec_blocks." + name
        out = F.conv_transpose2d(out, sd[p + ".0.weight"], None, 2, 1, 1)
        out = torch.cat([out, enc[-2 - i]], dim=1)
        out = F.conv2d(out, sd[p + ".1.weight"], None, 1, 1)
        out = _basic_block(out, sd, p + ".2.0", 1)
        dec.append(out)
    return (F.conv2d(dec[0], sd["final_conv_1.weight"], None, 1, 1), F.conv2d(dec[1], sd["final_conv_2.weight"], None, 1, 1),
            F.conv2d(dec[2], sd["final_conv_3.weight"], None, 1, 1))

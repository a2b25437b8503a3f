"""Row f4: FeatExt (code/utils/my_utils.py:693-708) and the channels-last scene feature store.
CPU: the oracle restatement against the golden output of the live reference (seeded random weights under the reference's
key names).  GPU: the native kernels (csrc/featext.cu) against the golden fixture and against the oracle at other shapes,
and the store fed by FeatExt against the tensor path of get_feat_loss_corr."""
import numpy as np
import pytest
import torch

from mvsdf_b200 import synth
from mvsdf_b200.featext import B200FeatExt
from oracle import featext_oracle as FO


def _golden_inputs(g):
    seed = int(g["meta_seed"])
    n, H, W = [int(v) for v in g["meta_shape"]]
    model = B200FeatExt(seed=seed)
    sd = model.state_dict()
    got_sha = synth.state_dict_checksum({k: v for k, v in sd.items() if v.dtype.is_floating_point})
    assert got_sha == str(g["meta_weights_sha"]), "seeded FeatExt weights drifted from the golden fixture"
    x = torch.randn(n, 3, H, W, generator=torch.Generator().manual_seed(seed + 100))
    return model, sd, x


def test_oracle_matches_reference_golden(golden):
    g = golden("featext_seed5")
    _, sd, x = _golden_inputs(g)
    with torch.no_grad():
        o8, o4, o2 = FO.featext_forward(sd, x)
    for got, key in ((o8, "out_eighth"), (o4, "out_quarter"), (o2, "out_half")):
        ref = torch.from_numpy(g[key])
        assert got.shape == ref.shape
        assert (got - ref).abs().max().item() <= 1e-5 * ref.abs().max().item(), key


def test_state_dict_has_the_reference_key_names():
    """The slice of utils/vismvsnet.pt the reference loads (module.feat_ext.* keys, my_utils.py:702-703) must load as is."""
    keys = set(B200FeatExt().state_dict().keys())
    for k in ("init_conv.0.weight", "init_conv.1.running_var", "unet.enc_blocks.2d2_0.0.downsample.0.weight",
              "unet.enc_blocks.2d8_2.1.bn2.num_batches_tracked", "unet.dec_blocks.2d16_3.0.weight", "unet.dec_blocks.2d8_4.2.0.conv1.weight",
              "final_conv_1.weight", "final_conv_3.weight"):
        assert k in keys, k
    assert len(keys) == 127


@pytest.mark.gpu
def test_native_featext_matches_reference_golden(golden):
    from tests.helpers import gate
    g = golden("featext_seed5")
    model, sd, x = _golden_inputs(g)
    dev = torch.device("cuda:0")
    model = model.to(dev)
    o8, o4, o2 = model(x.to(dev))
    for got, key in ((o8, "out_eighth"), (o4, "out_quarter"), (o2, "out_half")):
        ref = torch.from_numpy(g[key])
        assert tuple(got.shape) == tuple(ref.shape)
        gate("featext_rel_of_max_" + key, (got.cpu() - ref).abs().max().item() / ref.abs().max().item(), 2e-5)
    # the finest map is produced channels-last: exactly what the feature store keeps
    nhwc = model.forward_nhwc(x.to(dev))
    assert nhwc.is_contiguous() and torch.equal(nhwc.permute(0, 3, 1, 2), o2)


@pytest.mark.gpu
@pytest.mark.parametrize("n,H,W,seed", [(1, 16, 16, 1), (3, 40, 72, 2), (2, 200, 136, 3)])
def test_native_featext_matches_oracle_at_other_shapes(n, H, W, seed):
    from tests.helpers import gate
    dev = torch.device("cuda:0")
    model = B200FeatExt(seed=seed)
    sd = model.state_dict()
    x = torch.randn(n, 3, H, W, generator=torch.Generator().manual_seed(seed))
    with torch.no_grad():
        refs = FO.featext_forward(sd, x)
    outs = model.to(dev)(x.to(dev))
    for got, ref, name in zip(outs, refs, ("eighth", "quarter", "half")):
        gate("featext_rel_of_max_" + name, (got.cpu() - ref).abs().max().item() / ref.abs().max().item(), 2e-5)


@pytest.mark.gpu
def test_featext_fills_the_scene_store_used_by_the_feature_loss():
    """scene_dataset.py:138-149 + loss.py:115-165 through the store: maps written by FeatExt (channels-last, indexed per
    (image, view)) give the same feature loss as the reference's tensor interface fed with the same maps in NCHW."""
    from mvsdf_b200.loss import B200IDRLoss
    dev = torch.device("cuda:0")
    scene = synth.make_scene(64, 64, n_images=2, n_src=2, seed=4)
    n_views = 4
    fe = B200FeatExt(seed=7).to(dev)
    imgs = torch.randn(n_views, 3, 64, 64, generator=torch.Generator().manual_seed(9))
    loss_a, loss_b = B200IDRLoss(), B200IDRLoss()
    maps = fe.fill_store(loss_a.store, imgs, batch=3, device=dev)               # [4, 32, 32, 32] channels-last, resident
    assert tuple(maps.shape) == (n_views, 32, 32, 32)
    feats_nchw = maps.permute(0, 3, 1, 2).contiguous()
    src_idx = torch.tensor([[(i + 1 + s) % n_views for s in range(2)] for i in range(2)])
    pts = (torch.rand(300, 3, generator=torch.Generator().manual_seed(1)) - 0.5).to(dev)
    offs = torch.tensor([0, 170, 300], dtype=torch.int32, device=dev)
    mask = torch.ones(2 * scene["uv"].shape[1], dtype=torch.bool, device=dev)
    common = (scene["cam"].to(dev), )
    a = loss_a.get_feat_loss_corr(pts, None, None, scene["cam"].to(dev), None, scene["src_cams"].to(dev), scene["size"][:1].to(dev),
                                  scene["center"][:1].to(dev), mask, mask, hit_offsets=offs, feat_index=torch.arange(2), src_index=src_idx)
    b = loss_b.get_feat_loss_corr(pts, None, feats_nchw[:2], scene["cam"].to(dev), feats_nchw[src_idx.to(dev)], scene["src_cams"].to(dev),
                                  scene["size"][:1].to(dev), scene["center"][:1].to(dev), mask, mask, hit_offsets=offs)
    assert torch.isfinite(a) and float(a) == float(b)
    assert loss_a.store.restacks == 0 and loss_b.store.restacks > 0

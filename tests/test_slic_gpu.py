"""GPU parity of the SLIC kernel (csrc/slic_kernels.cu) and of the "slic" / "none" segmentation paths of FeatureExtractor.

Bar: SLIC is integer work end to end, so the kernel's labels must equal the numpy oracle's (oracle/slic.py) exactly, for
every pixel, at every tested size / cluster count / compactness / iteration count.  (The oracle itself is a restatement
of an un-vendored C++ package — see its header for what that pins and what it does not.)
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("h,w,k,m,iters,batch", [
    (448, 448, 100, 10.0, 10, 3),      # the reference's configuration at the bench image size
    (224, 224, 100, 10.0, 10, 2),
    (96, 160, 30, 5.0, 4, 2),          # non-square, ragged tiles
    (64, 64, 7, 20.0, 10, 1),          # 3 x 3 grid: every pixel sees every centre
    (37, 91, 12, 10.0, 1, 1),          # single round, image smaller than the tiles' multiple
    (448, 448, 400, 1.0, 10, 1),       # many small superpixels, colour-dominated distance
])
def test_slic_labels_bit_exact_vs_oracle(h, w, k, m, iters, batch):
    from oracle import slic
    from wild_visual_navigation_b200 import ops

    imgs = np.stack([slic.synthetic_image(h, w, seed=10 * b + h % 7) for b in range(batch)])
    got = ops.slic(torch.from_numpy(imgs).cuda(), k, m, iters).cpu().numpy()
    S, nx, ny = slic.grid_geometry(h, w, k)
    assert got.shape == (batch, h, w) and got.min() >= 0 and got.max() < nx * ny
    for b in range(batch):
        want, _ = slic.slic(slic.float_image_to_u8(imgs[b]), k, m, iters)
        bad = int((got[b] != want).sum())
        assert bad == 0, f"frame {b}: {bad} of {h * w} labels differ"


def test_slic_edge_inputs():
    """Constant image (every distance ties on colour: pure Voronoi cells of the grid), saturated values, one cluster."""
    from oracle import slic
    from wild_visual_navigation_b200 import ops

    for img in (np.zeros((3, 80, 80), np.float32), np.ones((3, 80, 80), np.float32),
                np.full((3, 80, 80), 0.999999, np.float32)):
        got = ops.slic(torch.from_numpy(img)[None].cuda(), 16, 10.0, 10)[0].cpu().numpy()
        want, _ = slic.slic(slic.float_image_to_u8(img), 16, 10.0, 10)
        assert np.array_equal(got, want) and len(np.unique(got)) == 16
    one = slic.synthetic_image(40, 40, seed=3)
    got = ops.slic(torch.from_numpy(one)[None].cuda(), 1, 10.0, 3)[0]
    assert int(got.max()) == 0


def test_feature_extractor_slic_default_vs_oracle():
    """FeatureExtractor's constructor default (segmentation_type="slic", feature_extractor.py:20-27,88-95,221-225):
    segments = the oracle's SLIC labels renumbered without gaps; graph + pooled features from those segments."""
    from oracle import slic, wvn_path
    from oracle.dino_vit import ViTConfig, synthetic_state_dict
    from wild_visual_navigation_b200.feature_extractor import FeatureExtractor

    cfg = ViTConfig.from_name("vit_small", 8, 224)
    sd = synthetic_state_dict(cfg, seed=4)
    fe = FeatureExtractor("cuda", feature_type="dino", input_size=224, state_dict=sd, max_batch=2)
    assert fe.segmentation_type == "slic"
    imgs = np.stack([slic.synthetic_image(224, 224, seed=s) for s in (1, 2)])
    img = torch.from_numpy(imgs).cuda()
    edges, feat, seg, center, dense = fe.extract(img[:1], return_dense_features=True)
    want, _ = slic.slic(slic.float_image_to_u8(imgs[0]), 100, 10.0, 10)
    seg_ref = torch.from_numpy(slic.compact_labels(want)).cuda()
    assert torch.equal(seg, seg_ref)
    n = int(seg_ref.max()) + 1
    assert feat.shape == (n, 384) and center.shape == (n, 2)
    assert torch.equal(edges, wvn_path.adjacency_list(seg[None, None]).T)
    assert (center - wvn_path.centers(seg[None, None])).abs().max() < 1e-3
    assert rel_l2(feat, wvn_path.sparsify_features(dense, seg)) <= 1e-4
    assert not torch.isnan(feat).any()
    # batched form: both frames, padded rows
    r = fe.extract_batch(img)
    ns = r["n_segments"].tolist()
    for b in range(2):
        w, _ = slic.slic(slic.float_image_to_u8(imgs[b]), 100, 10.0, 10)
        assert torch.equal(r["seg"][b].cpu(), torch.from_numpy(slic.compact_labels(w))) and ns[b] == len(np.unique(w))


def test_feature_extractor_pixelwise_segmentation():
    """segmentation_type="none" (segment_pixelwise, feature_extractor.py:179-196; sparsify_features returns the dense
    features unchanged, :389-396): every pixel a node, 4-neighbour edges, (y, x) centres."""
    from oracle import wvn_path
    from oracle.dino_vit import ViTConfig, synthetic_state_dict
    from wild_visual_navigation_b200.feature_extractor import FeatureExtractor

    cfg = ViTConfig.from_name("vit_small", 8, 224)
    sd = synthetic_state_dict(cfg, seed=4)
    fe = FeatureExtractor("cuda", segmentation_type="none", feature_type="dino", input_size=224, state_dict=sd)
    img = torch.rand(1, 3, 224, 224, generator=torch.Generator().manual_seed(5)).cuda()
    edges, feat, seg, center, dense = fe.extract(img, return_dense_features=True)
    H = W = 224
    assert torch.equal(seg, torch.arange(H * W, device="cuda").reshape(H, W))
    assert edges.shape == (2, H * (W - 1) + (H - 1) * W)
    assert edges[:, 0].tolist() == [0, 1] and edges[:, H * (W - 1)].tolist() == [0, W]
    assert center.shape == (H * W, 2) and center[W + 3].tolist() == [1, 3]
    assert feat.shape == (1, 384, H, W) and dense is feat
    assert rel_l2(feat, wvn_path.dino_inference(img, {k: v.cuda() for k, v in sd.items()}, cfg)) <= 2e-2

"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/wvn_b200.h declares, the package imports without a GPU, and the host-side containers
behave like the reference's."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_lib():
    import __graft_entry__ as g

    if not os.path.exists(os.path.join(ROOT, "wild_visual_navigation_b200", "libwvn_b200.so")):
        g.build()
    from wild_visual_navigation_b200 import _C

    return _C


def test_library_exports_every_declared_symbol(built_lib):
    hdr = open(os.path.join(ROOT, "include", "wvn_b200.h")).read()
    declared = set(re.findall(r"\b(wvn_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 28
    l = built_lib.lib()
    for name in declared:
        assert hasattr(l, name), f"libwvn_b200.so lacks {name}"
    assert declared == set(built_lib.SIGNATURES), declared ^ set(built_lib.SIGNATURES)
    assert l.wvn_version() >= 100


def test_no_gpu_means_loud_failure(built_lib):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert built_lib.lib().wvn_check_device() == -3
    with pytest.raises(built_lib.WvnError):
        built_lib.require_device()
    from wild_visual_navigation_b200 import DinoInterface

    with pytest.raises(Exception):
        DinoInterface("cpu")


def test_data_batch_semantics():
    from wild_visual_navigation_b200 import Batch, Data

    d1 = Data(x=torch.zeros(3, 4), y=torch.ones(3), y_valid=torch.tensor([True, False, True]),
              edge_index=torch.tensor([[0, 1], [1, 2]]))
    d2 = Data(x=torch.ones(2, 4), y=torch.zeros(2), y_valid=torch.tensor([False, False]),
              edge_index=torch.tensor([[0], [1]]))
    b = Batch.from_data_list([d1, d2])
    assert b.x.shape == (5, 4) and b.y.shape == (5,) and b.ba == 5
    assert b.ptr.tolist() == [0, 3, 5] and b.batch.tolist() == [0, 0, 0, 1, 1]
    assert b.edge_index.tolist() == [[0, 1, 3], [1, 2, 4]]
    assert Batch.from_data_list([]) is None


def test_simple_mlp_state_dict_layout_and_quirk():
    from wild_visual_navigation_b200 import SimpleMLP, get_model

    hs = [256, 32, 1]
    m = SimpleMLP(384, hs, True)
    assert hs == [256, 32, 385]  # the reference mutates the caller's list (simple_mlp.py:21-22)
    assert m.nr_sigmoid_layers == 1 and m.output_features == 385
    sd = m.state_dict()
    assert [tuple(v.shape) for v in sd.values()] == [(256, 384), (256,), (32, 256), (32,), (385, 32), (385,)]
    assert m.flat_params.numel() == 119489
    # parameters are views into the flat buffer, and load_state_dict writes through
    new = {k: torch.full_like(v, 0.5) for k, v in sd.items()}
    m.load_state_dict(new)
    assert torch.all(m.flat_params == 0.5)
    m2 = get_model({"name": "SimpleMLP", "simple_mlp_cfg": {"input_size": 90, "hidden_sizes": [256, 32, 1], "reconstruction": True}})
    assert m2.output_features == 91


def test_checkpoint_keys_match_reference_made_checkpoint(golden_dir):
    """A checkpoint written by the reference's own classes (tests/golden/checkpoint_ref.pt, make_golden.py) and the
    state dicts this package produces have the same keys — in both directions strict load_state_dict works."""
    from wild_visual_navigation_b200 import SimpleMLP, TraversabilityLoss

    ck = torch.load(os.path.join(golden_dir, "checkpoint_ref.pt"), weights_only=False)
    assert set(ck) == {"step", "model_state_dict", "optimizer_state_dict", "traversability_loss_state_dict", "loss"}
    model = SimpleMLP(16, [8, 4, 1], True)
    loss = TraversabilityLoss(0.03, 0.5, 0.0, True, model, "latest_measurement", 0.5)
    assert list(model.state_dict()) == list(ck["model_state_dict"])
    assert list(loss.state_dict()) == list(ck["traversability_loss_state_dict"])
    model.load_state_dict(ck["model_state_dict"])                      # strict
    loss.load_state_dict(ck["traversability_loss_state_dict"])         # strict: needs the _model.* keys
    assert torch.equal(model.flat_params[: 8 * 16].view(8, 16), ck["model_state_dict"]["layers.0.weight"])
    assert torch.equal(loss._confidence_generator.mean, ck["traversability_loss_state_dict"]["_confidence_generator.mean"])
    assert len(ck["optimizer_state_dict"]["state"]) == 6 and "exp_avg" in ck["optimizer_state_dict"]["state"][0]


def test_confidence_generator_state_dict_keys():
    from wild_visual_navigation_b200 import ConfidenceGenerator, TraversabilityLoss, SimpleMLP

    cg = ConfidenceGenerator(0.5, "latest_measurement")
    assert set(cg.state_dict()) == {"mean", "var", "std"}
    assert cg.var.shape == (1, 1)
    loss = TraversabilityLoss(0.03, 0.5, 0.0, True, SimpleMLP(8, [4, 2, 1], True), "latest_measurement", 0.5)
    # the reference registers the model as a sub-module of the loss (loss.py:74), so its keys ride in this state dict
    assert list(loss.state_dict()) == ["_model.layers.0.weight", "_model.layers.0.bias", "_model.layers.2.weight",
                                       "_model.layers.2.bias", "_model.layers.4.weight", "_model.layers.4.bias",
                                       "_confidence_generator.mean", "_confidence_generator.var",
                                       "_confidence_generator.std"]
    x = torch.rand(10)
    assert torch.equal(cg.inference_without_update(x.to("meta")) if False else cg.inference_without_update(x), cg.inference_without_update(x))


def test_product_weight_generators_match_the_oracles():
    """The product must not import oracle/, so its seeded random-init generators (used when no checkpoint is given) are
    written twice; this ties the two copies together bit for bit."""
    import torch

    from oracle.dino_vit import ViTConfig, synthetic_state_dict
    from oracle.stego_head import synthetic_head
    from wild_visual_navigation_b200.feature_extractor import weights

    for name, patch, image in (("vit_small", 8, 224), ("vit_base", 8, 64)):
        cfg = ViTConfig.from_name(name, patch, image)
        for seed, attn_std in ((1, 0.09), (7, 0.2)):
            want = synthetic_state_dict(cfg, seed=seed, attn_std=attn_std)
            got = weights.synthetic_dino_state_dict(cfg.dim, cfg.depth, cfg.mlp_dim, patch, cfg.pretrain_grid, seed=seed,
                                                    attn_std=attn_std)
            assert list(got) == list(want)
            assert all(torch.equal(got[k], want[k]) for k in want)
    a, b = weights.synthetic_stego_head(384, 90, 32, 27, seed=3), synthetic_head(384, 90, 32, 27, seed=3)
    assert list(a) == list(b) and all(torch.equal(a[k], b[k]) for k in b)

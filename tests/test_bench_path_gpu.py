"""GPU parity tests of EXACTLY what bench.py times: ``HotPathStep.step`` at the benchmark's configuration
(stego segmentation with the per-image k-means + dino features, 448x448, B = 32 frames in one ViT chunk of 32, per-pixel
MLP, fused train step on the padded pooled rows) against the oracle chain run frame by frame on the same GPU (eager fp32 PyTorch in the
reference's order of operations, oracle/pipeline.py) — plus the holes VERDICT r1 listed: segment-wise prediction,
grid / random segmentation, the wire format on device tensors.

Tolerances (reference arithmetic fp32; the tensor-core path uses bf16 operands with fp32 accumulation):
  ViT tokens                 rel-L2 <= 2e-2 per batch, cosine >= 0.99 per token
  STEGO code                 rel-L2 <= 3e-2 vs the fp32 oracle
  segment ids                >= 99 % of the pixels on average (>= 90 % on every frame) vs the oracle's k-means +
                             nearest-centroid labelling of OUR code
  pooled features            rel-L2 <= 2e-2 vs the oracle's definition evaluated on OUR segmentation and the ORACLE's
                             dense features (so the error is the ViT's, not the pooling's)
  centers / edges            exact definition on our segmentation: centers 1e-3 px, edge list equal
  trav / conf maps           mean abs <= 1e-2, max abs <= 6e-2 end to end (tokens differ by ~1e-2 rel)
  train step                 2e-5 rel on loss terms and updated parameters, fed the SAME pooled rows
"""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


@pytest.fixture(scope="module", autouse=True)
def _fp32_reference_math():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield


def _to(sd, dev):
    return {k: v.to(dev) for k, v in sd.items()}


@pytest.fixture(scope="module")
def bench_setup():
    import bench
    from wild_visual_navigation_b200 import HotPathStep

    cfg, sd, hd = bench.make_weights()
    hp = HotPathStep("cuda", sd, hd, batch=bench.BATCH, input_size=bench.IMG, chunk=32, flip_tta=False,
                     run_clustering=True, n_image_clusters=bench.K_IMAGE_CLUSTERS)
    return cfg, sd, hd, hp


def test_bench_step_end_to_end_vs_oracle(bench_setup):
    """Two consecutive bench steps (B = 32, chunk = 32): every output of the first, and the second's maps / train step
    with the weights and confidence statistics the first one produced."""
    import bench
    from oracle import pipeline, stego_head, wvn_path

    cfg, sd, hd, hp = bench_setup
    B, S = bench.BATCH, bench.IMG
    sdc, hdc = _to(sd, "cuda"), _to(hd, "cuda")
    smax = hp.smax
    y_all, yv_all = wvn_path.synthetic_supervision(B * smax, seed=2)
    y_dev, yv_dev = y_all.cuda(), yv_all.cuda()
    mlp_sd = {k: v.detach().clone() for k, v in hp.te._model.state_dict().items()}
    opt_state = None
    mean, std = hp.cg.mean.data.clone(), hp.cg.std.data.clone()
    frames = bench.synthetic_images(2, B, seed=100)
    for it in range(2):
        img = frames[it].cuda()
        r = hp.step(img, y_dev, yv_dev)
        torch.cuda.synchronize()
        ns = r["n_segments"].tolist()
        # the segmentation is held to the oracle's clustering of OUR code (Lloyd iterations amplify the ~1e-2 difference
        # between the bf16 tensor-core head and the fp32 oracle, see test_stego_per_image_kmeans_default)
        code = hp.fe._stego.code_tokens.transpose(1, 2).reshape(B, 90, S // 8, S // 8)
        cl_same = stego_head.kmeans_predict(code, stego_head.image_kmeans(code, bench.K_IMAGE_CLUSTERS, bench.KMEANS_ITERS), (S, S))
        assert r["tokens"].shape == (B, 3136, 384) and r["seg"].shape == (B, S, S) and r["feat"].shape == (B, smax, 384)
        tok_err, seg_agree, feat_err, trav_mean, trav_max, conf_mean, conf_max, code_err = [], [], [], [], [], [], [], []
        for b in range(B):
            f = pipeline.frame_features(img[b : b + 1], sdc, cfg, hdc, n_image_clusters=bench.K_IMAGE_CLUSTERS,
                                        kmeans_iters=bench.KMEANS_ITERS)
            tok_ref = f["fmap"][0].permute(1, 2, 0).reshape(-1, 384)
            cos = torch.nn.functional.cosine_similarity(r["tokens"][b], tok_ref, dim=-1)
            assert cos.min() >= 0.99, (b, cos.min().item())
            tok_err.append(rel_l2(r["tokens"][b], tok_ref))
            seg_agree.append((r["seg"][b] == wvn_path.relabel(cl_same[b])).float().mean().item())
            code_err.append(rel_l2(code[b], stego_head.head_code(f["fmap"], hdc)[0]))
            n = ns[b]
            assert n == int(r["seg"][b].max()) + 1
            if b in (0, 13, 31) or it == 0 and b % 8 == 0:
                seg = r["seg"][b]
                feat_ref = wvn_path.sparsify_features(f["dense"], seg)
                feat_err.append(rel_l2(r["feat"][b, :n], feat_ref))
                assert (r["centers"][b, :n] - wvn_path.centers(seg[None, None])).abs().max() < 1e-3
                ne = int(r["n_edges"][b])
                assert ne >= 0 and torch.equal(r["edges"][b, :ne], wvn_path.adjacency_list(seg[None, None]))
            t_ref, c_ref = wvn_path.pixel_inference(f["dense"], _to(mlp_sd, "cuda"), mean, std, 0.5)
            dt, dc = (r["trav"][b] - t_ref).abs(), (r["conf"][b] - c_ref).abs()
            trav_mean.append(dt.mean().item()); trav_max.append(dt.max().item())
            conf_mean.append(dc.mean().item()); conf_max.append(dc.max().item())
        print(f"bench step {it}: tokens rel_l2 max {max(tok_err):.2e}; seg agreement min {min(seg_agree):.4f}; pooled feat rel_l2 "
              f"max {max(feat_err):.2e}; trav |d| mean/max {max(trav_mean):.2e}/{max(trav_max):.2e}; conf |d| mean/max "
              f"{max(conf_mean):.2e}/{max(conf_max):.2e}")
        print(f"bench step {it}: STEGO code rel_l2 max {max(code_err):.2e}")
        assert max(tok_err) <= 2e-2 and max(code_err) <= 3e-2
        # Lloyd iterations amplify even summation-order differences (fp32 sums in another order can move one boundary
        # patch, which moves two centroids, ...): most frames agree to > 99.99 %, an occasional frame drifts by a few
        # percent of its pixels.  >= 99 % on average, >= 90 % on every frame; the kernel itself is held to the oracle
        # on identical inputs (exact labels) by test_stego_kmeans_kernel_vs_oracle_on_identical_code.
        assert min(seg_agree) >= 0.90 and sum(seg_agree) / len(seg_agree) >= 0.99
        assert max(feat_err) <= 2e-2
        assert max(trav_mean) <= 1e-2 and max(trav_max) <= 6e-2
        assert max(conf_mean) <= 1e-2 and max(conf_max) <= 6e-2
        # ---- the train step on OUR pooled rows (what `feat[mask]` gives) vs the reference's autograd + Adam
        x = torch.cat([r["feat"][b, : ns[b]] for b in range(B)]).cpu()
        n = x.shape[0]
        mlp_sd = {k: v.cpu() for k, v in mlp_sd.items()}
        mlp_sd, opt_state, ref = wvn_path.train_step(mlp_sd, opt_state, x, y_all[:n], yv_all[:n], lr=1e-3)
        m = hp.te._trainer.metrics.tolist()
        got = torch.cat([v.reshape(-1) for v in hp.te._model.state_dict().values()]).cpu()
        want = torch.cat([v.reshape(-1) for v in mlp_sd.values()])
        print(f"bench step {it}: rows {n}; loss {m[0]:.7f} vs {ref['loss_total']:.7f}; param rel {rel_l2(got, want):.2e}")
        assert abs(m[0] - ref["loss_total"]) <= 2e-5 * max(1.0, abs(ref["loss_total"]))
        assert abs(m[1] - ref["loss_trav"]) <= 2e-5 and abs(m[2] - ref["loss_reco"]) <= 2e-5 * max(1.0, abs(ref["loss_reco"]))
        assert abs(m[4] - ref["mean"]) <= 2e-5 * max(1.0, abs(ref["mean"])) and abs(m[5] - ref["std"]) <= 2e-5 * max(1.0, abs(ref["std"]))
        assert rel_l2(got, want) <= 2e-5
        assert (r["confidence_rows"][:n].cpu() - ref["confidence"]).abs().max() <= 1e-4
        mean, std = torch.tensor([ref["mean"]], device="cuda"), torch.tensor([ref["std"]], device="cuda")
        assert torch.allclose(hp.cg.mean.data, mean, rtol=2e-5) and torch.allclose(hp.cg.std.data, std, rtol=2e-5)


def test_padded_step_equals_compacted_step():
    """train_on_padded(feat, n_rows) == step(feat[mask]) : the in-kernel row compaction is the boolean-mask gather."""
    from wild_visual_navigation_b200 import ops

    torch.manual_seed(3)
    G, S, D = 5, 16, 384
    feat = torch.randn(G, S, D, device="cuda")
    n_rows = torch.tensor([16, 3, 0, 9, 1], device="cuda", dtype=torch.int32)
    feat[1, 3:] = float("nan")  # padding rows may hold anything
    mask = torch.arange(S, device="cuda")[None, :] < n_rows[:, None]
    n = int(mask.sum())
    yv = torch.rand(G * S, device="cuda") < 0.4
    yv[:2] = True
    y = torch.where(yv, torch.rand(G * S, device="cuda").clamp(min=0.001), torch.zeros(G * S, device="cuda"))
    p0 = torch.randn(ops.lib().wvn_mlp_param_count(D, 256, 32), device="cuda") * 0.05
    a = ops.MlpTrainer(p0.clone(), D, 256, 32, max_rows=G * S)
    b = ops.MlpTrainer(p0.clone(), D, 256, 32, max_rows=G * S)
    c = ops.MlpTrainer(p0.clone(), D, 256, 32, max_rows=G * S, legacy=True)
    for _ in range(3):
        ca = a.step_padded(feat, n_rows, y, yv)[:n].clone()
        cb = b.step(feat[mask], y[:n], yv[:n]).clone()
        cc = c.step(feat[mask], y[:n], yv[:n]).clone()
        assert rel_l2(a.params, b.params) <= 1e-6 and (ca - cb).abs().max() <= 1e-5   # shared-memory atomics reorder sums
        assert rel_l2(a.params, c.params) <= 2e-5 and (ca - cc).abs().max() <= 1e-5   # fused vs round-1 kernels
        assert (a.metrics - b.metrics).abs().max() <= 1e-5 and (a.metrics - c.metrics).abs().max() <= 2e-5
        assert int(a.step_counter) == int(c.step_counter)


def test_predict_segments_vs_oracle():
    """Segment-wise prediction mode (node :324-327): MLP on the pooled rows, scattered back through seg."""
    from oracle import wvn_path
    from wild_visual_navigation_b200 import ConfidenceGenerator, SimpleMLP, TraversabilityInference
    from wild_visual_navigation_b200.feature_extractor import DinoInterface
    from oracle.dino_vit import ViTConfig, synthetic_state_dict

    cfg = ViTConfig.from_name("vit_small", 8, 224)
    di = DinoInterface("cuda", input_size=224, backbone_type="vit_small", patch_size=8,
                       state_dict=synthetic_state_dict(cfg, seed=4), max_batch=1)
    torch.manual_seed(42)
    model = SimpleMLP(384, [256, 32, 1], True).cuda()
    with torch.no_grad():
        model.flat_params.mul_(2.0)
    cg = ConfidenceGenerator(std_factor=0.5, method="latest_measurement").cuda()
    cg.mean[0], cg.std[0] = 0.8, 0.3
    ti = TraversabilityInference(di, model, cg)
    feat = torch.randn(37, 384, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    seg = torch.randint(0, 37, (224, 224), device="cuda", generator=torch.Generator(device="cuda").manual_seed(2))
    trav, conf = ti.predict_segments(feat, seg)
    msd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    pred = wvn_path.mlp_forward(feat, msd)
    lr = ((pred[:, 1:] - feat) ** 2).mean(1)
    c_ref = wvn_path.confidence_inference(lr, cg.mean.data, cg.std.data, 0.5)
    assert trav.shape == conf.shape == (224, 224)
    print("segment-wise |d| trav/conf", (trav - pred[:, 0][seg]).abs().max().item(), (conf - c_ref[seg]).abs().max().item())
    assert (trav - pred[:, 0][seg]).abs().max() <= 2e-2 and (conf - c_ref[seg]).abs().max() <= 3e-2


@pytest.mark.parametrize("seg_type", ["grid", "random"])
def test_grid_and_random_segmentation_vs_reference_definitions(seg_type):
    """segment_grid (feature_extractor.py:198-219) / segment_random (:227-235) + sparsify_features on those maps."""
    from oracle import wvn_path
    from oracle.dino_vit import ViTConfig, synthetic_state_dict
    from wild_visual_navigation_b200.feature_extractor import FeatureExtractor

    cfg = ViTConfig.from_name("vit_small", 8, 224)
    sd = synthetic_state_dict(cfg, seed=4)
    fe = FeatureExtractor("cuda", segmentation_type=seg_type, feature_type="dino", input_size=224, state_dict=sd,
                          max_batch=1)
    img = torch.rand(1, 3, 224, 224, generator=torch.Generator().manual_seed(5)).cuda()
    torch.manual_seed(7)
    edges, feat, seg, center, dense = fe.extract(img, return_dense_features=True, cell_size=32, n_random_pixels=50)
    dense_ref = wvn_path.dino_inference(img, _to(sd, "cuda"), cfg)
    assert rel_l2(dense, dense_ref) <= 2e-2
    if seg_type == "grid":
        # reference: cells of cell_size x cell_size numbered row-major, edges + centers from SegmentExtractor
        ys, xs = torch.arange(224, device="cuda") // 32, torch.arange(224, device="cuda") // 32
        seg_ref = ys[:, None] * 7 + xs[None, :]
        assert torch.equal(seg, seg_ref) and feat.shape == (49, 384)
        assert torch.equal(edges, wvn_path.adjacency_list(seg[None, None]).T)
        assert (center - wvn_path.centers(seg[None, None])).abs().max() < 1e-3
        assert rel_l2(feat, wvn_path.sparsify_features(dense, seg)) <= 1e-4
    else:
        # reference: n_random_pixels distinct pixels labelled 0..n-1, everything else -1; features = the dense feature
        # at that pixel; no edges / centers (feature_extractor.py:121-123)
        assert edges is None and center is None and feat.shape == (50, 384)
        assert int((seg >= 0).sum()) == 50 and sorted(seg[seg >= 0].tolist()) == list(range(50))
        ys, xs = torch.where(seg >= 0)
        order = seg[ys, xs].argsort()
        ref = dense[0][:, ys[order], xs[order]].T
        assert rel_l2(feat, ref) <= 1e-4


def test_wire_format_roundtrip_on_device_outputs():
    """ImageFeatures encode / decode (wvn_feature_extractor_node.py:373-393, wvn_learning_node.py:651-656) fed with
    the device tensors ``extract`` returns."""
    from oracle.dino_vit import ViTConfig, synthetic_state_dict
    from oracle.stego_head import synthetic_head
    from wild_visual_navigation_b200.feature_extractor import FeatureExtractor
    from wild_visual_navigation_b200.utils import decode_image_features, encode_image_features

    cfg = ViTConfig.from_name("vit_small", 8, 224)
    fe = FeatureExtractor("cuda", segmentation_type="stego", feature_type="stego", input_size=224,
                          state_dict=synthetic_state_dict(cfg, seed=6), head_state_dict=synthetic_head(384, 90, 32, 27, seed=3),
                          flip_tta=False, run_clustering=False, max_batch=1)
    img = torch.rand(1, 3, 224, 224, generator=torch.Generator().manual_seed(8)).cuda()
    edges, feat, seg, center, _ = fe.extract(img)
    msg = encode_image_features(feat, seg, header={"seq": 3, "frame_id": "cam"})
    assert [d["size"] for d in msg["features"]["layout"]["dim"]] == [feat.shape[0], 90]
    feat2, seg2 = decode_image_features(msg, device="cuda")
    assert feat2.dtype == torch.float32 and torch.equal(feat2, feat)
    assert seg2.dtype == torch.int32 and torch.equal(seg2.long(), seg)


@pytest.mark.parametrize("size", [448, 224])
def test_stego_feature_per_pixel_inference_d90(size):
    """The ROS default ``feature_type: "stego"`` (default.yaml:21-22): 90-d features through the per-pixel head —
    the fused kernel at 448 (operands zero-padded to 128 columns), the interp + 3-GEMM path at 224 (geometry the fused
    kernel does not tile) — and the segment-wise rows path, against the oracle on identical code tokens."""
    from oracle import wvn_path
    from oracle.dino_vit import ViTConfig, synthetic_state_dict
    from oracle.stego_head import synthetic_head
    from wild_visual_navigation_b200 import ConfidenceGenerator, SimpleMLP, TraversabilityInference
    from wild_visual_navigation_b200.feature_extractor import FeatureExtractor

    cfg = ViTConfig.from_name("vit_small", 8, size)
    fe = FeatureExtractor("cuda", segmentation_type="stego", feature_type="stego", input_size=size,
                          state_dict=synthetic_state_dict(cfg, seed=6), head_state_dict=synthetic_head(384, 90, 32, 27, seed=3),
                          flip_tta=False, max_batch=2)
    assert fe.feature_dim == 90
    torch.manual_seed(42)
    model = SimpleMLP(90, [256, 32, 1], True).cuda()
    with torch.no_grad():
        model.flat_params.mul_(2.0)
    cg = ConfidenceGenerator(std_factor=0.5, method="latest_measurement").cuda()
    ti = TraversabilityInference(fe._dino, model, cg)          # wvn_mlp_infer_create(90, ...) must construct
    img = torch.rand(2, 3, size, size, generator=torch.Generator().manual_seed(9)).cuda()
    r = fe.extract_batch(img)
    code = r["tokens"]                                           # (2, P, 90) fp32 code at patch resolution
    g = size // 8
    assert code.shape == (2, g * g, 90)
    msd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    with torch.no_grad():
        xs = code[0, ::5]
        lr = ((wvn_path.mlp_forward(xs, msd)[:, 1:] - xs) ** 2).mean(1)
        cg.mean[0], cg.std[0] = lr.mean() - 0.5 * lr.std(), lr.std()
    trav, conf = ti.predict_from_tokens(code, size)
    for b in range(2):
        dense = torch.nn.functional.interpolate(code[b : b + 1].reshape(1, g, g, 90).permute(0, 3, 1, 2), (size, size),
                                                mode="bilinear", align_corners=True)
        t_ref, c_ref = wvn_path.pixel_inference(dense, msd, cg.mean.data, cg.std.data, 0.5)
        dt, dc = (trav[b] - t_ref).abs().max().item(), (conf[b] - c_ref).abs().max().item()
        print(f"D=90 per-pixel @ {size}: max |d| trav {dt:.2e} conf {dc:.2e}; conf mean {c_ref.mean().item():.3f}")
        assert dt <= 2e-2 and dc <= 3e-2
        assert 0.02 < c_ref.mean() < 0.98
    # segment-wise mode on the pooled 90-d rows
    n = int(r["n_segments"][0])
    t_rows, c_rows = ti.predict_segments(r["feat"][0, :n], r["seg"][0])
    pred = wvn_path.mlp_forward(r["feat"][0, :n], msd)
    assert (t_rows - pred[:, 0][r["seg"][0]]).abs().max() <= 2e-2


def test_estimator_loads_reference_made_checkpoint(golden_dir, tmp_path):
    """``load_checkpoint`` on a file written by the reference's own SimpleMLP / TraversabilityLoss / torch Adam
    (tests/golden/checkpoint_ref.pt): strict state-dict loads incl. the loss' ``_model.*`` keys, Adam moments and
    step restored; training continues from it and the next save has the same layout."""
    from wild_visual_navigation_b200 import TraversabilityEstimator
    from wild_visual_navigation_b200.traversability_estimator.traversability_estimator import MissionNode, default_params

    ck_path = os.path.join(golden_dir, "checkpoint_ref.pt")
    ck = torch.load(ck_path, weights_only=False)
    params = default_params()
    params["model"]["simple_mlp_cfg"] = {"input_size": 16, "hidden_sizes": [8, 4, 1], "reconstruction": True}
    te = TraversabilityEstimator(params=params, device="cuda", min_samples_for_training=0)
    te.load_checkpoint(ck_path)
    assert te.step == ck["step"] == 2
    for k, v in ck["model_state_dict"].items():
        assert torch.equal(te._model.state_dict()[k].cpu(), v), k
    assert torch.equal(te._traversability_loss._confidence_generator.mean.cpu(),
                       ck["traversability_loss_state_dict"]["_confidence_generator.mean"])
    assert int(te._trainer.step_counter) == 2
    exp_avg0 = ck["optimizer_state_dict"]["state"][0]["exp_avg"].reshape(-1)
    assert torch.equal(te._trainer.exp_avg[: exp_avg0.numel()].cpu(), exp_avg0)
    batch = torch.load(os.path.join(golden_dir, "checkpoint_ref_batch.pt"))
    te.add_mission_node(MissionNode(batch["x"].cuda(), batch["y"].cuda(), batch["y_valid"].cuda()))
    out = te.train()
    assert out["loss_total"] > 0 and int(te._trainer.step_counter) == 3
    te.save_checkpoint(str(tmp_path), "ck.pt")
    ck2 = torch.load(os.path.join(tmp_path, "ck.pt"), weights_only=False)
    assert list(ck2["traversability_loss_state_dict"]) == list(ck["traversability_loss_state_dict"])
    assert list(ck2["model_state_dict"]) == list(ck["model_state_dict"])


def test_bench_eager_and_cpu_legs_run():
    """bench.py's comparison legs (the oracle port as eager fp32 PyTorch on the GPU, with and without SDPA, and on the CPU)
    must keep running when the oracle changes: one small step of each."""
    import bench

    for device, sdpa in (("cuda", False), ("cuda", True)):
        fps, ms = bench.oracle_frames_per_s("c3", 1, 1, 0, device, sdpa=sdpa)
        assert fps > 0 and ms > 0

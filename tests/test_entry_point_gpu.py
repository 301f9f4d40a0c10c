"""End-to-end runs of the reference-shaped entry point (gen3c_b200/inference/gen3c_single_image.py) on a small network
with the weight-free tokenizer and a synthetic depth predictor: BASELINE config 3 (single chunk, left pan) and config 5
(autoregressive extension with per-chunk cache update + depth alignment) in miniature.  What is checked: the plumbing
(shapes, conditioning frames, cache growth, determinism), not image quality — the weights are random."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

NET = dict(model_channels=256, num_blocks=2, num_heads=2, adaln_lora_dim=32)   # context dim stays 1024 (T5 width)


def _args(tmp_path, **over):
    from gen3c_b200.inference import gen3c_single_image as m

    import cv2

    img = (np.random.RandomState(0).rand(90, 160, 3) * 255).astype(np.uint8)
    path = str(tmp_path / "in.png")
    cv2.imwrite(path, img)
    argv = ["--synthetic", "--prompt", "a room", "--input_image_path", path, "--disable_guardrail",
            "--disable_prompt_upsampler", "--disable_prompt_encoder", "--height", "128", "--width", "256", "--num_steps", "3",
            "--video_save_folder", str(tmp_path / "out"), "--seed", "3"]
    args = m.create_parser().parse_args(argv)
    for k, v in over.items():
        setattr(args, k, v)
    return m, args


def _pipeline(args):
    from gen3c_b200.gen3c_pipeline import Gen3cPipeline

    return Gen3cPipeline(inference_type="video2world", checkpoint_dir=args.checkpoint_dir, checkpoint_name="Gen3C-Cosmos-7B",
                         enable_prompt_upsampler=False, disable_guardrail=True, disable_prompt_encoder=True,
                         guidance=args.guidance, num_steps=args.num_steps, height=args.height, width=args.width, fps=args.fps,
                         num_video_frames=121, seed=args.seed, synthetic=True, net_kwargs=NET)


def test_single_chunk_left_pan(tmp_path):
    m, args = _args(tmp_path, foreground_masking=True, save_buffer=True)
    pipe = _pipeline(args)
    (path, video), = m.demo(args, pipeline=pipe)
    assert video.dtype == np.uint8 and video.shape == (121, 128, 256 * 2, 3)   # warp buffer strip + generated video
    assert os.path.exists(path) or os.path.exists(os.path.splitext(path)[0] + ".npy")
    # deterministic: same seed, same output
    (_, again), = m.demo(args, pipeline=pipe)
    assert np.abs(video.astype(int) - again.astype(int)).max() <= 1
    assert pipe.model.net.last_launch_count() > 0


def test_autoregressive_two_chunks_updates_the_cache(tmp_path):
    m, args = _args(tmp_path, num_video_frames=241, trajectory="clockwise")
    pipe = _pipeline(args)
    seen = {}
    from gen3c_b200 import cache_3d

    orig = cache_3d.Cache3D_Buffer.update_cache

    def spy(self, *a, **k):
        seen["n_before"] = self.input_image.shape[2]
        orig(self, *a, **k)
        seen["n_after"] = self.input_image.shape[2]
        seen["aligned"] = k.get("depth_alignment", True)

    cache_3d.Cache3D_Buffer.update_cache = spy
    try:
        (_, video), = m.demo(args, pipeline=pipe)
    finally:
        cache_3d.Cache3D_Buffer.update_cache = orig
    assert video.shape == (241, 128, 256, 3)
    assert seen == {"n_before": 1, "n_after": 2, "aligned": True}   # ring grew; the default (aligned) path ran


def test_pipeline_refuses_out_of_scope_models():
    from gen3c_b200.gen3c_pipeline import Gen3cPipeline

    kw = dict(inference_type="video2world", checkpoint_dir="checkpoints", checkpoint_name="Gen3C-Cosmos-7B", synthetic=True,
              net_kwargs=NET)
    with pytest.raises(NotImplementedError):
        Gen3cPipeline(enable_prompt_upsampler=True, disable_guardrail=True, **kw)
    with pytest.raises(NotImplementedError):
        Gen3cPipeline(enable_prompt_upsampler=False, disable_guardrail=False, **kw)
    with pytest.raises(FileNotFoundError):
        Gen3cPipeline(inference_type="video2world", checkpoint_dir="/nonexistent", checkpoint_name="Gen3C-Cosmos-7B",
                      enable_prompt_upsampler=False, disable_guardrail=True, disable_prompt_encoder=True, net_kwargs=NET)


def test_checkpoint_round_trip_through_model_pt(tmp_path):
    """A model.pt in the reference's layout ({"model": {net.*, conditioner.*, net...._extra_state}}) loads into the
    engine's network; the forward equals the source network's; stale derived copies are refreshed by the reload."""
    from gen3c_b200 import inference_utils as iu
    from gen3c_b200.dit import VideoExtendGeneralDIT
    from gen3c_b200.model_gen3c import DiffusionGen3CModel

    torch.manual_seed(0)
    src = VideoExtendGeneralDIT(**NET)
    with torch.no_grad():
        for k, p in src.state_dict(keep_vars=True).items():
            if k != "pos_embedder.seq":
                p.copy_((0.05 * torch.randn(p.shape, device=p.device)).to(p.dtype) + (1.0 if p.dim() == 1 else 0.0))
    ck = {"net." + k: v.cpu() for k, v in src.state_dict().items()}
    ck["net.blocks.block0.blocks.0.block.attn.attn_op._extra_state"] = torch.zeros(1)       # TE FP8 blob
    ck["conditioner.embedders.text.dummy"] = torch.zeros(3)
    ck["logvar.0.freqs"] = torch.zeros(4)
    path = str(tmp_path / "model.pt")
    torch.save({"model": ck}, path)
    dst = DiffusionGen3CModel(net=VideoExtendGeneralDIT(**NET))

    def fwd(net):
        g = torch.Generator(device="cuda").manual_seed(1)
        bf = torch.bfloat16
        return net(x=torch.randn(1, 16, 2, 16, 32, device="cuda", generator=g).to(bf),
                   timesteps=torch.tensor([0.5], device="cuda", dtype=bf),
                   crossattn_emb=torch.randn(1, 128, 1024, device="cuda", generator=g).to(bf),
                   padding_mask=torch.zeros(1, 1, 16, 32, device="cuda", dtype=bf),
                   condition_video_input_mask=torch.zeros(1, 1, 2, 16, 32, device="cuda", dtype=bf),
                   condition_video_pose=torch.randn(1, 64, 2, 16, 32, device="cuda", generator=g).to(bf))

    before = fwd(dst.net).float()           # zero weights: registers the (all-zero) derived copies in the engine
    res = iu.load_network_model(dst, path)  # in-place load_state_dict: same addresses, bumped versions
    assert res.missing_keys == [] and res.incorrect_shapes == []
    assert sorted(res.unexpected_keys) == ["conditioner.embedders.text.dummy", "logvar.0.freqs"]
    after, want = fwd(dst.net).float(), fwd(src).float()
    assert float(before.abs().max()) == 0.0 and float(want.abs().max()) > 0
    assert torch.equal(after, want)


def _persistent_args(tmp_path):
    from gen3c_b200.inference import gen3c_persistent as pm

    argv = ["--synthetic", "--prompt", "a room", "--trajectory", "none", "--video_save_name", "", "--disable_guardrail",
            "--disable_prompt_upsampler", "--disable_prompt_encoder", "--height", "128", "--width", "256", "--num_steps", "2",
            "--video_save_folder", str(tmp_path / "out"), "--seed", "5"]
    return pm, pm.create_parser().parse_args(argv)


def test_persistent_model_single_image_seed_and_cameras(tmp_path):
    """Gen3cPersistentModel (server-side caller): seed from one image by value, then serve a 121-camera request with
    estimated depths returned (reference gen3c_persistent.py:138-515)."""
    from gen3c_b200.camera_utils import generate_camera_trajectory

    pm, args = _persistent_args(tmp_path)
    model = pm.Gen3cPersistentModel(args, pipeline=_pipeline(args))
    img = np.random.RandomState(1).rand(1, 90, 160, 3).astype(np.float32)
    w2c, focal, pp, res = model.seed_model_from_values(img, None, np.eye(4, dtype=np.float32)[None], np.ones((1, 2), np.float32),
                                                       np.full((1, 2), 0.5, np.float32), np.array([[160, 90]]))
    assert w2c.shape == (1, 4, 4) and focal.shape == (1, 2) and pp.shape == (1, 2) and list(res[0]) == [256, 128]
    assert model.seeding_image.shape == (1, 3, 1, 128, 256) and model.model_was_seeded
    K = np.array([[focal[0, 0], 0, pp[0, 0]], [0, focal[0, 1], pp[0, 1]], [0, 0, 1]], dtype=np.float32)
    w2cs, Ks = generate_camera_trajectory("right", torch.eye(4), torch.from_numpy(K), 121, 0.2, "center_facing", device="cpu")
    out = model.inference_on_cameras(w2cs[0].numpy(), Ks[0].numpy(), fps=24, return_estimated_depths=True)
    assert out["video"].shape == (1, 121, 3, 128, 256) and out["video"].dtype == np.uint8
    assert out["rendered_warp_images"].shape == (1, 121, 1, 3, 128, 256)
    d = out["predicted_depth"]
    assert d.shape == (121, 1, 128, 256) and np.isnan(d[:-1]).all() and np.isfinite(d[-1]).all()
    assert model.get_cache_input_depths().shape == (1, 1, 128, 256)
    model.clear_cache()
    assert model.cache is None and not model.model_was_seeded
    with pytest.raises(AssertionError):   # persistent mode takes images by value
        bad = pm.create_parser().parse_args(["--prompt", "x", "--trajectory", "none", "--video_save_name", "",
                                             "--input_image_path", "a.png"])
        pm.validate_args(bad)


def test_persistent_model_multiframe_seed_uses_cache4d(tmp_path):
    from gen3c_b200.cache_3d import Cache4D

    pm, args = _persistent_args(tmp_path)
    model = pm.Gen3cPersistentModel(args, pipeline=_pipeline(args))
    n, h, w = 121, 128, 256
    rs = np.random.RandomState(2)
    imgs = rs.rand(n, h, w, 3).astype(np.float32)
    depths = (2.0 + rs.rand(n, h, w)).astype(np.float32)
    masks = np.ones((n, h, w), np.float32)
    w2cs = np.tile(np.eye(4, dtype=np.float32)[None], (n, 1, 1))
    w2cs[:, 0, 3] = np.linspace(0, 0.1, n)
    focal = np.full((n, 2), 220.0, np.float32)
    pp = np.full((n, 2), 0.5, np.float32)
    model.seed_model_from_values(imgs, depths, w2cs, focal, pp, np.tile([[w, h]], (n, 1)), masks)
    assert isinstance(model.cache, Cache4D) and model.cache.input_frame_count() == n
    K = np.zeros((n, 3, 3), np.float32)
    K[:, 0, 0] = K[:, 1, 1] = 220.0
    K[:, 0, 2], K[:, 1, 2], K[:, 2, 2] = 0.5 * w, 0.5 * h, 1.0
    out = model.inference_on_cameras(w2cs, K, fps=24)
    assert out["video"].shape == (1, n, 3, h, w) and out["predicted_depth"] is None
    with pytest.raises(NotImplementedError):
        model.seed_model_from_values(imgs[:2], None, w2cs[:2], focal[:2], pp[:2], np.tile([[w, h]], (2, 1)))

"""Denoise-loop harness: restated DDIM schedule (CPU) and the 25-step cosine bar on the GPU
(north_star: final 25-step DDIM latents cosine >= 0.999 against the reference arithmetic for a fixed seed)."""
import pytest
import torch

from neurons_b200 import sampler
from oracle import motion_oracle as mo
from tests import helpers


def test_ddim_timesteps_and_identities():
    sch = sampler.DDIMSchedule()
    ts = sch.timesteps(25)
    assert ts[0] == 961 and ts[-1] == 1 and len(ts) == 25 and ts[0] - ts[1] == 40       # 961, 921, ..., 1
    g = torch.Generator().manual_seed(0)
    x0, n = torch.randn(2, 4, 3, 4, 4, generator=g), torch.randn(2, 4, 3, 4, 4, generator=g)
    xt = sch.add_noise(x0, n, 961)
    # with the true noise as the prediction one DDIM step lands exactly on the same x0 / noise mix at the previous timestep
    assert torch.allclose(sch.step(n, 961, xt, 25), sch.add_noise(x0, n, 921), atol=1e-5)
    # last step goes to alpha = 1: returns x0
    assert torch.allclose(sch.step(n, 1, sch.add_noise(x0, n, 1), 25), x0, atol=1e-5)


def test_denoise_loop_cfg_and_order():
    calls = []

    def den(x2, t, ctx):
        calls.append((t, x2.shape[0]))
        return 0.1 * x2
    lat = torch.ones(1, 4, 2, 2, 2)
    out = sampler.denoise(den, lat, None, sampler.DDIMSchedule(), num_inference_steps=5, guidance_scale=8.5)
    assert [c[0] for c in calls] == sampler.DDIMSchedule().timesteps(5) and all(c[1] == 2 for c in calls)     # CFG doubles the batch
    assert out.shape == lat.shape and torch.isfinite(out).all()


def _stack_and_params(channels, seed):
    stack = sampler.MotionStack(channels, seed=seed)
    cfgs, params = [], []
    for i, c in enumerate(stack.module_channels()):
        cfg = mo.MotionConfig(c, 8, 1, 2, True, 24)
        cfgs.append(cfg)
        params.append({k: helpers.round_bf16(v) for k, v in mo.make_params(cfg, 100 + i).items()})
    return stack, cfgs, params


@pytest.mark.gpu
@pytest.mark.timeout(900)
def test_ddim_25_step_cosine_vs_oracle():
    dev = "cuda:0"
    stack, cfgs, params = _stack_and_params((64, 128), seed=3)
    mods = [helpers.mirror_module(c, p, dev, torch.bfloat16) for c, p in zip(cfgs, params)]
    g = torch.Generator().manual_seed(7)
    lat0 = torch.randn(1, 4, 8, 8, 8, generator=g)
    noise = torch.randn(1, 4, 8, 8, 8, generator=g)
    ctx = torch.randn(2, 16, generator=g)                      # uncond | cond "text" embeddings
    sch = sampler.DDIMSchedule()

    def den_gpu(x2, t, c):
        return stack(lambda i, h: mods[i](h.to(torch.bfloat16), None, None).float(), x2, t, c)

    def den_ref(x2, t, c):      # reference arithmetic in fp32 on the same bf16-rounded weights
        return stack(lambda i, h: mo.forward_reference_order(params[i], h, cfgs[i]), x2, t, c)

    out_gpu = sampler.denoise(den_gpu, lat0.to(dev), ctx.to(dev), sch, 25, 8.5, noise=noise.to(dev), low_strength=0.3).float().cpu()
    out_ref = sampler.denoise(den_ref, lat0, ctx, sch, 25, 8.5, noise=noise, low_strength=0.3)
    assert torch.isfinite(out_gpu).all() and torch.isfinite(out_ref).all()
    cos = torch.nn.functional.cosine_similarity(out_gpu.flatten(), out_ref.flatten(), dim=0).item()
    assert cos >= 0.999, cos


@pytest.mark.gpu
@pytest.mark.timeout(1500)
def test_ddim_25_step_cosine_at_unet_channel_widths():
    """The same bar on a stack at the UNet's real widths -- 320 / 640 / 1280 channels, 10 motion modules (the fused C = 320 kernel, the
    d_h = 80 fused QKV + attention kernel and the d_h = 160 stand-alone attention path all take part), CFG batch 2, 8 frames, 8x8 latent --
    against the reference arithmetic (oracle, fp32 on the same bf16-rounded weights) through all 25 DDIM steps."""
    dev = "cuda:0"
    stack, cfgs, params = _stack_and_params((320, 640, 1280), seed=5)
    mods = [helpers.mirror_module(c, p, dev, torch.bfloat16) for c, p in zip(cfgs, params)]
    g = torch.Generator().manual_seed(17)
    lat0 = torch.randn(1, 4, 8, 8, 8, generator=g)
    noise = torch.randn(1, 4, 8, 8, 8, generator=g)
    ctx = torch.randn(2, 16, generator=g)
    sch = sampler.DDIMSchedule()

    def den_gpu(x2, t, c):
        return stack(lambda i, h: mods[i](h.to(torch.bfloat16), None, None).float(), x2, t, c)

    def den_ref(x2, t, c):
        return stack(lambda i, h: mo.forward_reference_order(params[i], h, cfgs[i]), x2, t, c)

    out_gpu = sampler.denoise(den_gpu, lat0.to(dev), ctx.to(dev), sch, 25, 8.5, noise=noise.to(dev)).float().cpu()
    out_ref = sampler.denoise(den_ref, lat0, ctx, sch, 25, 8.5, noise=noise)
    assert torch.isfinite(out_gpu).all() and torch.isfinite(out_ref).all()
    cos = torch.nn.functional.cosine_similarity(out_gpu.flatten(), out_ref.flatten(), dim=0).item()
    assert cos >= 0.999, cos


@pytest.mark.gpu
@pytest.mark.timeout(1500)
def test_ddim_25_step_cosine_spatial_plus_motion_blocks():
    """The call order of the UNet's CrossAttn blocks (unet_blocks.py:409-411) through the whole loop: every stage of the stack is a spatial
    Transformer3DModel (self-attention over h*w, text cross-attention, GEGLU) followed by the motion module that consumes
    its [B,F,C,H,W]-storage view, at 320 / 640 channels, CFG batch 2, 8 frames, 16x16 latent (so the d_h = 40 self-attention runs on the
    tcgen05 kernel at the first level: 256 keys), bf16 on the GPU against the reference arithmetic of both modules (oracles, fp32 on the
    same bf16-rounded weights) over all 25 DDIM steps."""
    import neurons_b200 as nb
    from oracle import spatial_oracle as so
    dev = "cuda:0"
    stack, cfgs, params = _stack_and_params((320, 640), seed=9)
    mods = [helpers.mirror_module(c, p, dev, torch.bfloat16) for c, p in zip(cfgs, params)]
    scfgs = [so.SpatialConfig(c.channels, 8, 1, 768, True) for c in cfgs]
    sparams = [{k: helpers.round_bf16(v) for k, v in so.make_params(sc, 40 + i).items()} for i, sc in enumerate(scfgs)]
    smods = []
    for sc, sp in zip(scfgs, sparams):
        m = nb.Transformer3DModel(num_attention_heads=8, attention_head_dim=sc.head_dim, in_channels=sc.channels, cross_attention_dim=768,
                                  unet_use_cross_frame_attention=False, unet_use_temporal_attention=False)
        m.load_state_dict(sp, strict=True)
        smods.append(m.eval().to(dev).to(torch.bfloat16))
    g = torch.Generator().manual_seed(23)
    lat0 = torch.randn(1, 4, 8, 16, 16, generator=g)
    noise = torch.randn(1, 4, 8, 16, 16, generator=g)
    ctx = torch.randn(2, 16, generator=g)
    ehs = helpers.round_bf16(torch.randn(2, 77, 768, generator=g))          # uncond | cond text states
    ehs_dev = ehs.to(dev).to(torch.bfloat16)
    sch = sampler.DDIMSchedule()

    def den_gpu(x2, t, c):
        def block(i, h):
            y = smods[i](h.to(torch.bfloat16), encoder_hidden_states=ehs_dev).sample
            return mods[i](y, None, None).float()
        return stack(block, x2, t, c)

    def den_ref(x2, t, c):
        def block(i, h):
            y = so.forward_reference_order(sparams[i], h, ehs, scfgs[i])
            return mo.forward_reference_order(params[i], y, cfgs[i])
        return stack(block, x2, t, c)

    out_gpu = sampler.denoise(den_gpu, lat0.to(dev), ctx.to(dev), sch, 25, 8.5, noise=noise.to(dev)).float().cpu()
    out_ref = sampler.denoise(den_ref, lat0, ctx, sch, 25, 8.5, noise=noise)
    assert torch.isfinite(out_gpu).all() and torch.isfinite(out_ref).all()
    cos = torch.nn.functional.cosine_similarity(out_gpu.flatten(), out_ref.flatten(), dim=0).item()
    assert cos >= 0.999, cos


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("n", [4 * 16 * 32 * 32, 1003, 8])
@pytest.mark.parametrize("with_cfg", [True, False])
def test_fused_cfg_ddim_step_matches_schedule_step(dtype, n, with_cfg):
    """nmm_cfg_ddim_step (one elementwise kernel) against DDIMSchedule.step + the CFG combine in fp64."""
    from neurons_b200 import ops
    dev = "cuda:0"
    g = torch.Generator().manual_seed(n)
    x, eu, ec = (torch.randn(n, generator=g).to(dtype) for _ in range(3))
    sch = sampler.DDIMSchedule()
    for t in (961, 481, 1):
        a_t, a_prev = sch.alphas(t, 25)
        eps = eu.double() + 8.5 * (ec.double() - eu.double()) if with_cfg else eu.double()
        ref = sch.step(eps, t, x.double(), 25)
        out = ops.cfg_ddim_step(x.to(dev).clone(), eu.to(dev), ec.to(dev) if with_cfg else None, 8.5, a_t, a_prev)
        scale = ref.abs().max().item()
        tol = (2e-6 if dtype == torch.float32 else 2 ** -8 * 1.01) * scale
        assert (out.double().cpu() - ref).abs().max().item() <= tol


@pytest.mark.gpu
def test_denoise_fused_step_matches_unfused_and_keeps_input():
    from neurons_b200 import ops  # noqa: F401
    dev = "cuda:0"
    g = torch.Generator().manual_seed(11)
    lat0 = torch.randn(1, 4, 8, 8, 8, generator=g).to(dev)
    keep = lat0.clone()
    w = torch.randn(4, 4, generator=g).to(dev) * 0.3

    def den(x2, t, c):
        return torch.einsum("oc,bcfhw->bofhw", w, x2) * (t / 1000.0)

    sch = sampler.DDIMSchedule()
    a = sampler.denoise(den, lat0, None, sch, 25, 8.5)
    b = sampler.denoise(den, lat0, None, sch, 25, 8.5, fused_step=True)
    assert torch.equal(lat0, keep)
    assert (a - b).abs().max().item() <= 1e-5 * a.abs().max().item()


def _reference_next_step():
    """The reference's own DDIM update, animatediff/utils/util.py:211-222 (`next_step`, the inversion direction of the same eta = 0
    update), compiled from the reference file where it lies (util.py itself imports packages this image lacks).  None off-box."""
    import ast
    from oracle import ref_shim
    root = ref_shim.reference_root()
    if root is None:
        return None
    import os
    path = os.path.join(root, "animatediff", "utils", "util.py")
    if not os.path.isfile(path):
        return None
    tree = ast.parse(open(path).read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "next_step")
    import numpy as np
    from typing import Union
    ns = {"torch": torch, "np": np, "Union": Union}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), ns)
    return ns["next_step"]


@pytest.mark.skipif(_reference_next_step() is None, reason="reference tree not mounted")
def test_ddim_update_pinned_to_reference_next_step():
    """PIN for the scheduler piece of row N2: diffusers' DDIMScheduler is not installable here, but the reference carries the same
    eta = 0 update in its own tree (util.py:211-222, used for DDIM inversion): x(t - d) -> x(t) with the alphas_cumprod indexing and
    the final_alpha_cumprod convention of the scheduler object it is handed.  Our forward step t -> t - d followed by the reference's
    step t - d -> t with the same epsilon must return the input, for every timestep of the 25-step schedule (incl. the last, which
    uses final_alpha_cumprod), and the reference step run on OUR schedule object must equal our formula with the alphas exchanged."""
    from types import SimpleNamespace
    next_step = _reference_next_step()
    sch = sampler.DDIMSchedule()
    n = 25
    fake = SimpleNamespace(config=SimpleNamespace(num_train_timesteps=sch.num_train_timesteps), num_inference_steps=n,
                           alphas_cumprod=sch.alphas_cumprod, final_alpha_cumprod=torch.tensor(sch.final_alpha_cumprod, dtype=torch.float64))
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 4, 3, 5, 5, generator=g, dtype=torch.float64)
    eps = torch.randn(2, 4, 3, 5, 5, generator=g, dtype=torch.float64)
    ts = sch.timesteps(n)
    assert ts[0] == 961 and ts[-1] == 1 and len(ts) == n
    for t in ts:
        x_prev = sch.step(eps, t, x, n)
        back = next_step(eps, t, x_prev, fake)                       # reference: (t - 40) -> t
        assert (back - x).abs().max().item() <= 1e-12, t
        a_t, a_prev = sch.alphas(t, n)
        direct = (a_t ** 0.5) * (x - ((1 - a_prev) ** 0.5) * eps) / (a_prev ** 0.5) + ((1 - a_t) ** 0.5) * eps
        assert (next_step(eps, t, x, fake) - direct).abs().max().item() <= 1e-12, t

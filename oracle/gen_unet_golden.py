"""TEST INFRASTRUCTURE ONLY -- capture the motion modules' inputs INSIDE one reference UNet / SparseControlNet step.

Run in the build container (needs /root/reference):   python -m oracle.gen_unet_golden
Writes tests/golden/unet_step_inputs.pt; it travels to the GPU box (the reference tree does not).

What it does (VERDICT r1 "parity breadth"):
  * builds the UNMODIFIED reference UNet3DConditionModel (SD1.5 topology, inference-v3.yaml kwargs) and SparseControlNetModel
    (sparsectrl/latent_condition.yaml kwargs) through oracle/unet_shim.py, random init (seed 0);
  * gives every VanillaTemporalModule the name-keyed synthetic weights of oracle.motion_oracle.make_params(cfg, seed) (bf16-rounded;
    proj_out not zero), so that the GPU box can rebuild them from the seed alone;
  * runs ONE denoising-step forward at a small latent (CFG batch 2, 8 frames, 8x8 latent; timestep 961) with forward hooks on the
    20 + 8 motion modules and records each module's INPUT (the activation the surrounding resnet / spatial transformer hands over,
    in the [B,F,C,H,W]-storage view the call site passes) rounded to bf16;
  * checks, call by call, that oracle.forward_reference_order on (recorded input, seeded weights) reproduces the hooked OUTPUT of the
    reference module (fp32; <= 2e-5) -- the pin -- and stores only the inputs + per-call statistics.
The GPU test (tests/test_unet_activations.py) feeds each recorded input to the CUDA path and compares with the oracle: the kernels see the
heavy-tailed activation statistics of a real UNet position instead of N(0, 1).
"""
from __future__ import annotations

import os
import sys

import torch

from . import motion_oracle as mo
from . import unet_shim

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "unet_step_inputs.pt")
LATENT, FRAMES, BATCH, SEED0 = 8, 8, 2, 100


def motion_modules(model):
    return [(n, m) for n, m in model.named_modules() if type(m).__name__ == "VanillaTemporalModule"]


def cfg_of(m) -> mo.MotionConfig:
    tt = m.temporal_transformer
    blk = tt.transformer_blocks[0]
    pe = blk.attention_blocks[0].pos_encoder
    return mo.MotionConfig(tt.norm.num_channels, blk.attention_blocks[0].heads, len(tt.transformer_blocks), len(blk.attention_blocks), pe is not None,
                           int(pe.pe.shape[1]) if pe is not None else 0)


def seed_weights(model, seed0):
    out = []
    for i, (name, m) in enumerate(motion_modules(model)):
        cfg = cfg_of(m)
        params = {k: v.to(torch.bfloat16).float() for k, v in mo.make_params(cfg, seed0 + i).items()}
        missing, unexpected = m.load_state_dict(params, strict=False)
        assert not unexpected and all(k.endswith("pos_encoder.pe") for k in missing), (missing, unexpected)
        out.append((name, cfg, params))
    return out


def capture(model, run, seed0, tag):
    mods = seed_weights(model, seed0)
    rec = {}
    hooks = []
    for name, m in motion_modules(model):
        def pre(mod, args, kwargs, name=name):
            x = args[0] if args else kwargs["input_tensor"]
            xb = x.detach().to(torch.bfloat16)
            rec[name] = {"x": xb, "stride": tuple(x.stride())}
            # feed the module the bf16-rounded activation, so that (recorded input, seeded weights) -> output is exactly reproducible
            return ((xb.float().as_strided(x.shape, x.stride()) if False else xb.float().reshape(x.shape),) + tuple(args[1:]), kwargs) if args else None
        def post(mod, args, out, name=name):
            rec[name]["y"] = out.detach()
        hooks.append(m.register_forward_pre_hook(pre, with_kwargs=True))
        hooks.append(m.register_forward_hook(post))
    with torch.no_grad():
        run(model)
    for h in hooks:
        h.remove()
    calls = []
    for i, (name, cfg, params) in enumerate(mods):
        r = rec[name]
        x = r["x"].float()
        ref = mo.forward_reference_order(params, x, cfg)
        err = (ref - r["y"]).abs().max().item()
        assert err <= 2e-5, f"{tag} {name}: oracle vs hooked reference output {err}"
        xa = x.abs()
        calls.append(dict(model=tag, name=name, index=i, seed=seed0 + i, channels=cfg.channels, attn_blocks=cfg.attn_blocks, max_len=cfg.max_len,
                          x=r["x"].contiguous(), x_absmax=float(xa.max()), x_std=float(x.std()), y_absmax=float(r["y"].abs().max()),
                          x_kurtosis=float(((x - x.mean()) ** 4).mean() / x.var() ** 2), oracle_vs_reference=err))
        print(f"{tag:10s} {i:2d} C={cfg.channels:4d} A={cfg.attn_blocks} x{tuple(x.shape)} |x|max {calls[-1]['x_absmax']:7.2f} std {calls[-1]['x_std']:5.2f} "
              f"kurt {calls[-1]['x_kurtosis']:6.1f} |y|max {calls[-1]['y_absmax']:7.2f}  oracle-vs-reference {err:.1e}", flush=True)
    return calls


def main():
    if not unet_shim.available():
        sys.exit("reference tree not found")
    torch.manual_seed(0)
    g = torch.Generator().manual_seed(7)
    sample = torch.randn(BATCH, 4, FRAMES, LATENT, LATENT, generator=g)
    ctx = torch.randn(BATCH, 77, 768, generator=g)
    t = torch.tensor([961])
    unet = unet_shim.build_unet(0)
    calls = capture(unet, lambda m: m(sample, t, ctx), SEED0, "unet")
    del unet
    cn = unet_shim.build_controlnet(0)
    cond = torch.randn(BATCH, 4, FRAMES, LATENT, LATENT, generator=g)
    mask = torch.zeros(BATCH, 1, FRAMES, LATENT, LATENT)
    mask[:, :, 0] = 1.0                      # keyframe conditioning at frame 0 (scripts/neuroclips_video_enhance.py)
    calls += capture(cn, lambda m: m(sample, t, encoder_hidden_states=ctx, controlnet_cond=cond, conditioning_mask=mask, return_dict=False),
                     SEED0 + 50, "controlnet")
    torch.save({"meta": dict(latent=LATENT, frames=FRAMES, batch=BATCH, timestep=961, note="inputs of the 20 + 8 motion modules inside one reference step"),
                "calls": calls}, OUT)
    print("wrote", OUT, os.path.getsize(OUT) / 1e6, "MB")


if __name__ == "__main__":
    main()

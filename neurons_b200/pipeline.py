"""Stage-5 video-enhancement harness around the hot path: the per-clip inputs the reference reads, the SparseCtrl
ControlNet residual plumbing of its denoising loop, and the clip loop sharded over ranks.

Mirrors (host-side glue only; every model stays whatever the caller passes -- in NEURONS the reference UNet3DConditionModel /
SparseControlNetModel with `neurons_b200.patch()` / `patch_spatial()` applied):
    scripts/neuroclips_video_enhance.py:39-40     get_original_index: clip i of rank r is original clip r + i * world
    scripts/neuroclips_video_enhance.py:47-57     cccat: 6 blurry frames -> 16 frames (2 blends between neighbours)
    scripts/neuroclips_video_enhance.py:171-191   stage-3 inputs: video_subj0{S}_all_recons.pt (key frames), recon_videos.pt (blurry
                                                  videos [1200, 6, 3, 224, 224]), pred_test_caption(_self).pt (captions)
    animatediff/pipelines/pipeline_neuroclips.py:433-483   the loop: per step ControlNet(latent_model_input, t, ctx, cond, mask) ->
                                                  (down residuals, mid residual) -> UNet(..., down_block_additional_residuals=,
                                                  mid_block_additional_residual=) -> CFG -> DDIM step
    pipeline_neuroclips.py:445-466                controlnet_cond / conditioning_mask: zeros over all frames, the key-frame latents and a
                                                  mask of ones at `controlnet_image_index`
The VAE, CLIP text encoder and GIF writer are outside the path and are passed in as callables / skipped.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Callable, Iterator, List, Optional, Sequence, Tuple

import torch

from . import sampler
from .sharding import shard_indices


def cccat(a: torch.Tensor) -> torch.Tensor:
    """[b, T, ...] -> [b, 3 T - 2, ...]: frame 0, then for every neighbour pair the 0.67/0.33 and 0.33/0.67 blends and the next frame
    (scripts/neuroclips_video_enhance.py:47-57): 6 blurry frames become the 16 frames the motion modules see."""
    out = [a[:, 0].unsqueeze(1)]
    for i in range(a.size(1) - 1):
        out.append((0.67 * a[:, i] + 0.33 * a[:, i + 1]).unsqueeze(1))
        out.append((0.33 * a[:, i] + 0.67 * a[:, i + 1]).unsqueeze(1))
        out.append(a[:, i + 1].unsqueeze(1))
    return torch.cat(out, dim=1)


@dataclass
class Stage3Inputs:
    """What stage 3 leaves for the enhancement stage (scripts/neuroclips_video_enhance.py:171-191)."""
    keyframes: torch.Tensor         # video_subj0{S}_all_recons.pt   [clips, 3, H, W] in [0, 1]
    blurry: torch.Tensor            # recon_videos.pt                [clips, 6, 3, 224, 224] in [0, 1]
    captions: Sequence[str]         # pred_test_caption(.pt | _self.pt)

    @staticmethod
    def load(outdir: str, subj: int = 1, self_captions: bool = False, clips: int = 1200) -> "Stage3Inputs":
        key = torch.load(os.path.join(outdir, f"video_subj0{subj}_all_recons.pt"), map_location="cpu")
        blurry = torch.load(os.path.join(outdir, "recon_videos.pt"), map_location="cpu")
        if blurry.numel() != clips * 6 * 3 * blurry.shape[-2] * blurry.shape[-1]:
            raise ValueError(f"recon_videos.pt holds {tuple(blurry.shape)}, not {clips} clips of 6 RGB frames")
        blurry = blurry.reshape(clips, 6, 3, blurry.shape[-2], blurry.shape[-1]).float()          # :181
        caps = torch.load(os.path.join(outdir, "pred_test_caption_self.pt" if self_captions else "pred_test_caption.pt"), map_location="cpu",
                          weights_only=False)
        if len(key) != clips or len(caps) != clips:
            raise ValueError(f"stage-3 inputs disagree on the clip count: {len(key)} key frames, {len(caps)} captions, expected {clips}")
        return Stage3Inputs(key, blurry, caps)

    def __len__(self) -> int:
        return len(self.keyframes)

    def shard(self, rank: int, world: int) -> Iterator[Tuple[int, torch.Tensor, torch.Tensor, str]]:
        """(original index, key frame [3,H,W], blurry clip [6,3,h,w], caption) of this rank's clips, in the reference's order."""
        for idx in shard_indices(len(self), rank, world):
            yield idx, self.keyframes[idx], self.blurry[idx], str(self.captions[idx])


def controlnet_condition(controlnet_images: torch.Tensor, image_index: Sequence[int], video_length: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """pipeline_neuroclips.py:445-463.  controlnet_images [b, c, n, h, w] (key-frame latents) -> (controlnet_cond [b, c, L, h, w] with the
    images at `image_index` and zeros elsewhere, conditioning_mask [b, 1, L, h, w] with ones at `image_index`)."""
    assert controlnet_images.dim() == 5                                            # :446
    assert controlnet_images.shape[2] >= len(image_index)                          # :461
    shape = list(controlnet_images.shape)
    shape[2] = video_length
    cond = torch.zeros(shape, device=controlnet_images.device, dtype=controlnet_images.dtype)
    mshape = list(shape)
    mshape[1] = 1
    mask = torch.zeros(mshape, device=controlnet_images.device, dtype=controlnet_images.dtype)
    idx = list(image_index)
    cond[:, :, idx] = controlnet_images[:, :, :len(idx)]
    mask[:, :, idx] = 1
    return cond, mask


class NeuroclipsDenoiser:
    """The noise predictor of one loop iteration (pipeline_neuroclips.py:439-475) as the callable `sampler.denoise` drives:
    eps = UNet(x2, t, ctx, down/mid residuals from the SparseCtrl ControlNet run on the same x2, t, ctx)."""

    def __init__(self, unet, controlnet=None, controlnet_images: Optional[torch.Tensor] = None, controlnet_image_index: Sequence[int] = (0,),
                 conditioning_scale: float = 1.0):
        self.unet, self.controlnet = unet, controlnet
        self.images, self.index, self.scale = controlnet_images, tuple(controlnet_image_index), conditioning_scale
        self._cond = None

    def __call__(self, latent_model_input: torch.Tensor, t: int, text_embeddings: torch.Tensor) -> torch.Tensor:
        down = mid = None
        if self.controlnet is not None and self.images is not None:               # :441
            if self._cond is None or self._cond[0].shape[2] != latent_model_input.shape[2]:
                self._cond = controlnet_condition(self.images.to(latent_model_input.device), self.index, latent_model_input.shape[2])
            cond, mask = self._cond
            down, mid = self.controlnet(latent_model_input, t, encoder_hidden_states=text_embeddings, controlnet_cond=cond,
                                        conditioning_mask=mask, conditioning_scale=self.scale, guess_mode=False, return_dict=False)   # :464-471
        out = self.unet(latent_model_input, t, encoder_hidden_states=text_embeddings, down_block_additional_residuals=down,
                        mid_block_additional_residual=mid)                        # :469-475
        return out.sample if hasattr(out, "sample") else out[0]


def enhance_clip(denoiser: Callable, latents: torch.Tensor, text_embeddings: torch.Tensor, schedule: Optional[sampler.DDIMSchedule] = None,
                 num_inference_steps: int = 25, guidance_scale: float = 8.5, low_strength: float = 0.3, seed: int = 0,
                 fused_step: bool = False) -> torch.Tensor:
    """One clip of the enhancement stage: the VAE latents of the interpolated blurry video are noised to the `low_strength` start
    timestep (pipeline_neuroclips.py:410-423, seed as scripts/neuroclips_video_enhance.py:284-288) and denoised over ALL timesteps
    (:433) with classifier-free guidance.  text_embeddings = cat([uncond, cond]) as the pipeline builds them."""
    schedule = schedule or sampler.DDIMSchedule()
    g = torch.Generator(device=latents.device).manual_seed(seed)
    noise = torch.randn(latents.shape, generator=g, device=latents.device, dtype=latents.dtype)
    return sampler.denoise(denoiser, latents, text_embeddings, schedule, num_inference_steps, guidance_scale, noise=noise,
                           low_strength=low_strength, fused_step=fused_step)


def enhance_shard(inputs: Stage3Inputs, rank: int, world: int, encode_latents: Callable[[torch.Tensor], torch.Tensor],
                  encode_text: Callable[[str], torch.Tensor], make_denoiser: Callable[[torch.Tensor], Callable], **kw) -> List[Tuple[int, torch.Tensor]]:
    """This rank's clips through `enhance_clip` (the loop of scripts/neuroclips_video_enhance.py:246-327 without file output).
    encode_latents: frames [n, 3, H, W] in [0, 1] -> latents [n, 4, h, w] (the reference: vae.encode(2 x - 1).sample() * 0.18215);
    encode_text: caption -> cat([uncond, cond]) embeddings [2, 77, 768]; make_denoiser(key-frame latents [1, 4, 1, h, w]) -> callable.
    Returns [(original clip index, final latents [1, 4, 16, h, w])]; decode / gather (sharding.gather_clips) is the caller's."""
    out = []
    for idx, key, blurry, caption in inputs.shard(rank, world):
        motion = cccat(blurry.unsqueeze(0))[0]                                     # [16, 3, H, W]
        lat = encode_latents(motion).unsqueeze(0).permute(0, 2, 1, 3, 4)           # "(b f) c h w -> b c f h w"
        key_lat = encode_latents(key.unsqueeze(0)).unsqueeze(0).permute(0, 2, 1, 3, 4)
        out.append((idx, enhance_clip(make_denoiser(key_lat), lat, encode_text(caption), **kw)))
    return out

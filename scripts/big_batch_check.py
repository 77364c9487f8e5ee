"""Development check: BASELINE configs[3] sizes (CFG batch 8, 16 frames) through the module with and without the fused kernels."""
import os, sys, torch
sys.path.insert(0, os.getcwd())
import neurons_b200 as nb
dev = torch.device("cuda", 0)
kw = dict(num_attention_heads=8, num_transformer_block=1, attention_block_types=("Temporal_Self", "Temporal_Self"), temporal_position_encoding=True,
          temporal_position_encoding_max_len=24, temporal_attention_dim_div=1, zero_initialize=False)
torch.manual_seed(0)
for C, side in [(320, 32), (640, 16), (1280, 8), (1280, 4), (320, 64)]:
    with torch.device(dev):
        m = nb.get_motion_module(C, "Vanilla", kw).to(torch.bfloat16).eval()
    x = torch.randn(8, 16, C, side, side, device=dev, dtype=torch.bfloat16).permute(0, 2, 1, 3, 4)
    with torch.no_grad():
        y1 = m(x, None, None).float()
        os.environ["NMM_NO_ATTN_FUSE"] = "1"; os.environ["NMM_NO_GN_FUSE"] = "1"
        y2 = m(x, None, None).float()
        del os.environ["NMM_NO_ATTN_FUSE"]; del os.environ["NMM_NO_GN_FUSE"]
    torch.cuda.synchronize()
    print(C, side, tuple(y1.shape), "finite", bool(torch.isfinite(y1).all()), "max|y|", round(y1.abs().max().item(), 3), "fused-vs-plain", round((y1 - y2).abs().max().item(), 4))

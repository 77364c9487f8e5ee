import sys, torch
sys.path.insert(0, '/root/repo')
import neurons_b200 as nb
from neurons_b200 import lib as nlib
kw = dict(num_attention_heads=8, num_transformer_block=1, attention_block_types=("Temporal_Self", "Temporal_Self"),
          temporal_position_encoding=True, temporal_position_encoding_max_len=24, temporal_attention_dim_div=1, zero_initialize=False)
dev = torch.device('cuda', 0)
for C, side in ((320, 64), (640, 32), (1280, 16)):
    with torch.no_grad():
        with torch.device(dev):
            m = nb.get_motion_module(C, "Vanilla", kw).eval()
        x = torch.randn(1, 8, C, side, side, device=dev).permute(0, 2, 1, 3, 4)
        m(x, None, None); m(x, None, None)
        torch.cuda.synchronize()
        nlib.profile_begin(); m(x, None, None); p = nlib.profile_end()
    print(C, side, {k: (v['launches'], round(v['total_ms'] * 1e3, 1)) for k, v in p.items() if v['launches']})

"""Short single-GPU workload for ncu captures (not a test): one render call, no counters."""
import sys
sys.path.insert(0, '.')
from tests.conftest import load_agpt
agpt = load_agpt()
cfg = int(sys.argv[1]); level = int(sys.argv[2]); spp = int(sys.argv[3]); flags = int(sys.argv[4]) if len(sys.argv) > 4 else 0
d = agpt.config_defaults(cfg)
hs = agpt.HostScene(cfg, level); ctx = agpt.Context(0); hs.upload(ctx)
ctx.set_film(d['width'], d['height'])
ctx.render(0, spp, d['max_depth'], d['depth_arg'], flags)
s = ctx.stats(); print('ms', s.ms_render, 'Mrays/s', s.rays / s.ms_render / 1e3)

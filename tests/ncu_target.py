"""Short single-GPU workload for ncu captures (not a test): one agpt_render call of a configuration --
with `16` samples per pixel on cfg 3 this is exactly one bench.py step.  Also writes the ray counts of the
kernel classes beside the capture (gpurun_out/ncu_target_stats.json), for profiles/ncu_step_summary.py."""
import json
import os
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
os.makedirs('gpurun_out', exist_ok=True)
json.dump({"config": cfg, "level": level, "spp": spp, "width": d['width'], "height": d['height'], "paths": s.paths,
           "rays_closest_kernel": s.rays_closest + s.rays_mis, "rays_any_kernel": s.rays_shadow, "waves": s.waves,
           "launches_closest": s.launches_closest, "launches_any": s.launches_any, "launches_shade": s.launches_shade, "ms_render": s.ms_render},
          open(os.environ.get('AGPT_NCU_STATS', 'gpurun_out/ncu_target_stats.json'), 'w'))

"""Ad-hoc perf probe (not a test): time agpt_render on a configuration and print stats."""
import sys, time, json
import numpy as np
sys.path.insert(0, '.')
from tests.conftest import load_agpt
agpt = load_agpt()

def run(config, level, spp, flags=0, reps=2, W=None, H=None, depth=None):
    d = agpt.config_defaults(config)
    W = W or d['width']; H = H or d['height']
    depth = d['max_depth'] if depth is None else depth
    t = time.time(); hs = agpt.HostScene(config, level); tb = time.time() - t
    ctx = agpt.Context(0)
    t = time.time(); hs.upload(ctx); tu = time.time() - t
    ctx.set_film(W, H)
    ctx.render(0, 1, depth, d['depth_arg'], 0)   # warm-up
    for f in ([0] * reps + [agpt.FLAG_TIMING, agpt.FLAG_COUNTERS]):
        ctx.clear(); ctx.reset_stats()
        ctx.render(0, spp, depth, d['depth_arg'], f)
        s = ctx.stats()
        rays = s.rays
        out = dict(cfg=config, flags=f, ms=round(s.ms_render, 2), Mrays_s=round(rays / s.ms_render / 1e3, 1), Mpaths_s=round(s.paths / s.ms_render / 1e3, 1),
                   rays_per_path=round(rays / s.paths, 2), closest=s.rays_closest, shadow=s.rays_shadow, mis=s.rays_mis, waves=s.waves, launches=s.kernel_launches,
                   ms_closest=round(s.ms_trace_closest, 2), ms_any=round(s.ms_trace_any, 2), ms_shade=round(s.ms_shade, 2))
        if f & agpt.FLAG_COUNTERS:
            rc = s.rays_closest + s.rays_mis
            out.update(closest_visits_per_ray=round(s.node_visits[0] / rc, 2), closest_tri_per_ray=round(s.tri_tests[0] / rc, 2),
                       closest_analytic_per_ray=round(s.analytic_tests[0] / rc, 2), closest_bytes_per_ray=round(s.algorithmic_bytes(0) / rc, 1),
                       any_visits_per_ray=round(s.node_visits[1] / max(s.rays_shadow, 1), 2), any_bytes_per_ray=round(s.algorithmic_bytes(1) / max(s.rays_shadow, 1), 1))
        print(json.dumps(out), flush=True)
    print(f"build {tb:.2f}s upload {tu:.2f}s scene {ctx.scene_bytes()/1e6:.1f} MB", flush=True)
    ctx.close()

if __name__ == '__main__':
    cfg = int(sys.argv[1]); level = int(sys.argv[2]); spp = int(sys.argv[3])
    depth = int(sys.argv[4]) if len(sys.argv) > 4 else None
    run(cfg, level, spp, depth=depth)

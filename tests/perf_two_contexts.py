"""Ad-hoc probe (not a test): two contexts on ONE GPU rendering interleaved sample subsets from two host threads,
against one context rendering the same samples in sequence -- do the kernels of two independent wavefronts fill
each other's tails?"""
import sys, time, threading
sys.path.insert(0, '.')
from tests.conftest import load_agpt
agpt = load_agpt()
cfg = int(sys.argv[1]); spp = int(sys.argv[2]); nctx = int(sys.argv[3]) if len(sys.argv) > 3 else 2
d = agpt.config_defaults(cfg)
hs = agpt.HostScene(cfg, 0)
ctxs = [agpt.Context(0) for _ in range(nctx)]
for c in ctxs:
    hs.upload(c); c.set_film(d['width'], d['height']); c.render(0, 1, d['max_depth'], d['depth_arg'])     # warm-up
def one(c, first, count, stride):
    c.render(first, count, d['max_depth'], d['depth_arg'], 0, stride)
for rep in range(3):
    t0 = time.perf_counter(); one(ctxs[0], 0, spp, 1); t1 = time.perf_counter()
    ths = [threading.Thread(target=one, args=(c, g, spp // nctx, nctx)) for g, c in enumerate(ctxs)]
    t2 = time.perf_counter()
    for t in ths: t.start()
    for t in ths: t.join()
    t3 = time.perf_counter()
    print(f"cfg {cfg} {spp} spp: one context {1e3 * (t1 - t0):.1f} ms, {nctx} contexts side by side {1e3 * (t3 - t2):.1f} ms", flush=True)

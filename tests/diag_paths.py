"""Diagnostic (not a test): per-path radiance GPU vs reference, classified by size of difference."""
import sys
import numpy as np
sys.path.insert(0, '.')
from tests.conftest import load_agpt
agpt = load_agpt()
from oracle import ref_binding as ref
cfg = int(sys.argv[1]); level = int(sys.argv[2]); W = int(sys.argv[3]); H = int(sys.argv[4]); ns = int(sys.argv[5])
d = agpt.config_defaults(cfg)
hs = agpt.HostScene(cfg, level); rs = ref.RefScene(cfg, level)
ctx = agpt.Context(0); hs.upload(ctx); ctx.set_film(W, H)
ys, xs, ss = np.meshgrid(np.arange(H), np.arange(W), np.arange(ns), indexing='ij')
xs = xs.ravel().astype(np.int32); ys = ys.ravel().astype(np.int32); ss = ss.ravel().astype(np.int32)
g = ctx.li_pixels(xs, ys, ss, d['max_depth'], d['depth_arg'])
c = rs.li_pixels(W, H, xs, ys, ss, d['max_depth'], d['depth_arg'])
exact = np.all(g.view(np.uint32) == c.view(np.uint32), axis=1)
diff = np.abs(g.astype(np.float64) - c).max(axis=1)
scale = np.maximum(np.abs(c).max(axis=1), 1e-6)
rel = diff / scale
print(f"paths {len(xs)} exact {exact.mean():.4f} rel<1e-5 {(rel < 1e-5).mean():.4f} rel<1e-3 {(rel<1e-3).mean():.5f} rel>1e-2 {(rel>1e-2).mean():.6f}")
big = np.argsort(-diff)[:25]
for i in big:
    print(int(xs[i]), int(ys[i]), int(ss[i]), 'gpu', g[i], 'cpu', c[i], 'diff', diff[i])
# how much of the squared error do the big ones carry
order = np.argsort(-diff)
sq = diff[order] ** 2
print('share of squared error in top 10/100/1000 paths:', sq[:10].sum() / sq.sum(), sq[:100].sum() / sq.sum(), sq[:1000].sum() / sq.sum())
np.savez('gpurun_out/diag_paths.npz', xs=xs, ys=ys, ss=ss, g=g, c=c)

import os, sys, torch, time
sys.path.insert(0, os.getcwd())
import bench
from tensoflow_b200 import ops
dev = torch.device('cuda:0')
cfg = dict(bench.SHAPE_CFG); cfg['rays'] = 2048
field, var = bench.build_shape(cfg, dev)
from tensoflow_b200 import synthetic
rays = {k: v.to(dev) for k, v in synthetic.make_rays(cfg['rays'], seed=0).items()}
t0, t1, idx = synthetic.uniform_samples(rays['rays_o'], rays['dirs'], field.aabb, cfg['samples'])
mid = (t0+t1)*0.5
pts = rays['rays_o'][idx] + rays['dirs'][idx]*mid[:,None]
lv = torch.rand(pts.shape[0], device=dev)*3-0.5
for it in range(2):
    with torch.no_grad():
        for _ in range(2): field.stencil(pts, lv)
        torch.cuda.synchronize(); a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3): field.stencil(pts, lv)
        b.record(); torch.cuda.synchronize()
print('TF_TC_DEBUG', os.environ.get('TF_TC_DEBUG','0'), 'samples', pts.shape[0], 'ms per fwd call', a.elapsed_time(b)/3)

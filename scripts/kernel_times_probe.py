import sys; sys.path.insert(0,".")
import torch, ctypes as C, bench
from tensoflow_b200 import _lib, synthetic
dev=torch.device("cuda:0"); cfg=dict(bench.SHAPE_CFG); cfg["rays"]=2048
field,var=bench.build_shape(cfg,dev); params=list(field.parameters())+[var]; rays=synthetic.make_rays(2048,seed=50,device=dev)
lib=_lib.load()
for i in range(3):
    for p in params: p.grad=None
    bench.shape_step(field,var,rays,cfg)
torch.cuda.synchronize(); lib.tf_kernel_timing_reset(); lib.tf_kernel_timing_enable(1)
a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True); a.record()
for i in range(3):
    for p in params: p.grad=None
    bench.shape_step(field,var,rays,cfg)
b.record(); torch.cuda.synchronize()
print("step_ms", round(a.elapsed_time(b)/3,3))
for name in ("sdf_stencil_fwd_tc","sdf_stencil_bwd_tc","xty_tc","linear_tc","linear_tc_bwd"):
    t,n=C.c_double(0),C.c_int32(0); lib.tf_kernel_timing_read(name.encode(),C.byref(t),C.byref(n)); print(name, round(t.value/3,3), n.value//3)

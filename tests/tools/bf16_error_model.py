"""Error model for bf16 activation storage (CPU only, no GPU): rounds the fp64 oracle's activations and/or
gradients to bf16 at the points where the CUDA path stores them and reports how far outputs and per-variable
gradients move.  Evidence for the bf16 gradient tolerances in tests/test_gpu_parity.py ("bf16 tolerance").
    python tests/tools/bf16_error_model.py
"""
import numpy as np, torch, sys
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
from oracle import models as om, tf_ops as T
from tests import common as C
def mk(fwd_dt, bwd_dt):
    class R(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x): return x.to(fwd_dt).to(x.dtype) if fwd_dt else x
        @staticmethod
        def backward(ctx, g): return g.to(bwd_dt).to(g.dtype) if bwd_dt else g
    return R.apply
def run(cfg, x, dy, conv_r, other_r):
    o = om.create_model(cfg, torch.float64)
    w = om.init_variables(o.var_specs, 7); rng = np.random.RandomState(8)
    w = [a + rng.normal(0, 0.05, a.shape).astype(np.float32) if a.ndim == 1 else a for a in w]; o.load(w)
    names_conv = ['conv2d','conv2d_transpose']; names_other=['avg_pool2','activation','leaky_relu','upsample2','reflection_pad']
    orig = {k:getattr(T,k) for k in names_conv+names_other}; relu = torch.relu
    wrap = lambda f,r: (lambda *a, **k: r(f(*a, **k)))
    for k in names_conv: setattr(T,k,wrap(orig[k],conv_r))
    for k in names_other: setattr(T,k,wrap(orig[k],other_r))
    torch.relu = wrap(relu, other_r)
    try:
        xt = torch.from_numpy(x).double().requires_grad_(True)
        y = o.forward(other_r(xt))
        g = torch.autograd.grad(y, [xt]+o.variables, other_r(torch.from_numpy(dy).double()))
    finally:
        for k,v in orig.items(): setattr(T,k,v)
        torch.relu = relu
    return y.detach().numpy(), [t.numpy() for t in g]
bf, h = torch.bfloat16, torch.float16
idt = mk(None,None)
variants = dict(all_bf16=(mk(bf,bf),mk(bf,bf)), fwd_only=(mk(bf,None),mk(bf,None)), bwd_only=(mk(None,bf),mk(None,bf)),
   convout_fp16=(mk(h,h),mk(bf,bf)), convout_fp32=(idt,mk(bf,bf)), convout_fp32_gradsfp32=(idt, mk(bf,None)))
for name,cfg,size in (('SMALL_RESNET',C.SMALL_RESNET,32),('FIX_RESNET64',C.FIX_RESNET,64),('SMALL_UNET',C.SMALL_UNET,40)):
    rng = np.random.RandomState(5)
    x = rng.uniform(-1, 1, (2, size, size, 3)).astype(np.float32)
    yshape = tuple(om.create_model(cfg, torch.float64)(x).shape)
    dy = rng.normal(0, 1, yshape).astype(np.float32)
    y0,g0 = run(cfg,x,dy,idt,idt)
    for vn,(cr,orr) in variants.items():
        y1,g1 = run(cfg,x,dy,cr,orr)
        scale = max(np.linalg.norm(r) for r in g0[1:])
        errs = [np.linalg.norm(a-b)/max(np.linalg.norm(b),0.05*scale) for a,b in zip(g1[1:],g0[1:])]
        print(name, vn, 'y %.4f dx %.4f maxvar %.4f medvar %.4f' % (C.rel_l2(y1,y0), C.rel_l2(g1[0],g0[0]), max(errs), np.median(errs)))
